// Two-lanes-per-row LDPC kernel instantiations, part B.
#include "ldpc_v2l.cuh"

namespace s2 {
const VariantL kLdpc2lVariantsB[] = {V2LR(17), V2LU(20), V2LU(25), V2LU(28)};
const int kLdpc2lVariantsB_n = (int)(sizeof(kLdpc2lVariantsB) / sizeof(kLdpc2lVariantsB[0]));
}  // namespace s2
