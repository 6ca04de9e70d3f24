// Two-lanes-per-row LDPC kernel instantiations, part A.
#include "ldpc_v2l.cuh"

namespace s2 {
const VariantL kLdpc2lVariantsA[] = {V2LU(8), V2LU(9), V2LR(11), V2LU(12), V2LU(16)};
const int kLdpc2lVariantsA_n = (int)(sizeof(kLdpc2lVariantsA) / sizeof(kLdpc2lVariantsA[0]));
}  // namespace s2
