// Second-generation LDPC kernel instantiations, part B (several translation units so that they compile in parallel).
#include "ldpc_v2.cuh"

namespace s2 {
const Variant2 kLdpc2VariantsB[] = {V2R(11), V2U(12), V2U(16)};
const int kLdpc2VariantsB_n = (int)(sizeof(kLdpc2VariantsB) / sizeof(kLdpc2VariantsB[0]));
}  // namespace s2
