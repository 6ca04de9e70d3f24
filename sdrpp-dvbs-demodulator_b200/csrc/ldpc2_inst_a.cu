// Second-generation LDPC kernel instantiations, part A (several translation units so that they compile in parallel).
#include "ldpc_v2.cuh"

namespace s2 {
const Variant2 kLdpc2VariantsA[] = {V2B(2), V2U(3), V2U(4), V2B(5), V2U(8), V2U(9)};
const int kLdpc2VariantsA_n = (int)(sizeof(kLdpc2VariantsA) / sizeof(kLdpc2VariantsA[0]));
}  // namespace s2
