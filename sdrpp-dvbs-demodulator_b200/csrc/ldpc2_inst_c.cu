// Second-generation LDPC kernel instantiations, part C (several translation units so that they compile in parallel).
#include "ldpc_v2.cuh"

namespace s2 {
const Variant2 kLdpc2VariantsC[] = {V2R(17), V2U(20)};
const int kLdpc2VariantsC_n = (int)(sizeof(kLdpc2VariantsC) / sizeof(kLdpc2VariantsC[0]));
}  // namespace s2
