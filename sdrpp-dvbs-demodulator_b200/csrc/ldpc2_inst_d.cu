// Second-generation LDPC kernel instantiations, part D (several translation units so that they compile in parallel).
#include "ldpc_v2.cuh"

namespace s2 {
const Variant2 kLdpc2VariantsD[] = {V2U(25), V2U(28)};
const int kLdpc2VariantsD_n = (int)(sizeof(kLdpc2VariantsD) / sizeof(kLdpc2VariantsD[0]));
}  // namespace s2
