// LDPC kernel instantiations, part D (split over several translation units so that they compile in parallel).
#include "ldpc_kernels.cuh"

namespace s2 {
const Variant kLdpcVariantsD[] = {VU(25), VU(28)};
const int kLdpcVariantsD_n = (int)(sizeof(kLdpcVariantsD) / sizeof(kLdpcVariantsD[0]));
}  // namespace s2
