// LDPC kernel instantiations, part C (split over several translation units so that they compile in parallel).
#include "ldpc_kernels.cuh"

namespace s2 {
const Variant kLdpcVariantsC[] = {VR(17), VU(20)};
const int kLdpcVariantsC_n = (int)(sizeof(kLdpcVariantsC) / sizeof(kLdpcVariantsC[0]));
}  // namespace s2
