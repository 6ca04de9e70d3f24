// LDPC kernel instantiations, part B (split over several translation units so that they compile in parallel).
#include "ldpc_kernels.cuh"

namespace s2 {
const Variant kLdpcVariantsB[] = {VR(11), VU(12), VU(16)};
const int kLdpcVariantsB_n = (int)(sizeof(kLdpcVariantsB) / sizeof(kLdpcVariantsB[0]));
}  // namespace s2
