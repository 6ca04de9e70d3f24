// LDPC kernel instantiations, part A (split over several translation units so that they compile in parallel).
#include "ldpc_kernels.cuh"

namespace s2 {
const Variant kLdpcVariantsA[] = {VB(2), VU(3), VU(4), VB(5), VU(8), VU(9)};
const int kLdpcVariantsA_n = (int)(sizeof(kLdpcVariantsA) / sizeof(kLdpcVariantsA[0]));
}  // namespace s2
