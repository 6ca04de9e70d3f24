// TEST INFRASTRUCTURE (oracle/_ref build only): the two VOLK kernels dvbs2_pl_sync.cpp:112-113 calls, as VOLK's
// generic (portable C) implementations define them: element-wise conjugate and element-wise complex product.
// (VOLK's machine-specific kernels may round the product differently, e.g. with FMA; unpinned.)
#pragma once
#include <complex>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <sstream>
#include <stdexcept>
#include <string>
typedef std::complex<float> lv_32fc_t;
static inline void volk_32fc_conjugate_32fc(lv_32fc_t* out, const lv_32fc_t* in, unsigned int n) {
    for (unsigned int i = 0; i < n; ++i) out[i] = std::conj(in[i]);
}
static inline void volk_32fc_x2_multiply_32fc(lv_32fc_t* out, const lv_32fc_t* a, const lv_32fc_t* b, unsigned int n) {
    for (unsigned int i = 0; i < n; ++i) {
        const float ar = a[i].real(), ai = a[i].imag(), br = b[i].real(), bi = b[i].imag();
        out[i] = lv_32fc_t(ar * br - ai * bi, ar * bi + ai * br);
    }
}

// ---- row 8(f)-4 (dvbs/viterbi/cc_decoder.cpp:61-94): the reference asks VOLK which implementations of
// volk_8u_x4_conv_k7_r2_8u exist and runs "spiral" / "neonspiral" when there is one, else the generic kernel it bundles
// (dvbs/viterbi/volk_k7_r2_generic_fixed.h).  VOLK is not under /root/reference and not in this image: this stand-in
// reports no implementation, so the harness pins the BUNDLED GENERIC kernel -- the only one whose source is there.
struct volk_func_desc {
    const char** impl_names;
    const int* impl_deps;
    const bool* impl_alignment;
    size_t n_impls;
};
static inline volk_func_desc volk_8u_x4_conv_k7_r2_8u_get_func_desc(void) { return volk_func_desc{nullptr, nullptr, nullptr, 0}; }
static inline void volk_8u_x4_conv_k7_r2_8u_manual(unsigned char*, unsigned char*, unsigned char*, unsigned char*, unsigned int, unsigned int,
                                                   unsigned char*, const char*) { __builtin_trap(); }
