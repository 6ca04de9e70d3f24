// TEST INFRASTRUCTURE (oracle/_ref build only): the two VOLK kernels dvbs2_pl_sync.cpp:112-113 calls, as VOLK's
// generic (portable C) implementations define them: element-wise conjugate and element-wise complex product.
// (VOLK's machine-specific kernels may round the product differently, e.g. with FMA; unpinned.)
#pragma once
#include <complex>
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
