// TEST INFRASTRUCTURE (oracle/_ref build only): stand-in for SDR++ core's <dsp/types.h>.
// The reference's constellation.cpp uses complex_t::{operator*(float), operator/(float),
// operator-, operator*(complex_t), amplitude(), conj(), phase()} (constellation.cpp:15,156,
// 210-212,221,260).  SDR++ core (AlexandreRouma/SDRPlusPlus, unpinned "master" in the
// reference's CI) is not under /root/reference, so parity of this shim with the real header is
// unpinned; it restates the obvious component-wise definitions.
#pragma once
#include <math.h>
namespace dsp {
    struct complex_t {
        float re, im;
        complex_t operator*(const float b) const { return complex_t{re * b, im * b}; }
        complex_t operator/(const float b) const { return complex_t{re / b, im / b}; }
        complex_t operator*(const complex_t& b) const {
            return complex_t{(re * b.re) - (im * b.im), (im * b.re) + (re * b.im)};
        }
        complex_t operator+(const complex_t& b) const { return complex_t{re + b.re, im + b.im}; }
        complex_t operator-(const complex_t& b) const { return complex_t{re - b.re, im - b.im}; }
        complex_t& operator+=(const complex_t& b) { re += b.re; im += b.im; return *this; }   // dvbs2_pl_sync.cpp:176-190
        complex_t& operator-=(const complex_t& b) { re -= b.re; im -= b.im; return *this; }
        complex_t conj() const { return complex_t{re, -im}; }
        float phase() const { return atan2f(im, re); }
        float amplitude() const { return sqrtf((re * re) + (im * im)); }
    };
}
