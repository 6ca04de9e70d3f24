// Host-side construction of everything the kernels read as tables, plus the in-tree DVB-S2
// transmitter used to make synthetic input (BB scramble -> BCH -> LDPC -> interleave -> map).
// Written from EN 302 307; the reference behaviour each item must agree with is cited inline.
#pragma once
#include <cstdint>
#include <vector>

#include "s2_codes.h"

namespace s2 {

// GF(2^m) exp/log tables with the conventions the BCH decoder relies on (log[0] = N, exp[N] = 0;
// reference: bch/galois_field.hh:149-163).  m = 16: x^16+x^5+x^3+x^2+1, m = 14: x^14+x^5+x^3+x+1.
struct GfHost {
    int m, N;
    std::vector<uint16_t> log, exp;
    uint32_t mul(uint32_t a, uint32_t b) const {
        if (!a || !b) return 0;
        int s = log[a] + log[b];
        return exp[s >= N ? s - N : s];
    }
};
const GfHost& gf_host(int m);

// Binary BCH code over GF(2^m) with design distance 2t+1 (EN 302 307 5.3.1, tables 6a/6b).
struct BchHost {
    int m, t, np;                       // np = m * t parity bits
    std::vector<uint8_t> gen;           // g(x) coefficients, gen[k] = coeff of x^k, degree np
    std::vector<uint16_t> crc;          // [t][256]: (idx * x^m) mod minpoly(alpha^(2k+1))
    std::vector<uint16_t> basis;        // [t][16]:  alpha^((2k+1) * b)
};
const BchHost& bch_host(int m, int t);
// systematic encode in place: frame holds kbch data bits MSB-first, np parity bits are appended
void bch_encode(const BchHost& code, uint8_t* frame, int kbch);

// BB scrambler sequence 1 + x^14 + x^15, init 100101010000000 (EN 302 307 5.2.2;
// reference: bbframe_descramble.cpp:122-136), MSB first, 64800 bits.
const std::vector<uint8_t>& bb_prbs();

// PL scrambling sequence Rn (EN 302 307 5.5.4; reference: S2Scrambling, dvbs2/codings/s2_scrambling.cpp:9-28) for
// Gold code `codenum`: 2 bits per symbol position after the PLHEADER, `count` positions.
std::vector<uint8_t> pl_scrambling_rn(int codenum, int count);

// LDPC systematic encode (EN 302 307 5.3.2): data_bits[K] 0/1 -> code_bits[N] 0/1
void ldpc_encode_bits(int code, const uint8_t* data_bits, uint8_t* code_bits);

// Constellations as the reference demapper sees them (common/dsp/demod/constellation.cpp:19-150):
// point index = bit-inverted DVB-S2 label, amplitudes pre-multiplied by the demapper's own scale.
struct ConstellationHost {
    Constellation type;
    int bits, states;
    float amp, sca, prescale;
    float re[32], im[32];
};
ConstellationHost make_constellation(Constellation type, float g1, float g2);
// LLRs of one received sample exactly as constellation_t::demod_soft_calc produces them (:205-261)
void demap_calc(const ConstellationHost& c, float re, float im, int8_t* bits);
// 256x256 LUT of constellation_t::make_lut (:272-291), one uint32 per cell (byte k = soft bit k)
std::vector<uint32_t> demap_lut(const ConstellationHost& c);
// phase_error of constellation_t::demod_soft_calc (:209-231,258-260) for one sample, and the 256x256 table of it that
// make_lut stores per cell (:284-288): what S2PLLBlock::process reads through demod_soft_lut (dvbs2_pll.cpp:45,49)
float demap_phase_error(const ConstellationHost& c, float re, float im);
std::vector<float> demap_phase_lut(const ConstellationHost& c);
// transmit point for a group of `bits` code bits (first bit = MSB of the label), unit-energy scale the
// reference's own modulator uses (constellation_t::mod, :156-158)
void map_symbol(const ConstellationHost& c, const uint8_t* code_bits, float* re_im);
// position in the interleaved (transmit-order) stream of code bit n: inverse of
// S2Deinterleaver::deinterleave (s2_deinterleaver.cpp:72-136)
int interleaved_position(const ModcodCfg& cfg, int n);

}  // namespace s2
