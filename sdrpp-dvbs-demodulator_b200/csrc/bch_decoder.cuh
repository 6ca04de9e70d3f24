// K3/K4/K5 -- BCH decode (syndromes, Berlekamp-Massey, root search, Forney check, correction),
// BB descramble and BBFRAME output for sm_100a.
//
// Reference semantics reproduced (SURVEY.md spec S-BCH):
//   bch/bose_chaudhuri_hocquenghem_decoder.hh:40-143, bch/reed_solomon_error_correction.hh:34-316,
//   bch/galois_field.hh:122-362, codings/bbframe_bch.cpp:380-405 (dispatch),
//   codings/bbframe_descramble.cpp:122-143, module_dvbs2_demod.cpp:357-366 (repack/copy).
// Different construction: one warp per frame; syndromes come from 2t/2 CRC-style remainders modulo
// the minimal polynomials (one table lookup per byte and syndrome) over 32 lane-chunks instead of
// (kbch+NP)*2t Horner steps; the root search only walks the nbch positions that exist in the
// shortened code; Berlekamp-Massey runs one coefficient per lane.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace s2 {

struct GfDev {            // GF(2^m) tables in global memory
    int m, N;             // N = 2^m - 1
    const uint16_t* log;  // [2^m], log[0] = N
    const uint16_t* exp;  // [2^m], exp[N] = 0
};

struct BchDev {
    GfDev gf;
    int t;                 // correctable errors; 2t syndromes
    int nbch, kbch;        // bits
    int prefix;            // shortening: full-length position of codeword bit 0 ( = 2^m-1 - nbch )
    const uint16_t* crc;   // [t][256] remainder tables modulo the minimal polynomial of alpha^(2k+1)
    const uint16_t* basis; // [t][16]  alpha^((2k+1) b): maps a remainder to the field element r(alpha^(2k+1))
    const uint8_t* prbs;   // [8100] BB scrambler sequence, MSB first
};

struct BchArgs {
    BchDev code;
    uint8_t* hard;             // [nframes][hard_stride] codewords (corrected in place)
    int hard_stride;           // bytes, multiple of 16
    int nframes;
    const int16_t* ldpc_iters; // [nframes] or nullptr
    const uint64_t* tags;      // [nframes] or nullptr (then tag = tag_base + frame index)
    unsigned long long tag_base;
    uint8_t* bb_out;           // [nframes][kbch/8] descrambled BBFRAMEs, or nullptr
    void* results;             // dvbs2fec_result[nframes], or nullptr
    int16_t* corr_out;         // [nframes] or nullptr
    int descramble;            // 0: leave bb_out scrambled (stage-level BBFrameBCH::decode parity)
};

int bch_launch(const BchArgs& args, cudaStream_t stream);

// stand-alone descrambler (BBFrameDescrambler::work): frames [n][stride] bytes, first kbch/8 XORed in place
int descramble_launch(uint8_t* frames, int stride, int nframes, int kbch, const uint8_t* prbs, cudaStream_t stream);

}  // namespace s2
