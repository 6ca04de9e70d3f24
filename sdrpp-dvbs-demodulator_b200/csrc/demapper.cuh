// K1 -- symbols-to-soft demapper + bit deinterleaver for sm_100a.
//
// Reference semantics reproduced: S2BBToSoft::process (dvbs2/dvbs2_bb_to_soft.cpp:7-33),
// constellation_t::demod_soft_lut (common/dsp/demod/constellation.cpp:293-322) and
// S2Deinterleaver::deinterleave (dvbs2/codings/s2_deinterleaver.cpp:72-136).
//   * QPSK / 8PSK / 16APSK: the 256x256 LUT is built on the host by the same float expressions
//     (host_tables.cpp) and gathered here -> bit-exact int8 LLRs;
//   * 32APSK has no LUT in the reference (per-symbol expf/logf); the device evaluates the same
//     formula with CUDA's libm, so LLRs may differ by an LSB from glibc's (tolerance in the tests);
//   * pilots: the intended rule (skip 36 symbols after every 16 slots); the reference's own loop
//     mis-handles pilots-on (SURVEY.md note N2), so exact parity is defined for pilots off.
//   * optional PL descrambling (S2Scrambling::descramble, dvbs2/codings/s2_scrambling.cpp:37-58, as driven by
//     S2PLLBlock::process, dvbs2_pll.cpp:37-44): the symbol is turned by -Rn[position] * 90 degrees first;
// One thread per payload symbol; LLR bytes land directly at their deinterleaved position.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace s2 {

struct DemapDev {
    int constellation;     // 0 QPSK, 1 8PSK, 2 16APSK, 3 32APSK
    int bits;
    int nsym;              // payload symbols per frame = N / bits
    int N;
    int reversed_cols;     // 8PSK 3/5
    int pilots;
    int plframe_syms;      // stride of one PLFRAME in complex samples (header + payload + pilots)
    const uint32_t* lut;   // [256*256] packed soft bits (null for 32APSK)
    const uint8_t* rn;     // PL scrambling sequence per position after the header (incl. pilots), or null: input is descrambled
    float amp, prescale, sca;
    float pts[64];         // 32APSK points (re, im), demapper scale
};

int demap_launch(const DemapDev& d, const float* plframes, int nframes, int8_t* llr_out, cudaStream_t stream);
// from LUT coordinates (two bytes per payload symbol, pilots and header already removed); not for 32APSK
int demap_idx_launch(const DemapDev& d, const uint8_t* idx, int nframes, int8_t* llr_out, cudaStream_t stream);

}  // namespace s2
