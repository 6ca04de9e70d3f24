// K7: PL frame synchronisation, PLHEADER demodulation / PLS decoding and the coarse frequency error detector on the
// device (row 8(f)-3, upstream of the demapper).
//
// Replaces S2PLSyncBlock::process (dvbs2/dvbs2_pl_sync.cpp:80-165), S2PLHDRDemod::process
// (dvbs2/dvbs2_plhdr_demod.cpp:33-67) and dvbs2_pilot_coarse_fed (dvbs2/dvbs2_fed.h:7-48) as DVBS2Demod::process
// drives them per block of symbols (dvbs2/module_dvbs2_demod.cpp:300-316).  Float arithmetic follows the reference
// operation by operation (no fused multiply-add), so the correlation maxima -- and with them the frame positions --
// are the reference's own.
#pragma once
#include <cstdint>

#include <cuda_runtime.h>

namespace s2 {

constexpr int kPlHeader = 90;   // SOF 26 + PLS code 64

// S2PLSyncBlock's members, device-resident between calls.  The symbols gathered but not yet delivered (in_buffer /
// correlation_buffer of the reference) are the first `pend` entries of the current work buffer.
struct PlSyncState {
    int state;              // in_buffer_state: 0 gathering a window, 1 gathering best_pos more symbols
    int best_pos;
    int pend;               // symbols carried over from the previous call
    int current_position;   // (public member)
    double best_match;      // (public member)
    int nframes;            // frames delivered by the last call
    int cur;                // work buffer that holds the carried symbols
};

struct PlSyncArgs {
    float2* work;           // work buffer of this call: [pend carried][count new]   (device)
    float2* work_next;      // receives the symbols carried into the next call
    const float2* append;   // device-buffer calls: the new symbols, copied behind the carried ones by the first kernel
                            // (only the device knows exactly how many those are); null: they are already in place
    int count;              // new symbols in this call
    int rfs;                // raw_frame_size
    float* metric;          // scratch: correlation magnitude per position
    int* starts;            // scratch: first symbol of every delivered frame in `work`
    int max_frames;
    float2* out;            // delivered frames, rfs symbols each
    PlSyncState* st;
    int* nframes_out;       // optional device int
};
int plsync_launch(const PlSyncArgs& a, int pend_upper_bound, cudaStream_t stream);   // three launches (four with append)

// S2PLHDRDemod: the phase loop runs through the frames in order (one warp); PLS decoding per frame
struct PlHdrState {
    float alpha, beta, phase, freq;
};
struct PlHdrResult {
    int modcod, shortframes, pilots, pls;   // detect_modcod, detect_shortframes, detect_pilots; pls = best_header
};
struct PlTables {           // s2_sof::symbols, s2_plscodes::symbols / codewords (dvbs2/s2_defs.h:16-87), built on the host
    float2 sof[26];
    float2 pls[128][64];
    unsigned long long codewords[128];
};
int plhdr_launch(const float2* frames, int nframes, int rfs, float2* headers_out, PlHdrResult* res, PlHdrState* st,
                 const PlTables* tab, cudaStream_t stream);
// dvbs2_pilot_coarse_fed for nframes frames: a thread per frame.  rn: PL scrambling sequence (may be null without pilots)
int fed_launch(const float2* frames, int nframes, int rfs, int pilots, int pls_code, const uint8_t* rn, const PlTables* tab,
               float* err_out, cudaStream_t stream);

// K8 (pl_pll.cu): S2PLLBlock::process for consecutive frames of one stream, one warp, speculative in blocks of 32 symbols
struct PllState {
    float alpha, beta, phase, freq, error;   // loop coefficients, pcl.phase, pcl.freq, S2PLLBlock::error
    unsigned rounds;                          // evaluation rounds of the last call (diagnostic: 2 per 32 symbols is the floor)
};
struct PllArgs {
    const float2* frames;   // nframes frames at stride rfs
    float2* out;            // same stride; the first `total` symbols of each are written
    int nframes, rfs;
    int total;              // (frame_slot_count + 1) * 90 + pilot_cnt * 36 (dvbs2_pll.cpp:38)
    int pilot_cnt;          // S2PLLBlock::update (dvbs2_pll.h:47-59)
    float divisor;          // (float)(frame_slot_count + 1) * 90 + pilot_cnt * 36 (:82)
    int pls_code;
    const uint8_t* rn;      // PL scrambling sequence
    const float* perr_lut;  // 256 x 256 phase errors of the demapper's cells, null for 32APSK
    const PlTables* tab;
    PllState* st;
    float* state_out;       // optional: phase, freq, error after every frame
    float amp, prescale;    // 32APSK only: demapper scale and points
    int states;
    float pts[64];
};
// mode 0: six warps taking the blocks of 32 symbols in turn, each starting from what its predecessor has published (default);
// 1: one thread, one symbol at a time (the yardstick the tests hold the speculative kernels against); 2: one warp, block
// after block (the first generation)
int pll_launch(const PllArgs& a, int mode, cudaStream_t stream);
// njobs independent streams, one CTA each; d_jobs in device memory
int pll_launch_multi(const PllArgs* d_jobs, int njobs, cudaStream_t stream);

}  // namespace s2
