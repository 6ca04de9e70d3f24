// K6: BBFRAME -> MPEG-TS packets on the device (row 8(f)-1, downstream of the BCH/descramble kernel).
//
// Replaces the TS branch of BBFrameTSParser::work (dvbs2/bbframe_ts_parser.cpp:100-212): BBHEADER CRC-8 and
// field checks, resynchronisation through SYNCD, cutting the data-field byte stream into 188-byte units across
// frame (and call) boundaries, and re-inserting the 0x47 sync byte.  GSE frames (ts_gs = 01, :215-389) are
// counted and reported, their payload is left to the caller.
#pragma once
#include <cstdint>

#include <cuda_runtime.h>

namespace s2 {

// Device-resident parser state: BBFrameTSParser's private members (bbframe_ts_parser.h:77-89) plus the public
// counters (:72-76).  The unfinished unit is double-buffered so that the copy kernel can still read the carry
// the call started with while the plan kernel already stores the carry the call ends with.
struct TsState {
    unsigned count;      // bytes of the unfinished 188-byte unit
    int synched;
    int cur;             // which unit[] buffer holds the carry
    int entry_buf;       // buffer that held it when the current call started
    int have_header;
    int last_bb_cnt, last_bb_proc;
    int gse_frames;      // accepted frames of the last call with ts_gs = 01
    int produced;        // bytes written by the last call
    int pad_;
    uint8_t last_header[16];
    uint8_t unit[2][192];
};

struct TsPlan {          // what one frame contributes to the output
    int out_off;         // byte offset of its first packet in the output
    int src_off;         // offset in the frame where its own bytes start being consumed
    int head_src;        // frame holding the first unit's leading bytes (-1: TsState::unit[entry_buf])
    int head_src_off;
    short npk;           // packets
    short head;          // leading bytes of the first unit that come from the carry (0: none)
};

struct TsArgs {
    const uint8_t* bb;   // cnt frames of kb bytes (device)
    int cnt, kb, max_dfl;
    uint8_t* out;        // device
    int out_cap;
    TsState* state;
    TsPlan* plan;        // cnt entries of scratch
    uint32_t* meta;      // cnt words of scratch
    int* produced_out;   // optional device int
};

int ts_launch(const TsArgs& a, cudaStream_t stream);  // three launches; returns a cudaError_t value

}  // namespace s2
