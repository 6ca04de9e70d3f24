// K6: BBFRAME -> MPEG-TS packets on the device (row 8(f)-1, downstream of the BCH/descramble kernel).
//
// Replaces the TS branch of BBFrameTSParser::work (dvbs2/bbframe_ts_parser.cpp:100-212): BBHEADER CRC-8 and
// field checks, resynchronisation through SYNCD, cutting the data-field byte stream into 188-byte units across
// frame (and call) boundaries, and re-inserting the 0x47 sync byte; and its GSE branch (ts_gs = 01, :213-389): the
// walk over the GSE packets of a data field, reassembly of fragmented PDUs in three FragID slots with CRC-32, and
// the GRE wrapping of every delivered PDU (see the GSE section of ts_parser.cu).
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
    int produced;        // bytes written by the last call (kTsNoSpace: GSE output did not fit, nothing written)
    int phase;           // 0: the plan kernel committed the call; 1: GSE frames seen, call deferred to the GSE pass;
                         // 2: committed by the final plan; 3: final plan found too little room
    uint8_t last_header[16];
    uint8_t unit[2][192];
};

constexpr int kTsNoSpace = -28;
constexpr int kGseBuf = 65536 + 4096;   // reassembly buffer per slot (reference: 65536; slack for the last fragment's CRC-32)

// GSE reassembly state (bbframe_ts_parser.h:91-98), device-resident like TsState.  The bytes of unfinished PDUs
// live in GseBuffers, double-buffered: a call reads the heads it inherited from buf[cur] while it saves what it
// leaves behind into buf[cur ^ 1].
struct GseSlot {
    int on, id, proto, ctr;
    uint32_t crc;
};
struct GseSave {         // what the copy kernel must put into the new buffer of a slot at the end of a call
    int head;            // first fragment descriptor of this call's part of the chain (-1: none)
    int carry;           // bytes inherited from the previous call that stay in front of it
    int active;
    int cont;            // the chain continues one begun in an earlier call
};
struct GseState {
    GseSlot slot[3];
    GseSave save[3];
    uint32_t entry_crc[3];   // running CRC-32 of the slots when the call began
    int cur, old;        // buffer set holding the inherited heads: for the next call / of the call in flight
    int last_crc_err;    // last_gse_crc_err (bbframe_ts_parser.h:73)
    int pdus, crc_errors, malformed, dropped;   // counters of the last call
    int ndesc, ndesc_wanted;   // packets walked / found (more than the pool holds: kTsNoSpace)
    int out_bytes;       // GSE bytes the last call emits
};
struct GseDesc {         // one GSE packet of the call, in stream order (16 bytes)
    uint32_t src;        // byte offset of its first header byte in the BBFRAME array
    uint16_t len;        // data bytes (header, label and the like taken off as the reference does, :234-252,288-310,338,366)
    uint8_t kind;        // 0 complete PDU, 1 first, 2 middle, 3 last fragment
    uint8_t hdr;         // bytes from src to the data
    uint16_t proto;
    uint8_t fragid, label6;
    uint32_t frame;
};
struct GseOut {          // what the sequential pass decides per packet (16 bytes)
    int before;          // GSE bytes of this call in front of it
    int emit;            // bytes it puts out (GRE header included), 0: none
    int pos;             // fragments: position of its data in the reassembly buffer; -1: dropped
    int link;            // last fragment: first descriptor of the chain (-2 once its CRC-32 has failed); others: unused
};
struct GseWork {         // per-call scratch; null desc = GSE pass not available (frames are only counted)
    int* doff;           // [cnt + 1] first descriptor of every frame
    uint8_t* entry_sync; // [cnt] parser in sync in front of the frame?
    int* before;         // [cnt + 1] GSE bytes in front of every frame; [cnt] = all of them
    GseDesc* desc;
    GseOut* out;
    uint32_t* crc0;      // zero-start CRC-32 of the data (first fragment: the finished start value; last: xor received)
    uint32_t* xpow;      // x^(8 len) mod P
    int* nxt;            // next fragment of the chain
    int* aux;            // last fragment: slot | inherited bytes << 2 when its chain began in an earlier call, else -1;
                         // protocol type of the PDU in aux2
    int* aux2;
    int cap;             // descriptors the pool holds
    GseState* state;
    uint8_t* buf;        // [2][3][kGseBuf]
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
    GseWork gse;         // gse.state == nullptr: GSE frames are counted only
    int mode;            // plan kernel: 0 commit (defer when GSE frames show up and gse.state is set), 2 final plan after the GSE pass
    int copy_phase;      // copy kernel: runs when TsState::phase equals this
};

int ts_launch(const TsArgs& a, cudaStream_t stream);  // three launches; returns a cudaError_t value
// GSE pass of a deferred call (TsState::phase == 1), all on `stream`; every kernel is a no-op otherwise.
int gse_launch_count(const TsArgs& a, cudaStream_t stream);   // packets per frame -> gse.doff, GseState::ndesc_wanted
int gse_launch_rest(const TsArgs& a, cudaStream_t stream);    // descriptors, CRCs, reassembly, final plan, copies

}  // namespace s2
