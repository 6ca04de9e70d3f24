// K6: BBFRAME -> MPEG-TS packets.  See ts_parser.cuh.
//
// Three kernels.  ts_header_kernel checks the BBHEADERs, a thread per frame.  ts_plan_kernel (one CTA) works out
// what every frame contributes; that depends on the parser state the frames before it left behind (in sync or not, bytes of an
// unfinished unit), which a three-phase scan over composable state maps delivers without walking the frames
// one by one.  ts_copy_kernel (one CTA per frame, a warp per packet) moves the bytes.
// HBM-bound: every BBFRAME byte is read once and every TS byte written once.
//
// GSE (ts_gs = 01): the packets of a data field are a linked list (every header says where the next one starts)
// and fragments are reassembled in three FragID slots, so the reference's loop (:213-389) is sequential twice
// over.  Here only what must be sequential is: a thread per frame walks its packet headers (count, then fill a
// compact descriptor array), a thread per fragment computes the CRC-32 of its own bytes from a ZERO start value
// plus x^(8 len) mod P -- CRC-32 is linear, crc(s, B) = s x^(8|B|) + crc(0, B) -- and one warp then replays the
// slot logic over the descriptors in stream order with a 32-step polynomial multiplication per fragment instead
// of touching any payload byte.  The payload moves afterwards, a CTA per delivered PDU.  A call that contains GSE
// frames is deferred by the plan kernel (phase 1) and planned again once the GSE byte counts are known.
#include "ts_parser.cuh"

#include <algorithm>

namespace s2 {
namespace {

constexpr int kHeaderThreads = 128;
constexpr int kPlanThreads = 256;
constexpr int kCopyThreads = 128;

// check_crc8(bbf, 80) (bbframe_ts_parser.cpp:66-80): bit-serial, LSB-first register, polynomial 0xAB
__device__ inline unsigned bbheader_crc8(const uint8_t* h) {
    unsigned crc = 0;
    for (int n = 0; n < 80; ++n) {
        unsigned b = ((h[n >> 3] >> (7 - (n & 7))) & 1u) ^ (crc & 1u);
        crc >>= 1;
        if (b) crc ^= 0xABu;
    }
    return crc;
}

enum { kInvalid = 0, kTs = 1, kGse = 2, kOther = 3 };

// One frame's verdict in a word: kind | data-field bytes << 2 | resync skip (SYNCD/8 + 1) << 15
// | bit 30: GSE frame whose field is walked (no ISSY, no NPD, UPL = 0, :216)
__device__ inline uint32_t pack_meta(int kind, int dfl, int syncd, int plain) {
    return (uint32_t)kind | ((uint32_t)(dfl >> 3) << 2) | ((uint32_t)((syncd >> 3) + 1) << 15) | ((uint32_t)plain << 30);
}
__device__ inline int meta_skip(uint32_t m) { return (int)((m >> 15) & 0x7FFFu); }

// Parser state seen from outside a frame: -1 = out of sync, else in sync with `st` bytes of an unfinished unit.
// (The byte count of an out-of-sync parser is never read again: resynchronising zeroes it, :163.)
struct Step {
    int next;       // state after the frame
    int npk;        // packets it emits
    int off;        // offset in the frame where its bytes start being consumed
    int tail_off;   // where its own unfinished unit starts, -1 if it leaves none of its own
};
__device__ inline Step ts_step(uint32_t m, int st) {
    const int kind = m & 3;
    Step r{-1, 0, 10, -1};
    if (kind == kInvalid) return r;
    int left = (m >> 2) & 0x1FFF, count = st;
    if (st < 0) {   // enter just past the first sync byte (:157-168)
        const int skip = meta_skip(m);
        r.off += skip;
        left -= skip;
        count = 0;
    }
    r.next = count;
    if (kind != kTs) return r;
    if (left >= 188) {   // (:176-201) the carried bytes are completed first, then whole units
        const int total = left + count;
        r.npk = total / 188;
        r.next = total - 188 * r.npk;
        if (r.next > 0) r.tail_off = r.off + 188 * r.npk - count;
    } else if (left > 0) {   // (:203-207) too short to complete anything: the carry is replaced
        r.next = left;
        r.tail_off = r.off;
    }
    return r;
}

// Exact sequential walk, used when the output may run out of room (:176,208-211): one thread, in frame order.
__device__ void ts_plan_serial(const TsArgs& a, const uint32_t* meta, int st, int& o_out, int& processed_out, int& gse_out,
                               int& last_valid_out, int& st_out, int& car_src_out, int& car_off_out) {
    int o = 0, processed = 0, gse = 0, last_valid = -1, car_src = -1, car_off = 0;
    bool stop = false;
    for (int f = 0; f < a.cnt; ++f) {
        TsPlan p{o, 0, -1, 0, 0, 0};
        const uint32_t m = meta[f];
        const int kind = m & 3;
        if (!stop) {
            if (kind == kInvalid) {
                st = -1;
            } else {
                int left = (m >> 2) & 0x1FFF, off = 10, count = st;
                if (st < 0) {
                    const int skip = meta_skip(m);
                    off += skip;
                    left -= skip;
                    count = 0;
                }
                last_valid = f;
                ++processed;
                gse += kind == kGse;
                if (kind == kTs) {
                    const int room = a.out_cap - o;
                    int consumed = 0;
                    p.src_off = off;
                    if (left >= 188 && room > 188) {
                        int fit = (room - 189) / 188 + 1;
                        if (count > 0) {
                            p.head = (short)count;
                            p.head_src = car_src;
                            p.head_src_off = car_off;
                            consumed = 188 - count;
                            left -= consumed;
                            count = 0;
                            p.npk = 1;
                            --fit;
                        }
                        const int whole = min(left / 188, fit);
                        p.npk = (short)(p.npk + whole);
                        left -= 188 * whole;
                        consumed += 188 * whole;
                    }
                    o += 188 * p.npk;
                    if (left >= 188) {
                        // only when the room test stopped the loop.  The reference then copies `left` bytes into its
                        // 188-byte packet_reassembly and carries count >= 188 into the next call (:203-207: a buffer
                        // overrun, then a negative memcpy length) -- undefined there; here the parser drops out of
                        // sync and picks up again at the next frame's SYNCD.
                        count = -1;
                    } else if (left > 0) {
                        count = left;
                        car_src = f;
                        car_off = off + consumed;
                    }
                    if (a.out_cap - o <= 188) stop = true;
                }
                st = count;
            }
        }
        a.plan[f] = p;
    }
    o_out = o; processed_out = processed; gse_out = gse; last_valid_out = last_valid; st_out = st;
    car_src_out = car_src; car_off_out = car_off;
}

// What a run of frames does to the parser state is one of a small family of maps -- out of sync -> a constant;
// in sync with c bytes -> out of sync ("kill") or (a c + b) mod 188 with a in {0,1} -- which is closed under
// composition.  So the state in front of every frame follows from a scan: every thread summarises its run of
// frames (by probing it with three entry states), a block-wide scan composes the summaries, every thread
// replays its run.  A map is packed as (image of "out of sync") + 1 | b << 9 | a << 17 | kill << 18.
__device__ inline uint32_t map_pack(int from_unsync, int a, int b, int kill) {
    return (uint32_t)(from_unsync + 1) | (uint32_t)b << 9 | (uint32_t)a << 17 | (uint32_t)kill << 18;
}
__device__ inline int map_apply(uint32_t m, int st) {
    if (st < 0) return (int)(m & 0x1FF) - 1;
    if ((m >> 18) & 1) return -1;
    return (int)(((m >> 17) & 1) * st + ((m >> 9) & 0xFF)) % 188;
}
__device__ inline uint32_t map_compose(uint32_t f, uint32_t g) {   // f first, then g
    const int from_unsync = map_apply(g, (int)(f & 0x1FF) - 1);
    if ((f >> 18) & 1) {   // in sync -> out of sync -> wherever g sends "out of sync"
        const int gu = (int)(g & 0x1FF) - 1;
        return gu < 0 ? map_pack(from_unsync, 0, 0, 1) : map_pack(from_unsync, 0, gu, 0);
    }
    if ((g >> 18) & 1) return map_pack(from_unsync, 0, 0, 1);
    const int fa = (f >> 17) & 1, fb = (f >> 9) & 0xFF, ga = (g >> 17) & 1, gb = (g >> 9) & 0xFF;
    return map_pack(from_unsync, ga & fa, (ga * fb + gb) % 188, 0);
}
constexpr uint32_t kMapIdentity = 0u | 0u << 9 | 1u << 17;   // out of sync stays out of sync, c -> c

// block-wide scans over the kPlanThreads per-thread values: exclusive prefix in the return value, total in `total`
template <typename T, typename Op>
__device__ inline T block_scan_exclusive(T v, T identity, Op op, T* s_warp, T& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kPlanThreads / 32;
    T incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const T o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl = op(o, incl);
    }
    __syncthreads();   // s_warp may still be read from a previous scan
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    T wpre = identity;
    T run = identity;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {   // 8 warp totals: every thread folds them itself
        if (w == warp) wpre = run;
        run = op(run, s_warp[w]);
    }
    total = run;
    T excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
    if (lane == 0) excl = identity;
    return op(wpre, excl);
}

// (0) BBHEADER checks, one thread per frame (:122-150): CRC-8, DFL <= kbch-80, SYNCD < DFL-8 as signed ints,
//     DFL a whole number of bytes
__global__ void __launch_bounds__(kHeaderThreads) ts_header_kernel(const TsArgs a) {
    const int f = blockIdx.x * kHeaderThreads + threadIdx.x;
    if (f >= a.cnt) return;
    const uint8_t* h = a.bb + (size_t)f * a.kb;
    uint8_t b[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) b[i] = h[i];
    const int dfl = (b[4] << 8) | b[5], syncd = (b[7] << 8) | b[8];
    int kind = kInvalid;
    if (bbheader_crc8(b) == 0 && dfl <= a.max_dfl && syncd < dfl - 8 && (dfl & 7) == 0) {
        const int ts_gs = b[0] >> 6;
        kind = ts_gs == 3 ? kTs : ts_gs == 1 ? kGse : kOther;
    }
    const int plain = kind == kGse && !((b[0] >> 3) & 1) && !((b[0] >> 2) & 1) && ((b[2] << 8) | b[3]) == 0;
    a.meta[f] = pack_meta(kind, dfl, syncd, plain);
}

constexpr int kMetaShared = 8192;   // frames whose verdict words are staged in shared memory

__global__ void __launch_bounds__(kPlanThreads) ts_plan_kernel(const TsArgs a) {
    __shared__ uint32_t s_meta[kMetaShared];
    __shared__ uint32_t s_w32[kPlanThreads / 32];
    __shared__ unsigned long long s_w64[kPlanThreads / 32];
    __shared__ int s_fin[8];
    const int tid = threadIdx.x;
    TsState* S = a.state;
    const uint32_t* M = a.meta;
    const bool final = a.mode == 2;          // second planning of a call that was deferred to the GSE pass
    if (final && S->phase != 1) return;      // (uniform: phase is only written behind the barriers below)
    if (a.cnt <= kMetaShared) {
        for (int f = tid; f < a.cnt; f += kPlanThreads) s_meta[f] = a.meta[f];
        M = s_meta;
        __syncthreads();
    }
    const int st0 = S->synched ? (int)S->count : -1;
    const int per = (a.cnt + kPlanThreads - 1) / kPlanThreads;
    const int f0 = min(a.cnt, tid * per), f1 = min(a.cnt, f0 + per);
    // (1) summarise the run
    uint32_t mymap;
    {
        int su = -1, s0 = 0, s1 = 1;
        for (int f = f0; f < f1; ++f) {
            const uint32_t m = M[f];
            su = ts_step(m, su).next;
            s0 = ts_step(m, s0).next;
            s1 = ts_step(m, s1).next;
        }
        mymap = map_pack(su, s0 >= 0 && s1 != s0, max(s0, 0), s0 < 0);
    }
    // (2) compose the summaries: state in front of each run, and after the last one
    uint32_t total_map;
    const uint32_t pre_map = block_scan_exclusive(mymap, kMapIdentity, map_compose, s_w32, total_map);
    const int entry = map_apply(pre_map, st0);
    const int final_state = map_apply(total_map, st0);
    // (3) replay for the run's totals
    int npk = 0, valid = 0, gse = 0, gse_plain = 0, lastv = -1;
    unsigned long long writer = 0;   // (frame + 1) << 32 | offset of the carry it leaves; 0: none
    {
        int st = entry;
        for (int f = f0; f < f1; ++f) {
            const uint32_t m = M[f];
            const Step r = ts_step(m, st);
            if ((m & 3) != kInvalid) {
                ++valid;
                lastv = f;
                gse += (m & 3) == kGse;
                gse_plain += (m >> 30) & 1;
            }
            npk += r.npk;
            if (r.tail_off >= 0) writer = ((unsigned long long)(f + 1) << 32) | (unsigned)r.tail_off;
            st = r.next;
        }
    }
    // (4) prefixes over the runs: packets before a run, last frame before it that left a carry; totals
    auto add = [](uint32_t x, uint32_t y) { return x + y; };
    auto umax = [](uint32_t x, uint32_t y) { return x > y ? x : y; };
    auto later = [](unsigned long long x, unsigned long long y) { return y ? y : x; };   // y is the later run
    uint32_t npk_total, valid_total, gse_total, lastv_total;
    unsigned long long writer_total;
    const uint32_t npk_before = block_scan_exclusive((uint32_t)npk, 0u, add, s_w32, npk_total);
    block_scan_exclusive((uint32_t)valid, 0u, add, s_w32, valid_total);
    block_scan_exclusive((uint32_t)gse, 0u, add, s_w32, gse_total);
    uint32_t gse_plain_total;   // GSE frames whose field is walked (the others put nothing out, :216,386-388)
    block_scan_exclusive((uint32_t)gse_plain, 0u, add, s_w32, gse_plain_total);
    block_scan_exclusive((uint32_t)(lastv + 1), 0u, umax, s_w32, lastv_total);
    const unsigned long long writer_before = block_scan_exclusive(writer, 0ull, later, s_w64, writer_total);
    if (!final && a.gse.state && gse_plain_total > 0) {
        // GSE frames: what they put out is known only after the GSE pass.  Leave the state untouched, say which
        // frames are entered in sync (a frame entered out of sync is walked from SYNCD/8 + 1, :157-168) and defer.
        int st = entry;
        for (int f = f0; f < f1; ++f) {
            a.gse.entry_sync[f] = st >= 0;
            st = ts_step(M[f], st).next;
        }
        if (tid == 0) {
            S->phase = 1;
            S->produced = 0;
            if (a.produced_out) *a.produced_out = 0;
        }
        return;
    }
    const int og_total = final ? a.gse.before[a.cnt] : 0;
    // enough room for everything (no test of :176/:208 can fail)?  Otherwise one thread redoes it in frame order.
    // With GSE output in the call there is no such fallback: the reference writes PDUs without looking at the
    // room (it overruns the buffer), so the call is refused as a whole (kTsNoSpace) with the state advanced.
    const bool roomy = a.out_cap - 188 * (int)npk_total - og_total > 188 && !(final && a.gse.state->ndesc < 0);
    if (tid == 0) {
        S->entry_buf = S->cur;
        s_fin[0] = final_state;
        s_fin[1] = 188 * (int)npk_total;
        s_fin[2] = (int)valid_total;
        s_fin[3] = (int)gse_total;
        s_fin[4] = (int)lastv_total - 1;
        s_fin[5] = (int)(writer_total >> 32) - 1;
        s_fin[6] = (int)(writer_total & 0xFFFFFFFFu);
        if (!roomy && !final) ts_plan_serial(a, M, st0, s_fin[1], s_fin[2], s_fin[3], s_fin[4], s_fin[0], s_fin[5], s_fin[6]);
    }
    // (5) replay once more, now writing what every frame contributes
    if (roomy || final) {
        int st = entry, o = (int)npk_before;
        int wsrc = (int)(writer_before >> 32) - 1, woff = (int)(writer_before & 0xFFFFFFFFu);
        for (int f = f0; f < f1; ++f) {
            const Step r = ts_step(M[f], st);
            TsPlan p{188 * o + (final ? a.gse.before[f] : 0), r.off, -1, 0, (short)r.npk, 0};
            if (r.npk > 0 && st > 0) {   // first unit starts with the carried bytes (st <= 0: entered clean or resynced)
                p.head = (short)st;
                p.head_src = wsrc;
                p.head_src_off = woff;
            }
            a.plan[f] = p;
            o += r.npk;
            if (r.tail_off >= 0) {
                wsrc = f;
                woff = r.tail_off;
            }
            st = r.next;
        }
    }
    __syncthreads();
    // (6) commit the state the call ends with
    if (tid == 0) {
        const int st = s_fin[0];
        S->synched = st >= 0;
        if (st >= 0) S->count = (unsigned)st;   // out of sync: the stale count is dead (zeroed on resync)
        S->last_bb_cnt = a.cnt;
        S->last_bb_proc = s_fin[2];
        S->gse_frames = s_fin[3];
        const int produced = roomy ? s_fin[1] + og_total : (final ? kTsNoSpace : s_fin[1]);
        S->produced = produced;
        S->phase = final ? (roomy ? 2 : 3) : 0;
        if (a.produced_out) *a.produced_out = produced;
        if (s_fin[4] >= 0) {
            S->have_header = 1;
            for (int i = 0; i < 10; ++i) S->last_header[i] = a.bb[(size_t)s_fin[4] * a.kb + i];
        }
        if (s_fin[5] >= 0) S->cur ^= 1;
    }
    __syncthreads();
    if (s_fin[5] >= 0 && tid < min(max(s_fin[0], 0), 188))
        S->unit[S->entry_buf ^ 1][tid] = a.bb[(size_t)s_fin[5] * a.kb + s_fin[6] + tid];
}

__global__ void __launch_bounds__(kCopyThreads) ts_copy_kernel(const TsArgs a) {
    const int f = blockIdx.x;
    if (a.state->phase != a.copy_phase) return;
    const TsPlan p = a.plan[f];
    if (p.npk == 0) return;
    const uint8_t* fr = a.bb + (size_t)f * a.kb;
    const uint8_t* carry = nullptr;
    if (p.head) carry = p.head_src < 0 ? a.state->unit[a.state->entry_buf] : a.bb + (size_t)p.head_src * a.kb + p.head_src_off;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool words = (((uintptr_t)a.out + (uintptr_t)p.out_off) & 3) == 0;   // (GSE bytes in front of a frame may be any number)
    const uint8_t* bb_end = a.bb + (size_t)a.cnt * a.kb;
    for (int u = warp; u < p.npk; u += kCopyThreads / 32) {
        uint8_t* dst = a.out + p.out_off + 188 * u;
        // frame offset of the unit's byte 0; the first `head` bytes of unit 0 lie in the carry instead
        const int ustart = p.src_off + 188 * u - p.head;
        const int head = u == 0 ? p.head : 0;
        // output byte k is unit byte k-1 (k = 0: the sync byte); the unit's 188th byte, the next CRC slot, is dropped
        const uintptr_t s0 = (uintptr_t)(fr + ustart - 1);
        if (words && head == 0 && reinterpret_cast<const uint8_t*>((s0 & ~(uintptr_t)3) + 192) <= bb_end) {
            // aligned 32-bit loads, realigned with a funnel shift
            const uint32_t* al = reinterpret_cast<const uint32_t*>(s0 & ~(uintptr_t)3);
            const int sh = 8 * (int)(s0 & 3);
            for (int w = lane; w < 47; w += 32) {
                uint32_t v = __funnelshift_r(__ldg(al + w), __ldg(al + w + 1), sh);
                if (w == 0) v = (v & 0xFFFFFF00u) | 0x47u;
                reinterpret_cast<uint32_t*>(dst)[w] = v;
            }
            continue;
        }
        for (int w = lane; w < 47; w += 32) {
            uint32_t v = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = 4 * w + k - 1;
                const uint32_t byte = i < 0 ? 0x47u : (i < head ? carry[i] : fr[ustart + i]);
                v |= byte << (8 * k);
            }
            if (words) {
                reinterpret_cast<uint32_t*>(dst)[w] = v;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) dst[4 * w + k] = (uint8_t)(v >> (8 * k));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ GSE pass
constexpr int kGseThreads = 128;
constexpr uint32_t kCrc32Poly = 0x04C11DB7u;

// The walk over one data field (:214-389): packet headers only.  FILL = false counts, FILL = true also writes the
// descriptors.  A packet whose header or data would lie outside the BBFRAME array (the reference reads on, with
// whatever follows its input) ends the walk and is reported as malformed.
template <bool FILL>
__device__ int gse_walk(const TsArgs& a, int f, GseDesc* out, int* malformed) {
    const uint32_t m = a.meta[f];
    if ((m & 3) != kGse || !((m >> 30) & 1)) return 0;
    const size_t total = (size_t)a.cnt * a.kb;
    const size_t field = (size_t)f * a.kb + 10 + (a.gse.entry_sync[f] ? 0 : meta_skip(m));
    const unsigned df_bytes = (m >> 2) & 0x1FFFu;   // the full DFL/8 even behind a resync skip (:214)
    unsigned at = 0;
    int n = 0;
    while (at < df_bytes) {
        const size_t g = field + at;
        if (g + 3 > total) { *malformed = 1; break; }
        const unsigned h1 = a.bb[g], h2 = a.bb[g + 1];
        const unsigned start = h1 >> 7, end = (h1 >> 6) & 1, label6 = (h1 & 0x30u) == 0;   // (:219: "(h1 & 0x30) >> 2" is 0, 4, 8 or 12)
        if (!start && !end && label6) break;                                               // padding (:220-222)
        unsigned len = ((h1 & 0x0Fu) << 8) | h2, hdr, kind;
        if (start && end) { kind = 0; hdr = 4; len -= 2; }
        else if (start) { kind = 1; hdr = 7; len -= 5; }
        else { kind = end ? 3 : 2; hdr = 3; len -= 1; }
        if (kind < 2 && label6) { hdr += 6; len -= 6; }
        len &= 0xFFFFu;                                                                    // uint16_t gse_len
        if (g + hdr + len > total || (kind == 3 && len < 4)) { *malformed = 1; break; }
        if (FILL) {
            GseDesc d;
            d.src = (uint32_t)g;
            d.len = (uint16_t)len;
            d.kind = (uint8_t)kind;
            d.hdr = (uint8_t)hdr;
            d.proto = kind == 0 ? (uint16_t)((a.bb[g + 2] << 8) | a.bb[g + 3]) : kind == 1 ? (uint16_t)((a.bb[g + 5] << 8) | a.bb[g + 6]) : 0;
            d.fragid = a.bb[g + 2];
            d.label6 = (uint8_t)label6;
            d.frame = (uint32_t)f;
            out[n] = d;
        }
        ++n;
        at += hdr + len;
    }
    return n;
}

__global__ void __launch_bounds__(kGseThreads) gse_count_kernel(const TsArgs a) {
    if (a.state->phase != 1) return;
    const int f = blockIdx.x * kGseThreads + threadIdx.x;
    if (f >= a.cnt) return;
    int bad = 0;
    a.gse.doff[f] = gse_walk<false>(a, f, nullptr, &bad);
}

// counts -> first descriptor of every frame, in place; doff[cnt] = GseState::ndesc_wanted = all of them
__global__ void __launch_bounds__(kPlanThreads) gse_scan_kernel(const TsArgs a) {
    __shared__ uint32_t s_w32[kPlanThreads / 32];
    if (a.state->phase != 1) return;
    const int tid = threadIdx.x;
    const int per = (a.cnt + kPlanThreads - 1) / kPlanThreads;
    const int f0 = min(a.cnt, tid * per), f1 = min(a.cnt, f0 + per);
    uint32_t mine = 0;
    for (int f = f0; f < f1; ++f) mine += (uint32_t)a.gse.doff[f];
    uint32_t total;
    auto add = [](uint32_t x, uint32_t y) { return x + y; };
    uint32_t at = block_scan_exclusive(mine, 0u, add, s_w32, total);
    for (int f = f0; f < f1; ++f) {
        const uint32_t n = (uint32_t)a.gse.doff[f];
        a.gse.doff[f] = (int)at;
        at += n;
    }
    if (tid == 0) {
        a.gse.doff[a.cnt] = (int)total;
        a.gse.state->ndesc_wanted = (int)total;
        a.gse.state->ndesc = (int)total <= a.gse.cap ? (int)total : -1;   // -1: the pool is too small, the call is refused
        a.gse.state->malformed = 0;
    }
}

__global__ void __launch_bounds__(kGseThreads) gse_fill_kernel(const TsArgs a) {
    if (a.state->phase != 1 || a.gse.state->ndesc < 0) return;
    const int f = blockIdx.x * kGseThreads + threadIdx.x;
    if (f >= a.cnt) return;
    int bad = 0;
    gse_walk<true>(a, f, a.gse.desc + a.gse.doff[f], &bad);
    if (bad) atomicAdd(&a.gse.state->malformed, 1);
}

// crc32_checksum (:96-101) over a fragment's own bytes.  First fragment: from the all-ones start value over total
// length, protocol type, label and data, which lie back to back from byte 3 of the packet (:320-327).  Later
// fragments: from zero, together with x^(8 n) mod P, so that the sequential pass can continue the running value
// without the bytes; the last one has the four received CRC bytes (:343-346) folded in: the PDU is good when the
// continued value comes out as zero.
__global__ void __launch_bounds__(kGseThreads) gse_crc_kernel(const TsArgs a) {
    __shared__ uint32_t tab[256];
    if (a.state->phase != 1) return;
    const int nd = a.gse.state->ndesc;
    for (int i = threadIdx.x; i < 256; i += kGseThreads) {   // crc32_init (:85-94)
        uint32_t c = (uint32_t)i << 24;
#pragma unroll
        for (int b = 0; b < 8; ++b) c = (c & 0x80000000u) ? (c << 1) ^ kCrc32Poly : (c << 1);
        tab[i] = c;
    }
    __syncthreads();
    for (int d = blockIdx.x * kGseThreads + threadIdx.x; d < nd; d += gridDim.x * kGseThreads) {
        const GseDesc e = a.gse.desc[d];
        if (e.kind == 0) continue;
        uint32_t crc, xp = 1u;
        const uint8_t* p;
        int n;
        if (e.kind == 1) {
            crc = 0xFFFFFFFFu;
            p = a.bb + e.src + 3;
            n = (int)e.hdr - 3 + (int)e.len;
        } else {
            crc = 0u;
            p = a.bb + e.src + e.hdr;
            n = e.kind == 3 ? (int)e.len - 4 : (int)e.len;
        }
        for (int i = 0; i < n; ++i) {
            crc = (crc << 8) ^ tab[((crc >> 24) ^ p[i]) & 0xFFu];
            xp = (xp << 8) ^ tab[xp >> 24];
        }
        if (e.kind == 3) crc ^= ((uint32_t)p[n] << 24) | ((uint32_t)p[n + 1] << 16) | ((uint32_t)p[n + 2] << 8) | p[n + 3];
        a.gse.crc0[d] = crc;
        a.gse.xpow[d] = xp;
    }
}

// a(x) b(x) mod P(x), bit k = x^k
__device__ inline uint32_t crc32_mulmod(uint32_t x, uint32_t y) {
    uint32_t r = 0;
#pragma unroll 8
    for (int i = 31; i >= 0; --i) {
        r = (r << 1) ^ ((r & 0x80000000u) ? kCrc32Poly : 0u);
        if ((y >> i) & 1u) r ^= x;
    }
    return r;
}

// The slot logic of :313-334 (first), :363-377 (middle) and :335-362 (last fragment) over the descriptors in stream
// order.  One warp: the lanes fetch 32 descriptors at a time, every lane replays the fragments among them (complete
// PDUs do not touch the slots and are skipped by a ballot) with warp-uniform slot state and keeps the verdict of its
// own.  Which slot a fragment lands in, where its bytes go and which fragments form a chain does not depend on any
// CRC (a last fragment frees its slot either way, :341), so the CRCs are left to the next kernel.
__global__ void __launch_bounds__(32) gse_assemble_kernel(const TsArgs a) {
    if (a.state->phase != 1) return;
    GseState* G = a.gse.state;
    const int lane = threadIdx.x;
    const int nd = max(G->ndesc, 0);
    int on[3], id[3], proto[3], ctr[3], head[3], tail[3], carry[3], cont[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        on[r] = G->slot[r].on; id[r] = G->slot[r].id; proto[r] = G->slot[r].proto; ctr[r] = G->slot[r].ctr;
        head[r] = tail[r] = -1;
        cont[r] = on[r];                 // the chain continues one begun in an earlier call
        carry[r] = on[r] ? ctr[r] : 0;   // with this many bytes in the reassembly buffer
    }
    if (lane < 3) G->entry_crc[lane] = G->slot[lane].crc;
    int dropped = 0;
    for (int base = 0; base < nd; base += 32) {
        const int d = base + lane;
        const bool have = d < nd;
        uint32_t w0 = 0, w1 = 0;
        if (have) {
            const GseDesc e = a.gse.desc[d];
            w0 = (uint32_t)e.kind | (uint32_t)e.fragid << 8 | (uint32_t)e.proto << 16;
            w1 = e.len;
            a.gse.nxt[d] = -1;
        }
        __syncwarp();
        GseOut mine{0, 0, -1, -1};
        int myaux = -1, myaux2 = 0;
        if (have && (w0 & 0xFF) == 0) mine.emit = ((((w0 >> 16) == 0x0800u) || ((w0 >> 16) == 0x86DDu)) ? 4 : 2) + (int)w1;
        unsigned frag = __ballot_sync(0xFFFFFFFFu, have && (w0 & 0xFF) != 0);
        while (frag) {
            const int k = __ffs(frag) - 1;
            frag &= frag - 1;
            const uint32_t v0 = __shfl_sync(0xFFFFFFFFu, w0, k);
            const int len = (int)__shfl_sync(0xFFFFFFFFu, w1, k);
            const int kind = v0 & 0xFF, fid = (v0 >> 8) & 0xFF, pr = (int)(v0 >> 16), me = base + k;
            GseOut o{0, 0, -1, -1};
            int aux = -1, aux2 = 0;
            if (kind == 1) {
                int r = -1;
#pragma unroll
                for (int rr = 2; rr >= 0; --rr)
                    if (!on[rr] || id[rr] == fid) r = rr;
                if (r < 0) ++dropped;
#pragma unroll
                for (int rr = 0; rr < 3; ++rr)
                    if (rr == r) {
                        on[rr] = 1; id[rr] = fid; proto[rr] = pr; ctr[rr] = len;
                        head[rr] = tail[rr] = me;
                        carry[rr] = 0;
                        cont[rr] = 0;
                        o.pos = 0;
                    }
            } else {
                int r = -1;
#pragma unroll
                for (int rr = 2; rr >= 0; --rr)
                    if (on[rr] && id[rr] == fid) r = rr;
                bool taken = false;
#pragma unroll
                for (int rr = 0; rr < 3; ++rr)
                    if (rr == r && ctr[rr] + len <= kGseBuf) {
                        taken = true;
                        o.pos = ctr[rr];
                        if (tail[rr] >= 0) {
                            if (lane == 0) a.gse.nxt[tail[rr]] = me;
                        } else {
                            head[rr] = me;
                        }
                        tail[rr] = me;
                        if (kind == 2) {
                            ctr[rr] += len;
                        } else {
                            on[rr] = 0;
                            ctr[rr] += len - 4;
                            o.emit = ((proto[rr] == 0x0800 || proto[rr] == 0x86DD) ? 4 : 2) + ctr[rr];   // if the CRC-32 agrees
                            o.link = head[rr];
                            aux = cont[rr] ? (rr | carry[rr] << 2) : -1;
                            aux2 = proto[rr];
                        }
                    }
                if (!taken) ++dropped;
            }
            if (lane == k) {
                mine = o;
                myaux = aux;
                myaux2 = aux2;
            }
        }
        if (have) {
            a.gse.out[d] = mine;
            a.gse.aux[d] = myaux;
            a.gse.aux2[d] = myaux2;
        }
        __syncwarp();
    }
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            G->slot[r] = GseSlot{on[r], id[r], proto[r], ctr[r], G->slot[r].crc};
            G->save[r] = GseSave{head[r], carry[r], on[r], cont[r]};
        }
        G->old = G->cur;
        G->cur ^= 1;
        G->dropped = dropped;
    }
}

// running CRC-32 of a chain of fragments from their zero-start CRCs: crc(s, B) = s x^(8|B|) + crc(0, B)
__device__ inline uint32_t gse_chain_crc(const TsArgs& a, uint32_t crc, int x, int stop_at) {
    for (; x >= 0; x = a.gse.nxt[x]) {
        crc = crc32_mulmod(crc, a.gse.xpow[x]) ^ a.gse.crc0[x];
        if (x == stop_at) break;
    }
    return crc;
}

// A thread per last fragment: the CRC-32 over its chain decides whether the PDU goes out (:343-361).  Three more
// threads bring the running CRCs of the slots that stay open up to date.
__global__ void __launch_bounds__(kGseThreads) gse_verify_kernel(const TsArgs a) {
    if (a.state->phase != 1) return;
    GseState* G = a.gse.state;
    const int nd = max(G->ndesc, 0);
    for (int d = blockIdx.x * kGseThreads + threadIdx.x; d < nd + 3; d += gridDim.x * kGseThreads) {
        if (d >= nd) {
            const int r = d - nd;
            const GseSave sv = G->save[r];
            if (!sv.active) continue;
            if (sv.cont) G->slot[r].crc = gse_chain_crc(a, G->entry_crc[r], sv.head, -1);
            else G->slot[r].crc = gse_chain_crc(a, a.gse.crc0[sv.head], a.gse.nxt[sv.head], -1);
            continue;
        }
        if (a.gse.desc[d].kind != 3) continue;
        const GseOut o = a.gse.out[d];
        if (o.pos < 0) continue;   // no slot was waiting for it
        const int aux = a.gse.aux[d];
        const uint32_t crc = aux >= 0 ? gse_chain_crc(a, G->entry_crc[aux & 3], o.link, d)
                                      : gse_chain_crc(a, a.gse.crc0[o.link], a.gse.nxt[o.link], d);
        if (crc != 0) {   // (the received CRC-32 is folded into the last fragment's value)
            a.gse.out[d].emit = 0;
            a.gse.out[d].link = -2;   // marks the failure for the counters
        }
    }
}

// GSE bytes in front of every packet and every frame (a scan over what the packets put out), counters, and
// last_gse_crc_err = the verdict of the last reassembly that ended in this call
__global__ void __launch_bounds__(kPlanThreads) gse_offsets_kernel(const TsArgs a) {
    __shared__ uint32_t s_w32[kPlanThreads / 32];
    if (a.state->phase != 1) return;
    GseState* G = a.gse.state;
    const int nd = max(G->ndesc, 0), tid = threadIdx.x;
    const int per = (nd + kPlanThreads - 1) / kPlanThreads;
    const int d0 = min(nd, tid * per), d1 = min(nd, d0 + per);
    uint32_t bytes = 0, pdus = 0, errs = 0, last_end = 0;
    for (int d = d0; d < d1; ++d) {
        const GseOut o = a.gse.out[d];
        bytes += (uint32_t)o.emit;
        pdus += o.emit > 0;
        errs += o.link == -2;
        if (a.gse.desc[d].kind == 3 && o.pos >= 0) last_end = (uint32_t)(2 * d + 2 + (o.link == -2));
    }
    auto add = [](uint32_t x, uint32_t y) { return x + y; };
    auto umax = [](uint32_t x, uint32_t y) { return x > y ? x : y; };
    uint32_t total, pdus_total, errs_total, last_total;
    uint32_t at = block_scan_exclusive(bytes, 0u, add, s_w32, total);
    block_scan_exclusive(pdus, 0u, add, s_w32, pdus_total);
    block_scan_exclusive(errs, 0u, add, s_w32, errs_total);
    block_scan_exclusive(last_end, 0u, umax, s_w32, last_total);
    for (int d = d0; d < d1; ++d) {
        a.gse.out[d].before = (int)at;
        at += (uint32_t)a.gse.out[d].emit;
    }
    __syncthreads();
    // GSE bytes in front of every frame = in front of its first descriptor
    for (int f = tid; f <= a.cnt; f += kPlanThreads) {
        const int first = (f < a.cnt && G->ndesc >= 0) ? a.gse.doff[f] : nd;
        a.gse.before[f] = first < nd ? a.gse.out[first].before : (int)total;
    }
    if (tid == 0) {
        if (last_total) G->last_crc_err = (int)(last_total & 1u);
        G->pdus = (int)pdus_total;
        G->crc_errors = (int)errs_total;
        G->out_bytes = (int)total;
    }
}

__device__ inline void cta_copy(uint8_t* dst, const uint8_t* src, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}
// data bytes of the fragments of one chain, each to its place in the reassembled PDU
__device__ inline void gse_copy_chain(const TsArgs& a, uint8_t* base, int first, int stop_at) {
    for (int x = first; x >= 0; x = a.gse.nxt[x]) {
        const GseDesc e = a.gse.desc[x];
        cta_copy(base + a.gse.out[x].pos, a.bb + e.src + e.hdr, e.kind == 3 ? (int)e.len - 4 : (int)e.len);
        if (x == stop_at) break;
    }
}

// One CTA per delivered PDU: 2-byte zero GRE header, the protocol type when it is IPv4/IPv6 (:262-271,349-358), the
// payload.  Three more CTAs save what the open slots carry into the next call.
__global__ void __launch_bounds__(256) gse_copy_kernel(const TsArgs a) {
    const int phase = a.state->phase;
    if (phase != 2 && phase != 3) return;   // 3: no room for the output, but the slots moved on: their heads are still saved
    const GseState* G = a.gse.state;
    const int nd = max(G->ndesc, 0);
    const uint8_t* oldbuf = a.gse.buf + (size_t)G->old * 3 * kGseBuf;
    uint8_t* newbuf = a.gse.buf + (size_t)(G->old ^ 1) * 3 * kGseBuf;
    for (int d = blockIdx.x; d < nd + 3; d += gridDim.x) {
        if (d >= nd) {
            const int r = d - nd;
            const GseSave sv = G->save[r];
            if (!sv.active) continue;
            if (sv.carry > 0) cta_copy(newbuf + (size_t)r * kGseBuf, oldbuf + (size_t)r * kGseBuf, sv.carry);
            gse_copy_chain(a, newbuf + (size_t)r * kGseBuf, sv.head, -1);
            continue;
        }
        const GseOut o = a.gse.out[d];
        if (o.emit == 0 || phase != 2) continue;
        const GseDesc e = a.gse.desc[d];
        uint8_t* dst = a.out + a.plan[e.frame].out_off + (o.before - a.gse.before[e.frame]);
        const int proto = e.kind == 0 ? (int)e.proto : a.gse.aux2[d];
        const int hl = (proto == 0x0800 || proto == 0x86DD) ? 4 : 2;
        if (threadIdx.x < hl) dst[threadIdx.x] = threadIdx.x < 2 ? 0 : (threadIdx.x == 2 ? (uint8_t)(proto >> 8) : (uint8_t)proto);
        if (e.kind == 0) {
            cta_copy(dst + hl, a.bb + e.src + e.hdr, e.len);
        } else {
            const int aux = a.gse.aux[d];
            if (aux >= 0) cta_copy(dst + hl, oldbuf + (size_t)(aux & 3) * kGseBuf, aux >> 2);
            gse_copy_chain(a, dst + hl, o.link, d);
        }
    }
}

}  // namespace

int ts_launch(const TsArgs& a, cudaStream_t stream) {
    if (a.cnt > 0) ts_header_kernel<<<(a.cnt + kHeaderThreads - 1) / kHeaderThreads, kHeaderThreads, 0, stream>>>(a);
    ts_plan_kernel<<<1, kPlanThreads, 0, stream>>>(a);   // also for cnt == 0: the counters are per call
    if (a.cnt > 0) ts_copy_kernel<<<a.cnt, kCopyThreads, 0, stream>>>(a);
    return (int)cudaGetLastError();
}

int gse_launch_count(const TsArgs& a, cudaStream_t stream) {
    if (a.cnt > 0) gse_count_kernel<<<(a.cnt + kGseThreads - 1) / kGseThreads, kGseThreads, 0, stream>>>(a);
    gse_scan_kernel<<<1, kPlanThreads, 0, stream>>>(a);
    return (int)cudaGetLastError();
}

int gse_launch_rest(const TsArgs& a, cudaStream_t stream) {
    const int per_frame = (a.cnt + kGseThreads - 1) / kGseThreads;
    const int by_desc = std::max(1, std::min((a.gse.cap + kGseThreads - 1) / kGseThreads, 1184));
    if (a.cnt > 0) gse_fill_kernel<<<per_frame, kGseThreads, 0, stream>>>(a);
    gse_crc_kernel<<<by_desc, kGseThreads, 0, stream>>>(a);
    gse_assemble_kernel<<<1, 32, 0, stream>>>(a);
    gse_verify_kernel<<<by_desc, kGseThreads, 0, stream>>>(a);
    gse_offsets_kernel<<<1, kPlanThreads, 0, stream>>>(a);
    TsArgs b = a;
    b.mode = 2;
    b.copy_phase = 2;
    ts_plan_kernel<<<1, kPlanThreads, 0, stream>>>(b);
    if (a.cnt > 0) ts_copy_kernel<<<a.cnt, kCopyThreads, 0, stream>>>(b);
    gse_copy_kernel<<<std::max(1, std::min(a.gse.cap + 3, 2368)), 256, 0, stream>>>(b);
    return (int)cudaGetLastError();
}

}  // namespace s2
