// K6: BBFRAME -> MPEG-TS packets.  See ts_parser.cuh.
//
// Three kernels.  ts_header_kernel checks the BBHEADERs, a thread per frame.  ts_plan_kernel (one CTA) works out
// what every frame contributes; that depends on the parser state the frames before it left behind (in sync or not, bytes of an
// unfinished unit), which a three-phase scan over composable state maps delivers without walking the frames
// one by one.  ts_copy_kernel (one CTA per frame, a warp per packet) moves the bytes.
// HBM-bound: every BBFRAME byte is read once and every TS byte written once.
#include "ts_parser.cuh"

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
__device__ inline uint32_t pack_meta(int kind, int dfl, int syncd) { return (uint32_t)kind | ((uint32_t)(dfl >> 3) << 2) | ((uint32_t)((syncd >> 3) + 1) << 15); }

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
        const int skip = (int)(m >> 15);
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
                    const int skip = (int)(m >> 15);
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
    a.meta[f] = pack_meta(kind, dfl, syncd);
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
    int npk = 0, valid = 0, gse = 0, lastv = -1;
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
    block_scan_exclusive((uint32_t)(lastv + 1), 0u, umax, s_w32, lastv_total);
    const unsigned long long writer_before = block_scan_exclusive(writer, 0ull, later, s_w64, writer_total);
    // enough room for everything (no test of :176/:208 can fail)?  Otherwise one thread redoes it in frame order.
    const bool roomy = a.out_cap - 188 * (int)npk_total > 188;
    if (tid == 0) {
        S->entry_buf = S->cur;
        s_fin[0] = final_state;
        s_fin[1] = 188 * (int)npk_total;
        s_fin[2] = (int)valid_total;
        s_fin[3] = (int)gse_total;
        s_fin[4] = (int)lastv_total - 1;
        s_fin[5] = (int)(writer_total >> 32) - 1;
        s_fin[6] = (int)(writer_total & 0xFFFFFFFFu);
        if (!roomy) ts_plan_serial(a, M, st0, s_fin[1], s_fin[2], s_fin[3], s_fin[4], s_fin[0], s_fin[5], s_fin[6]);
    }
    // (5) replay once more, now writing what every frame contributes
    if (roomy) {
        int st = entry, o = (int)npk_before;
        int wsrc = (int)(writer_before >> 32) - 1, woff = (int)(writer_before & 0xFFFFFFFFu);
        for (int f = f0; f < f1; ++f) {
            const Step r = ts_step(M[f], st);
            TsPlan p{188 * o, r.off, -1, 0, (short)r.npk, 0};
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
        S->produced = s_fin[1];
        if (a.produced_out) *a.produced_out = s_fin[1];
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
    const TsPlan p = a.plan[f];
    if (p.npk == 0) return;
    const uint8_t* fr = a.bb + (size_t)f * a.kb;
    const uint8_t* carry = nullptr;
    if (p.head) carry = p.head_src < 0 ? a.state->unit[a.state->entry_buf] : a.bb + (size_t)p.head_src * a.kb + p.head_src_off;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool words = ((uintptr_t)a.out & 3) == 0;
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

}  // namespace

int ts_launch(const TsArgs& a, cudaStream_t stream) {
    if (a.cnt > 0) ts_header_kernel<<<(a.cnt + kHeaderThreads - 1) / kHeaderThreads, kHeaderThreads, 0, stream>>>(a);
    ts_plan_kernel<<<1, kPlanThreads, 0, stream>>>(a);   // also for cnt == 0: the counters are per call
    if (a.cnt > 0) ts_copy_kernel<<<a.cnt, kCopyThreads, 0, stream>>>(a);
    return (int)cudaGetLastError();
}

}  // namespace s2
