// K6: BBFRAME -> MPEG-TS packets.  See ts_parser.cuh.
//
// Two kernels.  ts_plan_kernel (one CTA): all threads check BBHEADERs in parallel, 256 frames at a time; one
// thread then walks the 256 verdicts in order, because what a frame contributes depends on the parser state the
// frames before it left behind (in sync or not, bytes of an unfinished unit, room left in the output) -- a few
// register operations per frame.  ts_copy_kernel (one CTA per frame, a warp per packet) moves the bytes.
// HBM-bound: every BBFRAME byte is read once and every TS byte written once.
#include "ts_parser.cuh"

namespace s2 {
namespace {

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

__global__ void __launch_bounds__(kPlanThreads) ts_plan_kernel(const TsArgs a) {
    __shared__ uint16_t s_dfl[kPlanThreads], s_syncd[kPlanThreads];
    __shared__ uint8_t s_kind[kPlanThreads];
    __shared__ int s_carry[3];   // frame, offset, bytes of the carry this call ends with
    const int tid = threadIdx.x;
    TsState* S = a.state;
    // parser state, live in thread 0 only
    unsigned count = 0;
    int synched = 0, car_src = -1, car_off = 0, o = 0, processed = 0, gse = 0, last_valid = -1;
    bool stop = false;
    if (tid == 0) {
        count = S->count;
        synched = S->synched;
        S->entry_buf = S->cur;
    }
    for (int base = 0; base < a.cnt; base += kPlanThreads) {
        const int f = base + tid;
        if (f < a.cnt) {
            const uint8_t* h = a.bb + (size_t)f * a.kb;
            const int dfl = (h[4] << 8) | h[5], syncd = (h[7] << 8) | h[8];
            int kind = kInvalid;
            // (:122-150) CRC-8, DFL <= kbch-80, SYNCD < DFL-8 as signed ints, DFL a whole number of bytes
            if (bbheader_crc8(h) == 0 && dfl <= a.max_dfl && syncd < dfl - 8 && (dfl & 7) == 0) {
                const int ts_gs = h[0] >> 6;
                kind = ts_gs == 3 ? kTs : ts_gs == 1 ? kGse : kOther;
            }
            s_kind[tid] = (uint8_t)kind;
            s_dfl[tid] = (uint16_t)dfl;
            s_syncd[tid] = (uint16_t)syncd;
        }
        __syncthreads();
        if (tid == 0) {
            const int m = min(kPlanThreads, a.cnt - base);
            for (int k = 0; k < m; ++k) {
                TsPlan p{o, 0, -1, 0, 0, 0};
                const int kind = s_kind[k];
                if (!stop) {
                    if (kind == kInvalid) {
                        synched = 0;
                    } else {
                        int left = s_dfl[k] >> 3, off = 10;
                        if (!synched) {   // enter just past the first sync byte (:157-168)
                            const int skip = (s_syncd[k] >> 3) + 1;
                            off += skip;
                            left -= skip;
                            count = 0;
                            synched = 1;
                        }
                        last_valid = base + k;
                        ++processed;
                        gse += kind == kGse;
                        if (kind == kTs) {   // (:173-212)
                            const int room = a.out_cap - o;
                            int consumed = 0;
                            p.src_off = off;
                            if (left >= 188 && room > 188) {
                                int fit = (room - 189) / 188 + 1;
                                if (count > 0) {
                                    p.head = (short)count;
                                    p.head_src = car_src;
                                    p.head_src_off = car_off;
                                    consumed = 188 - (int)count;
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
                            if (left > 0) {
                                count = (unsigned)left;
                                car_src = base + k;
                                car_off = off + consumed;
                            }
                            if (a.out_cap - o <= 188) stop = true;
                        }
                    }
                }
                a.plan[base + k] = p;
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        S->count = count;
        S->synched = synched;
        S->last_bb_cnt = a.cnt;
        S->last_bb_proc = processed;
        S->gse_frames = gse;
        S->produced = o;
        if (a.produced_out) *a.produced_out = o;
        if (last_valid >= 0) {
            S->have_header = 1;
            for (int i = 0; i < 10; ++i) S->last_header[i] = a.bb[(size_t)last_valid * a.kb + i];
        }
        s_carry[0] = car_src;
        s_carry[1] = car_off;
        s_carry[2] = (int)min(count, 188u);
        if (car_src >= 0) S->cur ^= 1;
    }
    __syncthreads();
    if (s_carry[0] >= 0 && tid < s_carry[2])
        S->unit[S->entry_buf ^ 1][tid] = a.bb[(size_t)s_carry[0] * a.kb + s_carry[1] + tid];
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
    for (int u = warp; u < p.npk; u += kCopyThreads / 32) {
        uint8_t* dst = a.out + p.out_off + 188 * u;
        // frame offset of the unit's byte 0; the first `head` bytes of unit 0 lie in the carry instead
        const int ustart = p.src_off + 188 * u - p.head;
        const int head = u == 0 ? p.head : 0;
        for (int w = lane; w < 47; w += 32) {
            uint32_t v = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = 4 * w + k - 1;   // unit byte behind output byte 4w+k; the unit's 188th byte is dropped
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
    ts_plan_kernel<<<1, kPlanThreads, 0, stream>>>(a);   // also for cnt == 0: the counters are per call
    if (a.cnt > 0) ts_copy_kernel<<<a.cnt, kCopyThreads, 0, stream>>>(a);
    return (int)cudaGetLastError();
}

}  // namespace s2
