// The LDPC kernels (templates).  Included by the instantiation units ldpc_inst_*.cu only; see ldpc_decoder.cuh for
// the interface and DESIGN.md for the design.
#pragma once
#include "ldpc_decoder.cuh"
#include "ldpc_variants.h"

namespace s2 {
namespace {


__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
// two int8 (bytes 0,1 / bytes 2,3 of w) -> sign-extended s16x2
__device__ __forceinline__ uint32_t unpack01(uint32_t w) { return prmt(w, 0, 0x9180); }
__device__ __forceinline__ uint32_t unpack23(uint32_t w) { return prmt(w, 0, 0xB3A2); }
// s16x2 holding int8-range values -> two bytes in the low half
__device__ __forceinline__ uint32_t pack1(uint32_t x) { return prmt(x, 0, 0x4420); }
__device__ __forceinline__ uint32_t pack2(uint32_t x0, uint32_t x1) { return prmt(x0, x1, 0x6420); }

constexpr uint32_t kM128 = 0xFF80FF80u, kP127 = 0x007F007Fu, kP32 = 0x00200020u, kP31 = 0x001F001Fu;
constexpr uint32_t kM1 = 0xFFFFFFFFu, kBig = 0x7FFF7FFFu;

__device__ __forceinline__ uint32_t sat8x2(uint32_t x) { return __vmins2(__vmaxs2(x, kM128), kP127); }

// per-byte "is zero" of a packed pair: bit 7 (frame A) / bit 15 (frame B) set iff that byte == 0
__device__ __forceinline__ uint32_t zero_bytes(uint32_t w) {
    return ~(((w & 0x7F7Fu) + 0x7F7Fu) | w) & 0x8080u;
}

// bits [o, o+32) of the 360-periodic extension of a 360-bit vector stored in 12 words (+1 zero word)
__device__ __forceinline__ uint32_t win360(const uint32_t* H, int o) {
    int w = o >> 5, s = o & 31;
    uint32_t r = __funnelshift_r(H[w], H[w + 1], s);
    int over = o + 32 - 360;
    if (over > 0) r |= H[0] << (32 - over);
    return r;
}


// Hard decisions + zero test of 8 consecutive offset-binary LLR pairs (one uint4: A0 B0 A1 B1 ...).
// Returns bits 0-7 = "A_k negative", bits 8-15 = "B_k negative"; clears bit 7+16f of every byte position in
// nzacc that sees a zero LLR (nzacc is ANDed: bytes 0,2 belong to frame A, bytes 1,3 to frame B).
__device__ __forceinline__ uint32_t harvest8(uint4 w, uint32_t& nzacc) {
    const uint32_t k80 = 0x80808080u;
    // |llr| per byte; +0x7F sets bit 7 of a byte iff |llr| >= 1 (|llr| <= 128: no carry between bytes)
    nzacc &= __vabsdiffu4(w.x, k80) + 0x7F7F7F7Fu;
    nzacc &= __vabsdiffu4(w.y, k80) + 0x7F7F7F7Fu;
    nzacc &= __vabsdiffu4(w.z, k80) + 0x7F7F7F7Fu;
    nzacc &= __vabsdiffu4(w.w, k80) + 0x7F7F7F7Fu;
    // gather the four A bytes / four B bytes of two words, then movemask by multiplication
    uint32_t a0 = prmt(w.x, w.y, 0x6420), b0 = prmt(w.x, w.y, 0x7531);
    uint32_t a1 = prmt(w.z, w.w, 0x6420), b1 = prmt(w.z, w.w, 0x7531);
    auto mask4 = [](uint32_t x) {   // bit k = (byte k of x is a negative LLR) = (offset-binary byte < 128)
        return (((~x & 0x80808080u) >> 7) * 0x00204081u >> 21) & 0xFu;
    };
    return mask4(a0) | (mask4(a1) << 4) | (mask4(b0) << 8) | (mask4(b1) << 12);
}
// bit planes of n360 groups of 360 LLR pairs starting at src (shared or global, 16-byte aligned):
// plane A at H[g*13 words], plane B at H[(gstride + g)*13 words], byte k of a group = bits 8k..8k+7
template <bool GLOBAL>
__device__ __forceinline__ void harvest_planes(const uint4* src, int n360, uint32_t* H, int gstride, int tid,
                                               uint32_t& nzacc) {
    uint8_t* Hb = reinterpret_cast<uint8_t*>(H);
    for (int t = tid; t < n360 * 45; t += kLdpcThreads) {
        uint4 w = GLOBAL ? __ldcg(src + t) : src[t];
        uint32_t bits = harvest8(w, nzacc);
        int g = t / 45, k = t - g * 45;
        Hb[g * (kBitWords * 4) + k] = (uint8_t)bits;
        Hb[(gstride + g) * (kBitWords * 4) + k] = (uint8_t)(bits >> 8);
    }
}

// ---- register formats of the row update -------------------------------------------------------
// LLRs are kept OFFSET-BINARY everywhere inside the kernel (shared memory, workspace, registers):
// u = llr + 128 in [0, 255], one frame per 16-bit lane.  With that bias the int8 saturation of the
// reference (vqadd / vqsub) is the single instruction  max(min(a + b, 255), 0)  = VIADDMNMX.RELU.
__device__ __forceinline__ uint32_t unpack_u01(uint32_t w) { return prmt(w, 0, 0x4140); }  // zero-extended bytes 0,1
__device__ __forceinline__ uint32_t sat_add_u8x2(uint32_t u, uint32_t d) { return __viaddmin_s16x2_relu(u, d, 0x00FF00FFu); }

// One check row for both frames of the pair.
//   voff[] : byte offsets into the shared LLR array of this row's data links (hoisted out of the level loop)
//   msg[]  : this row's CNT+2 message slots as packed int8, two slots per word (in/out)
//   pown   : packed offset-binary parity LLR pty[i][j]          (in/out)
//   psec   : packed offset-binary parity LLR of the second link (in/out, ignored when !has2)
// EXACT: every one of the CNT data links exists (cnt == CNT); otherwise links c >= cnt are skipped.
// BOTH : both frames still iterate; otherwise only frame `lf` (0/1) may change, the other keeps its LLRs.
//
// Maths (SURVEY.md spec S-LDPC, algorithms.hh:235-256,273-276), per 16-bit lane:
//   t = sat8(link - msg)                        -> tu = t + 128
//   a = |t|   (VABSDIFF4 against 128; the reference's  max(|max(t,-127)|-1, 0)  is a monotone function
//              of a, so the two smallest are found on a itself and the "-1, >= 0, <= 32" is applied after)
//   ex_k = min of a over the other links        (prefix/suffix minima; == "a_k == min0 ? min1 : min0")
//   mag_k = clamp(ex_k - 1, 0, 32);  msg_k' = sign * mag_k limited to [-32, 31];  link' = sat8(t + msg_k')
// MID  : second pass of a "middle" row of a chained layer (see chain_walk): link `xlink` takes its t from `xsaved`
//        (computed by row_pre from the LLR as it was before the layer) and its LLR is not written back, because
//        the row that shares the bit comes later in row order and writes the final value.
template <int CNT, bool EXACT, bool BOTH, bool MID = false>
__device__ __forceinline__ void row_update(uint8_t* __restrict__ vbytes, const int (&voff)[CNT], int cnt,
                                           uint32_t (&msg)[(CNT + 3) / 2], uint32_t& pown, uint32_t& psec,
                                           bool has2, int lf, int xlink = -1, uint32_t xsaved = 0) {
    constexpr int D = CNT + 2;
    constexpr uint32_t kNeutralT = 0x00FF00FFu;   // t = +127: never the minimum that matters, sign +
    uint32_t tu[D], a[D], mo_keep[D];
    // sign bookkeeping: bit 7 of tu is set for t >= 0.  For link k the product of the OTHER signs is
    // negative iff bit7(sx ^ tu_k) ^ parity(D + 1), sx = XOR of all tu (skipped links count as +).
    uint32_t sx = ((D + 1) & 1) ? 0x00800080u : 0u;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        uint32_t u;
        bool present = true;
        if (c < CNT) {
            present = EXACT || c < cnt;
            // (MID: the bit behind xlink is being written by the row that shares it -- not read at all here)
            u = (present && !(MID && c == xlink)) ? unpack_u01(*reinterpret_cast<const uint16_t*>(vbytes + voff[c])) : 0u;
        } else if (c == CNT) {
            u = unpack_u01(pown);
        } else {
            present = has2;
            u = unpack_u01(psec);
        }
        uint32_t mw = msg[c >> 1];
        uint32_t m = (c & 1) ? unpack23(mw) : unpack01(mw);
        uint32_t x = sat_add_u8x2(u, __vsub2(0u, m));        // t = sat8(link - msg), offset binary
        if (!(EXACT && c < CNT) && c != CNT) x = present ? x : kNeutralT;
        if (MID && c < CNT && c == xlink) x = xsaved;
        tu[c] = x;
        a[c] = __vabsdiffu4(x, 0x00800080u);                 // |t| in [0, 128]
        sx ^= x;
    }
    // exclusive minima of a[] by prefix/suffix scans
    uint32_t suf[D];
    suf[D - 1] = a[D - 1];
#pragma unroll
    for (int c = D - 2; c >= 1; --c) suf[c] = __vminu2(suf[c + 1], a[c]);
    uint32_t pre = a[0];
#pragma unroll
    for (int c = 0; c < D; ++c) {
        uint32_t ex;
        if (c == 0) {
            ex = suf[1];
        } else if (c == D - 1) {
            ex = pre;
        } else {
            ex = __vminu2(pre, suf[c + 1]);
            pre = __vminu2(pre, a[c]);
        }
        if (c < CNT && !EXACT && c >= cnt) continue;
        const uint32_t x = tu[c];
        uint32_t om = __viaddmin_s16x2_relu(ex, kM1, kP32);          // clamp(|t|min - 1, 0, 32)
        uint32_t neg = prmt(sx ^ x, 0, 0xAA88);                      // 0xFFFF in lanes whose outgoing sign is -
        uint32_t mo = __viaddmin_s16x2(om ^ neg, neg & 0x00010001u, kP31);  // +-om, limited to [-32, 31]
        uint32_t pk = pack1(sat_add_u8x2(x, mo));                    // link' = sat8(t + msg')
        if (c < CNT) {
            if (MID && c == xlink) {
                // not written: the bit's final value comes from the later row that shares it
            } else if (BOTH)
                *reinterpret_cast<uint16_t*>(vbytes + voff[c]) = (uint16_t)pk;
            else
                vbytes[voff[c] + lf] = (uint8_t)(pk >> (8 * lf));
        } else {
            uint32_t& dst = (c == CNT) ? pown : psec;
            if (BOTH) {
                dst = pk;
            } else {
                uint32_t keep = lf ? 0x00FFu : 0xFF00u;
                dst = (dst & keep) | (pk & ~keep & 0xFFFFu);
            }
        }
        mo_keep[c] = mo;
    }
    // new messages go back two slots per word, packed pairwise
#pragma unroll
    for (int c = 0; c < D; c += 2) {
        bool ok0 = EXACT || c >= CNT || c < cnt;
        bool ok1 = (c + 1 < D) && (EXACT || c + 1 >= CNT || c + 1 < cnt);
        uint32_t lo = ok0 ? mo_keep[c] : 0u, hi = ok1 ? mo_keep[c + 1 < D ? c + 1 : c] : 0u;
        msg[c >> 1] = pack2(lo, hi);
    }
}

// row_update with the smallest register footprint: instead of prefix/suffix minima (three arrays of D words) it
// keeps the two smallest |t| and picks per link ("|t| == min0 ? min1 : min0", the reference's own form,
// algorithms.hh:242-255).  A few more instructions per link, about 2 D fewer live registers: what lets the
// kernels with many links per row fit two CTAs on an SM without spilling.  No chained-layer modes.
template <int CNT, bool EXACT, bool BOTH>
__device__ __forceinline__ void row_update_lean(uint8_t* __restrict__ vbytes, const int (&voff)[CNT], int cnt,
                                                uint32_t (&msg)[(CNT + 3) / 2], uint32_t& pown, uint32_t& psec,
                                                bool has2, int lf) {
    constexpr int D = CNT + 2;
    constexpr uint32_t kNeutralT = 0x00FF00FFu;
    uint32_t tu[D];
    uint32_t sx = ((D + 1) & 1) ? 0x00800080u : 0u;
    uint32_t m1 = 0x00FF00FFu, m2 = 0x00FF00FFu;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        uint32_t u;
        bool present = true;
        if (c < CNT) {
            present = EXACT || c < cnt;
            u = present ? unpack_u01(*reinterpret_cast<const uint16_t*>(vbytes + voff[c])) : 0u;
        } else if (c == CNT) {
            u = unpack_u01(pown);
        } else {
            present = has2;
            u = unpack_u01(psec);
        }
        const uint32_t mw = msg[c >> 1];
        const uint32_t m = (c & 1) ? unpack23(mw) : unpack01(mw);
        uint32_t x = sat_add_u8x2(u, __vsub2(0u, m));
        if (!(EXACT && c < CNT) && c != CNT) x = present ? x : kNeutralT;
        tu[c] = x;
        const uint32_t a = __vabsdiffu4(x, 0x00800080u);
        m2 = __vminu2(m2, __vmaxu2(m1, a));
        m1 = __vminu2(m1, a);
        sx ^= x;
    }
    uint32_t prev = 0;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        uint32_t mo = 0;
        if (!(c < CNT && !EXACT && c >= cnt)) {
            const uint32_t x = tu[c];
            const uint32_t a = __vabsdiffu4(x, 0x00800080u);
            const uint32_t eq = __vcmpeq2(a, m1);
            const uint32_t ex = (m2 & eq) | (m1 & ~eq);
            const uint32_t om = __viaddmin_s16x2_relu(ex, kM1, kP32);
            const uint32_t neg = prmt(sx ^ x, 0, 0xAA88);
            mo = __viaddmin_s16x2(om ^ neg, neg & 0x00010001u, kP31);
            const uint32_t pk = pack1(sat_add_u8x2(x, mo));
            if (c < CNT) {
                if (BOTH)
                    *reinterpret_cast<uint16_t*>(vbytes + voff[c]) = (uint16_t)pk;
                else
                    vbytes[voff[c] + lf] = (uint8_t)(pk >> (8 * lf));
            } else {
                uint32_t& dst = (c == CNT) ? pown : psec;
                if (BOTH) {
                    dst = pk;
                } else {
                    const uint32_t keep = lf ? 0x00FFu : 0xFF00u;
                    dst = (dst & keep) | (pk & ~keep & 0xFFFFu);
                }
            }
        }
        if (c & 1)
            msg[c >> 1] = pack2(prev, mo);
        else if (c == D - 1)
            msg[c >> 1] = pack2(mo, 0u);
        prev = mo;
    }
}

// ---- chained layers ---------------------------------------------------------------------------
// A layer in which exactly two links X, Y fall into the same 360-bit group makes row j and row j+d share one bit
// (row j's X bit is row j+d's Y bit, d = (shift_Y - shift_X) mod 360 taken <= 180).  The reference visits rows in
// order, so row j+d must see that bit as row j left it.  Instead of one CTA barrier per dependency level (up to
// 360/d of them, with one warp working and eleven waiting), such a layer runs in three phases:
//   P  rows j < d ("sources", nobody before them) are updated normally; rows with a successor (j >= d, j+d < 360:
//      "middle") reduce everything that does not depend on the late Y input to four words (row_pre);
//   C  thread c < d walks its chain c+d, c+2d, ...: from the predecessor's X output it gets the row's Y input
//      and from that, in ten instructions, the row's X output, which it leaves in the LLR array;
//   F  every row j >= d is updated normally -- its Y bit now holds what the row before it in the chain produced
//      -- except that a middle row takes the t of its X link from phase P and leaves the X bit alone (MID).
// Rows j+d >= 360 ("sinks") have two late inputs (Y from the chain, X from source j+d-360) and no successor.
template <int CNT, bool EXACT>
__device__ __forceinline__ uint32_t row_pre(const uint8_t* __restrict__ vbytes, const int (&voff)[CNT], int cnt,
                                            const uint32_t (&msg)[(CNT + 3) / 2], uint32_t pown, uint32_t psec, bool has2,
                                            int X, int Y, uint4* slot) {
    constexpr int D = CNT + 2;
    constexpr uint32_t kNeutralT = 0x00FF00FFu;
    uint32_t sx = ((D + 1) & 1) ? 0x00800080u : 0u;   // as in row_update, without the t of X and Y
    uint32_t rmin = 0x00FF00FFu;                       // min |t| over the links other than X and Y
    uint32_t tux = 0, my = 0;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        const uint32_t mw = msg[c >> 1];
        const uint32_t m = (c & 1) ? unpack23(mw) : unpack01(mw);
        if (c < CNT && c == Y) {
            my = m;
            continue;
        }
        uint32_t u;
        bool present = true;
        if (c < CNT) {
            present = EXACT || c < cnt;
            u = present ? unpack_u01(*reinterpret_cast<const uint16_t*>(vbytes + voff[c])) : 0u;
        } else if (c == CNT) {
            u = unpack_u01(pown);
        } else {
            present = has2;
            u = unpack_u01(psec);
        }
        uint32_t x = sat_add_u8x2(u, __vsub2(0u, m));
        if (!(EXACT && c < CNT) && c != CNT) x = present ? x : kNeutralT;
        if (c < CNT && c == X) {
            tux = x;
            continue;
        }
        sx ^= x;
        rmin = __vminu2(rmin, __vabsdiffu4(x, 0x00800080u));
    }
    *slot = make_uint4(rmin, sx, tux, my);
    return tux;
}

// phase C for the chain that starts at source row c0; lkx = link word of X ((group << 16) | shift)
__device__ __forceinline__ void chain_walk(uint8_t* __restrict__ vbytes, const uint4* __restrict__ scratch, uint32_t lkx,
                                           int c0, int d) {
    uint8_t* grp = vbytes + 2 * 360 * (int)(lkx >> 16);
    int m = c0 - (int)(lkx & 0xFFFFu);                 // X bit of row c0 inside the group
    m += (m < 0) ? 360 : 0;
    uint32_t u = unpack_u01(*reinterpret_cast<const uint16_t*>(grp + 2 * m));   // what the source left there
    for (int r = c0 + d; r + d < 360; r += d) {
        m += d;                                        // X bit of row r
        m -= (m >= 360) ? 360 : 0;
        const uint4 sc = scratch[r];                   // rmin, sx, t_X, old message of Y
        const uint32_t ty = sat_add_u8x2(u, __vsub2(0u, sc.w));
        const uint32_t ex = __vminu2(sc.x, __vabsdiffu4(ty, 0x00800080u));       // min |t| over all links but X
        const uint32_t om = __viaddmin_s16x2_relu(ex, kM1, kP32);
        const uint32_t neg = prmt(sc.y ^ ty, 0, 0xAA88);                         // sign over all links but X
        const uint32_t mo = __viaddmin_s16x2(om ^ neg, neg & 0x00010001u, kP31);
        const uint32_t pk = pack1(sat_add_u8x2(sc.z, mo));
        *reinterpret_cast<uint16_t*>(grp + 2 * m) = (uint16_t)pk;
        u = unpack_u01(pk);
    }
}

// Streamed input: block until the copy engine has delivered `need` frames.  Kept out of line so that the decoder
// around the call site compiles to the same code as without it.
__device__ __noinline__ void wait_arrived(const unsigned int* arrived_ptr, unsigned need) {
    const volatile unsigned* arrived = arrived_ptr;
    const long long t0 = clock64();
    while (*arrived < need) {
        __nanosleep(500);
        if (clock64() - t0 > (10ll << 30)) __trap();   // ~5 s: the host never sent the data
    }
}

// STREAMED: the input copy is still running when the kernel starts (LdpcArgs::arrived); a separate instantiation,
// because the waiting code in the pair hand-out measurably perturbs the scheduling of the resident-input kernel.
// CHAINS: chained layers run in three phases (see chain_walk) instead of level by level; again its own
// instantiation, chosen per code where it measurably pays (ldpc_chains_pay_off).
// OCC: resident CTAs per SM the register allocation aims at (one more than the default pays for the codes whose
// shared memory lets the extra CTA in, and costs the others: ldpc_ctas_wanted3).
template <int CNT, bool UNIFORM, bool STREAMED, bool CHAINS, int OCC = (CNT <= 9 ? 2 : 1)>
__global__ void __launch_bounds__(kLdpcThreads, OCC) ldpc_pair_kernel(const __grid_constant__ LdpcParams p) {
    constexpr int SLOTS = CNT + 2;
    constexpr int MW = (SLOTS + 1) / 2;     // message words per row in registers
    constexpr int SG = (SLOTS + 7) / 8;     // uint4 groups per row in the workspace
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint16_t* vdata = reinterpret_cast<uint16_t*>(smem_raw);
    uint32_t* HD = reinterpret_cast<uint32_t*>(smem_raw + (size_t)p.K * 2);  // [2][ngroups][13]
    uint32_t* HP = HD + 2 * p.ngroups * kBitWords;                           // [2][q][13]
    __shared__ unsigned int s_pair;
    __shared__ int s_bad[2];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int j = tid;
    const bool active = j < 360;
    const int q = p.q, K = p.K, N = p.N, R = p.R;
    uint4* wmsg = reinterpret_cast<uint4*>(p.workspace + (size_t)blockIdx.x * p.ws_stride);
    uint16_t* wpty = reinterpret_cast<uint16_t*>(p.workspace + (size_t)blockIdx.x * p.ws_stride +
                                                 (size_t)q * SG * 360 * 16);
    const int npairs = (p.nframes + 1) >> 1;
    // four words per row handed from phase P to phase C of a chained layer, behind the bit planes
    uint4* chain_scratch = reinterpret_cast<uint4*>(smem_raw + ((((size_t)K * 2 + (size_t)2 * (p.ngroups + q) * kBitWords * 4) + 15) & ~(size_t)15));

    // zero the bit planes once: bytes 45..51 of every 360-bit group are never written and must read as 0
    for (int x = tid; x < 2 * (p.ngroups + q) * kBitWords; x += kLdpcThreads) HD[x] = 0;

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            const unsigned np = atomicAdd(p.work_counter, 1u);
            s_pair = np;
            if (STREAMED && np < (unsigned)npairs) {
                // streamed input: the copy engine raises *arrived behind every piece of frames it has delivered
                // (same stream, so the data is in memory before the count); pairs are handed out in frame order,
                // so every waiter is waiting for a copy that is already queued
                wait_arrived(p.arrived, min(2u * np + 2u, (unsigned)p.nframes));
                __threadfence();
            }
        }
        __syncthreads();
        const unsigned pair = s_pair;
        if (pair >= (unsigned)npairs) break;
        const int fa = 2 * pair, fb = 2 * pair + 1;
        const bool hasB = fb < p.nframes;
        const int8_t* inA = p.llr_in + (size_t)fa * N;
        const int8_t* inB = p.llr_in + (size_t)(hasB ? fb : fa) * N;

        // ---- load: systematic LLRs -> shared (A in even bytes, B in odd), parity LLRs -> workspace,
        //      permuted to layered order pty[360 i + j] = v[K + q j + i] (layered_decoder.hh:124-126)
        for (int x = tid; x < K / 8; x += kLdpcThreads) {
            // streamed: the copy engine is still writing other frames of this buffer -> L2-coherent loads
            uint2 a = STREAMED ? __ldcg(reinterpret_cast<const uint2*>(inA) + x) : __ldg(reinterpret_cast<const uint2*>(inA) + x);
            uint2 b = STREAMED ? __ldcg(reinterpret_cast<const uint2*>(inB) + x) : __ldg(reinterpret_cast<const uint2*>(inB) + x);
            uint4 o;
            o.x = prmt(a.x, b.x, 0x5140) ^ 0x80808080u;   // interleave A/B bytes, to offset binary
            o.y = prmt(a.x, b.x, 0x7362) ^ 0x80808080u;
            o.z = prmt(a.y, b.y, 0x5140) ^ 0x80808080u;
            o.w = prmt(a.y, b.y, 0x7362) ^ 0x80808080u;
            reinterpret_cast<uint4*>(vdata)[x] = o;
        }
        // The parity part is a q x 360 transpose.  It goes through shared memory in tiles of 32 columns (the
        // bit-plane area is free at this point): 64-bit coalesced reads of v[K + q j + i], byte scatter into the
        // tile, then rows of 32 pairs (64 B) out to the workspace -- instead of 2R single-byte loads and R
        // isolated 2-byte stores.
        {
            uint8_t* tile = reinterpret_cast<uint8_t*>(HD);           // [q][32] pairs of bytes
            const uint2* srcA = reinterpret_cast<const uint2*>(inA + K);
            const uint2* srcB = reinterpret_cast<const uint2*>(inB + K);
            for (int jj0 = 0; jj0 < 360; jj0 += 32) {
                const int ncol = min(32, 360 - jj0);
                const int nvec = ncol * q / 8;                        // q * 32 and q * 8 are multiples of 8
                for (int x = tid; x < nvec; x += kLdpcThreads) {
                    const int base = (q * jj0) / 8 + x;
                    uint2 a = STREAMED ? __ldcg(srcA + base) : __ldg(srcA + base), b = STREAMED ? __ldcg(srcB + base) : __ldg(srcB + base);
                    int jr = (8 * x) / q, ii = 8 * x - jr * q;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        uint32_t va = ((k < 4 ? a.x : a.y) >> (8 * (k & 3))) & 0xFFu;
                        uint32_t vb = ((k < 4 ? b.x : b.y) >> (8 * (k & 3))) & 0xFFu;
                        *reinterpret_cast<uint16_t*>(tile + 2 * (ii * 32 + jr)) = (uint16_t)((va | (vb << 8)) ^ 0x8080u);
                        if (++ii == q) {
                            ii = 0;
                            ++jr;
                        }
                    }
                }
                __syncthreads();
                for (int x = tid; x < q * 32; x += kLdpcThreads) {
                    const int ii = x >> 5, jr = x & 31;
                    if (jr < ncol) __stcg(&wpty[360 * ii + jj0 + jr], *reinterpret_cast<const uint16_t*>(tile + 2 * x));
                }
                __syncthreads();
            }
            // the tile lived in the bit-plane area: restore the all-zero state the planes rely on
            for (int x = tid; x < 2 * (p.ngroups + q) * kBitWords; x += kLdpcThreads) HD[x] = 0;
        }
        __syncthreads();

        int live = hasB ? 3 : 1;      // bit f set: frame f still iterating
        int resA = -1, resB = -1;
        uint32_t nzacc = 0xFFFFFFFFu; // bit 7 of byte f (+16) cleared once frame f shows a zero LLR
        // hard decisions + zero test of the parity part, straight from the workspace
        harvest_planes<true>(reinterpret_cast<const uint4*>(wpty), q, HP, q, tid, nzacc);

        for (int n = 0;; ++n) {
            // ---- hard decisions + zero test of the systematic part
            harvest_planes<false>(reinterpret_cast<const uint4*>(vdata), p.ngroups, HD, p.ngroups, tid, nzacc);
            if (tid < 2) s_bad[tid] = 0;
            __syncthreads();
            // ---- LDPCDecoder::bad (layered_decoder.hh:28-45) on bit planes: a row is bad when its sign
            //      product is not positive, i.e. odd parity of hard decisions or any zero LLR
            if (~nzacc & 0x00800080u) atomicOr(&s_bad[0], 1);
            if (~nzacc & 0x80008000u) atomicOr(&s_bad[1], 1);
            nzacc = 0xFFFFFFFFu;
            for (int task = tid; task < q * 12; task += kLdpcThreads) {
                int ii = task / 12, w = task - ii * 12;
                const uint32_t* L = &p.links[p.layer_off[ii]];
                int cnt = p.layer_off[ii + 1] - p.layer_off[ii];
                uint32_t sA = HP[ii * kBitWords + w], sB = HP[(q + ii) * kBitWords + w];
                if (ii > 0) {
                    sA ^= HP[(ii - 1) * kBitWords + w];
                    sB ^= HP[(q + ii - 1) * kBitWords + w];
                } else {  // row (0,j) uses pty[q-1][j-1], row (0,0) has no second parity link
                    const uint32_t* ta = &HP[(q - 1) * kBitWords];
                    const uint32_t* tb = &HP[(2 * q - 1) * kBitWords];
                    sA ^= (ta[w] << 1) | (w ? ta[w - 1] >> 31 : 0u);
                    sB ^= (tb[w] << 1) | (w ? tb[w - 1] >> 31 : 0u);
                }
                for (int c = 0; c < cnt; ++c) {
                    uint32_t lk = L[c];
                    int o = 32 * w - (int)(lk & 0xFFFFu);
                    o += (o < 0) ? 360 : 0;
                    int g = lk >> 16;
                    sA ^= win360(&HD[g * kBitWords], o);
                    sB ^= win360(&HD[(p.ngroups + g) * kBitWords], o);
                }
                if (w == 11) {
                    sA &= 0xFFu;
                    sB &= 0xFFu;
                }
                if (sA) atomicOr(&s_bad[0], 1);
                if (sB) atomicOr(&s_bad[1], 1);
            }
            __syncthreads();
            // ---- while (bad() && --trials >= 0) update();  (layered_decoder.hh:127-128), per frame
            {
                int badA = s_bad[0], badB = s_bad[1];
                if ((live & 1) && !badA) { resA = n; live &= ~1; }
                if ((live & 2) && !badB) { resB = n; live &= ~2; }
                if (n == p.max_trials) live = 0;  // frames still bad after max_trials updates report -1
            }
            if (!live) break;

            // ---- LDPCDecoder::update (layered_decoder.hh:46-74): one pass over all layers
            uint32_t msg[MW];
#pragma unroll
            for (int x = 0; x < MW; ++x) msg[x] = 0;
            uint4 nxt[SG];
            uint32_t pown = 0, psec = 0, pnext = 0;
            const bool first = (n == 0);
            if (active) {
                pown = __ldcg(&wpty[j]);
                if (j > 0) psec = __ldcg(&wpty[360 * (q - 1) + j - 1]);
                if (!first) {
#pragma unroll
                    for (int s = 0; s < SG; ++s) nxt[s] = __ldcg(&wmsg[(size_t)s * 360 + j]);
                }
            }
            int mylev_next = (active && p.layer_nlev[0] > 1) ? p.row_level[j] : 0;
            for (int i = 0; i < q; ++i) {
                if (active) {
                    if (!first) {
#pragma unroll
                        for (int s = 0; s < SG; ++s) {
                            if (4 * s + 0 < MW) msg[4 * s + 0] = nxt[s].x;
                            if (4 * s + 1 < MW) msg[4 * s + 1] = nxt[s].y;
                            if (4 * s + 2 < MW) msg[4 * s + 2] = nxt[s].z;
                            if (4 * s + 3 < MW) msg[4 * s + 3] = nxt[s].w;
                        }
                    } else {
#pragma unroll
                        for (int x = 0; x < MW; ++x) msg[x] = 0;
                    }
                    if (i + 1 < q) {  // prefetch the next layer's row while this one computes
                        pnext = __ldcg(&wpty[360 * (i + 1) + j]);
                        if (!first) {
#pragma unroll
                            for (int s = 0; s < SG; ++s)
                                nxt[s] = __ldcg(&wmsg[((size_t)(i + 1) * SG + s) * 360 + j]);
                        }
                    }
                }
                const int loff = p.layer_off[i];
                const int cnt = UNIFORM ? CNT : (int)p.layer_off[i + 1] - loff;
                const int nlev = p.layer_nlev[i];
                const int mylev = mylev_next;
                if (i + 1 < q && active && p.layer_nlev[i + 1] > 1) mylev_next = p.row_level[(i + 1) * 360 + j];
                else mylev_next = 0;
                const bool has2 = (i | j) != 0;
                int voff[CNT];   // byte offset of each data link's LLR pair: 2 * (360 g + (j - shift) mod 360)
#pragma unroll
                for (int c = 0; c < CNT; ++c) {
                    uint32_t lk = p.links[loff + ((UNIFORM || c < cnt) ? c : 0)];
                    int m = j - (int)(lk & 0xFFFFu);
                    m += (m < 0) ? 360 : 0;
                    voff[c] = 2 * ((int)(lk >> 16) * 360 + m);
                }
                uint8_t* vbytes = reinterpret_cast<uint8_t*>(vdata);
                const int lf = (live == 2) ? 1 : 0;
                const int chain = CHAINS ? p.layer_chain[i] : 0;
                if (CHAINS && chain && live == 3) {   // chained layer, both frames iterating: phases P, C, F (see chain_walk)
                    const int d = (chain >> 6) & 0x1FF, first = chain & 31;
                    const int X = first + ((chain >> 5) & 1), Y = first + 1 - ((chain >> 5) & 1);
                    const bool source = j < d, middle = !source && j + d < 360;
                    uint32_t xsaved = 0;
                    if (active) {
                        if (source)
                            row_update<CNT, UNIFORM, true>(vbytes, voff, cnt, msg, pown, psec, has2, 0);
                        else if (middle)
                            xsaved = row_pre<CNT, UNIFORM>(vbytes, voff, cnt, msg, pown, psec, has2, X, Y, &chain_scratch[j]);
                    }
                    __syncthreads();
                    if (source) chain_walk(vbytes, chain_scratch, p.links[loff + X], j, d);
                    __syncthreads();
                    if (active && !source) {
                        if (middle)
                            row_update<CNT, UNIFORM, true, true>(vbytes, voff, cnt, msg, pown, psec, has2, 0, X, xsaved);
                        else
                            row_update<CNT, UNIFORM, true>(vbytes, voff, cnt, msg, pown, psec, has2, 0);
                    }
                    __syncthreads();
                } else
                for (int lvl = 0; lvl < nlev; ++lvl) {
                    if (active && mylev == lvl) {
                        constexpr bool LEAN = CNT > 9 && OCC == 2 && !CHAINS;   // two CTAs where one used to be
                        if (LEAN) {
                            if (live == 3)
                                row_update_lean<CNT, UNIFORM, true>(vbytes, voff, cnt, msg, pown, psec, has2, 0);
                            else
                                row_update_lean<CNT, UNIFORM, false>(vbytes, voff, cnt, msg, pown, psec, has2, lf);
                        } else if (live == 3)
                            row_update<CNT, UNIFORM, true>(vbytes, voff, cnt, msg, pown, psec, has2, 0);
                        else
                            row_update<CNT, UNIFORM, false>(vbytes, voff, cnt, msg, pown, psec, has2, lf);
                    }
                    // barriers separate dependency levels, and layers only where a later layer touches a bit group
                    // that a layer since the last barrier also touches (rows on disjoint bits commute)
                    if (lvl + 1 < nlev || p.layer_sync[i]) __syncthreads();
                }
                // write the row's messages back; retire the parity LLR that just got its last update
                if (active) {
#pragma unroll
                    for (int s = 0; s < SG; ++s) {
                        uint4 o;
                        o.x = (4 * s + 0 < MW) ? msg[4 * s + 0] : 0u;
                        o.y = (4 * s + 1 < MW) ? msg[4 * s + 1] : 0u;
                        o.z = (4 * s + 2 < MW) ? msg[4 * s + 2] : 0u;
                        o.w = (4 * s + 3 < MW) ? msg[4 * s + 3] : 0u;
                        __stcg(&wmsg[((size_t)i * SG + s) * 360 + j], o);
                    }
                    if (i == 0) {
                        if (j > 0) __stcg(&wpty[360 * (q - 1) + j - 1], (uint16_t)psec);  // updated again in layer q-1
                    } else {
                        __stcg(&wpty[360 * (i - 1) + j], (uint16_t)psec);
                    }
                }
                psec = pown;
                pown = pnext;
            }
            // pty[q-1][j]: its second link was served in layer 0, the own link just now -> final
            if (active) __stcg(&wpty[360 * (q - 1) + j], (uint16_t)psec);
            __syncthreads();
            // hard decisions + zero test of the parity part after this pass
            harvest_planes<true>(reinterpret_cast<const uint4*>(wpty), q, HP, q, tid, nzacc);
        }

        // ---- results: iteration counts, MSB-first hard decisions of the K systematic bits
        if (tid == 0) {
            p.iters_out[fa] = (int16_t)resA;
            if (hasB) p.iters_out[fb] = (int16_t)resB;
        }
        const int kbytes = K / 8;
        for (int x = tid; x < 2 * kbytes; x += kLdpcThreads) {
            int f = x >= kbytes, b = x - f * kbytes;
            if (f && !hasB) break;
            int g = b / 45, k = b - g * 45;
            uint32_t word = HD[(f * p.ngroups + g) * kBitWords + (k >> 2)];
            uint32_t byte = (word >> (8 * (k & 3))) & 0xFFu;
            p.hard_out[(size_t)(f ? fb : fa) * p.hard_stride + b] = (uint8_t)(__brev(byte) >> 24);
        }
        if (p.llr_out) {
            for (int x = tid; x < K; x += kLdpcThreads) {
                uint32_t v = vdata[x] ^ 0x8080u;
                p.llr_out[(size_t)fa * N + x] = (int8_t)(v & 0xFF);
                if (hasB) p.llr_out[(size_t)fb * N + x] = (int8_t)(v >> 8);
            }
            for (int x = tid; x < R; x += kLdpcThreads) {
                int jj = x / q, ii = x - jj * q;
                uint32_t v = __ldcg(&wpty[360 * ii + jj]) ^ 0x8080u;
                p.llr_out[(size_t)fa * N + K + x] = (int8_t)(v & 0xFF);
                if (hasB) p.llr_out[(size_t)fb * N + K + x] = (int8_t)(v >> 8);
            }
        }
    }
}

}  // namespace
}  // namespace s2

#define K2(c, u, st, ch) {ldpc_pair_kernel<c, u, st, ch>, ldpc_pair_kernel<c, u, st, ch, (c <= 9 ? 3 : 2)>}
#define K8(c, u) {{K2(c, u, false, false), K2(c, u, false, true)}, {K2(c, u, true, false), K2(c, u, true, true)}}
#define N8 {{{nullptr, nullptr}, {nullptr, nullptr}}, {{nullptr, nullptr}, {nullptr, nullptr}}}
#define VU(c) {c, K8(c, true), N8}
#define VB(c) {c, K8(c, true), K8(c, false)}
#define VR(c) {c, N8, K8(c, false)}
