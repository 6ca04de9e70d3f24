// K8: the payload phase loop on the device (the part with loop state).
//
// Replaces S2PLLBlock::process (dvbs2/dvbs2_pll.cpp:34-86): every symbol of a frame is turned back by the loop's
// phase, the header symbols are compared with the known SOF / PLS symbols, the payload symbols with the nearest
// constellation point (the phase_error of the demapper's LUT cell, constellation.cpp:293-322) or -- in the stretches
// the reference's pilot counter marks -- with the nearest diagonal, the error advances a second-order loop
// (PhaseControlLoop::advance), and the PL-descrambled symbol goes out.
//
// The loop is a recurrence over the symbols: phase[n+1] = F(phase[n], freq[n], in[n]).  The reference walks it one
// symbol at a time; nothing about it is associative.  What makes it parallel all the same is that, for a payload
// symbol, the error is a TABLE CELL: piecewise constant in the phase.  One warp takes 32 symbols at a time:
//   1. every lane has an error value for its symbol (zero to begin with);
//   2. all lanes run the 32 loop updates with those errors -- the reference's own float operations in the reference's
//      order, a few dependent instructions per symbol -- and lane n keeps the phase the loop had BEFORE symbol n;
//   3. every lane evaluates its symbol with that phase (sin/cos, rotation, double-precision cell index, table
//      look-up or atan2: the expensive part, now 32 wide);
//   4. if every lane got the error it already had, phases and errors satisfy the recurrence for all 32 symbols, and
//      the recurrence has exactly one solution: the sequential one.  Otherwise back to 2 with the new errors.
// Lane k is certainly right after k + 1 rounds, so the procedure ends; in practice the first guess (no error: the
// phase extrapolated by the loop frequency) already lands nearly every symbol in its final cell and two rounds do.
// Header and pilot symbols (atan2 of a continuous quantity) go through the same rounds; they settle as soon as the
// float error values stop changing.
//
// The result is what the sequential loop computes with the SAME per-symbol float operations.  Those are CUDA's
// sinf / cosf / atan2f instead of glibc's, so output symbols and loop state agree with the reference to float
// rounding, not bit for bit (tests/test_gpu_pll.py states the bar); every other operation is the reference's (no fused
// multiply-add, double-precision cell index with x86 truncation).
#include "pl_sync.cuh"

#include <climits>

namespace s2 {
namespace {

__device__ __forceinline__ float2 cmul_rn(float2 a, float2 b) {   // complex_t::operator*
    return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fadd_rn(__fmul_rn(a.y, b.x), __fmul_rn(a.x, b.y)));
}
__device__ __forceinline__ float2 conj2(float2 a) { return make_float2(a.x, -a.y); }

// int(double) as the x86 host does it (cvttsd2si): out-of-range and NaN give INT_MIN
__device__ __forceinline__ int host_trunc(double v) {
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
    return __double2int_rz(v);
}
// s / 1.5 in double, correctly rounded, without the division sequence: with y = RN(1 / 1.5) and q0 = RN(s y), the
// residual r = s - 1.5 q0 is exact in one fused multiply-add and RN(q0 + r y) is the correctly rounded quotient
// (Markstein's theorem; s is a float widened to double, far from the overflow and underflow thresholds).
__device__ __forceinline__ double div_1p5(double s) {
    const double y = 0.66666666666666663;   // RN(2/3)
    const double q0 = __dmul_rn(s, y);
    const double r = __fma_rn(-1.5, q0, s);
    return __fma_rn(r, y, q0);
}
__device__ __forceinline__ int lut_index(float s) {   // constellation.cpp:295-310
    int x = host_trunc(__dadd_rn(__dmul_rn(div_1p5((double)s), 256.0), 128.0));
    return min(max(x, 0), 255);
}

// PhaseControlLoop<float>::advance with the limits of S2PLLBlock::init (dvbs2_pll.cpp:11): frequency within
// +-0.01 pi, phase wrapped into [-pi, pi]
__device__ __forceinline__ void advance(float& phase, float& freq, float alpha, float beta, float error) {
    const float pi = 3.1415926535f;
    const float fmax = __fmul_rn(0.01f, pi), fmin = __fmul_rn(-0.01f, pi);
    freq = __fadd_rn(freq, __fmul_rn(beta, error));
    if (freq > fmax) freq = fmax;
    else if (freq < fmin) freq = fmin;
    phase = __fadd_rn(phase, __fadd_rn(freq, __fmul_rn(alpha, error)));
    const float delta = __fsub_rn(pi, -pi);
    while (phase > pi) phase = __fsub_rn(phase, delta);
    while (phase < -pi) phase = __fadd_rn(phase, delta);
}

// One symbol of S2PLLBlock::process (:39-79) given the loop phase before it: the error it feeds the loop, and in `o`
// what goes to out[i].
__device__ __forceinline__ float symbol_error(const PllArgs& a, int i, float2 x, float phase, float2& o) {
    float sn, cs;
    sincosf(-phase, &sn, &cs);
    const float2 t = cmul_rn(x, make_float2(cs, sn));
    if (i < kPlHeader) {
        const float2 ref = i < 26 ? a.tab->sof[i] : a.tab->pls[a.pls_code][i - 26];
        const float2 m = cmul_rn(t, conj2(ref));
        o = (i & 1) ? make_float2(-t.x, t.y) : make_float2(t.y, t.x);
        return atan2f(m.y, m.x);
    }
    const int k = i - kPlHeader;
    float2 d = t;
    switch (a.rn[k]) {   // S2Scrambling::descramble (s2_scrambling.cpp:37-58); the sequence restarts with every frame
    case 3: d = make_float2(-t.y, t.x); break;
    case 2: d = make_float2(-t.x, -t.y); break;
    case 1: d = make_float2(t.y, -t.x); break;
    default: break;
    }
    o = d;
    // The reference's pilot counter (:47-70) is a function of the position alone: 1439 symbols judged against the
    // constellation, then 35 against the diagonals (the 1440th data symbol and 34 more), period 1474.
    if (a.pilot_cnt && (k % 1474) >= 1439) {
        const float2 ideal = make_float2(d.x > 0.f ? 0.707f : -0.707f, d.y > 0.f ? 0.707f : -0.707f);
        const float2 m = cmul_rn(d, conj2(ideal));
        return __fdiv_rn(atan2f(m.y, m.x), 10.0f);
    }
    if (a.perr_lut) return __ldg(&a.perr_lut[lut_index(t.x) * 256 + lut_index(t.y)]);
    // 32APSK: demod_soft_calc per symbol (constellation.cpp:209-231,258-260)
    const float re = __fmul_rn(__fmul_rn(t.x, a.amp), a.prescale), im = __fmul_rn(__fmul_rn(t.y, a.amp), a.prescale);
    float best = 3.402823466e+38f, cr = 0.f, ci = 0.f;
    for (int s = 0; s < a.states; ++s) {
        const float dr = __fsub_rn(re, a.pts[2 * s]), di = __fsub_rn(im, a.pts[2 * s + 1]);
        const float dist = __fsqrt_rn(__fadd_rn(__fmul_rn(dr, dr), __fmul_rn(di, di)));
        if (dist < best) {
            best = dist;
            cr = a.pts[2 * s];
            ci = a.pts[2 * s + 1];
        }
    }
    const float2 m = cmul_rn(make_float2(re, im), make_float2(cr, -ci));
    return atan2f(m.y, m.x);
}

// The 32 loop updates of a block with the errors every lane holds (all lanes compute the same values).  Fast form:
// frequency clamp and phase wrap are left out and checked afterwards -- without them an update is three independent
// additions -- and the block is redone with the full PhaseControlLoop::advance if any of the 32 states left the range
// (a wrap comes once per 2 pi / freq symbols, the clamp only while the loop is pulling in).
struct ScanOut { float ph_in, ph_end, fr_end, es_end; };
template <bool FULL>
__device__ __forceinline__ ScanOut scan_block(float phase, float freq, float errsum, float alpha, float beta, float e_cur, int nv, int lane) {
    const unsigned full = 0xFFFFFFFFu;
    const float pi = 3.1415926535f;
    const float fmax = __fmul_rn(0.01f, pi), fmin = __fmul_rn(-0.01f, pi);
    ScanOut r;
    r.ph_in = phase;
    float ph = phase, fr = freq, es = errsum;
    const float be = __fmul_rn(beta, e_cur), ae = __fmul_rn(alpha, e_cur);   // this lane's increments
    if (FULL) {
        float flo = fr, fhi = fr, plo = ph, phi = ph;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const float ek = __shfl_sync(full, e_cur, k), bk = __shfl_sync(full, be, k), ak = __shfl_sync(full, ae, k);
            if (lane == k) r.ph_in = ph;
            es = __fadd_rn(es, ek);
            fr = __fadd_rn(fr, bk);
            ph = __fadd_rn(ph, __fadd_rn(fr, ak));
            flo = fminf(flo, fr); fhi = fmaxf(fhi, fr);
            plo = fminf(plo, ph); phi = fmaxf(phi, ph);
        }
        // NaN-safe: a comparison with NaN is false -> slow path
        if (flo >= fmin && fhi <= fmax && plo >= -pi && phi <= pi) {
            r.ph_end = ph; r.fr_end = fr; r.es_end = es;
            return r;
        }
        ph = phase; fr = freq; es = errsum;
    }
    for (int k = 0; k < nv; ++k) {
        const float ek = __shfl_sync(full, e_cur, k);
        if (lane == k) r.ph_in = ph;
        es = __fadd_rn(es, ek);
        advance(ph, fr, alpha, beta, ek);
    }
    r.ph_end = ph; r.fr_end = fr; r.es_end = es;
    return r;
}

__device__ __forceinline__ void pll_stream(const PllArgs& a) {
    const int lane = threadIdx.x;
    const unsigned full = 0xFFFFFFFFu;
    float phase = a.st->phase, freq = a.st->freq;
    const float alpha = a.st->alpha, beta = a.st->beta;
    float err_avg = a.st->error;
    unsigned rounds = 0;
    for (int f = 0; f < a.nframes; ++f) {
        const float2* in = a.frames + (size_t)f * a.rfs;
        float2* out = a.out + (size_t)f * a.rfs;
        float errsum = 0.f;
        float2 xn = lane < a.total ? in[lane] : make_float2(0.f, 0.f);
        for (int base = 0; base < a.total; base += 32) {
            const int nv = min(32, a.total - base), i = base + lane;
            const bool valid = lane < nv;
            const float2 x = xn;
            if (i + 32 < a.total) xn = in[i + 32];     // next block's symbols arrive while this one settles
            float e_cur = 0.f;
            float2 o = make_float2(0.f, 0.f);
            ScanOut sc;
            for (int round = 0;; ++round) {
                sc = (nv == 32) ? scan_block<true>(phase, freq, errsum, alpha, beta, e_cur, nv, lane)
                                : scan_block<false>(phase, freq, errsum, alpha, beta, e_cur, nv, lane);
                const float e_new = valid ? symbol_error(a, i, x, sc.ph_in, o) : 0.f;
                const bool same = __all_sync(full, __float_as_uint(e_new) == __float_as_uint(e_cur));
                ++rounds;
                if (round > 0 && same) break;
                e_cur = e_new;
            }
            phase = sc.ph_end; freq = sc.fr_end; errsum = sc.es_end;
            if (valid) out[i] = o;
        }
        err_avg = __fdiv_rn(errsum, a.divisor);
        if (lane == 0 && a.state_out) {
            a.state_out[3 * f] = phase;
            a.state_out[3 * f + 1] = freq;
            a.state_out[3 * f + 2] = err_avg;
        }
    }
    if (lane == 0) {
        a.st->phase = phase;
        a.st->freq = freq;
        a.st->error = err_avg;
        a.st->rounds = rounds;
    }
}

// ---- the same idea a CTA wide (second generation) -------------------------------------------------------------------------
// 256 symbols per block, a thread each.  The two halves of a round cost very different things: the loop updates are a chain
// of two dependent additions per symbol that ONE thread walks at a handful of cycles per symbol; the evaluation (sin/cos,
// rotation, double-precision cell index, look-up or atan2) is several hundred cycles of latency however many symbols are
// evaluated at once.  So the block is made as wide as the evaluation can be (eight warps), the walk keeps what it computed
// (phase, frequency and error sum BEFORE every symbol, in shared memory), and a round only redoes what can have changed:
// the walk resumes at the first symbol whose error changed, and only symbols from there on are evaluated again.  The fixed
// point is the same as the warp-wide kernel's and the sequential walk's -- per-symbol operations and their order are
// untouched -- and is reached in fewer, cheaper rounds per symbol.
constexpr int kWide = 256;
struct WideShared {
    float e[kWide];                                       // the error every symbol currently feeds the loop
    float ph[kWide + 1], fr[kWide + 1], es[kWide + 1];    // loop phase / frequency / error sum before symbol k; [nv] = after the block
    int jmin[2];                                          // first symbol whose error changed in this round
    int slow;                                             // the walk left the range where clamp and wrap are no-ops: use advance()
};

// one thread: loop updates for symbols [j, nv) from the state kept at j; returns false if a fast walk left the range
__device__ __forceinline__ bool wide_walk(WideShared& S, int j, int nv, float alpha, float beta, bool slow) {
    const float pi = 3.1415926535f;
    const float fmax = __fmul_rn(0.01f, pi), fmin = __fmul_rn(-0.01f, pi);
    float ph = S.ph[j], fr = S.fr[j], es = S.es[j];
    if (slow) {
        for (int k = j; k < nv; ++k) {
            const float ek = S.e[k];
            es = __fadd_rn(es, ek);
            advance(ph, fr, alpha, beta, ek);
            S.ph[k + 1] = ph; S.fr[k + 1] = fr; S.es[k + 1] = es;
        }
        return true;
    }
    float flo = fr, fhi = fr, plo = ph, phi = ph;
#pragma unroll 8
    for (int k = j; k < nv; ++k) {
        const float ek = S.e[k];
        es = __fadd_rn(es, ek);
        fr = __fadd_rn(fr, __fmul_rn(beta, ek));
        ph = __fadd_rn(ph, __fadd_rn(fr, __fmul_rn(alpha, ek)));
        flo = fminf(flo, fr); fhi = fmaxf(fhi, fr);
        plo = fminf(plo, ph); phi = fmaxf(phi, ph);
        S.ph[k + 1] = ph; S.fr[k + 1] = fr; S.es[k + 1] = es;
    }
    return flo >= fmin && fhi <= fmax && plo >= -pi && phi <= pi;      // (NaN: false)
}

__device__ __forceinline__ void pll_stream_wide(const PllArgs& a) {
    __shared__ WideShared S;
    const int tid = threadIdx.x;
    float phase = a.st->phase, freq = a.st->freq;
    const float alpha = a.st->alpha, beta = a.st->beta;
    float err_avg = a.st->error;
    unsigned rounds = 0;
    for (int f = 0; f < a.nframes; ++f) {
        const float2* in = a.frames + (size_t)f * a.rfs;
        float2* out = a.out + (size_t)f * a.rfs;
        float errsum = 0.f;
        for (int base = 0; base < a.total; base += kWide) {
            const int nv = min(kWide, a.total - base), i = base + tid;
            const bool valid = tid < nv;
            const float2 x = valid ? in[i] : make_float2(0.f, 0.f);
            float e_cur = 0.f;
            float2 o = make_float2(0.f, 0.f);
            S.e[tid] = 0.f;
            if (tid == 0) { S.ph[0] = phase; S.fr[0] = freq; S.es[0] = errsum; S.slow = 0; }
            int j = 0;
            bool slow = false;
            for (int round = 0;; ++round) {
                __syncthreads();      // errors of the previous round (or the zeros) are in place
                if (tid == 0) {
                    S.jmin[round & 1] = INT_MAX;
                    if (!wide_walk(S, j, nv, alpha, beta, slow)) S.slow = 1;
                }
                __syncthreads();
                if (!slow && S.slow) {      // once per block at most: start over with the full advance()
                    slow = true;
                    j = 0;
                    continue;
                }
                if (valid && tid >= j) {
                    const float e_new = symbol_error(a, i, x, S.ph[tid], o);
                    if (__float_as_uint(e_new) != __float_as_uint(e_cur)) {
                        e_cur = e_new;
                        S.e[tid] = e_new;
                        atomicMin(&S.jmin[round & 1], tid);
                    }
                }
                ++rounds;
                __syncthreads();
                j = S.jmin[round & 1];
                if (j == INT_MAX) break;      // every symbol got the error it already had: the sequential solution
            }
            phase = S.ph[nv]; freq = S.fr[nv]; errsum = S.es[nv];
            if (valid) out[i] = o;
            __syncthreads();      // S is reused by the next block
        }
        err_avg = __fdiv_rn(errsum, a.divisor);
        if (tid == 0 && a.state_out) {
            a.state_out[3 * f] = phase;
            a.state_out[3 * f + 1] = freq;
            a.state_out[3 * f + 2] = err_avg;
        }
    }
    if (tid == 0) {
        a.st->phase = phase;
        a.st->freq = freq;
        a.st->error = err_avg;
        a.st->rounds = rounds;
    }
}

// The same loop walked one symbol at a time by one thread, as the reference walks it: the yardstick for the speculative
// kernel (tests: both produce the same bits) and for its timing.
__global__ void __launch_bounds__(32) pll_sequential_kernel(const __grid_constant__ PllArgs a) {
    if (threadIdx.x != 0) return;
    float phase = a.st->phase, freq = a.st->freq, err_avg = a.st->error;
    const float alpha = a.st->alpha, beta = a.st->beta;
    for (int f = 0; f < a.nframes; ++f) {
        const float2* in = a.frames + (size_t)f * a.rfs;
        float2* out = a.out + (size_t)f * a.rfs;
        float errsum = 0.f;
        for (int i = 0; i < a.total; ++i) {
            float2 o;
            const float e = symbol_error(a, i, in[i], phase, o);
            out[i] = o;
            errsum = __fadd_rn(errsum, e);
            advance(phase, freq, alpha, beta, e);
        }
        err_avg = __fdiv_rn(errsum, a.divisor);
        if (a.state_out) {
            a.state_out[3 * f] = phase;
            a.state_out[3 * f + 1] = freq;
            a.state_out[3 * f + 2] = err_avg;
        }
    }
    a.st->phase = phase;
    a.st->freq = freq;
    a.st->error = err_avg;
    a.st->rounds = 0;
}

__global__ void __launch_bounds__(32) pll_kernel(const __grid_constant__ PllArgs a) { pll_stream(a); }
__global__ void __launch_bounds__(kWide) pll_wide_kernel(const __grid_constant__ PllArgs a) { pll_stream_wide(a); }
// several independent streams (transponders) at once: a CTA each
__global__ void __launch_bounds__(kWide) pll_multi_kernel(const PllArgs* jobs) { pll_stream_wide(jobs[blockIdx.x]); }

}  // namespace

int pll_launch(const PllArgs& a, int mode, cudaStream_t stream) {
    if (a.nframes > 0) {
        if (mode == 1) pll_sequential_kernel<<<1, 32, 0, stream>>>(a);
        else if (mode == 2) pll_kernel<<<1, 32, 0, stream>>>(a);
        else pll_wide_kernel<<<1, kWide, 0, stream>>>(a);
    }
    return (int)cudaGetLastError();
}
int pll_launch_multi(const PllArgs* d_jobs, int njobs, cudaStream_t stream) {
    if (njobs > 0) pll_multi_kernel<<<njobs, kWide, 0, stream>>>(d_jobs);
    return (int)cudaGetLastError();
}

}  // namespace s2
