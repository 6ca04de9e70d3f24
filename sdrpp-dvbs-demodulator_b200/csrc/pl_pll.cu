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

// ---- the blocks of a stream in a pipeline (second generation) --------------------------------------------------------------
// A block settles in about four rounds, but only the first of them moves its end state by much: the later ones correct a
// table cell here and there and confirm.  So the next block need not wait: kPipe warps take the blocks of a stream in turn
// and work in ticks.  In a tick every warp reads what its predecessor published in the tick before (the loop state at the
// end of the predecessor's block, and whether that is final), runs one round of its own block from it -- unless neither
// the start state nor its errors have changed -- and publishes where its block ends.  A block is done when its errors
// reproduce themselves from a start state declared final; what comes out of it is then computed from the final start state
// and confirmed by a round of its own: by induction from the first block, the sequential result, as before.  Reads and
// writes of the records are in different halves of a tick, with a CTA barrier between the halves: no warp ever waits for
// another one except at those barriers.
constexpr int kPipe = 6, kLocal = 1;
struct Rec { float ph, fr, es; int blk, fin; };      // loop state after block `blk`
struct PipeShared {
    Rec mail[kPipe];      // what a warp's current block looked like after its last round
    Rec fmail[kPipe];     // the final state of the last block a warp finished (its successor may still need it when the warp has moved on)
    unsigned rounds;
};

__device__ __forceinline__ void pll_stream_pipe(const PllArgs& a) {
    __shared__ PipeShared S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pw = (warp + kPipe - 1) % kPipe;
    const unsigned full = 0xFFFFFFFFu;
    const float alpha = a.st->alpha, beta = a.st->beta;
    const float ph0 = a.st->phase, fr0 = a.st->freq;
    if (threadIdx.x < kPipe) { S.mail[threadIdx.x].blk = -1; S.fmail[threadIdx.x].blk = -1; }
    if (threadIdx.x == 0) S.rounds = 0;
    const int bpf = (a.total + 31) / 32, nblocks = a.nframes * bpf;
    unsigned rounds = 0;
    int g = warp;      // my block: global index over the frames of the call
    int f = 0, base = 0, nv = 0, i = 0;
    bool valid = false, have = false;
    float2 x = make_float2(0.f, 0.f), o = make_float2(0.f, 0.f);
    float e_cur = 0.f, uph = 0.f, ufr = 0.f, ues = 0.f;
    ScanOut sc = {0.f, 0.f, 0.f, 0.f}, pub = {0.f, 0.f, 0.f, 0.f};      // the last walk; what the record says about the block's end
    auto open_block = [&]() {
        if (g >= nblocks) return;
        f = g / bpf; base = (g - f * bpf) * 32;
        nv = min(32, a.total - base); i = base + lane;
        valid = lane < nv;
        x = valid ? a.frames[(size_t)f * a.rfs + i] : make_float2(0.f, 0.f);
        e_cur = 0.f;
        have = false;
    };
    open_block();
    __syncthreads();
    for (;;) {
        // first half of the tick: read
        const bool active = g < nblocks;
        bool got = false;
        float sph = ph0, sfr = fr0, ses = 0.f;
        int pfin = 1;
        if (active) {
            if (g == 0) got = true;
            else {
                Rec r = S.mail[pw];
                if (r.blk != g - 1) { r = S.fmail[pw]; r.fin = 1; }
                if (r.blk == g - 1) { sph = r.ph; sfr = r.fr; ses = r.es; pfin = r.fin; got = true; }
            }
            if (base == 0) ses = 0.f;      // a frame starts: S2PLLBlock::process sums the error per frame
        }
        __syncthreads();
        // second half: one round, publish
        bool done = false;
        if (active && got) {
            const bool same_start = __float_as_uint(sph) == __float_as_uint(uph) && __float_as_uint(sfr) == __float_as_uint(ufr) &&
                                    __float_as_uint(ses) == __float_as_uint(ues);
            bool same = have && same_start;      // same start, and the last round found the errors it was given: nothing to redo
            // (kLocal > 1: several rounds on one start state before publishing -- measured, no gain)
            for (int r = 0; r < kLocal && !same; ++r) {
                sc = (nv == 32) ? scan_block<true>(sph, sfr, ses, alpha, beta, e_cur, nv, lane)
                                : scan_block<false>(sph, sfr, ses, alpha, beta, e_cur, nv, lane);
                const float e_new = valid ? symbol_error(a, i, x, sc.ph_in, o) : 0.f;
                same = __all_sync(full, __float_as_uint(e_new) == __float_as_uint(e_cur));
                ++rounds;
                uph = sph; ufr = sfr; ues = ses;
                have = same;      // the errors in e_cur reproduce themselves from (uph, ufr, ues)
                if (!same) {
                    e_cur = e_new;
                    // Walk once more with the new errors, for the record only: where the block ends with them is, more
                    // often than not, where it will end for good, and the block behind gets to see that a tick earlier.
                    const ScanOut s2 = (nv == 32) ? scan_block<true>(sph, sfr, ses, alpha, beta, e_cur, nv, lane)
                                                  : scan_block<false>(sph, sfr, ses, alpha, beta, e_cur, nv, lane);
                    pub = s2;
                } else
                    pub = sc;
            }
            done = same && pfin;
            if (lane == 0) {
                const Rec r = {pub.ph_end, pub.fr_end, pub.es_end, g, done ? 1 : 0};
                S.mail[warp] = r;
                if (done) S.fmail[warp] = r;
            }
            if (done) {
                if (valid) a.out[(size_t)f * a.rfs + i] = o;
                if (lane == 0 && base + 32 >= a.total) {      // the frame's last block
                    const float err_avg = __fdiv_rn(sc.es_end, a.divisor);
                    if (a.state_out) {
                        a.state_out[3 * f] = sc.ph_end;
                        a.state_out[3 * f + 1] = sc.fr_end;
                        a.state_out[3 * f + 2] = err_avg;
                    }
                    if (g == nblocks - 1) {
                        a.st->phase = sc.ph_end;
                        a.st->freq = sc.fr_end;
                        a.st->error = err_avg;
                    }
                }
                g += kPipe;
                open_block();
            }
        }
        if (__syncthreads_and(g >= nblocks)) break;
    }
    if (lane == 0) atomicAdd(&S.rounds, rounds);
    __syncthreads();
    if (threadIdx.x == 0) a.st->rounds = S.rounds;
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
__global__ void __launch_bounds__(32 * kPipe) pll_pipe_kernel(const __grid_constant__ PllArgs a) { pll_stream_pipe(a); }
// several independent streams (transponders) at once: a CTA of kPipe warps each
__global__ void __launch_bounds__(32 * kPipe) pll_multi_kernel(const PllArgs* jobs) { pll_stream_pipe(jobs[blockIdx.x]); }

}  // namespace

int pll_launch(const PllArgs& a, int mode, cudaStream_t stream) {
    if (a.nframes > 0) {
        if (mode == 1) pll_sequential_kernel<<<1, 32, 0, stream>>>(a);
        else if (mode == 2) pll_kernel<<<1, 32, 0, stream>>>(a);
        else pll_pipe_kernel<<<1, 32 * kPipe, 0, stream>>>(a);
    }
    return (int)cudaGetLastError();
}
int pll_launch_multi(const PllArgs* d_jobs, int njobs, cudaStream_t stream) {
    if (njobs > 0) pll_multi_kernel<<<njobs, 32 * kPipe, 0, stream>>>(d_jobs);
    return (int)cudaGetLastError();
}

}  // namespace s2
