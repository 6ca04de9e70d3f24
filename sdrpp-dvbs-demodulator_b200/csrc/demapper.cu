// K1 -- see demapper.cuh.
#include "demapper.cuh"

#include <climits>

namespace s2 {
namespace {

// int(double) as the x86 host does it (cvttsd2si): out-of-range and NaN give INT_MIN
__device__ __forceinline__ int host_trunc(double v) {
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
    return __double2int_rz(v);
}
__device__ __forceinline__ int lut_index(float s) {
    int x = host_trunc(((double)s / 1.5) * 256 + 128);
    return min(max(x, 0), 255);
}
__device__ __forceinline__ int halving_clamp(float x) {
    while (x < -127 || x > 127) {
        x *= 0.5f;
        if (!isfinite(x)) return 0;   // (int8_t)(int)inf/nan on the host: low byte of INT_MIN
    }
    return (int)x;
}

// LLR bytes of symbol s to their deinterleaved places (S2Deinterleaver::deinterleave, s2_deinterleaver.cpp:72-136)
__device__ __forceinline__ void store_llrs(const DemapDev& d, int8_t* o, int s, const int8_t (&b)[5]) {
    if (d.constellation == 0) {           // QPSK: the pair is swapped, no column interleave
        *reinterpret_cast<uint16_t*>(o + 2 * s) = (uint16_t)((uint8_t)b[1] | ((uint16_t)(uint8_t)b[0] << 8));
        return;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k)
        if (k < d.bits) {
            int col = (d.reversed_cols) ? (2 - k) : k;
            o[col * d.nsym + s] = b[k];
        }
}

__global__ void __launch_bounds__(256) demap_kernel(const __grid_constant__ DemapDev d, const float2* __restrict__ in,
                                                    int nframes, int8_t* __restrict__ out) {
    const int frame = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= d.nsym) return;
    const int raw = 90 + s + (d.pilots ? 36 * (s / 1440) : 0);
    float2 v = __ldg(&in[(size_t)frame * d.plframe_syms + raw]);
    if (d.rn) {   // exact: quarter turns only swap and negate
        const int r = __ldg(&d.rn[raw - 90]);
        const float2 u = v;
        if (r == 1) v = make_float2(u.y, -u.x);
        else if (r == 2) v = make_float2(-u.x, -u.y);
        else if (r == 3) v = make_float2(-u.y, u.x);
    }
    int8_t* o = out + (size_t)frame * d.N;
    int8_t b[5];
    if (d.constellation != 3) {
        uint32_t w = __ldg(&d.lut[lut_index(v.x) * 256 + lut_index(v.y)]);
        b[0] = (int8_t)(w & 0xFF);
        b[1] = (int8_t)((w >> 8) & 0xFF);
        b[2] = (int8_t)((w >> 16) & 0xFF);
        b[3] = (int8_t)(w >> 24);
    } else {
        float re = (v.x * d.amp) * d.prescale, im = (v.y * d.amp) * d.prescale;
        float acc[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) acc[k] = 0.f;
        for (int i = 0; i < 32; ++i) {
            float dr = re - d.pts[2 * i], di = im - d.pts[2 * i + 1];
            float e = expf(-sqrtf(__fadd_rn(__fmul_rn(dr, dr), __fmul_rn(di, di))));
#pragma unroll
            for (int jb = 0; jb < 5; ++jb) {
                if ((i >> jb) & 1) acc[2 * jb + 1] += e;
                else acc[2 * jb] += e;
            }
        }
#pragma unroll
        for (int jb = 0; jb < 5; ++jb)
            b[4 - jb] = (int8_t)halving_clamp((logf(acc[2 * jb + 1]) - logf(acc[2 * jb])) * d.sca);
    }
    store_llrs(d, o, s, b);
}

// Same gather from the two 8-bit LUT coordinates per symbol that the host computed (dvbs2fec_quantize_plframes:
// the reference's own index arithmetic, constellation.cpp:295-310, after pilot removal and PL descrambling):
// 2 bytes per symbol over PCIe instead of 8, identical LLRs.
__global__ void __launch_bounds__(256) demap_idx_kernel(const __grid_constant__ DemapDev d, const uchar2* __restrict__ idx,
                                                        int nframes, int8_t* __restrict__ out) {
    const int frame = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= d.nsym) return;
    const uchar2 xy = __ldg(&idx[(size_t)frame * d.nsym + s]);
    const uint32_t w = __ldg(&d.lut[(int)xy.x * 256 + (int)xy.y]);
    int8_t b[5] = {(int8_t)(w & 0xFF), (int8_t)((w >> 8) & 0xFF), (int8_t)((w >> 16) & 0xFF), (int8_t)(w >> 24), 0};
    store_llrs(d, out + (size_t)frame * d.N, s, b);
}

}  // namespace

int demap_launch(const DemapDev& d, const float* plframes, int nframes, int8_t* llr_out, cudaStream_t stream) {
    if (nframes <= 0) return 0;
    dim3 grid((d.nsym + 255) / 256, nframes);
    demap_kernel<<<grid, 256, 0, stream>>>(d, reinterpret_cast<const float2*>(plframes), nframes, llr_out);
    return (int)cudaGetLastError();
}

int demap_idx_launch(const DemapDev& d, const uint8_t* idx, int nframes, int8_t* llr_out, cudaStream_t stream) {
    if (nframes <= 0) return 0;
    if (!d.lut) return (int)cudaErrorInvalidValue;   // 32APSK has no LUT
    dim3 grid((d.nsym + 255) / 256, nframes);
    demap_idx_kernel<<<grid, 256, 0, stream>>>(d, reinterpret_cast<const uchar2*>(idx), nframes, llr_out);
    return (int)cudaGetLastError();
}

}  // namespace s2
