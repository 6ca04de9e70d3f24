// Issue-rate microbenchmark for the instructions the LDPC kernel is built from (sm_100a), plus shared-memory
// wavefront rates.  One CTA of 512 threads (4 warps per sub-partition) per SM; every thread runs 8 independent
// dependency chains of the instruction under test, so a pipe that accepts one warp instruction every r cycles per
// sub-partition shows 4/r warp-instructions per clock per SM.  Cycles come from clock64() inside the kernel
// (slowest CTA), so the figures do not depend on the SM clock.  Mixes of two instructions show whether they
// share a pipe (rates add up to the single-instruction rate) or issue side by side (up to 4 per clock per SM).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes [out.json]
//
// The SASS of every mode was checked with cuobjdump (the instruction named is the one in the loop).
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define ITER 2048
#define NCHAIN 8

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) {
    uint32_t d;
    asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
}
__device__ __forceinline__ uint32_t h2(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }

enum Mode {
    M_LOP3, M_PRMT, M_SHF, M_IADD3, M_VIADD16, M_VIMNMX_S, M_VIMNMX_U, M_VIMNMX3, M_VIADDMNMX, M_VIADDMNMX_RELU,
    M_VABSDIFF4, M_ISETP_SEL, M_HMNMX2, M_IMAD, M_IMAD_IADD, M_HADD2, M_HFMA2, M_HFMA2_RELU, M_FFMA, M_VIBMAX,
    MX_VIMNMX_IMAD, MX_VIMNMX_HFMA2, MX_LOP3_IMAD, MX_VIADDMNMX_HADD2, MX_HMNMX2_IMAD, MX_PRMT_IMAD, MX_VIMNMX_LOP3,
    M_COUNT
};
static const char* kNames[M_COUNT] = {
    "LOP3", "PRMT", "SHF", "IADD3", "VIADD.16x2", "VIMNMX.S16x2", "VIMNMX.U16x2", "VIMNMX3.S16x2", "VIADDMNMX.S16x2",
    "VIADDMNMX.S16x2.RELU", "VABSDIFF4", "ISETP+SEL", "HMNMX2", "IMAD", "mad.lo(x,1,c) pairs -> IADD3", "HADD2", "HFMA2", "HFMA2.RELU",
    "FFMA", "VIBMAX(pred: VIMNMX+VIADD+P2R/4)", "VIMNMX+IMAD", "VIMNMX+HFMA2", "LOP3+IMAD", "VIADDMNMX+HADD2", "HMNMX2+IMAD", "PRMT+IMAD",
    "VIMNMX+LOP3"};

#define A2(ptx) { uint32_t r; asm volatile(ptx : "=r"(r) : "r"(x), "r"(c)); return r; }
#define A3(ptx) { uint32_t r; asm volatile(ptx : "=r"(r) : "r"(x), "r"(c), "r"(d)); return r; }
// one instruction of the kind under test: x = f(x, c, d); `i` = chain index, `r` = repeat index (both compile-time)
template <int MODE>
__device__ __forceinline__ uint32_t op(uint32_t x, uint32_t c, uint32_t d, int i, int rr) {
    switch (MODE) {
        case M_LOP3: A3("lop3.b32 %0, %1, %2, %3, 0xE8;")
        case M_PRMT: A2("prmt.b32 %0, %1, %2, 0x9180;")
        case M_SHF: A2("shf.r.wrap.b32 %0, %1, %2, 7;")
        case M_IADD3: A3("{.reg .b32 t; add.u32 t, %1, %2; add.u32 %0, t, %3;}")
        case M_VIADD16: A2("add.s16x2 %0, %1, %2;")
        case M_VIMNMX_S: if (rr & 1) A2("min.s16x2 %0, %1, %2;") else A2("max.s16x2 %0, %2, %1;")
        case M_VIMNMX_U: if (rr & 1) A2("min.u16x2 %0, %1, %2;") else A2("max.u16x2 %0, %2, %1;")
        case M_VIMNMX3: A3("{.reg .b32 t; min.s16x2 t, %1, %2; min.s16x2 %0, t, %3;}")
        case M_VIADDMNMX: A3("{.reg .b32 t; add.s16x2 t, %1, %2; min.s16x2 %0, t, %3;}")
        case M_VIADDMNMX_RELU: A3("{.reg .b32 t; add.s16x2 t, %1, %2; min.s16x2.relu %0, t, %3;}")
        case M_VABSDIFF4: return __vabsdiffu4(x, c);
        case M_ISETP_SEL: A3("{.reg .pred p; setp.lt.u32 p, %1, %2; selp.b32 %0, %3, %1, p;}")
        case M_HMNMX2: if (rr & 1) A2("min.f16x2 %0, %1, %2;") else A2("max.f16x2 %0, %2, %1;")
        case M_IMAD: A3("mad.lo.u32 %0, %1, %2, %3;")
        case M_IMAD_IADD: A2("mad.lo.u32 %0, %1, 1, %2;")
        case M_HADD2: A2("add.f16x2 %0, %1, %2;")
        case M_HFMA2: A3("fma.rn.f16x2 %0, %1, %2, %3;")
        case M_HFMA2_RELU: A3("fma.rn.relu.f16x2 %0, %1, %2, %3;")
        case M_FFMA: { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(__uint_as_float(x)), "f"(__uint_as_float(c)), "f"(__uint_as_float(d))); return __float_as_uint(r); }
        case M_VIBMAX: { bool ph, pl; uint32_t r = __vibmax_s16x2(x, c, &ph, &pl); return r + (ph ? 1u : 0u); }
        // mixes: even chains one instruction, odd chains the other
        case MX_VIMNMX_IMAD: if (i & 1) A3("mad.lo.u32 %0, %1, %2, %3;") else return op<M_VIMNMX_S>(x, c, d, i, rr);
        case MX_VIMNMX_HFMA2: if (i & 1) A3("fma.rn.f16x2 %0, %1, %2, %3;") else return op<M_VIMNMX_S>(x, c, d, i, rr);
        case MX_LOP3_IMAD: if (i & 1) A3("mad.lo.u32 %0, %1, %2, %3;") else A3("lop3.b32 %0, %1, %2, %3, 0xE8;")
        case MX_VIADDMNMX_HADD2: if (i & 1) A2("add.f16x2 %0, %1, %2;") else A3("{.reg .b32 t; add.s16x2 t, %1, %2; min.s16x2.relu %0, t, %3;}")
        case MX_HMNMX2_IMAD: if (i & 1) A3("mad.lo.u32 %0, %1, %2, %3;") else return op<M_HMNMX2>(x, c, d, i, rr);
        case MX_PRMT_IMAD: if (i & 1) A3("mad.lo.u32 %0, %1, %2, %3;") else A2("prmt.b32 %0, %1, %2, 0x9180;")
        case MX_VIMNMX_LOP3: if (i & 1) A3("lop3.b32 %0, %1, %2, %3, 0xE8;") else return op<M_VIMNMX_S>(x, c, d, i, rr);
    }
    return x;
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) pipe_kernel(uint32_t* out, long long* cycles, uint32_t seed) {
    uint32_t a[NCHAIN];
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) a[i] = seed * (threadIdx.x + 1) + i * 0x01010101u;
    const uint32_t c = seed | 0x00010001u, d = seed * 7u + 0x00030003u;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < NCHAIN; ++i) a[i] = op<MODE>(a[i], c, d, i, r);
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// ---- shared memory: wavefronts per clock per SM ------------------------------------------------------------------
// W = access width in bytes (2, 4, 8, 16); ST = stores instead of loads.  Consecutive lanes touch consecutive
// W-byte items (no bank conflicts), 8 independent accesses in flight per thread.
template <int W, bool ST>
__global__ void __launch_bounds__(512, 1) smem_kernel(uint32_t* out, long long* cycles, uint32_t seed) {
    extern __shared__ __align__(16) uint8_t buf[];
    constexpr uint32_t kBuf = 512 * 16 * 8;
    for (int x = threadIdx.x; x < (int)kBuf / 4; x += 512) reinterpret_cast<uint32_t*>(buf)[x] = x * seed;
    __syncthreads();
    uint32_t acc = 0;
    uint32_t off = threadIdx.x * W;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t o = (off + i * 512 * W) & (kBuf - 1);
            if (ST) {
                if (W == 2) *reinterpret_cast<volatile uint16_t*>(buf + o) = (uint16_t)(acc + i);
                if (W == 4) *reinterpret_cast<volatile uint32_t*>(buf + o) = acc + i;
                if (W == 8) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"((uint32_t)__cvta_generic_to_shared(buf + o)), "r"(acc), "r"(acc + i)); }
                if (W == 16) { asm volatile("st.shared.v4.u32 [%0], {%1, %2, %1, %2};" ::"r"((uint32_t)__cvta_generic_to_shared(buf + o)), "r"(acc), "r"(acc + i)); }
            } else {
                if (W == 2) acc += *reinterpret_cast<volatile uint16_t*>(buf + o);
                if (W == 4) acc += *reinterpret_cast<volatile uint32_t*>(buf + o);
                if (W == 8) { uint32_t p, q; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(p), "=r"(q) : "r"((uint32_t)__cvta_generic_to_shared(buf + o))); acc += p ^ q; }
                if (W == 16) { uint32_t p, q, r, s; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(p), "=r"(q), "=r"(r), "=r"(s) : "r"((uint32_t)__cvta_generic_to_shared(buf + o))); acc += p ^ q ^ r ^ s; }
            }
        }
        off = (off + 16 * 32) & (kBuf - 1);
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static int g_sms = 148;
static double max_cycles(const long long* h, int n) {
    long long m = 0;
    for (int i = 0; i < n; ++i) m = h[i] > m ? h[i] : m;
    return (double)m;
}

template <int MODE>
double run_pipe(uint32_t* d, long long* dc, long long* hc) {
    pipe_kernel<MODE><<<g_sms, 512>>>(d, dc, 12345u);
    pipe_kernel<MODE><<<g_sms, 512>>>(d, dc, 12345u);
    cudaMemcpy(hc, dc, g_sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double winst = 16.0 * ITER * 8 * NCHAIN;   // warp instructions per SM
    if (MODE == M_IMAD_IADD) winst *= 0.5;      // ptxas merges two of these adds into one IADD3
    return winst / max_cycles(hc, g_sms);
}
template <int W, bool ST>
double run_smem(uint32_t* d, long long* dc, long long* hc) {
    cudaFuncSetAttribute(smem_kernel<W, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 512 * 16 * 8);
    smem_kernel<W, ST><<<g_sms, 512, 512 * 16 * 8>>>(d, dc, 12345u);
    smem_kernel<W, ST><<<g_sms, 512, 512 * 16 * 8>>>(d, dc, 12345u);
    cudaMemcpy(hc, dc, g_sms * sizeof(long long), cudaMemcpyDeviceToHost);
    const double winst = 16.0 * ITER * 8;
    return winst / max_cycles(hc, g_sms);
}

template <int MODE>
void all_pipes(uint32_t* d, long long* dc, long long* hc, double* r) {
    r[MODE] = run_pipe<MODE>(d, dc, hc);
    if constexpr (MODE + 1 < M_COUNT) all_pipes<MODE + 1>(d, dc, hc, r);
}

int main(int argc, char** argv) {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    g_sms = prop.multiProcessorCount;
    uint32_t* d;
    long long *dc, *hc = new long long[g_sms];
    cudaMalloc(&d, (size_t)g_sms * 512 * 4);
    cudaMalloc(&dc, g_sms * sizeof(long long));
    double r[M_COUNT];
    all_pipes<0>(d, dc, hc, r);
    double s[8];
    s[0] = run_smem<2, false>(d, dc, hc); s[1] = run_smem<4, false>(d, dc, hc);
    s[2] = run_smem<8, false>(d, dc, hc); s[3] = run_smem<16, false>(d, dc, hc);
    s[4] = run_smem<2, true>(d, dc, hc);  s[5] = run_smem<4, true>(d, dc, hc);
    s[6] = run_smem<8, true>(d, dc, hc);  s[7] = run_smem<16, true>(d, dc, hc);
    const char* sn[8] = {"LDS.U16", "LDS.32", "LDS.64", "LDS.128", "STS.U16", "STS.32", "STS.64", "STS.128"};
    const int sw[8] = {2, 4, 8, 16, 2, 4, 8, 16};
    if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "CUDA error\n"); return 1; }
    FILE* f = argc > 1 ? fopen(argv[1], "w") : stdout;
    fprintf(f, "{\n \"gpu\": \"%s\", \"sms\": %d,\n \"unit\": \"warp instructions per clock per SM (4 sub-partitions)\",\n \"issue\": {\n", prop.name, g_sms);
    for (int i = 0; i < M_COUNT; ++i) fprintf(f, "  \"%s\": %.3f%s\n", kNames[i], r[i], i + 1 < M_COUNT ? "," : "");
    fprintf(f, " },\n \"shared_memory\": {\n");
    for (int i = 0; i < 8; ++i)
        fprintf(f, "  \"%s\": {\"warp_instr_per_clk_per_sm\": %.3f, \"bytes_per_clk_per_sm\": %.1f}%s\n", sn[i], s[i], s[i] * 32 * sw[i], i + 1 < 8 ? "," : "");
    fprintf(f, " }\n}\n");
    if (f != stdout) fclose(f);
    return 0;
}
