// Which issue pipe do the candidate instructions use on sm_100a?  Each kernel runs long unrolled chains of
// independent ops; mixes that overlap on different pipes finish in max() of the parts, same-pipe mixes in the sum.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define ITER 4096
template <int MODE>
__global__ void k(uint32_t* out, uint32_t seed) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 8 + i;
    uint32_t c = seed | 0x00010001u;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0 && (i & 1) == 0) { uint32_t lo = __vmins2(a[i], a[i + 1]) + 1u, hi = __vmaxs2(a[i], a[i + 1]) ^ c; a[i] = hi; a[i + 1] = lo; }   // 2 VIMNMX (+2 cheap)
            if (MODE == 1 && (i & 1) == 0) { __half2 x = *(__half2*)&a[i], y = *(__half2*)&a[i + 1]; __half2 lo = __hmin2(x, y), hi = __hmax2(x, y); a[i] = (*(uint32_t*)&hi) ^ c; a[i + 1] = (*(uint32_t*)&lo) + 1u; }  // 2 HMNMX2 (+2 cheap)
            if (MODE == 2) { __half2 h = __hadd2(*(__half2*)&a[i], *(__half2*)&c); a[i] = *(uint32_t*)&h; }  // HADD2
            if (MODE == 3) a[i] = a[i] * 3u + c;                                       // IMAD
            if (MODE == 4 && (i & 1) == 0) { uint32_t lo = a[i] + 1u, hi = a[i + 1] ^ c; a[i] = hi; a[i + 1] = lo; }   // the 2 cheap ops alone (IADD + LOP3)
            if (MODE == 5 && (i & 3) == 0) { uint32_t lo = __vmins2(a[i], a[i + 1]) + 1u, hi = __vmaxs2(a[i], a[i + 1]) ^ c; a[i] = hi; a[i + 1] = lo;
                                             __half2 x = *(__half2*)&a[i + 2], y = *(__half2*)&a[i + 3]; __half2 l2 = __hmin2(x, y), h2 = __hmax2(x, y); a[i + 2] = (*(uint32_t*)&h2) ^ c; a[i + 3] = (*(uint32_t*)&l2) + 1u; }
            if (MODE == 6) { a[i] = __vmins2(a[i], c); i++; __half2 h = __hadd2(*(__half2*)&a[i], *(__half2*)&c); a[i] = *(uint32_t*)&h; }  // VIMNMX + HADD2
            if (MODE == 7) { a[i] = __vmins2(a[i], c); i++; a[i] = a[i] * 3u + c; }   // VIMNMX + IMAD
            if (MODE == 8) { __half2 h = __hmin2(*(__half2*)&a[i], *(__half2*)&c); a[i] = *(uint32_t*)&h; i++; __half2 g = __hadd2(*(__half2*)&a[i], *(__half2*)&c); a[i] = *(uint32_t*)&g; }  // HMNMX2 + HADD2
            if (MODE == 9) { __half2 h = __hfma2_relu(*(__half2*)&a[i], *(__half2*)&c, *(__half2*)&c); a[i] = *(uint32_t*)&h; }  // HFMA2.RELU
            if (MODE == 10) a[i] = __viaddmin_s16x2_relu(a[i], c, 0x00ff00ffu);        // VIADDMNMX.RELU
            if (MODE == 11) { a[i] = __viaddmin_s16x2_relu(a[i], c, 0x00ff00ffu); i++; __half2 h = __hfma2_relu(*(__half2*)&a[i], *(__half2*)&c, *(__half2*)&c); a[i] = *(uint32_t*)&h; }
            if (MODE == 12) a[i] = __vabsdiffu4(a[i], c);                              // VABSDIFF4
            if (MODE == 13) { uint32_t d; asm("prmt.b32 %0,%1,%2,0x9180;" : "=r"(d) : "r"(a[i]), "r"(c)); a[i] = d; }  // PRMT
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
float run(uint32_t* d) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 4, 256>>>(d, 12345u);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 256>>>(d, 12345u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}
int main() {
    uint32_t* d; cudaMalloc(&d, 148 * 4 * 256 * 4);
    const char* names[] = {"VIMNMX.S16x2", "HMNMX2", "HADD2", "IMAD", "LOP3", "VIMNMX+HMNMX2", "VIMNMX+HADD2", "VIMNMX+IMAD",
                           "HMNMX2+HADD2", "HFMA2.RELU", "VIADDMNMX.RELU", "VIADDMNMX.RELU+HFMA2.RELU", "VABSDIFF4", "PRMT"};
    float t[14];
    t[0] = run<0>(d); t[1] = run<1>(d); t[2] = run<2>(d); t[3] = run<3>(d); t[4] = run<4>(d); t[5] = run<5>(d); t[6] = run<6>(d);
    t[7] = run<7>(d); t[8] = run<8>(d); t[9] = run<9>(d); t[10] = run<10>(d); t[11] = run<11>(d); t[12] = run<12>(d); t[13] = run<13>(d);
    // ops per kernel: 148*4 blocks * 8 warps * ITER * 8 warp-instr
    double winst = 148.0 * 4 * 8 * ITER * 8;
    for (int i = 0; i < 14; ++i)
        printf("%-28s %.3f ms  -> %.2f warp-instr/clk/SM at 1.965 GHz\n", names[i], t[i], winst / 148 / (t[i] * 1e-3 * 1.965e9));
    return 0;
}
