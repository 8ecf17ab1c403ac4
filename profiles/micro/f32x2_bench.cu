// Micro-benchmark: issue rate of packed fp32 (FFMA2 / FADD2 / FMUL2, sm_100a) against the scalar forms, per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_bench f32x2_bench.cu && ./f32x2_bench
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define CHAINS 8
template <int MODE>
__global__ void __launch_bounds__(1024, 1) kern(float* out, int iters, float seed) {
    float a[CHAINS], b[CHAINS];
    unsigned x[CHAINS], y[CHAINS];
    u64 p[CHAINS];
    const float m = seed * 0.999f, c = seed * 1e-3f;
    u64 m2, c2;
    {
        float2 t = make_float2(m, m); m2 = *reinterpret_cast<u64*>(&t);
        t = make_float2(c, c); c2 = *reinterpret_cast<u64*>(&t);
    }
    for (int k = 0; k < CHAINS; k++) {
        a[k] = threadIdx.x * 1e-3f + k; b[k] = a[k] + 0.5f; x[k] = threadIdx.x + k; y[k] = x[k] * 3u;
        float2 t = make_float2(a[k], b[k]); p[k] = *reinterpret_cast<u64*>(&t);
    }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < CHAINS; k++) {
            if (MODE == 0) { a[k] = fmaf(a[k], m, c); b[k] = fmaf(b[k], m, c); }                 // 2 FFMA
            if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[k]) : "l"(m2), "l"(c2));   // 1 FFMA2
            if (MODE == 2) { a[k] = a[k] + c; b[k] = b[k] + c; }                                 // 2 FADD
            if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[k]) : "l"(c2));          // 1 FADD2
            if (MODE == 4) { a[k] = fmaf(a[k], m, c); b[k] = fminf(b[k], a[k]); }                // FFMA + FMNMX (two pipes)
            if (MODE == 5) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[k]) : "l"(m2), "l"(c2)); b[k] = fminf(b[k], a[k] + i); }
            if (MODE == 6) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[k]) : "l"(m2));          // 1 FMUL2
            if (MODE == 7) {   // 1 FFMA2 + 2 ALU (LOP3): 3 issue slots if FFMA2 takes one, 4 if it takes two
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[k]) : "l"(m2), "l"(c2));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(y[k]), "r"(i));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[k]) : "r"(x[k]), "r"(i));
            }
            if (MODE == 8) {   // 2 FFMA + 2 ALU
                a[k] = fmaf(a[k], m, c); b[k] = fmaf(b[k], m, c);
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(y[k]), "r"(i));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[k]) : "r"(x[k]), "r"(i));
            }
            if (MODE == 9) {   // 2 ALU only
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(y[k]), "r"(i));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[k]) : "r"(x[k]), "r"(i));
            }
        }
    }
    float s = 0;
    for (int k = 0; k < CHAINS; k++) { float2 t = *reinterpret_cast<float2*>(&p[k]); s += a[k] + b[k] + t.x + t.y + (float)(x[k] ^ y[k]); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int ops_per_iter_per_chain) {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<MODE><<<148, 1024>>>(out, iters, 1.0f);
    cudaEventRecord(e0);
    kern<MODE><<<148, 1024>>>(out, iters, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double warp_inst = 32.0 * iters * CHAINS * ops_per_iter_per_chain;          // per SM (32 warps)
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-28s %.3f ms  %.2f warp-inst/clk/SM (nominal clock %d kHz)\n", name, ms, warp_inst / cycles, clk);
    cudaFree(out);
}
int main() {
    run<0>("2x FFMA", 2); run<1>("1x FFMA2", 1); run<2>("2x FADD", 2); run<3>("1x FADD2", 1);
    run<4>("FFMA + FMNMX", 2); run<5>("FFMA2 + FADD + FMNMX", 3); run<6>("1x FMUL2", 1);
    run<7>("FFMA2 + 2 LOP3", 3); run<8>("2 FFMA + 2 LOP3", 4); run<9>("2 LOP3", 2);
    return 0;
}
