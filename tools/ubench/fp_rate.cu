// fp_rate.cu -- peak issue rates of scalar vs packed FP32 on sm_100a (register-only loops, 8 independent chains)
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long pk;
__device__ __forceinline__ pk pk_add(pk a, pk b) { pk r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ pk pk_fma(pk a, pk b, pk c) { pk r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float s_add(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float s_fma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

template <int MODE> __global__ void __launch_bounds__(256) k(float *out, int iters, float seed)
{
    float x[16]; pk p[8];
    for (int i = 0; i < 16; ++i) x[i] = seed * (i + 1) + threadIdx.x;
    for (int i = 0; i < 8; ++i) { float lo = x[2 * i], hi = x[2 * i + 1]; asm("mov.b64 %0, {%1,%2};" : "=l"(p[i]) : "f"(lo), "f"(hi)); }
    const float c = seed * 0.5f; pk pc; asm("mov.b64 %0, {%1,%2};" : "=l"(pc) : "f"(c), "f"(c));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (MODE == 0) { for (int i = 0; i < 16; ++i) x[i] = s_add(x[i], c); }                 // 16 FADD
            if (MODE == 1) { for (int i = 0; i < 8; ++i) p[i] = pk_add(p[i], pc); }                // 8 FADD2
            if (MODE == 2) { for (int i = 0; i < 16; ++i) x[i] = s_fma(x[i], c, x[(i + 1) & 15]); }   // 16 FFMA 3-reg
            if (MODE == 3) { for (int i = 0; i < 8; ++i) p[i] = pk_fma(p[i], pc, p[(i + 1) & 7]); }   // 8 FFMA2
            if (MODE == 4) { for (int i = 0; i < 8; ++i) { p[i] = pk_add(p[i], pc); x[i] = s_add(x[i], c); } }   // mixed 8 + 8
        }
    }
    float s = 0; for (int i = 0; i < 16; ++i) s += x[i];
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, double lane_ops_per_iter, double instr_per_iter)
{
    float *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    k<MODE><<<148 * 8, 256>>>(out, 100, 1.0f);
    cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(out, iters, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double threads = 148.0 * 8 * 256, clk = 1.965e9 * ms * 1e-3;
    printf("%-28s %.3f ms  lane-ops/clk/SM %.1f  warp-instr/clk/SM %.2f (at 1965 MHz)\n", name, ms,
           threads * iters * lane_ops_per_iter / clk / 148, threads / 32 * iters * instr_per_iter / clk / 148);
    cudaFree(out);
}
int main()
{
    run<0>("FADD x16", 64, 64);
    run<1>("FADD2 x8", 64, 32);
    run<2>("FFMA x16 (3 reg)", 64, 64);
    run<3>("FFMA2 x8 (3 reg)", 64, 32);
    run<4>("FADD2 x8 + FADD x8", 96, 64);
    return 0;
}
