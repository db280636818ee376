// f32x2_probe.cu -- does packed FP32 (FADD2/FMUL2/FFMA2, PTX add/mul/fma.f32x2, sm_100+) buy issue slots for a
// radix-16 shared-memory FFT pass?  Two kernels run the same pass (16 LDS, butterfly, 15 twiddles from 2 loaded
// values + products, 16 STS) over a 64 KB buffer, ITER times:
//   scalar : 512 threads, one butterfly per thread on float2 (re, im)            -- the round-1 kernel's inner loop
//   packed : 256 threads, the same butterfly of TWO transforms per thread, element = (reA, reB, imA, imB)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o f32x2_probe f32x2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../spectrogram-rs_b200/csrc/sbx_pk.cuh"

using namespace sbx;

#define RS2 0.70710678118654752440f
#define C8 0.92387953251128675613f
#define S8 0.38268343236508977173f

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 csqr(float2 a) { return make_float2(a.x * a.x - a.y * a.y, 2.0f * (a.x * a.y)); }
__device__ __forceinline__ void bf4(float2 &a0, float2 &a1, float2 &a2, float2 &a3)
{
    float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
    a0 = cadd(t0, t2); a2 = csub(t0, t2);
    a1 = make_float2(t1.x + t3.y, t1.y - t3.x);
    a3 = make_float2(t1.x - t3.y, t1.y + t3.x);
}
__device__ __forceinline__ void bf16s(float2 (&v)[16])
{
    bf4(v[0], v[4], v[8], v[12]); bf4(v[1], v[5], v[9], v[13]); bf4(v[2], v[6], v[10], v[14]); bf4(v[3], v[7], v[11], v[15]);
    const float2 w1 = make_float2(C8, -S8), w3 = make_float2(S8, -C8), w9 = make_float2(-C8, S8);
    v[5] = cmul(v[5], w1);
    v[6] = make_float2((v[6].x + v[6].y) * RS2, (v[6].y - v[6].x) * RS2);
    v[7] = cmul(v[7], w3);
    v[9] = make_float2((v[9].x + v[9].y) * RS2, (v[9].y - v[9].x) * RS2);
    v[10] = make_float2(v[10].y, -v[10].x);
    v[11] = make_float2((v[11].y - v[11].x) * RS2, -(v[11].x + v[11].y) * RS2);
    v[13] = cmul(v[13], w3);
    v[14] = make_float2((v[14].y - v[14].x) * RS2, -(v[14].x + v[14].y) * RS2);
    v[15] = cmul(v[15], w9);
    bf4(v[0], v[1], v[2], v[3]); bf4(v[4], v[5], v[6], v[7]); bf4(v[8], v[9], v[10], v[11]); bf4(v[12], v[13], v[14], v[15]);
    float2 t;
    t = v[1]; v[1] = v[4]; v[4] = t;  t = v[2]; v[2] = v[8]; v[8] = t;  t = v[3]; v[3] = v[12]; v[12] = t;
    t = v[6]; v[6] = v[9]; v[9] = t;  t = v[7]; v[7] = v[13]; v[13] = t;  t = v[11]; v[11] = v[14]; v[14] = t;
}
__device__ __forceinline__ void tw16s(float2 (&v)[16], float2 w1, float2 w4)
{
    const float2 w2 = csqr(w1), w3 = cmul(w2, w1);
    v[1] = cmul(v[1], w1); v[2] = cmul(v[2], w2); v[3] = cmul(v[3], w3);
    const float2 w5 = cmul(w4, w1), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
    v[4] = cmul(v[4], w4); v[5] = cmul(v[5], w5); v[6] = cmul(v[6], w6); v[7] = cmul(v[7], w7);
    const float2 w8 = csqr(w4);
    v[8] = cmul(v[8], w8); v[9] = cmul(v[9], cmul(w8, w1)); v[10] = cmul(v[10], cmul(w8, w2)); v[11] = cmul(v[11], cmul(w8, w3));
    v[12] = cmul(v[12], cmul(w8, w4)); v[13] = cmul(v[13], cmul(w8, w5)); v[14] = cmul(v[14], cmul(w8, w6)); v[15] = cmul(v[15], cmul(w8, w7));
}

__global__ void __launch_bounds__(512, 2) scalar_kernel(float2 *g, const float2 *tw, int iters)
{
    extern __shared__ __align__(16) unsigned char raw[];
    float2 *z = reinterpret_cast<float2 *>(raw);
    const int t = threadIdx.x;
    for (int r = 0; r < 16; ++r) z[r * 512 + t] = g[(size_t)blockIdx.x * 8192 + r * 512 + t];
    __syncthreads();
    const float2 w1 = tw[t], w4 = tw[512 + t];
    for (int it = 0; it < iters; ++it) {
        float2 v[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = z[r * 512 + t];
        bf16s(v);
        tw16s(v, w1, w4);
#pragma unroll
        for (int r = 0; r < 16; ++r) z[r * 512 + ((t + 32) & 511)] = v[r];
        __syncthreads();
    }
    for (int r = 0; r < 16; ++r) g[(size_t)blockIdx.x * 8192 + r * 512 + t] = z[r * 512 + t];
}

__global__ void __launch_bounds__(256, 2) packed_kernel(float4 *g, const float2 *tw, int iters)
{
    extern __shared__ __align__(16) unsigned char raw[];
    float4 *z = reinterpret_cast<float4 *>(raw);
    const int t = threadIdx.x;
    for (int r = 0; r < 16; ++r) z[r * 256 + t] = g[(size_t)blockIdx.x * 4096 + r * 256 + t];
    __syncthreads();
    const float2 w1s = tw[t], w4s = tw[512 + t];
    const cpk w1 = cpk_dup(w1s), w4 = cpk_dup(w4s);
    for (int it = 0; it < iters; ++it) {
        cpk v[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = cpk_lds(z + r * 256 + t);
        pk_bf16(v);
        pk_tw16(v, w1, w4);
#pragma unroll
        for (int r = 0; r < 16; ++r) cpk_sts(z + r * 256 + ((t + 32) & 255), v[r]);
        __syncthreads();
    }
    for (int r = 0; r < 16; ++r) g[(size_t)blockIdx.x * 4096 + r * 256 + t] = z[r * 256 + t];
}

int main()
{
    const int ctas = 148 * 2 * 8, iters = 200;
    float2 *g; float2 *tw;
    cudaMalloc(&g, (size_t)ctas * 65536);
    cudaMalloc(&tw, 1024 * 8);
    float2 *h = (float2 *)malloc((size_t)ctas * 65536);
    for (size_t i = 0; i < (size_t)ctas * 8192; ++i) h[i] = make_float2((float)((i * 7919) % 1000) * 1e-3f - 0.5f, (float)((i * 104729) % 1000) * 1e-3f - 0.5f);
    float2 htw[1024];
    for (int i = 0; i < 1024; ++i) { double a = -6.283185307179586 * (i % 512) / 8192.0 * (i < 512 ? 1 : 4); htw[i] = make_float2((float)cos(a), (float)sin(a)); }
    cudaMemcpy(tw, htw, sizeof htw, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(scalar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    // correctness: one iteration each on the same data; packed element i of CTA c = (A.re, B.re, A.im, B.im)
    // where A = scalar elements [0, 4096) ... compare sums only (layouts differ): checksum of |z|^2
    for (int which = 0; which < 2; ++which) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaMemcpy(g, h, (size_t)ctas * 65536, cudaMemcpyHostToDevice);
            cudaEventRecord(e0);
            if (which == 0) scalar_kernel<<<ctas, 512, 65536>>>(g, tw, iters);
            else packed_kernel<<<ctas, 256, 65536>>>(reinterpret_cast<float4 *>(g), tw, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            cudaError_t err = cudaGetLastError();
            double pts = (double)ctas * 8192 * iters;
            printf("%s rep %d: %.3f ms, %.1f G complex points/s per pass, err=%s\n", which ? "packed" : "scalar", rep, ms, pts / ms * 1e-6, cudaGetErrorString(err));
        }
    }
    // numerical cross-check with 1 iteration on a small deterministic input where A == B == scalar input of CTA 0
    {
        float2 *ha = (float2 *)malloc(65536); float4 *hb = (float4 *)malloc(65536);
        // scalar CTA 0: elements r*512+t ; packed: element r*256+t of transform A uses scalar element r*512+t, B uses r*512+256+t
        cudaMemcpy(g, h, 65536, cudaMemcpyHostToDevice);
        scalar_kernel<<<1, 512, 65536>>>(g, tw, 1);
        cudaMemcpy(ha, g, 65536, cudaMemcpyDeviceToHost);
        for (int r = 0; r < 16; ++r) for (int t = 0; t < 256; ++t) {
            float2 A = h[r * 512 + t], B = h[r * 512 + t];
            hb[r * 256 + t] = make_float4(A.x, B.x, A.y, B.y);
        }
        cudaMemcpy(g, hb, 65536, cudaMemcpyHostToDevice);
        packed_kernel<<<1, 256, 65536>>>(reinterpret_cast<float4 *>(g), tw, 1);
        cudaMemcpy(hb, g, 65536, cudaMemcpyDeviceToHost);
        double worst = 0;
        for (int r = 0; r < 16; ++r) for (int t = 0; t < 256; ++t) {
            // scalar thread t wrote its outputs at ((t+32)&511); packed thread t at ((t+32)&255); threads t<224 map identically
            if (t >= 224) continue;
            float2 s = ha[r * 512 + t + 32]; float4 p = hb[r * 256 + t + 32];
            worst = fmax(worst, fabs(s.x - p.x)); worst = fmax(worst, fabs(s.x - p.y));
            worst = fmax(worst, fabs(s.y - p.z)); worst = fmax(worst, fabs(s.y - p.w));
        }
        printf("max |scalar - packed| after one pass: %.3e\n", worst);
    }
    return 0;
}
