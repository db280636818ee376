// stft_kernels.cuh -- sm_100a device code of the fused frame transform.
//
//   PCM window -> Hann -> (implicit) zero-pad -> M-point C2C FFT in shared memory -> L/R split,
//   2/N scale -> log-frequency resample -> dB step function -> LUT -> packed RGBA column
//
// One CTA transforms one complex sequence z[n] = a[n] + i b[n] of N windowed samples:
//   stereo PCM : a = left, b = right of ONE frame      (reference packing, src/fourier/fft.rs:57)
//   mono PCM   : a = frame 2c, b = frame 2c+1           (two frames share one transform; the
//                reference's (s,s) duplication gives L == R == |S[k]|*2/N)
// and the same conjugate-symmetry split (src/fourier/fft.rs:81-89) separates the two real
// sequences afterwards.  Long mono windows (real length >= 16384, Geom<.., REAL = true>) instead give
// every frame its own CTA: z[n] = x[2n] + i x[2n+1] halves the transform (the even/odd split of a real
// sequence), so the buffer is 64 KB instead of 128 KB per unit and two CTAs fit an SM.
// Other launch shapes of the same phases: a persistent grid with cp.async.bulk-staged PCM where one CTA owns an
// SM (fused_frames_persistent_kernel), several equally long recordings in one grid (KernelArgs::units_per_seq),
// and a lone stereo frame on four times the threads with a polled completion flag (Geom<.., WIDE>, done_flag).
//
// FFT structure (see tools/fft_model.py for the numpy model of the index math):
//   * the zero padding is never materialised: the radix-P DIF step over a sequence whose last
//     (P-1)N entries are zero degenerates to P sub-transforms of length N over z[n]*W_M^(n*q0);
//   * each sub-transform is an in-place radix-(2^rem,16,16,..) DIF, natural order in,
//     digit-reversed order out; bin kk = P*k + q0 lives at position q0*N + digitrev(k);
//   * 16 points per thread per pass in registers, one __syncthreads per pass;
//   * shared memory holds float2, XOR-swizzled at 16-byte granularity so that both the strided
//     passes (half-warp = one 128 B row) and the final stride-1 pass (thread = one 128 B row,
//     float4 accesses) are bank-conflict free.
#pragma once
#ifdef SBX_EMU
#include "emu_shim.h"      // tests/emu: host emulation of the phase functions below (test-only)
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "sbx_pk.cuh"

namespace sbx {

struct RowMap {            // one output row py (0 = lowest frequency) of the log-frequency bin map
    float f0;              // first sample frequency            (lin_space start)
    float step;            // (f1 - f0) / n                     (lin_space step)
    uint32_t n;            // samples averaged for this row
    uint32_t aux;          // first scratch slot of this row's samples within its chunk
};

struct RowChunk {          // consecutive rows whose interpolated samples fit the shared scratch area
    uint32_t row_begin, row_end;       // rows [row_begin, row_end)
    uint32_t samp_begin, samp_end;     // their entries in the per-sample tables
};

struct KernelArgs {
    const float *pcm;            // device PCM, CH floats per sample
    uint8_t *rgba;               // frames x H x 4 or nullptr
    float *mags_out;             // frames x K x 2 or nullptr
    const float *mags_in;        // colourise-only kernel: frames x K x 2
    const float *window;         // [N] periodic Hann, f32, host-computed (fft.rs:59-63)
    const float2 *wtw;           // [(P-1) N] window[n] * W_M^(n q0), q0 = 1..P-1 (sub-transform premultiply)
    const float2 *twp;           // per DIF pass p: TW_PLANES planes of S_p entries, plane j = W_(S_p R_p)^(2^j lo)
    const float2 *wq;            // q2 kernels: [S_0] W_M^lo (M = P N, the padded complex length), lo < the first pass's stride
    const float4 *wtw4;          // REAL: [(P-1) N/2] {w0 c, w1 s, w0 s, w1 c}: (w[2n], w[2n+1]) and (c, s) = W_(M/2)^(n q0)
    const float2 *rtw;           // REAL: [M/32] W_M^t, the even/odd recombination twiddle of bin t (M = real length)
    const RowMap *rowmap;        // [H] (host-side description; the kernels read rowinfo)
    const uint32_t *rowinfo;     // [H] n | first scratch slot << 16
    const RowChunk *chunks;      // [n_chunks]
    const uint2 *samp;           // per (row, sample): {floor(x) | scratch slot << 17, weight as f32 bits},
                                 //   x = index_of(f0 + s*step); weight = x - floor(x) (CUBIC) or
                                 //   (1 - cos(frac*pi))/2 (COSINE, host libm)
    const float *thr;            // [2][256] pairs (ts[i], ts[i+1]) of step thresholds with sentinels: colour index, alpha
    const uint32_t *lut;         // [256] packed r | g<<8 | b<<16 | 255<<24
    const uint4 *ctab;           // [256] {ts[i], ts[i+1], lut[i], 0}: the colour step and its colour in one 16-byte load
    const uint32_t *pan_lut;     // stereo schemes: [pan_levels] packed r | g<<8 | b<<16, indexed by floor(pan * pan_levels);
                                 //   = lut with 256 levels unless sbx_plan_set_pan_gradient installed a finer table
    uint32_t pan_levels;
    unsigned long long n_frames; // frames per sequence
    // several equally long sequences (planar channels) in one launch: units_per_seq CTAs each, seq_stride FLOATS apart
    // in pcm; sequence s writes frames [s * n_frames, (s+1) * n_frames).  units_per_seq = 0: one sequence.
    unsigned long long seq_stride;
    uint32_t units_per_seq;
    // live single-frame launches: the CTA publishes its completion to a polling host thread -- *done_count (device)
    // is incremented and the new value stored to *done_flag (pinned host memory) after a system-scope fence
    unsigned int *done_flag;
    unsigned int *done_count;
    uint32_t hop, rows, k_out, n_chunks;
    RowChunk chunk0;             // chunks[0] by value: most plans have a single chunk, and its descriptor would otherwise be
                                 // an L2 round trip that every thread waits for between the split and the resample
    float scale;                 // 2 / N     (fft.rs:92)
    float guess_a[2], guess_b[2];// step guess = log2(p + eps) * a + b   (colour, alpha)
    float db_eps;
    float step_margin;           // colour step: the guess is final when its fractional part lies in (margin, 1 - margin); >= 0.5: never
    float neg_zero;              // -0.0f, opaque to the compiler: stft_q2.cuh writes exactly rounded products as fma(x, y, -0.0)
    int interp;                  // 0 cubic, 1 cosine
    int scheme;                  // 0 mono gradient, 1 stereo pan gradient
    int taps_clamped;            // 1 if any cubic tap index needs clamping to [0, K-1] (spectrum edges)
};

#ifndef SBX_ROWSUM_PER_ROW
#define SBX_ROWSUM_PER_ROW 1     // colour phase: every row of a thread is summed with ITS OWN trip count (1), instead of all the
                                 // thread's rows in one loop of max(n) steps with predicated adds (0: round 1).  A thread's rows
                                 // are H/ROWS apart -- n = 1 in the low rows, up to 23 in the top ones -- so the fused loop spent
                                 // three of four slots predicated off; lanes of a warp hold ADJACENT rows, so per-row trip counts
                                 // barely diverge.  Same additions in the same order: bit-identical sums.
#endif
#ifndef SBX_ROWSUM_UNROLL
#define SBX_ROWSUM_UNROLL 4      // unroll factor of the per-row sum loops (loads issued ahead of the dependent adds)
#endif
#ifndef SBX_ROWSUM_PER_ROW_CLASSIC
#define SBX_ROWSUM_PER_ROW_CLASSIC 0   // the same choice for colour_rows of the classic / mixed-radix / multi-kernel paths (two rows per thread)
#endif
#ifndef SBX_PK_RESAMPLE
#define SBX_PK_RESAMPLE 1        // 1: the classic kernels resample / sum rows with packed (L, R) instructions too (pk_resample_colour).
                                 // Measured on a B200 with the per-row unrolled row sums (M frames/s, 0 -> 1): cfg5 N = 512 / 1024 /
                                 // 2048 / 4096 at P = 4: 113.3 -> 114.2, 84.0 -> 89.7, 48.7 -> 49.2, 25.0 -> 26.2; stereo N 4096 P 1:
                                 // 47.6 -> 49.6 (with the round-1 fused row-sum loop it lost, which is why it was off)
#endif
#ifndef SBX_STEP_FAST
#define SBX_STEP_FAST 1          // step_colour: accept the guessed step without the threshold load when it is clear of a boundary
#endif
#ifndef SBX_EPT_LOG2
#define SBX_EPT_LOG2 4           // log2 of the transform points each thread owns per pass (4: M/16 threads, 5: M/32)
#endif
#ifndef SBX_MIN_CTAS_256
#define SBX_MIN_CTAS_256 5       // resident-CTA target of the 256-thread kernels (M = 4096): 5 -> 48 registers, no spills; measured cfg2 53.7 / 55.3 / 50.7 M frames/s at 4 / 5 / 6
#endif
#ifndef SBX_MAX_THREADS
#define SBX_MAX_THREADS 1024     // CTA size cap: M = 16384 runs one 1024-thread CTA per SM (measured +7..12 % over 512)
#endif
constexpr int TW_PLANES = 4;     // twiddle planes stored per pass: W^(1,2,4,8 * lo)
#ifndef SBX_TW_LOADS
#define SBX_TW_LOADS 2           // planes actually loaded per butterfly (1: W^lo; 2: + W^4lo; 4: all); rest by products
#endif
constexpr uint32_t REAL_MIN_M = 8192, REAL_MAX_M = 32768;   // real transform lengths the REAL geometry is built for
// sample-table entries whose slot field is 0x7fff (sbx_tables.h: fillers that keep a lane group's taps inside one
// wavefront) compute a value but store nothing; lanes past the end of the table take the same form
constexpr uint32_t SAMP_NO_STORE = 0x7fffu << 17;
constexpr int ROWSUM_UNROLL = SBX_ROWSUM_UNROLL;
constexpr int THR_STRIDE = 512;  // floats per threshold set: 256 pairs (ts[i], ts[i+1]), ts = [-inf, thr[0..254], +inf]

// ----------------------------------------------------------------------------------------------
// complex helpers and small DFT butterflies (forward, exp(-i...))
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 mul_neg_i(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)
__device__ __forceinline__ float2 csqr(float2 a) { return make_float2(a.x * a.x - a.y * a.y, 2.0f * (a.x * a.y)); }

__device__ __forceinline__ float fast_sqrt(float x)
{
#ifdef SBX_EMU
    return sqrtf(x);
#else
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));     // MUFU.SQRT: ~1 ulp, sqrt(0) = 0
    return r;
#endif
}
__device__ __forceinline__ float fast_log2(float x)
{
#ifdef SBX_EMU
    return log2f(x);
#else
    return __log2f(x);
#endif
}

#define SBX_RSQRT2 0.70710678118654752440f
#define SBX_COS_PI_8 0.92387953251128675613f
#define SBX_SIN_PI_8 0.38268343236508977173f

__device__ __forceinline__ void bf2(float2 &a, float2 &b)
{
    float2 t = csub(a, b);
    a = cadd(a, b);
    b = t;
}

// 4-point DFT: y_q = sum_r a_r (-i)^(r q), outputs in natural order
__device__ __forceinline__ void bf4(float2 &a0, float2 &a1, float2 &a2, float2 &a3)
{
    float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = make_float2(t1.x + t3.y, t1.y - t3.x);
    a3 = make_float2(t1.x - t3.y, t1.y + t3.x);
}

template <int R> struct Butterfly;

template <> struct Butterfly<2> {
    static __device__ __forceinline__ void run(float2 (&v)[2]) { bf2(v[0], v[1]); }
};
template <> struct Butterfly<4> {
    static __device__ __forceinline__ void run(float2 (&v)[4]) { bf4(v[0], v[1], v[2], v[3]); }
};
template <> struct Butterfly<8> {
    // 8 = 2 x 4: n = a + 2a', k = b + 4c
    static __device__ __forceinline__ void run(float2 (&v)[8])
    {
        bf4(v[0], v[2], v[4], v[6]);            // a = 0 : t[0][b] in v[0],v[2],v[4],v[6]
        bf4(v[1], v[3], v[5], v[7]);            // a = 1 : t[1][b] in v[1],v[3],v[5],v[7]
        // t[1][b] *= W8^b
        v[3] = make_float2((v[3].x + v[3].y) * SBX_RSQRT2, (v[3].y - v[3].x) * SBX_RSQRT2);
        v[5] = mul_neg_i(v[5]);
        v[7] = make_float2((v[7].y - v[7].x) * SBX_RSQRT2, -(v[7].x + v[7].y) * SBX_RSQRT2);
        // X[b] = t0 + t1, X[b+4] = t0 - t1
        float2 x0 = cadd(v[0], v[1]), x4 = csub(v[0], v[1]);
        float2 x1 = cadd(v[2], v[3]), x5 = csub(v[2], v[3]);
        float2 x2 = cadd(v[4], v[5]), x6 = csub(v[4], v[5]);
        float2 x3 = cadd(v[6], v[7]), x7 = csub(v[6], v[7]);
        v[0] = x0; v[1] = x1; v[2] = x2; v[3] = x3; v[4] = x4; v[5] = x5; v[6] = x6; v[7] = x7;
    }
};
template <> struct Butterfly<16> {
    // 16 = 4 x 4: n = a + 4a', k = b + 4c;  X[b+4c] = sum_a (W16^(ab) sum_a' x[a+4a'] W4^(a'b)) W4^(ac)
    static __device__ __forceinline__ void run(float2 (&v)[16])
    {
        bf4(v[0], v[4], v[8], v[12]);           // a = 0: t[0][b] at v[4b]
        bf4(v[1], v[5], v[9], v[13]);           // a = 1: t[1][b] at v[1+4b]
        bf4(v[2], v[6], v[10], v[14]);          // a = 2
        bf4(v[3], v[7], v[11], v[15]);          // a = 3
        const float2 w1 = make_float2(SBX_COS_PI_8, -SBX_SIN_PI_8);
        const float2 w3 = make_float2(SBX_SIN_PI_8, -SBX_COS_PI_8);
        const float2 w9 = make_float2(-SBX_COS_PI_8, SBX_SIN_PI_8);
        // b = 1: W16^a
        v[5] = cmul(v[5], w1);
        v[6] = make_float2((v[6].x + v[6].y) * SBX_RSQRT2, (v[6].y - v[6].x) * SBX_RSQRT2);    // W16^2
        v[7] = cmul(v[7], w3);
        // b = 2: W16^(2a)
        v[9] = make_float2((v[9].x + v[9].y) * SBX_RSQRT2, (v[9].y - v[9].x) * SBX_RSQRT2);    // W16^2
        v[10] = mul_neg_i(v[10]);                                                              // W16^4
        v[11] = make_float2((v[11].y - v[11].x) * SBX_RSQRT2, -(v[11].x + v[11].y) * SBX_RSQRT2); // W16^6
        // b = 3: W16^(3a)
        v[13] = cmul(v[13], w3);
        v[14] = make_float2((v[14].y - v[14].x) * SBX_RSQRT2, -(v[14].x + v[14].y) * SBX_RSQRT2); // W16^6
        v[15] = cmul(v[15], w9);
        // outer 4-point DFTs over a for each b: inputs t[a][b] = v[a+4b], outputs X[b+4c] -> slot a=c
        bf4(v[0], v[1], v[2], v[3]);            // b = 0 -> X[0],X[4],X[8],X[12]
        bf4(v[4], v[5], v[6], v[7]);            // b = 1 -> X[1],X[5],X[9],X[13]
        bf4(v[8], v[9], v[10], v[11]);          // b = 2 -> X[2],X[6],X[10],X[14]
        bf4(v[12], v[13], v[14], v[15]);        // b = 3 -> X[3],X[7],X[11],X[15]
        // now v[c + 4b] = X[b + 4c]: transpose the 4x4 to natural order
        float2 t;
        t = v[1]; v[1] = v[4]; v[4] = t;
        t = v[2]; v[2] = v[8]; v[8] = t;
        t = v[3]; v[3] = v[12]; v[12] = t;
        t = v[6]; v[6] = v[9]; v[9] = t;
        t = v[7]; v[7] = v[13]; v[13] = t;
        t = v[11]; v[11] = v[14]; v[14] = t;
    }
};

// ---- cache-policy loads ------------------------------------------------------------------------
// PCM streams through once per CTA while ~100 KB of plan tables (window, twiddles, sample table, thresholds)
// are re-read by every CTA, so tables are loaded with L1::evict_last.  Measured on a B200 (cfg1, M frames/s):
// default/default 52.0, PCM evict_first 52.2, tables evict_last 52.9, PCM no_allocate + tables evict_last 52.5
// -- within noise: the kernel is issue- and shared-memory-bound, not L1-capacity-bound.
#ifndef SBX_PCM_POLICY
#define SBX_PCM_POLICY 0        // 0: default, 1: L1::evict_first, 2: L1::no_allocate
#endif
#ifndef SBX_TABLE_KEEP
#define SBX_TABLE_KEEP 1        // 1: tables with L1::evict_last
#endif
__device__ __forceinline__ float ldg_pcm(const float *p)
{
#if defined(SBX_EMU) || SBX_PCM_POLICY == 0
    return __ldg(p);
#elif SBX_PCM_POLICY == 1
    float v; asm volatile("ld.global.nc.L1::evict_first.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v;
#else
    float v; asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v;
#endif
}
__device__ __forceinline__ float2 ldg_pcm2(const float2 *p)
{
#if defined(SBX_EMU) || SBX_PCM_POLICY == 0
    return __ldg(p);
#elif SBX_PCM_POLICY == 1
    float2 v; asm volatile("ld.global.nc.L1::evict_first.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p)); return v;
#else
    float2 v; asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p)); return v;
#endif
}
__device__ __forceinline__ float ldg_tab(const float *p)
{
#if defined(SBX_EMU) || SBX_TABLE_KEEP == 0
    return __ldg(p);
#else
    float v; asm volatile("ld.global.nc.L1::evict_last.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v;
#endif
}
__device__ __forceinline__ float2 ldg_tab(const float2 *p)
{
#if defined(SBX_EMU) || SBX_TABLE_KEEP == 0
    return __ldg(p);
#else
    float2 v; asm volatile("ld.global.nc.L1::evict_last.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p)); return v;
#endif
}
__device__ __forceinline__ float4 ldg_tab(const float4 *p)
{
#if defined(SBX_EMU) || SBX_TABLE_KEEP == 0
    return __ldg(p);
#else
    float4 v;
    asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#endif
}
__device__ __forceinline__ uint2 ldg_tab(const uint2 *p)
{
#if defined(SBX_EMU) || SBX_TABLE_KEEP == 0
    return __ldg(p);
#else
    uint2 v; asm volatile("ld.global.nc.L1::evict_last.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p)); return v;
#endif
}
__device__ __forceinline__ uint4 ldg_tab(const uint4 *p)
{
#if defined(SBX_EMU) || SBX_TABLE_KEEP == 0
    return __ldg(p);
#else
    uint4 v;
    asm volatile("ld.global.nc.L1::evict_last.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
#endif
}
__device__ __forceinline__ uint32_t ldg_tab(const uint32_t *p)
{
#if defined(SBX_EMU) || SBX_TABLE_KEEP == 0
    return __ldg(p);
#else
    uint32_t v; asm volatile("ld.global.nc.L1::evict_last.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
#endif
}

// Big once-per-CTA tables (window, window * W, sample table: 30-50 KB each per CTA) in the q2 kernel, whose three
// resident CTAs leave only ~36 KB of L1: they stream through L2 anyway (L1 hit rate 30 %), so they can be kept from
// evicting the small hot tables (colour steps, row table, twiddles).  0: L1::evict_last like every table, 1: L1::evict_first,
// (measured, M frames/s at cfg1 / cfg2 geometry / cfg1 stereo: policy 0 64.3 / 55.7 / 32.8, 1: 63.5 / 55.9 / 32.7, 2: 64.6 / 57.5 / 32.9)
// 2: L1::no_allocate
#ifndef SBX_Q2_STREAM_POLICY
#define SBX_Q2_STREAM_POLICY 2
#endif
__device__ __forceinline__ float ldg_stream(const float *p)
{
#if defined(SBX_EMU) || SBX_Q2_STREAM_POLICY == 0
    return ldg_tab(p);
#elif SBX_Q2_STREAM_POLICY == 1
    float v; asm volatile("ld.global.nc.L1::evict_first.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v;
#else
    float v; asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v;
#endif
}
__device__ __forceinline__ float2 ldg_stream(const float2 *p)
{
#if defined(SBX_EMU) || SBX_Q2_STREAM_POLICY == 0
    return ldg_tab(p);
#elif SBX_Q2_STREAM_POLICY == 1
    float2 v; asm volatile("ld.global.nc.L1::evict_first.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p)); return v;
#else
    float2 v; asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p)); return v;
#endif
}
__device__ __forceinline__ float4 ldg_stream(const float4 *p)
{
#if defined(SBX_EMU) || SBX_Q2_STREAM_POLICY == 0
    return ldg_tab(p);
#elif SBX_Q2_STREAM_POLICY == 1
    float4 v;
    asm volatile("ld.global.nc.L1::evict_first.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#else
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#endif
}
__device__ __forceinline__ uint2 ldg_stream(const uint2 *p)
{
#if defined(SBX_EMU) || SBX_Q2_STREAM_POLICY == 0
    return ldg_tab(p);
#elif SBX_Q2_STREAM_POLICY == 1
    uint2 v; asm volatile("ld.global.nc.L1::evict_first.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p)); return v;
#else
    uint2 v; asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p)); return v;
#endif
}

// ----------------------------------------------------------------------------------------------
// compile-time geometry
// ----------------------------------------------------------------------------------------------
// REAL_: the complex sequence is the even/odd packing z[n] = x[2n] + i x[2n+1] of ONE real (mono) frame of 2 N
// samples, zero-padded to 2 M; addressed like stereo PCM (CH_ = 2) with hop/2 pairs between frames.
// WIDE_: single-frame (latency) variant -- up to four times the threads (M/4, at most 1024): the transform passes
// keep 16 points per thread (the extra warps idle there), the split, resample and colour phases spread over all.
template <int LOG2N_, int LOG2PAD_, int CH_, bool REAL_ = false, bool WIDE_ = false> struct Geom {
    static constexpr int LOG2N = LOG2N_, LOG2PAD = LOG2PAD_, CH = CH_;
    static constexpr bool REAL = REAL_, WIDE = WIDE_;
    static_assert(!REAL_ || CH_ == 2, "the real-sequence packing reads sample pairs");
    static constexpr int N = 1 << LOG2N, PAD = 1 << LOG2PAD, LOG2M = LOG2N + LOG2PAD, M = 1 << LOG2M;
    static constexpr int REM = LOG2N & 3;                  // first pass radix 2^REM when REM != 0
    static constexpr int NPASS = LOG2N / 4 + (REM ? 1 : 0);
    static constexpr int BASE_THREADS = (M >> SBX_EPT_LOG2) < 64 ? 64 : ((M >> SBX_EPT_LOG2) > SBX_MAX_THREADS ? SBX_MAX_THREADS : (M >> SBX_EPT_LOG2));
    static constexpr int WIDE_THREADS = (M / 4) > 1024 ? 1024 : (M / 4);
    static constexpr int THREADS = (WIDE_ && WIDE_THREADS > BASE_THREADS) ? WIDE_THREADS : BASE_THREADS;
    // register budget: 64 per thread (THREADS * MIN_CTAS = 1024).  Without the second launch bound ptxas aims
    // for three resident 512-thread CTAs at cfg1, caps the kernel at 40 registers and spills (measured: 49 vs
    // 52.5 M frames/s); the kernel needs 79 unconstrained and fits 64 without spilling.
    static constexpr int MIN_CTAS = (WIDE_ || THREADS >= 1024) ? 1 : (THREADS == 256 ? SBX_MIN_CTAS_256 : 1024 / THREADS);
    static constexpr int FRAMES_PER_CTA = CH == 1 ? 2 : 1;
    static constexpr int SMEM_BYTES = M * 8;
    static __host__ __device__ constexpr int log2r(int p) { return (REM != 0 && p == 0) ? REM : 4; }
    static __host__ __device__ constexpr int log2stride(int p)                 // log2 of the butterfly stride of pass p
    {
        int s = LOG2N;
        for (int i = 0; i <= p; ++i) s -= log2r(i);
        return s;
    }
    static __host__ __device__ constexpr int tw_offset(int p)                  // first entry of pass p in KernelArgs::twp
    {
        int o = 0;
        for (int i = 0; i < p; ++i) o += TW_PLANES << log2stride(i);
        return o;
    }
    static constexpr int TW_TOTAL = tw_offset(NPASS - 1);                      // the last pass has no twiddles
    static constexpr int SCRATCH = M / 2;                                      // resample scratch slots (upper half of the buffer)

    // bin kk = PAD*k + q0 lives at position q0*N + digitrev(k): a pure bit permutation of kk
    static __host__ __device__ constexpr int pos_of_bin(int kk)
    {
        const int q0 = kk & (PAD - 1);
        int k = kk >> LOG2PAD;
        int pos = q0 << LOG2N, shift = LOG2N;
        for (int p = 0; p < NPASS; ++p) {
            const int lr = log2r(p);
            shift -= lr;
            pos |= (k & ((1 << lr) - 1)) << shift;
            k >>= lr;
        }
        return pos;
    }
    static __host__ __device__ constexpr int pos_bit(int j)                    // position bit that bin bit j maps to
    {
        const int p = pos_of_bin(1 << j);
        int b = 0;
        while ((1 << b) < p) ++b;
        return b;
    }
    // Swizzle key (3 bits, XOR-linear in the position bits >= 4).  Key bit b = position bit 4+b, XORed
    // with one "external" position bit: the split phase reads positions that differ in the bits the
    // lowest bin bits map to (sub-transform select, first-pass digit), so those are folded in too.
    static __host__ __device__ constexpr int key_ext_src(int b)
    {
        bool used[3] = {false, false, false};
        int ext[3] = {-1, -1, -1};
        for (int j = 0; j < 3; ++j) {
            const int pb = pos_bit(j);
            if (pb >= 4 && pb <= 6) used[pb - 4] = true;
        }
        int next = 0;
        for (int j = 0; j < 3; ++j) {
            const int pb = pos_bit(j);
            if (pb > 6) {
                while (next < 3 && used[next]) ++next;
                if (next < 3) { ext[next] = pb; used[next] = true; }
            }
        }
        return ext[b];
    }
    static __host__ __device__ constexpr int zkey(int i)
    {
        int k = (i >> 4) & 7;
        constexpr int e0 = key_ext_src(0), e1 = key_ext_src(1), e2 = key_ext_src(2);
        if (e0 >= 0) k ^= (i >> (e0 >= 0 ? e0 : 0)) & 1;
        if (e1 >= 0) k ^= ((i >> (e1 >= 0 ? e1 : 0)) & 1) << 1;
        if (e2 >= 0) k ^= ((i >> (e2 >= 0 ? e2 : 0)) & 1) << 2;
        return k;
    }
    // element index -> physical float2 slot: permutes the eight 16-byte chunks of every 128-byte row.
    // XOR-linear: zsw(a | b) = zsw(a) ^ zsw(b) for bit-disjoint a, b.
    static __host__ __device__ constexpr int zsw(int i) { return i ^ (zkey(i) << 1); }
};

// Physical slot of element (base + off): `off` is a compile-time multiple of 16 whose bits are clear
// in `base`, `bx` = G::zsw(base).  One LOP3 (or nothing) per access instead of a full swizzle.
#define SBX_ZAT(G, bx, off) (((bx) ^ (G::zkey(off) << 1)) + (off))

// v[q] *= w^q for q = 1..R-1, where plane[j * plane_stride] = w^(2^j) are exactly rounded table
// values.  SBX_TW_LOADS of the planes are read; the other powers are products of at most three
// table values (loads are the scarce resource: the kernel is L1/shared-memory-pipe bound).
template <int R>
__device__ __forceinline__ void apply_twiddles(float2 (&v)[R], const float2 *plane, int plane_stride)
{
    const float2 w1 = ldg_tab(plane);
    v[1] = cmul(v[1], w1);
    if constexpr (R >= 4) {
        const float2 w2 = SBX_TW_LOADS >= 4 ? ldg_tab(plane + plane_stride) : csqr(w1);
        const float2 w3 = cmul(w2, w1);
        v[2] = cmul(v[2], w2);
        v[3] = cmul(v[3], w3);
        if constexpr (R >= 8) {
            const float2 w4 = SBX_TW_LOADS >= 2 ? ldg_tab(plane + 2 * plane_stride) : csqr(w2);
            const float2 w5 = cmul(w4, w1), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
            v[4] = cmul(v[4], w4);
            v[5] = cmul(v[5], w5);
            v[6] = cmul(v[6], w6);
            v[7] = cmul(v[7], w7);
            if constexpr (R >= 16) {
                const float2 w8 = SBX_TW_LOADS >= 4 ? ldg_tab(plane + 3 * plane_stride) : csqr(w4);
                v[8] = cmul(v[8], w8);
                v[9] = cmul(v[9], cmul(w8, w1));
                v[10] = cmul(v[10], cmul(w8, w2));
                v[11] = cmul(v[11], cmul(w8, w3));
                v[12] = cmul(v[12], cmul(w8, w4));
                v[13] = cmul(v[13], cmul(w8, w5));
                v[14] = cmul(v[14], cmul(w8, w6));
                v[15] = cmul(v[15], cmul(w8, w7));
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------
// DIF passes
// ----------------------------------------------------------------------------------------------

// Pass 0: PCM -> window -> sub-transform premultiply -> radix butterfly -> twiddle -> smem.
// STAGED = false: PCM is read from global memory (pa = this unit's first sample).
// STAGED = true : `pa` points at the unit's sample span in SHARED memory, put there by a bulk async
//                 copy (the persistent kernel); frame B of a mono pair is the same span shifted by hop.
// The R inputs of a butterfly are S samples apart: all addresses are one base pointer plus compile-time
// offsets.  Q0ZERO selects the window-only (q0 = 0) or window * W_M^(n q0) (q0 > 0) factor; q0 is
// warp-uniform, so the caller branches once per butterfly instead of predicating every load.
// SHIFT > 0 (mono pairs, global loads only): hop == SHIFT * S, so frame B's input r is frame A's input r + SHIFT
// and the pair needs R + SHIFT loads instead of 2 R.
template <class G, bool STAGED, bool Q0ZERO, int SHIFT = 0>
__device__ __forceinline__ void first_inputs(float2 (&v)[1 << G::log2r(0)], const KernelArgs &a, const float *pa_lo,
                                             const float *pb_lo, bool second_valid, int q0, int lo)
{
    constexpr int R = 1 << G::log2r(0), S = 1 << G::log2stride(0);
    const float *win_lo = a.window + lo;
    const float2 *wtw_lo = a.wtw + (Q0ZERO ? 0 : (q0 - 1) * G::N) + lo;
    float xs[SHIFT > 0 ? R + SHIFT : 1];
    if constexpr (SHIFT > 0) {
#pragma unroll
        for (int r = 0; r < R + SHIFT; ++r)
            xs[r] = (r < R || second_valid) ? ldg_pcm(pa_lo + r * S) : 0.0f;   // r >= R lies in frame B only
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float2 s;
        if constexpr (SHIFT > 0) {
            s.x = xs[r];
            s.y = second_valid ? xs[r + SHIFT] : 0.0f;
        } else if constexpr (STAGED) {
            if constexpr (G::CH == 2) {
                s = reinterpret_cast<const float2 *>(pa_lo)[r * S];
            } else {
                s.x = pa_lo[r * S];
                s.y = second_valid ? pb_lo[r * S] : 0.0f;
            }
        } else if constexpr (G::CH == 2) {
            s = ldg_pcm2(reinterpret_cast<const float2 *>(pa_lo) + r * S);
        } else {
            s.x = ldg_pcm(pa_lo + r * S);
            s.y = second_valid ? ldg_pcm(pb_lo + r * S) : 0.0f;
        }
        if constexpr (G::REAL) {
            if constexpr (Q0ZERO) {
                const float2 wn = ldg_tab(reinterpret_cast<const float2 *>(a.window) + lo + r * S);
                s.x *= wn.x;  s.y *= wn.y;                           // even and odd sample, each with its own window value
            } else {
                const float4 f = ldg_tab(a.wtw4 + (q0 - 1) * G::N + lo + r * S);
                s = make_float2(s.x * f.x - s.y * f.y, s.x * f.z + s.y * f.w);
            }
        } else if constexpr (Q0ZERO) {
            const float wn = ldg_tab(win_lo + r * S);
            s.x *= wn;  s.y *= wn;                                   // fft.rs:57-63 (complex * f32)
        } else {
            s = cmul(s, ldg_tab(wtw_lo + r * S));                    // window and W_M^(n q0) in one factor
        }
        v[r] = s;
    }
}

template <class G, bool STAGED = false>
__device__ __forceinline__ void pass_first(float2 *z, const KernelArgs &a, const float *pa, bool second_valid)
{
    constexpr int LR = G::log2r(0), R = 1 << LR, LS = G::log2stride(0), S = 1 << LS;
    constexpr int PER_SUB = G::N / R, COUNT = G::PAD * PER_SUB;
    for (int w = threadIdx.x; w < COUNT; w += G::THREADS) {
        const int q0 = w / PER_SUB, lo = w % PER_SUB;                // q0 is warp-uniform (PER_SUB % 32 == 0)
        // CH == 2: pa addresses interleaved (l,r) pairs; frame B of a mono pair starts hop samples later
        const float *pa_lo = pa + (size_t)lo * G::CH;
        const float *pb_lo = pa_lo + a.hop;
        float2 v[R];
        if constexpr (G::CH == 1 && !STAGED && R >= 8) {
            // the common overlaps: hop = S or 2 S (cfg1, cfg4: hop 512 = 2 * 256)
            const uint32_t hop = a.hop;
            if (hop == 2u * S) {
                if (q0 == 0) first_inputs<G, STAGED, true, 2>(v, a, pa_lo, pb_lo, second_valid, q0, lo);
                else first_inputs<G, STAGED, false, 2>(v, a, pa_lo, pb_lo, second_valid, q0, lo);
            } else if (hop == 1u * S) {
                if (q0 == 0) first_inputs<G, STAGED, true, 1>(v, a, pa_lo, pb_lo, second_valid, q0, lo);
                else first_inputs<G, STAGED, false, 1>(v, a, pa_lo, pb_lo, second_valid, q0, lo);
            } else {
                if (q0 == 0) first_inputs<G, STAGED, true>(v, a, pa_lo, pb_lo, second_valid, q0, lo);
                else first_inputs<G, STAGED, false>(v, a, pa_lo, pb_lo, second_valid, q0, lo);
            }
        } else {
            if (q0 == 0) first_inputs<G, STAGED, true>(v, a, pa_lo, pb_lo, second_valid, q0, lo);
            else first_inputs<G, STAGED, false>(v, a, pa_lo, pb_lo, second_valid, q0, lo);
        }
        Butterfly<R>::run(v);
        if constexpr (G::NPASS > 1) apply_twiddles<R>(v, a.twp + lo, S);   // W_N^(lo q)
        const int bx = G::zsw((q0 << G::LOG2N) + lo);
#pragma unroll
        for (int q = 0; q < R; ++q) z[SBX_ZAT(G, bx, q * S)] = v[q];
    }
}

// Pass P >= 1: in-place radix-16 butterflies at stride S = 2^LS; the last pass (S = 1) moves whole
// 128-byte rows with float4 accesses
// (tid: the thread's index within the G::THREADS threads that share the transform buffer z -- threadIdx.x, or the
// index within a thread group where one CTA holds several independent buffers, stft_large.cuh)
template <class G, int P>
__device__ __forceinline__ void pass_inplace_t(float2 *z, const KernelArgs &a, int tid)
{
    constexpr int LS = G::log2stride(P), S = 1 << LS;
    constexpr int COUNT = G::M / 16;
    constexpr int PER_SUB_LOG2 = G::LOG2N - 4;
    for (int w = tid; w < COUNT; w += G::THREADS) {
        float2 v[16];
        if constexpr (S == 1) {
            // row w holds elements 16w .. 16w+15
            float4 *row = reinterpret_cast<float4 *>(z) + (w << 3);
            const int key = G::zkey(w << 4);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 t = row[c ^ key];
                v[2 * c] = make_float2(t.x, t.y);
                v[2 * c + 1] = make_float2(t.z, t.w);
            }
            Butterfly<16>::run(v);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                row[c ^ key] = make_float4(v[2 * c].x, v[2 * c].y, v[2 * c + 1].x, v[2 * c + 1].y);
        } else {
            const int blk = w >> PER_SUB_LOG2, u = w & ((1 << PER_SUB_LOG2) - 1);
            const int lo = u & (S - 1), hi = u >> LS;
            const int bx = G::zsw((blk << G::LOG2N) + (hi << (LS + 4)) + lo);
#pragma unroll
            for (int r = 0; r < 16; ++r) v[r] = z[SBX_ZAT(G, bx, r * S)];
            Butterfly<16>::run(v);
            apply_twiddles<16>(v, a.twp + G::tw_offset(P) + lo, S);        // W_(16 S)^(lo q)
#pragma unroll
            for (int q = 0; q < 16; ++q) z[SBX_ZAT(G, bx, q * S)] = v[q];
        }
    }
}
template <class G, int P>
__device__ __forceinline__ void pass_inplace(float2 *z, const KernelArgs &a)
{
    pass_inplace_t<G, P>(z, a, (int)threadIdx.x);
}

// passes P .. END-1, each followed by a barrier
template <class G, int P, int END = G::NPASS> struct RunPasses {
    static __device__ __forceinline__ void run(float2 *z, const KernelArgs &a)
    {
        if constexpr (P < END) {
            pass_inplace<G, P>(z, a);
            __syncthreads();
            RunPasses<G, P + 1, END>::run(z, a);
        }
    }
};

// ----------------------------------------------------------------------------------------------
// L/R split (fft.rs:81-95): bins kk = 1 .. M/2-1, a = Z[kk], b = Z[M-kk]
//   L = |a + conj b| / 2 * (2/N),  R = |a - conj b| / 2 * (2/N)
// Thread t owns bins kk = t | (p << log2 THREADS).  pos_of_bin is a bit permutation and the swizzle is
// XOR-linear, so slot(kk) = zsw(pos(t)) ^ const(p); the partner M-kk = (T-t) | ((~p & mask) << log2 T)
// for t != 0 splits the same way.  split_read: pairs -> registers; (barrier); split_write: (L,R) is
// written over the lower half of the transform buffer in natural bin order (no swizzle).
// ----------------------------------------------------------------------------------------------
template <class G> struct SplitRegs {
    static constexpr int PAIRS = G::M / 2;
    static constexpr int PER_THREAD = PAIRS / G::THREADS;
    static_assert(PAIRS % G::THREADS == 0 && (G::THREADS & (G::THREADS - 1)) == 0, "split mapping");
    float2 mag[PER_THREAD];
};

template <class G> struct SplitConst {
    static __host__ __device__ constexpr int log2_threads() { int l = 0; while ((1 << l) < G::THREADS) ++l; return l; }
    static constexpr int LT = log2_threads();
    static constexpr int PMASK = (G::M >> LT) - 1;
    static __host__ __device__ constexpr int own(int p) { return G::zsw(G::pos_of_bin(p << LT)); }
    static __host__ __device__ constexpr int partner(int p) { return G::zsw(G::pos_of_bin(((~p) & PMASK) << LT)); }
    // thread 0: M - kk carries into the p field; its kk = 0 slot (DC, dropped) pairs with itself, or -- REAL -- is
    // used for bin M/2, the one bin that is its own partner
    static __host__ __device__ constexpr int partner_t0(int p)
    {
        return p == 0 ? (G::REAL ? G::zsw(G::pos_of_bin(G::M / 2)) : 0) : G::zsw(G::pos_of_bin(G::M - (p << LT)));
    }
    // the bit of p that maps to position bit 0: p and p | (1 << B0) sit in one 16-byte chunk
    static __host__ __device__ constexpr int pair_bit()
    {
        for (int b = 0; (1 << b) < (G::M / 2) / G::THREADS; ++b)
            if (G::pos_of_bin(1 << (LT + b)) == 1) return b;
        return -1;
    }
    static constexpr int B0 = pair_bit();
};

__device__ __forceinline__ float2 split_pair(float2 x, float2 y, float scale)
{
    const float lx = x.x + y.x, ly = x.y - y.y;   // a + conj b
    const float rx = x.x - y.x, ry = x.y + y.y;   // a - conj b
    const float l = fast_sqrt(lx * lx + ly * ly) * 0.5f;
    const float rr = fast_sqrt(rx * rx + ry * ry) * 0.5f;
    return make_float2(l * scale, rr * scale);
}

// REAL: a = Z[kk], b = Z[M-kk] of the packed sequence, w = W_(2M)^kk.  With E, O the spectra of the even and odd
// samples, a + conj b = 2 E and a - conj b = 2i O, and the real sequence's spectrum is X[kk] = E + w O,
// X[M-kk] = conj(E - w O).  Returns (|X[kk]|, |X[M-kk]|) * scale  (L == R == |X| * 2/N for (s,s) input).
__device__ __forceinline__ float2 split_pair_real(float2 x, float2 y, float2 w, float scale)
{
    const float sx = x.x + y.x, sy = x.y - y.y;   // a + conj b
    const float dx = x.x - y.x, dy = x.y + y.y;   // a - conj b
    const float tx = w.x * dy + w.y * dx;         // w * (-i d)
    const float ty = w.y * dy - w.x * dx;
    const float ox = sx + tx, oy = sy + ty, px = sx - tx, py = sy - ty;
    return make_float2(fast_sqrt(ox * ox + oy * oy) * 0.5f * scale, fast_sqrt(px * px + py * py) * 0.5f * scale);
}
// W_32^p = exp(-i pi p / 16), p = 0..7: the part of the recombination twiddle that depends on the bin's p field
__host__ __device__ constexpr float w32_cos(int p)
{
    return p == 0 ? 1.0f : p == 1 ? 0.98078528040323044913f : p == 2 ? 0.92387953251128675613f : p == 3 ? 0.83146961230254523708f
         : p == 4 ? 0.70710678118654752440f : p == 5 ? 0.55557023301960222474f : p == 6 ? 0.38268343236508977173f : 0.19509032201612826785f;
}
__host__ __device__ constexpr float w32_sin(int p) { return p == 0 ? 0.0f : w32_cos(8 - p); }

template <class G>
__device__ __forceinline__ void split_read(const float2 *z, const KernelArgs &a, SplitRegs<G> &r)
{
    using C = SplitConst<G>;
    constexpr int PER = SplitRegs<G>::PER_THREAD;
    const int tid = threadIdx.x;
    const int own = G::zsw(G::pos_of_bin(tid));
    const int par = G::zsw(G::pos_of_bin((G::THREADS - tid) & (G::THREADS - 1)));
    if constexpr (C::B0 >= 0) {
        // 16-byte accesses: (pa, pb = pa | 1<<B0) share a chunk; the partners (M-kk) share one too, swapped
#pragma unroll
        for (int pa = 0; pa < PER; ++pa) {
            if ((pa >> C::B0) & 1) continue;
            const int pb = pa | (1 << C::B0);
            const float4 x = *reinterpret_cast<const float4 *>(z + (own ^ C::own(pa)));
            float4 y = *reinterpret_cast<const float4 *>(z + (par ^ C::partner(pb)));   // .xy: partner of pb, .zw: of pa
            if (tid == 0) {                                                             // M - kk carries into the p field
                const float2 ya = z[C::partner_t0(pa)], yb = z[C::partner_t0(pb)];
                y = make_float4(yb.x, yb.y, ya.x, ya.y);
            }
            r.mag[pa] = split_pair(make_float2(x.x, x.y), make_float2(y.z, y.w), a.scale);
            r.mag[pb] = split_pair(make_float2(x.z, x.w), make_float2(y.x, y.y), a.scale);
        }
    } else {
#pragma unroll
        for (int p = 0; p < PER; ++p)
            r.mag[p] = split_pair(z[own ^ C::own(p)], z[tid == 0 ? C::partner_t0(p) : (par ^ C::partner(p))], a.scale);
    }
}

template <class G>
__device__ __forceinline__ void split_write(float2 *z, const SplitRegs<G> &r)
{
    const int tid = threadIdx.x;
    if constexpr (G::REAL) {
        // one magnitude per bin kk = 1 .. M-1 of the real spectrum, natural order, over the lower half of the buffer
        float *zf = reinterpret_cast<float *>(z);
#pragma unroll
        for (int p = 0; p < SplitRegs<G>::PER_THREAD; ++p) {
            const int kk = tid | (p << SplitConst<G>::LT);
            if (kk != 0) { zf[kk] = r.mag[p].x; zf[G::M - kk] = r.mag[p].y; }
            else zf[G::M / 2] = r.mag[p].x;                                        // thread 0's DC slot carries bin M/2
        }
    } else {
#pragma unroll
        for (int p = 0; p < SplitRegs<G>::PER_THREAD; ++p)
            if (p != 0 || tid != 0) z[tid | (p << SplitConst<G>::LT)] = r.mag[p];  // kk = 0 (DC) is dropped
    }
}

// ---- last pass fused with the split (THREADS == M/16, i.e. one row per thread) -----------------------
// The stride-1 pass works on whole rows and any thread may take any row, so thread t takes the row that holds
// ITS split bins: row pos_of_bin(t) >> 4, whose elements 0..7 are the bins t | p << log2 T (p = 0..7).  Those stay
// in registers; only elements 8..15 -- the partners some other thread needs -- go back to shared memory.
// Saves the write of the lower half rows and the split's read of them (one of five 64 KB round trips).
template <class G> struct LastSplit {
    static constexpr bool ENABLED = G::THREADS * 16 == G::M && G::NPASS >= 2 && SplitRegs<G>::PER_THREAD == 8 &&
                                    SplitConst<G>::B0 == 0;
};

template <class G>
__device__ __forceinline__ void last_pass_keep_own(float2 *z, SplitRegs<G> &r)
{
    const int w = G::pos_of_bin(threadIdx.x) >> 4;                            // my row
    float4 *row = reinterpret_cast<float4 *>(z) + (w << 3);
    const int key = G::zkey(w << 4);
    float2 v[16];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 t = row[c ^ key];
        v[2 * c] = make_float2(t.x, t.y);
        v[2 * c + 1] = make_float2(t.z, t.w);
    }
    Butterfly<16>::run(v);
#pragma unroll
    for (int p = 0; p < 8; ++p) r.mag[p] = v[p];                              // own bins, kept in registers
#pragma unroll
    for (int c = 4; c < 8; ++c)                                               // elements 8..15: somebody's partners
        row[c ^ key] = make_float4(v[2 * c].x, v[2 * c].y, v[2 * c + 1].x, v[2 * c + 1].y);
}

// after the barrier: r.mag[p] (own Z values) -> (L,R) magnitudes, partners read 16 bytes at a time
template <class G>
__device__ __forceinline__ void split_with_own(const float2 *z, const KernelArgs &a, SplitRegs<G> &r)
{
    using C = SplitConst<G>;
    const int tid = threadIdx.x;
    const int par = G::zsw(G::pos_of_bin((G::THREADS - tid) & (G::THREADS - 1)));
    float2 wt = make_float2(1.0f, 0.0f);
    if constexpr (G::REAL) wt = ldg_tab(a.rtw + tid);                                 // W_(2M)^tid
#pragma unroll
    for (int pa = 0; pa < 8; pa += 2) {
        const int pb = pa + 1;
        float4 y = *reinterpret_cast<const float4 *>(z + (par ^ C::partner(pb)));     // .xy: partner of pb, .zw: of pa
        if (tid == 0) {
            const float2 ya = z[C::partner_t0(pa)], yb = z[C::partner_t0(pb)];
            y = make_float4(yb.x, yb.y, ya.x, ya.y);
        }
        if constexpr (G::REAL) {
            // W_(2M)^kk, kk = tid | p << log2 T:  W_(2M)^tid * W_32^p
            const float2 wa = pa == 0 ? wt : cmul(wt, make_float2(w32_cos(pa), -w32_sin(pa)));
            const float2 wb = cmul(wt, make_float2(w32_cos(pb), -w32_sin(pb)));
            if (pa == 0 && tid == 0)                                                  // y.zw = Z[M/2]: |X[M/2]| = |Z[M/2]|
                r.mag[0] = make_float2(fast_sqrt(y.z * y.z + y.w * y.w) * a.scale, 0.0f);
            else
                r.mag[pa] = split_pair_real(r.mag[pa], make_float2(y.z, y.w), wa, a.scale);
            r.mag[pb] = split_pair_real(r.mag[pb], make_float2(y.x, y.y), wb, a.scale);
        } else {
            r.mag[pa] = split_pair(r.mag[pa], make_float2(y.z, y.w), a.scale);
            r.mag[pb] = split_pair(r.mag[pb], make_float2(y.x, y.y), a.scale);
        }
    }
}

// ----------------------------------------------------------------------------------------------
// log-frequency resample + colour  (interpolated_frequency_sample.rs:60-105, colorscheme.rs:55-71,
// simple_spectrogram.rs:141-161).  Everything here is written with explicitly rounded intrinsics
// (no FMA contraction) in the reference's operation order: given identical magnitudes the result is
// bit-identical to the CPU oracle.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float cubic1(float y0, float y1, float y2, float y3, float mu, float mu2, float mu3)
{
    const float a0 = __fadd_rn(__fsub_rn(__fsub_rn(y3, y2), y0), y1);
    const float a1 = __fsub_rn(__fsub_rn(y0, y1), a0);
    const float a2 = __fsub_rn(y2, y0);
    return __fadd_rn(__fadd_rn(__fmul_rn(a0, mu3), __fmul_rn(a1, mu2)), __fadd_rn(__fmul_rn(a2, mu), y1));
}

// Step function of p: the number of thresholds <= p.  ts = [-inf, thr[0..254], +inf], stored as the pairs
// tp[i] = (ts[i], ts[i+1]) so that one 8-byte load brackets the guess.  The guess log2(p + eps) * ga + gb is the
// same dB chain evaluated with the fast log, and is corrected against the exact host-computed thresholds, so
// the result is exact whatever the error of the guess.
__device__ __forceinline__ int step_index(const float *ts, float p, float eps, float ga, float gb)
{
    const float2 *tp = reinterpret_cast<const float2 *>(ts);
    const float t = fast_log2(p + eps) * ga + gb;
    int idx = (int)fminf(fmaxf(t, 0.0f), 255.0f);           // NaN -> 0
    const float2 lh = ldg_tab(tp + idx);                     // ts[idx] <= p < ts[idx+1]  <=>  step == idx
    if (p >= lh.y) {
        do { ++idx; } while (idx < 255 && p >= ldg_tab(tp + idx).y);
        if (idx > 255) idx = 255;
    } else if (p < lh.x) {
        do { --idx; } while (idx > 0 && p < ldg_tab(tp + idx).x);
    }
    return idx;
}

// Mono schemes: the packed RGBA of power p.  One 16-byte load brings the bracket of the guessed step and its
// colour; only a wrong guess (p within a rounding error of a threshold) takes the exact search.
__device__ __forceinline__ uint32_t step_colour(const KernelArgs &a, float p)
{
    const float t = fast_log2(p + a.db_eps) * a.guess_a[0] + a.guess_b[0];
#if SBX_STEP_FAST
    // The guess and the reference's f32 dB chain both follow T(p) = (10 log10(p + eps) - db_min) / (db_max - db_min) * 256
    // to within a bound the plan has verified against every exact threshold (sbx_tables.h: step_margin): away from the
    // integers by more than that, floor(guess) IS the step, and the pixel needs the 4-byte colour only.  NaN / inf
    // guesses fail both comparisons and take the exact path below.
    const float fl = floorf(t), fr = t - fl;
    if (fr > a.step_margin && fr < 1.0f - a.step_margin)
        return ldg_tab(a.lut + (int)fminf(fmaxf(fl, 0.0f), 255.0f));
#endif
    const int idx = (int)fminf(fmaxf(t, 0.0f), 255.0f);      // NaN -> 0
    const uint4 e = ldg_tab(a.ctab + idx);
    if (p >= __uint_as_float(e.y) || p < __uint_as_float(e.x))
        return ldg_tab(a.lut + step_index(a.thr, p, a.db_eps, a.guess_a[0], a.guess_b[0]));
    return e.z;
}

// One interpolated sample -- magnitude_at (ifs.rs:46-48).  mags is indexed by bin kk = j + 1; x1 = floor(x),
// w = the interpolation weight, both precomputed by the plan.  T = float2: both channels (or both frames of a
// mono pair); T = float: the single frame of the REAL geometry.
__device__ __forceinline__ float cubicv(float y0, float y1, float y2, float y3, float mu, float mu2, float mu3)
{
    return cubic1(y0, y1, y2, y3, mu, mu2, mu3);
}
__device__ __forceinline__ float2 cubicv(float2 y0, float2 y1, float2 y2, float2 y3, float mu, float mu2, float mu3)
{
    return make_float2(cubic1(y0.x, y1.x, y2.x, y3.x, mu, mu2, mu3), cubic1(y0.y, y1.y, y2.y, y3.y, mu, mu2, mu3));
}
__device__ __forceinline__ float mixv(float yl, float yh, float omt, float w) { return __fadd_rn(__fmul_rn(yl, omt), __fmul_rn(yh, w)); }
__device__ __forceinline__ float2 mixv(float2 yl, float2 yh, float omt, float w)
{
    return make_float2(mixv(yl.x, yh.x, omt, w), mixv(yl.y, yh.y, omt, w));
}

template <bool CLAMP, class T>
__device__ __forceinline__ T interp_sample(const T *mags, const KernelArgs &a, int x1, float w, int k_last)
{
    if (a.interp == 0) {                                                     // cubic_interpolate, ifs.rs:88-105
        const float mu = w;
        const float mu2 = __fmul_rn(mu, mu), mu3 = __fmul_rn(mu, mu2);
        T y0, y1, y2, y3;
        if constexpr (CLAMP) {
            y0 = mags[(x1 > 0 ? x1 - 1 : 0) + 1];
            y1 = mags[x1 + 1];
            y2 = mags[min(x1 + 1, k_last) + 1];
            y3 = mags[min(x1 + 2, k_last) + 1];
        } else {
            const T *m = mags + x1;
            y0 = m[0]; y1 = m[1]; y2 = m[2]; y3 = m[3];
        }
        return cubicv(y0, y1, y2, y3, mu, mu2, mu3);
    } else {                                                                 // cosine_interpolate, ifs.rs:78-86
        const float omt = __fsub_rn(1.0f, w);
        const T yl = mags[x1 + 1], yh = mags[min(x1 + 1, k_last) + 1];       // high = clamp(ceil x, low+1, K-1)
        return mixv(yl, yh, omt, w);
    }
}

// Phase A of a row chunk: one lane per (row, sample); consecutive lanes read consecutive bins.
// The sample-table loads of four iterations are issued together (latency hiding).
template <int THREADS, bool CLAMP, class T, int UNROLL = 4>
__device__ __forceinline__ void resample_samples_impl(const T *mags, T *scratch, const KernelArgs &a, const RowChunk ck)
{
    const int k_last = (int)a.k_out - 1;
    const uint32_t count = ck.samp_end - ck.samp_begin;
    const uint2 *tab = a.samp + ck.samp_begin;
    for (uint32_t g0 = threadIdx.x; g0 < count; g0 += UNROLL * THREADS) {
        uint2 sp[UNROLL];
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {
            const uint32_t g = g0 + i * THREADS;
            sp[i] = g < count ? ldg_tab(tab + g) : make_uint2(SAMP_NO_STORE, 0u);
        }
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {
            if (sp[i].x < SAMP_NO_STORE)                                       // a real entry (not a filler, not past the end)
                scratch[sp[i].x >> 17] = interp_sample<CLAMP>(mags, a, (int)(sp[i].x & 0x1ffffu), __uint_as_float(sp[i].y), k_last);
        }
    }
}
template <int THREADS, class T, int UNROLL = 4>
__device__ __forceinline__ void resample_samples(const T *mags, T *scratch, const KernelArgs &a, const RowChunk ck)
{
    if (a.taps_clamped) resample_samples_impl<THREADS, true, T, UNROLL>(mags, scratch, a, ck);
    else resample_samples_impl<THREADS, false, T, UNROLL>(mags, scratch, a, ck);
}

// Stereo-scheme gradient index of a mono pixel (l == r >= 0): floor(l / (|l| + |l|) * 256) with the reference's f32
// sum -- 128 while l + l is finite, 0 once it overflows (l > FLT_MAX / 2: l / inf = 0) and for l = 0 or NaN (NaN -> 0)
__device__ __forceinline__ uint32_t pan_index_equal(float l, uint32_t levels)
{
    // floor(0.5 * levels) = levels >> 1 (clamped below levels for every levels >= 1)
    return (l > 0.0f && __fadd_rn(l, l) <= 3.402823466e+38f) ? (levels >> 1) : 0u;
}

// Phase B: one thread per row: sequential mean of its samples (the reference's summation order,
// ifs.rs:71-73), dB step, LUT, packed RGBA store.  Two rows per thread are kept in flight (their
// loads and the dependent chains of the two colour look-ups overlap).
template <bool PAIR>
__device__ __forceinline__ void colour_one_row(float ax, float ay, uint32_t n, uint32_t py, const float *thr, const uint32_t *lut,
                                               const KernelArgs &a, uint32_t *outA, uint32_t *outB, bool second_valid)
{
    if (n > 1) {                                                             // x / 1.0f == x exactly
        const float nf = (float)n;
        ax = __fdiv_rn(ax, nf);
        ay = __fdiv_rn(ay, nf);
    }
    const uint32_t row = a.rows - 1u - py;                                   // simple_spectrogram.rs:150
    const float *thr_a = thr + THR_STRIDE;
    if constexpr (PAIR) {
        // mono: (ax, ay) are the L(=R) magnitudes of frames A and B
        const float pA = __fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ax, ax));
        const float pB = __fadd_rn(__fmul_rn(ay, ay), __fmul_rn(ay, ay));
        if (a.scheme == 0) {
            const uint32_t cA = step_colour(a, pA), cB = step_colour(a, pB);
            outA[row] = cA;
            if (second_valid) outB[row] = cB;
        } else {
            // l == r: pan = l / (|l| + |r|) = 0.5 while the f32 sum l + l is finite (l <= FLT_MAX / 2); beyond that the
            // reference divides by inf -> pan 0, and l = 0 or NaN gives NaN -> index 0 as well
            const uint32_t panA = pan_index_equal(ax, a.pan_levels), panB = pan_index_equal(ay, a.pan_levels);
            outA[row] = (ldg_tab(a.pan_lut + panA) & 0x00ffffffu) | ((uint32_t)step_index(thr_a, pA, a.db_eps, a.guess_a[1], a.guess_b[1]) << 24);
            if (second_valid)
                outB[row] = (ldg_tab(a.pan_lut + panB) & 0x00ffffffu) | ((uint32_t)step_index(thr_a, pB, a.db_eps, a.guess_a[1], a.guess_b[1]) << 24);
        }
    } else {
        const float p = __fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay));    // norm_sqr
        if (a.scheme == 0) {
            outA[row] = step_colour(a, p);
        } else {
            const double pan = (double)ax / (double)__fadd_rn(fabsf(ax), fabsf(ay));
            const double v = floor(pan * (double)a.pan_levels);
            const uint32_t idx = !(v >= 0.0) ? 0u : (v > (double)(a.pan_levels - 1u) ? a.pan_levels - 1u : (uint32_t)v);
            outA[row] = (ldg_tab(a.pan_lut + idx) & 0x00ffffffu) | ((uint32_t)step_index(thr_a, p, a.db_eps, a.guess_a[1], a.guess_b[1]) << 24);
        }
    }
}

// REAL geometry: one mono frame, ax = the row's summed L (== R) magnitude
__device__ __forceinline__ void colour_one_row_single(float ax, uint32_t n, uint32_t py, const KernelArgs &a, uint32_t *out)
{
    if (n > 1) ax = __fdiv_rn(ax, (float)n);
    const uint32_t row = a.rows - 1u - py;                                   // simple_spectrogram.rs:150
    const float p = __fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ax, ax));        // norm_sqr of (l, l)
    if (a.scheme == 0) {
        out[row] = step_colour(a, p);
    } else {
        const uint32_t pan = pan_index_equal(ax, a.pan_levels);              // l / (|l| + |l|), see colour_one_row
        out[row] = (ldg_tab(a.pan_lut + pan) & 0x00ffffffu) |
                   ((uint32_t)step_index(a.thr + THR_STRIDE, p, a.db_eps, a.guess_a[1], a.guess_b[1]) << 24);
    }
}

template <int THREADS, int ROWS = 2>
__device__ __forceinline__ void colour_rows_single(const float *scratch, const KernelArgs &a, const RowChunk ck, unsigned long long frame)
{
    uint32_t *out = reinterpret_cast<uint32_t *>(a.rgba) + frame * a.rows;
    for (uint32_t py0 = ck.row_begin + threadIdx.x; py0 < ck.row_end; py0 += ROWS * THREADS) {
        uint32_t n[ROWS], nmax = 0;
        const float *m[ROWS];
        float acc[ROWS];
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
            const uint32_t py = py0 + i * THREADS;
            const uint32_t r = py < ck.row_end ? ldg_tab(a.rowinfo + py) : 0u;       // n = 0: row absent
            n[i] = r & 0xffffu;
            m[i] = scratch + (r >> 16);
            acc[i] = 0.0f;
            nmax = n[i] > nmax ? n[i] : nmax;
        }
#if SBX_ROWSUM_PER_ROW
        (void)nmax;
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
            float sum = 0.0f;
            const float *mi = m[i];
#pragma unroll ROWSUM_UNROLL
            for (uint32_t s = 0; s < n[i]; ++s) sum = __fadd_rn(sum, mi[s]);
            acc[i] = sum;
        }
#else
        for (uint32_t s = 0; s < nmax; ++s) {
#pragma unroll
            for (int i = 0; i < ROWS; ++i)
                if (s < n[i]) acc[i] = __fadd_rn(acc[i], m[i][s]);
        }
#endif
#pragma unroll
        for (int i = 0; i < ROWS; ++i)
            if (n[i] != 0) colour_one_row_single(acc[i], n[i], py0 + i * THREADS, a, out);
    }
}

template <bool PAIR, int THREADS>
__device__ __forceinline__ void colour_rows(const float2 *scratch, const float *thr, const uint32_t *lut,
                                            const KernelArgs &a, const RowChunk ck, unsigned long long frame0, bool second_valid)
{
    uint32_t *outA = reinterpret_cast<uint32_t *>(a.rgba) + frame0 * a.rows;
    uint32_t *outB = outA + a.rows;
    for (uint32_t py0 = ck.row_begin + threadIdx.x; py0 < ck.row_end; py0 += 2 * THREADS) {
        const uint32_t py1 = py0 + THREADS;
        const bool has1 = py1 < ck.row_end;
        const uint32_t r0 = ldg_tab(a.rowinfo + py0), r1 = has1 ? ldg_tab(a.rowinfo + py1) : 0u;
        const uint32_t n0 = r0 & 0xffffu, n1 = r1 & 0xffffu;                 // n1 == 0 when the second row is absent
        const float2 *m0 = scratch + (r0 >> 16), *m1 = scratch + (r1 >> 16); // odd stride between rows
        float ax0 = 0.0f, ay0 = 0.0f, ax1 = 0.0f, ay1 = 0.0f;
#if SBX_ROWSUM_PER_ROW_CLASSIC
        for (uint32_t s = 0; s < n0; ++s) { const float2 v = m0[s]; ax0 = __fadd_rn(ax0, v.x); ay0 = __fadd_rn(ay0, v.y); }
        for (uint32_t s = 0; s < n1; ++s) { const float2 v = m1[s]; ax1 = __fadd_rn(ax1, v.x); ay1 = __fadd_rn(ay1, v.y); }
#else
        const uint32_t nmax = n0 > n1 ? n0 : n1;
        for (uint32_t s = 0; s < nmax; ++s) {
            if (s < n0) { const float2 v = m0[s]; ax0 = __fadd_rn(ax0, v.x); ay0 = __fadd_rn(ay0, v.y); }
            if (s < n1) { const float2 v = m1[s]; ax1 = __fadd_rn(ax1, v.x); ay1 = __fadd_rn(ay1, v.y); }
        }
#endif
        colour_one_row<PAIR>(ax0, ay0, n0, py0, thr, lut, a, outA, outB, second_valid);
        if (has1) colour_one_row<PAIR>(ax1, ay1, n1, py1, thr, lut, a, outA, outB, second_valid);
    }
}

// ---- log-frequency resample + row means on packed (L, R) pairs ------------------------------------------------------
// The same operations in the same order with the same roundings as interp_sample / colour_rows of stft_kernels.cuh --
// bit-identical results -- with both channels (or both frames of a mono pair) in the two lanes of one instruction.
// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (it honours .rn only for scalar f32), which would change the
// reference's rounding, so every product that feeds a sum is written fma(x, y, -0.0): one rounding, and nothing left
// to contract.  -0.0 comes from the kernel arguments (KernelArgs::neg_zero) so that it is not folded back into a multiply.
__device__ __forceinline__ pk pk_ld2(const float2 *p)
{
#ifdef SBX_EMU
    return pk_make(p->x, p->y);
#else
    return *reinterpret_cast<const pk *>(p);
#endif
}
__device__ __forceinline__ void pk_st2(float2 *p, pk v)
{
#ifdef SBX_EMU
    *p = make_float2(v.lo, v.hi);
#else
    *reinterpret_cast<pk *>(p) = v;
#endif
}
__device__ __forceinline__ pk pk_mul_exact(pk a, pk b, pk negzero) { return pk_fma(a, b, negzero); }

template <bool CLAMP>
__device__ __forceinline__ pk pk_interp_sample(const float2 *mags, const KernelArgs &a, int x1, float w, int k_last, pk negzero)
{
    if (a.interp == 0) {                                                     // cubic_interpolate, ifs.rs:88-105
        const pk m1 = pk_dup(w);
        const pk m2 = pk_mul(m1, m1), m3 = pk_mul(m1, m2);
        pk y0, y1, y2, y3;
        if constexpr (CLAMP) {
            y0 = pk_ld2(mags + (x1 > 0 ? x1 - 1 : 0) + 1);
            y1 = pk_ld2(mags + x1 + 1);
            y2 = pk_ld2(mags + min(x1 + 1, k_last) + 1);
            y3 = pk_ld2(mags + min(x1 + 2, k_last) + 1);
        } else {
            const float2 *m = mags + x1;
            y0 = pk_ld2(m); y1 = pk_ld2(m + 1); y2 = pk_ld2(m + 2); y3 = pk_ld2(m + 3);
        }
        const pk a0 = pk_add(pk_sub(pk_sub(y3, y2), y0), y1);
        const pk a1 = pk_sub(pk_sub(y0, y1), a0);
        const pk a2 = pk_sub(y2, y0);
        return pk_add(pk_add(pk_mul_exact(a0, m3, negzero), pk_mul_exact(a1, m2, negzero)), pk_add(pk_mul_exact(a2, m1, negzero), y1));
    } else {                                                                 // cosine_interpolate, ifs.rs:78-86
        const float omt = __fsub_rn(1.0f, w);
        const pk yl = pk_ld2(mags + x1 + 1), yh = pk_ld2(mags + min(x1 + 1, k_last) + 1);
        return pk_add(pk_mul_exact(yl, pk_dup(omt), negzero), pk_mul_exact(yh, pk_dup(w), negzero));
    }
}

// phase A: one lane per (row, sample), UNROLL table loads in flight per thread
template <int THREADS, int UNROLL, bool CLAMP>
__device__ __forceinline__ void pk_resample_samples_impl(const float2 *mags, float2 *scratch, const KernelArgs &a, const RowChunk ck)
{
    const int k_last = (int)a.k_out - 1;
    const pk negzero = pk_dup(a.neg_zero);
    const uint32_t count = ck.samp_end - ck.samp_begin;
    const uint2 *tab = a.samp + ck.samp_begin;
    for (uint32_t g0 = threadIdx.x; g0 < count; g0 += UNROLL * THREADS) {
        uint2 sp[UNROLL];
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {
            const uint32_t g = g0 + i * THREADS;
            sp[i] = g < count ? ldg_stream(tab + g) : make_uint2(SAMP_NO_STORE, 0u);
        }
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {
            if (sp[i].x < SAMP_NO_STORE)                                       // a real entry (not a filler, not past the end)
                pk_st2(scratch + (sp[i].x >> 17),
                         pk_interp_sample<CLAMP>(mags, a, (int)(sp[i].x & 0x1ffffu), __uint_as_float(sp[i].y), k_last, negzero));
        }
    }
}
template <int THREADS, int UNROLL>
__device__ __forceinline__ void pk_resample_samples(const float2 *mags, float2 *scratch, const KernelArgs &a, const RowChunk ck)
{
    if (a.taps_clamped) pk_resample_samples_impl<THREADS, UNROLL, true>(mags, scratch, a, ck);
    else pk_resample_samples_impl<THREADS, UNROLL, false>(mags, scratch, a, ck);
}

// phase B: one thread per row, ROWS rows in flight per thread: sequential sums (one FADD2 per sample), mean, colour
template <bool PAIR, int THREADS, int ROWS>
__device__ __forceinline__ void pk_colour_rows(const float2 *scratch, const KernelArgs &a, const RowChunk ck,
                                               unsigned long long frame0, bool second_valid)
{
    uint32_t *outA = reinterpret_cast<uint32_t *>(a.rgba) + frame0 * a.rows;
    uint32_t *outB = outA + a.rows;
    for (uint32_t py0 = ck.row_begin + threadIdx.x; py0 < ck.row_end; py0 += ROWS * THREADS) {
        uint32_t n[ROWS], nmax = 0;
        const float2 *m[ROWS];
        pk acc[ROWS];
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
            const uint32_t py = py0 + i * THREADS;
            const uint32_t r = py < ck.row_end ? ldg_tab(a.rowinfo + py) : 0u;       // n = 0: row absent
            n[i] = r & 0xffffu;
            m[i] = scratch + (r >> 16);
            acc[i] = pk_make(0.0f, 0.0f);
            nmax = n[i] > nmax ? n[i] : nmax;
        }
        // measured on a B200 (M frames/s, fused loop -> per-row loops): cfg1 67.4 -> 69.25, cfg1 geometry stereo 34.5 -> 35.7,
        // cfg3 34.65 -> 34.94; but cfg2 (N = 2048: 128-thread CTAs, eight rows per thread) 58.2 -> 57.1: per row from 256 threads up
        if constexpr (SBX_ROWSUM_PER_ROW != 0 && THREADS >= 256) {
#pragma unroll
            for (int i = 0; i < ROWS; ++i) {
                pk sum = pk_make(0.0f, 0.0f);
                const float2 *mi = m[i];
#pragma unroll ROWSUM_UNROLL
                for (uint32_t s = 0; s < n[i]; ++s) sum = pk_add(sum, pk_ld2(mi + s));
                acc[i] = sum;
            }
        } else {
            for (uint32_t s = 0; s < nmax; ++s) {
#pragma unroll
                for (int i = 0; i < ROWS; ++i)
                    if (s < n[i]) acc[i] = pk_add(acc[i], pk_ld2(m[i] + s));
            }
        }
#pragma unroll
        for (int i = 0; i < ROWS; ++i)
            if (n[i] != 0)
                colour_one_row<PAIR>(pk_lo(acc[i]), pk_hi(acc[i]), n[i], py0 + i * THREADS, a.thr, a.lut, a, outA, outB, second_valid);
    }
}

#ifndef SBX_EMU
template <int LOG2M, bool PAIR, int THREADS, int UNROLL, int ROWS>
__device__ __forceinline__ void pk_resample_colour(float2 *z, const KernelArgs &a, unsigned long long frame0, bool second_valid)
{
    float2 *scratch = z + (1 << (LOG2M - 1));                 // the upper half of the buffer is free again
    for (uint32_t c = 0; c < a.n_chunks; ++c) {
        RowChunk ck = a.chunk0;
        if (c != 0) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(a.chunks) + c);
            ck = RowChunk{raw.x, raw.y, raw.z, raw.w};
            __syncthreads();
        }
        pk_resample_samples<THREADS, UNROLL>(z, scratch, a, ck);
        __syncthreads();
        pk_colour_rows<PAIR, THREADS, ROWS>(scratch, a, ck, frame0, second_valid);
    }
}
#endif

// A row that averages more samples than the scratch area holds (only possible on the large path: few rows over a
// very long transform).  The row is resampled in parts of `cap` samples; between parts one thread continues
// the row's sequential sum, so the summation order is still the reference's.  Phase functions for the emulation:
// big_row_part_samples (all threads) / big_row_part_sum (thread 0), then big_row_finish (thread 0).
template <int THREADS, bool CLAMP>
__device__ __forceinline__ void big_row_part_samples_impl(const float2 *mags, float2 *scratch, const KernelArgs &a,
                                                          uint32_t samp_lo, uint32_t samp_hi)
{
    const int k_last = (int)a.k_out - 1;
    for (uint32_t g = samp_lo + threadIdx.x; g < samp_hi; g += THREADS) {
        const uint2 sp = ldg_tab(a.samp + g);
        scratch[g - samp_lo] = interp_sample<CLAMP>(mags, a, (int)(sp.x & 0x1ffffu), __uint_as_float(sp.y), k_last);
    }
}
template <int THREADS>
__device__ __forceinline__ void big_row_part_samples(const float2 *mags, float2 *scratch, const KernelArgs &a,
                                                     uint32_t samp_lo, uint32_t samp_hi)
{
    if (a.taps_clamped) big_row_part_samples_impl<THREADS, true>(mags, scratch, a, samp_lo, samp_hi);
    else big_row_part_samples_impl<THREADS, false>(mags, scratch, a, samp_lo, samp_hi);
}
__device__ __forceinline__ void big_row_part_sum(const float2 *scratch, uint32_t count, float2 &acc)
{
    if (threadIdx.x != 0) return;
    for (uint32_t s = 0; s < count; ++s) {
        acc.x = __fadd_rn(acc.x, scratch[s].x);
        acc.y = __fadd_rn(acc.y, scratch[s].y);
    }
}
template <bool PAIR>
__device__ __forceinline__ void big_row_finish(float2 acc, uint32_t n, uint32_t py, const KernelArgs &a,
                                               unsigned long long frame0, bool second_valid)
{
    if (threadIdx.x != 0) return;
    uint32_t *outA = reinterpret_cast<uint32_t *>(a.rgba) + frame0 * a.rows;
    colour_one_row<PAIR>(acc.x, acc.y, n, py, a.thr, a.lut, a, outA, outA + a.rows, second_valid);
}

// ----------------------------------------------------------------------------------------------
// kernels
// ----------------------------------------------------------------------------------------------
#ifndef SBX_EMU
template <int LOG2M, bool PAIR, int THREADS>
__device__ __forceinline__ void resample_colour(float2 *z, const KernelArgs &a, unsigned long long frame0, bool second_valid)
{
    const float *thr = a.thr;
    const uint32_t *lut = a.lut;
    float2 *scratch = z + (1 << (LOG2M - 1));                 // the upper half of the transform buffer is free
    for (uint32_t c = 0; c < a.n_chunks; ++c) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(a.chunks) + c);
        const RowChunk ck = {raw.x, raw.y, raw.z, raw.w};
        if (c != 0) __syncthreads();
        resample_samples<THREADS>(z, scratch, a, ck);
        __syncthreads();
        colour_rows<PAIR, THREADS>(scratch, thr, lut, a, ck, frame0, second_valid);
    }
}

template <int LOG2M, int THREADS, int UNROLL = 4, int ROWS = 2>
__device__ __forceinline__ void resample_colour_single(float2 *z, const KernelArgs &a, unsigned long long frame)
{
    const float *mags = reinterpret_cast<const float *>(z);   // M floats: the lower half of the buffer
    float *scratch = reinterpret_cast<float *>(z) + (1 << LOG2M);
    for (uint32_t c = 0; c < a.n_chunks; ++c) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(a.chunks) + c);
        const RowChunk ck = {raw.x, raw.y, raw.z, raw.w};
        if (c != 0) __syncthreads();
        resample_samples<THREADS, float, UNROLL>(mags, scratch, a, ck);
        __syncthreads();
        colour_rows_single<THREADS, ROWS>(scratch, a, ck, frame);
    }
}

// everything after pass 0: remaining passes, L/R split, optional magnitude store, resample + colour
template <class G>
__device__ __forceinline__ void frame_tail(float2 *z, const KernelArgs &a, unsigned long long frame0, bool second_valid)
{
    if constexpr (LastSplit<G>::ENABLED) {
        RunPasses<G, 1, G::NPASS - 1>::run(z, a);
        SplitRegs<G> regs;
        last_pass_keep_own<G>(z, regs);
        __syncthreads();
        split_with_own<G>(z, a, regs);
        __syncthreads();
        split_write<G>(z, regs);
    } else {
        RunPasses<G, 1>::run(z, a);
        SplitRegs<G> regs;
        split_read<G>(z, a, regs);
        __syncthreads();
        split_write<G>(z, regs);
    }
    __syncthreads();
    if constexpr (G::REAL) {
        static_assert(LastSplit<G>::ENABLED, "the real-sequence split lives in the fused last pass");
        if (a.mags_out != nullptr) {
            float2 *out = reinterpret_cast<float2 *>(a.mags_out) + frame0 * a.k_out;
            const float *zf = reinterpret_cast<const float *>(z);
            for (uint32_t j = threadIdx.x; j < a.k_out; j += G::THREADS) out[j] = make_float2(zf[j + 1], zf[j + 1]);
        }
        if (a.rgba != nullptr) resample_colour_single<G::LOG2M, G::THREADS>(z, a, frame0);
    } else {
        if (a.mags_out != nullptr) {
            float2 *out = reinterpret_cast<float2 *>(a.mags_out) + frame0 * a.k_out;
            for (uint32_t j = threadIdx.x; j < a.k_out; j += G::THREADS) {
                const float2 m = z[j + 1];
                if constexpr (G::CH == 1) {
                    out[j] = make_float2(m.x, m.x);
                    if (second_valid) out[a.k_out + j] = make_float2(m.y, m.y);
                } else {
                    out[j] = m;
                }
            }
        }
        if (a.rgba != nullptr) {
#if SBX_PK_RESAMPLE
            pk_resample_colour<G::LOG2M, G::CH == 1, G::THREADS, 4, 2>(z, a, frame0, second_valid);
#else
            resample_colour<G::LOG2M, G::CH == 1, G::THREADS>(z, a, frame0, second_valid);
#endif
        }
    }
}

template <class G>
__global__ void __launch_bounds__(G::THREADS, G::MIN_CTAS) fused_frames_kernel(const KernelArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *z = reinterpret_cast<float2 *>(smem_raw);

    unsigned long long unit = blockIdx.x, seq = 0;
    if (a.units_per_seq != 0) {                                   // planar sequences: CTA -> (sequence, unit within it)
        seq = unit / a.units_per_seq;
        unit -= seq * a.units_per_seq;
    }
    const unsigned long long frame0 = unit * G::FRAMES_PER_CTA;   // within the sequence
    const bool second_valid = G::CH == 1 && frame0 + 1 < a.n_frames;

    pass_first<G, false>(z, a, a.pcm + seq * a.seq_stride + frame0 * (unsigned long long)a.hop * G::CH, second_valid);
    __syncthreads();
    frame_tail<G>(z, a, seq * a.n_frames + frame0, second_valid);
    if (a.done_flag != nullptr) {
        // the barrier orders every thread's output stores before thread 0's fence; the fence is cumulative, so a host
        // thread that observes the flag also observes the column
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned int v = *a.done_count + 1u;
            *a.done_count = v;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned int *>(a.done_flag) = v;
        }
    }
}

// ---- persistent variant: grid = SMs x resident CTAs, PCM staged through shared memory by bulk async copies ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// cp.async.bulk (the TMA engine's 1-D form): global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    uint32_t done = 0, spins = 0;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 24)) __trap();      // a lost copy must fault, never hang the GPU
    } while (!done);
}

// Unit u = one frame (stereo) or one frame pair (mono).  The sample span of unit u+grid is copied into the
// stage buffer by one elected thread as soon as pass 0 of unit u has consumed it, so the copy overlaps
// passes 1.., the split and the resample of unit u.  Requires 16-byte aligned spans (checked by the host).
template <class G>
__global__ void __launch_bounds__(G::THREADS, G::MIN_CTAS) fused_frames_persistent_kernel(const KernelArgs a, uint32_t n_units)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *z = reinterpret_cast<float2 *>(smem_raw);
    float *stage = reinterpret_cast<float *>(smem_raw + (size_t)G::M * 8);
    __shared__ __align__(8) unsigned long long bar;

    constexpr uint32_t FPC = G::FRAMES_PER_CTA;
    const uint32_t full_bytes = (G::N + (FPC - 1) * a.hop) * G::CH * 4u;      // both frames of a pair
    const uint32_t one_bytes = G::N * G::CH * 4u;                               // a trailing half-filled pair
    auto unit_bytes = [&](uint32_t u) { return (unsigned long long)u * FPC + FPC <= a.n_frames ? full_bytes : one_bytes; };
    auto unit_src = [&](uint32_t u) { return a.pcm + (unsigned long long)u * FPC * a.hop * G::CH; };

    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    uint32_t u = blockIdx.x;
    if (threadIdx.x == 0 && u < n_units) {
        mbar_expect_tx(&bar, unit_bytes(u));
        bulk_copy_g2s(stage, unit_src(u), unit_bytes(u), &bar);
    }
    uint32_t parity = 0;
    for (; u < n_units; u += gridDim.x) {
        const unsigned long long frame0 = (unsigned long long)u * FPC;
        const bool second_valid = G::CH == 1 && frame0 + 1 < a.n_frames;
        mbar_wait(&bar, parity);
        parity ^= 1u;
        pass_first<G, true>(z, a, stage, second_valid);
        __syncthreads();                                    // stage consumed, pass 0 stored
        const uint32_t next = u + gridDim.x;
        if (threadIdx.x == 0 && next < n_units) {
            mbar_expect_tx(&bar, unit_bytes(next));
            bulk_copy_g2s(stage, unit_src(next), unit_bytes(next), &bar);
        }
        frame_tail<G>(z, a, frame0, second_valid);
        __syncthreads();                                    // the next unit's pass 0 overwrites z
    }
}

// magnitudes (frames x K x (L,R), device) -> RGBA columns: the resample/colour stage alone
template <int LOG2M>
__global__ void __launch_bounds__(256) colourise_kernel(const KernelArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *z = reinterpret_cast<float2 *>(smem_raw);                            // M float2 (mags + scratch)
    const unsigned long long frame = blockIdx.x;
    const float2 *in = reinterpret_cast<const float2 *>(a.mags_in) + frame * a.k_out;
    for (uint32_t j = threadIdx.x; j < a.k_out; j += 256) z[j + 1] = __ldg(in + j);
    __syncthreads();
    resample_colour<LOG2M, false, 256>(z, a, frame, false);
}
#endif  // SBX_EMU

}  // namespace sbx
