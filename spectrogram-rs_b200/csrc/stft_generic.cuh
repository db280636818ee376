// stft_generic.cuh -- windows that are not a power of two (the live app's N = (0.05 s * sr) as usize:
// 2400 @ 48 kHz, 2205 @ 44.1 kHz; src/widgets/gpu_spectrogram.rs:323, src/fourier/fft.rs:19-21).
//
// Same structure as the fused kernel -- P premultiplied sub-transforms of length N, in-place DIF,
// digit-reversed output, conjugate-symmetry split, tabulated resample, threshold dB step -- with the
// radix list chosen at run time from {16, 15, 12, 10, 9, 8, 7, 6, 5, 4, 3, 2}: the fewest passes, the radix with the
// smallest gcd(R, 16) last (N = 2400 = 16*10*15, 2205 = 15*7*7*3).
// This is the low-rate live path (819 frames/s in the reference), so it trades the fused kernel's
// swizzles and compile-time index math for generality: bin -> position comes from a host table, twiddles
// from per-pass planes (contiguous in the butterfly index), divisions by host-computed reciprocals.
// Measured on a B200, N = 2400 stereo: 25.4 M frames/s in batch, 24 us per live frame.  History: 12.0 M / 41 us with
// strided loads from the full W_M table and run-time divisions; 18.5 M / 26.6 us with five prime-radix passes
// (16*2*3*5*5); 20.6 M with literal-constant 3- and 5-point DFTs; 23.1 M with composite radices 6 / 9 / 10 / 12 / 15
// in registers (three passes); 24.0 M with the odd radix last (the stride-1 pass conflict-free); 25.4 M with pass 0
// reading the PCM itself (batches only).  Radices of 20 / 21
// in registers were measured in round 1 and rejected: 128 registers halve the occupancy (12.8 M frames/s).
// One CTA per frame (stereo) or frame pair (mono); the transform (M float2) and the magnitudes (M/2 float2) live
// in shared memory: M <= 19370.  Windows with a prime factor > 7 take the Bluestein form further down.
#pragma once
#include "stft_kernels.cuh"

namespace sbx {

#ifndef SBX_GENERIC_THREADS
#define SBX_GENERIC_THREADS 512     // measured on a B200, N = 2400: live p50 43.2 / 40.9 / 40.9 us and batch 12.0 / 12.1 / 8.8 M frames/s at 256 / 512 / 1024
#endif
constexpr int GENERIC_THREADS = SBX_GENERIC_THREADS;
#ifndef SBX_GENERIC_FUSED_FIRST_PASS
#define SBX_GENERIC_FUSED_FIRST_PASS 1   // pass 0 reads the PCM itself (generic_first_pass) instead of a staged copy
#endif
#ifndef SBX_GENERIC_PK_EPILOGUE
#define SBX_GENERIC_PK_EPILOGUE 0   // 1: resample / row sums on packed (L, R) pairs as in the q2 kernels
#endif
constexpr int GENERIC_MAX_PASSES = 12;
constexpr uint32_t BLUESTEIN_MAX_LEN = 16384;   // Bluestein convolution length L (128 KB of float2 in shared memory)

struct GenericArgs {
    KernelArgs k;                 // window, rowmap, chunks, samp, thr, lut, rgba, mags_out, pcm ... (wtw/twp unused)
    const float2 *twm;            // [M] W_M^t, f64-computed, rounded once (premultiply, odd-radix DFT constants)
    const float2 *twg;            // per pass p (all but the last): (R_p - 1) planes of S_p entries, plane q-1 = W_(S_p R_p)^(lo q)
    const uint32_t *pos_tab;      // [M] position of bin kk in the transform buffer
    uint32_t n, pad, m, npass;
    uint32_t radix[GENERIC_MAX_PASSES];
    uint32_t stride[GENERIC_MAX_PASSES];   // butterfly stride of each pass: N / (R_0 ... R_p)
    uint32_t per_sub[GENERIC_MAX_PASSES];  // butterflies per sub-transform: N / R_p
    uint32_t tw_off[GENERIC_MAX_PASSES];   // first entry of pass p in twg
    uint32_t inv_stride[GENERIC_MAX_PASSES];   // floor(2^32 / S_p) + 1: u / S_p == __umulhi(u, inv) for u < 2^16, S_p < 2^15
    // Bluestein (window with a prime factor > 7): n = m = the power-of-two convolution length L, pad = 1, and
    const float2 *bl_a;           // [N]  window[n] * conj(b[n]),  b[j] = exp(+i pi j^2 / M)
    const float2 *bl_bf;          // [L]  FFT_L(h) / L at the POSITION of each bin (digit-reversed order), h = b wrapped mod L
    const float2 *bl_c;           // [M]  conj(b[k])
    uint32_t bl_n, bl_m;          // the window N and the padded transform length M = P N
};

// which of the (at most four) equal runs of length `len` holds u, without a division: u in [0, pad * len)
__device__ __forceinline__ uint32_t run_of(uint32_t u, uint32_t len, uint32_t &within)
{
    const uint32_t r = (u >= len ? 1u : 0u) + (u >= 2u * len ? 1u : 0u) + (u >= 3u * len ? 1u : 0u);
    within = u - r * len;
    return r;
}
__device__ __forceinline__ uint32_t mulhi_u32(uint32_t a, uint32_t b)
{
#ifdef SBX_EMU
    return (uint32_t)(((unsigned long long)a * b) >> 32);
#else
    return __umulhi(a, b);
#endif
}

// W_R^j = exp(-2 pi i j / R) for the composite radices, as literals: after unrolling every index is a compile-time
// constant and the values become instruction immediates
__device__ __forceinline__ float2 wconst6(int j)
{
    constexpr float re[6] = {1.0f, 0.5f, -0.5f, -1.0f, -0.5f, 0.5f};
    constexpr float im[6] = {0.0f, -0.866025404f, -0.866025404f, 0.0f, 0.866025404f, 0.866025404f};
    return make_float2(re[j], im[j]);
}
__device__ __forceinline__ float2 wconst9(int j)
{
    constexpr float re[9] = {1.0f, 0.766044443f, 0.173648178f, -0.5f, -0.939692621f, -0.939692621f, -0.5f, 0.173648178f, 0.766044443f};
    constexpr float im[9] = {0.0f, -0.64278761f, -0.984807753f, -0.866025404f, -0.342020143f, 0.342020143f, 0.866025404f, 0.984807753f, 0.64278761f};
    return make_float2(re[j], im[j]);
}
__device__ __forceinline__ float2 wconst10(int j)
{
    constexpr float re[10] = {1.0f, 0.809016994f, 0.309016994f, -0.309016994f, -0.809016994f, -1.0f, -0.809016994f, -0.309016994f, 0.309016994f, 0.809016994f};
    constexpr float im[10] = {0.0f, -0.587785252f, -0.951056516f, -0.951056516f, -0.587785252f, 0.0f, 0.587785252f, 0.951056516f, 0.951056516f, 0.587785252f};
    return make_float2(re[j], im[j]);
}
__device__ __forceinline__ float2 wconst12(int j)
{
    constexpr float re[12] = {1.0f, 0.866025404f, 0.5f, 0.0f, -0.5f, -0.866025404f, -1.0f, -0.866025404f, -0.5f, 0.0f, 0.5f, 0.866025404f};
    constexpr float im[12] = {0.0f, -0.5f, -0.866025404f, -1.0f, -0.866025404f, -0.5f, 0.0f, 0.5f, 0.866025404f, 1.0f, 0.866025404f, 0.5f};
    return make_float2(re[j], im[j]);
}
__device__ __forceinline__ float2 wconst15(int j)
{
    constexpr float re[15] = {1.0f, 0.913545458f, 0.669130606f, 0.309016994f, -0.104528463f, -0.5f, -0.809016994f, -0.978147601f,
                              -0.978147601f, -0.809016994f, -0.5f, -0.104528463f, 0.309016994f, 0.669130606f, 0.913545458f};
    constexpr float im[15] = {0.0f, -0.406736643f, -0.743144825f, -0.951056516f, -0.994521895f, -0.866025404f, -0.587785252f, -0.207911691f,
                              0.207911691f, 0.587785252f, 0.866025404f, 0.994521895f, 0.951056516f, 0.743144825f, 0.406736643f};
    return make_float2(re[j], im[j]);
}
template <int R> __device__ __forceinline__ float2 wconst(int j)
{
    if constexpr (R == 6) return wconst6(j);
    else if constexpr (R == 9) return wconst9(j);
    else if constexpr (R == 10) return wconst10(j);
    else if constexpr (R == 12) return wconst12(j);
    else return wconst15(j);
}

// prime-length DFTs with literal constants, natural order in and out
template <int R> struct PrimDft;
template <> struct PrimDft<2> { static __device__ __forceinline__ void run(float2 (&v)[2]) { Butterfly<2>::run(v); } };
template <> struct PrimDft<4> { static __device__ __forceinline__ void run(float2 (&v)[4]) { Butterfly<4>::run(v); } };
template <> struct PrimDft<3> {
    static __device__ __forceinline__ void run(float2 (&v)[3])
    {
        // W_3 = -1/2 - i sqrt(3)/2:  X1 = a - s/2 - i (sqrt3/2) d,  X2 = a - s/2 + i (sqrt3/2) d,  s = b + c, d = b - c
        const float2 s = cadd(v[1], v[2]), d = csub(v[1], v[2]);
        const float2 m = make_float2(v[0].x - 0.5f * s.x, v[0].y - 0.5f * s.y);
        const float2 e = make_float2(0.866025404f * d.y, -0.866025404f * d.x);      // -i (sqrt3/2) d
        v[0] = cadd(v[0], s);
        v[1] = cadd(m, e);
        v[2] = csub(m, e);
    }
};
template <> struct PrimDft<5> {
    static __device__ __forceinline__ void run(float2 (&v)[5])
    {
        constexpr float c1 = 0.309016994f, c2 = -0.809016994f, s1 = 0.951056516f, s2 = 0.587785252f;
        const float2 t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]), t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
        const float2 a1 = make_float2(v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y);
        const float2 a2 = make_float2(v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y);
        const float2 b1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
        const float2 b2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
        v[0] = cadd(v[0], cadd(t1, t2));
        v[1] = make_float2(a1.x + b1.y, a1.y - b1.x);                                // a1 - i b1
        v[4] = make_float2(a1.x - b1.y, a1.y + b1.x);                                // a1 + i b1
        v[2] = make_float2(a2.x + b2.y, a2.y - b2.x);
        v[3] = make_float2(a2.x - b2.y, a2.y + b2.x);
    }
};

// R = R1 R2 by one Cooley-Tukey step in registers: n = R2 n1 + n2, k = k1 + R1 k2,
// W_R^(n k) = W_R1^(n1 k1) W_R^(n2 k1) W_R2^(n2 k2)
template <int R1, int R2> struct CompositeDft {
    static constexpr int R = R1 * R2;
    static __device__ __forceinline__ void run(float2 (&v)[R])
    {
        float2 t[R1][R2];
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) {
            float2 a[R1];
#pragma unroll
            for (int n1 = 0; n1 < R1; ++n1) a[n1] = v[R2 * n1 + n2];
            PrimDft<R1>::run(a);
#pragma unroll
            for (int k1 = 0; k1 < R1; ++k1) t[k1][n2] = a[k1];
        }
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
            float2 b[R2];
#pragma unroll
            for (int n2 = 0; n2 < R2; ++n2) b[n2] = (k1 * n2) % R == 0 ? t[k1][n2] : cmul(t[k1][n2], wconst<R>((k1 * n2) % R));
            PrimDft<R2>::run(b);
#pragma unroll
            for (int k2 = 0; k2 < R2; ++k2) v[k1 + R1 * k2] = b[k2];
        }
    }
};

// R-point DFT, natural order in and out.  Powers of two reuse the fused kernel's butterflies; 3 and 5 use literal
// constants; 6, 9, 10, 12 and 15 are one Cooley-Tukey step of those in registers (fewer barrier-separated passes:
// 2400 = 16 x 15 x 10 instead of 16 x 2 x 3 x 5 x 5); 7 is evaluated directly from the table (w[j] = W_7^j).
template <int R> struct SmallDft {
    static __device__ __forceinline__ void run(float2 (&v)[R], const float2 *twm, uint32_t m)
    {
        if constexpr (R == 2 || R == 4 || R == 8 || R == 16) {
            (void)twm; (void)m;
            Butterfly<R>::run(v);
        } else if constexpr (R == 3 || R == 5) {
            (void)twm; (void)m;
            PrimDft<R>::run(v);
        } else if constexpr (R == 6) {
            (void)twm; (void)m;
            CompositeDft<2, 3>::run(v);
        } else if constexpr (R == 9) {
            (void)twm; (void)m;
            CompositeDft<3, 3>::run(v);
        } else if constexpr (R == 10) {
            (void)twm; (void)m;
            CompositeDft<2, 5>::run(v);
        } else if constexpr (R == 12) {
            (void)twm; (void)m;
            CompositeDft<4, 3>::run(v);
        } else if constexpr (R == 15) {
            (void)twm; (void)m;
            CompositeDft<3, 5>::run(v);
        } else {
            float2 w[R];
#pragma unroll
            for (int j = 0; j < R; ++j) w[j] = __ldg(twm + (size_t)j * (m / R));
            float2 y[R];
#pragma unroll
            for (int q = 0; q < R; ++q) {
                float2 acc = v[0];
#pragma unroll
                for (int r = 1; r < R; ++r) {
                    const float2 t = cmul(v[r], w[(r * q) % R]);
                    acc.x += t.x;  acc.y += t.y;
                }
                y[q] = acc;
            }
#pragma unroll
            for (int q = 0; q < R; ++q) v[q] = y[q];
        }
    }
};

// phase 0: PCM -> window -> W_M^(n q0) premultiply -> shared memory, natural order per sub-transform
template <int CH>
__device__ __forceinline__ void generic_load(float2 *z, const GenericArgs &g, unsigned long long frame0, bool second_valid)
{
    const KernelArgs &a = g.k;
    const float *pa = a.pcm + frame0 * (unsigned long long)a.hop * CH;
    const float *pb = pa + a.hop;
    // every sample is read ONCE and feeds all P sub-transforms (the live path reads it from pinned host memory)
    for (uint32_t n = threadIdx.x; n < g.n; n += GENERIC_THREADS) {
        float2 s;
        if constexpr (CH == 2) {
            s = __ldg(reinterpret_cast<const float2 *>(pa) + n);
        } else {
            s.x = __ldg(pa + n);
            s.y = second_valid ? __ldg(pb + n) : 0.0f;
        }
        const float wn = __ldg(a.window + n);
        s.x *= wn;  s.y *= wn;                                            // fft.rs:57-63
        z[n] = s;
        for (uint32_t q0 = 1; q0 < g.pad; ++q0)
            z[(size_t)q0 * g.n + n] = cmul(s, __ldg(g.twm + n * q0));     // n q0 < N P = M
    }
}

// one in-place DIF pass of radix R at stride S over all P sub-transforms
template <int R>
__device__ __forceinline__ void generic_pass(float2 *z, const GenericArgs &g, uint32_t p, bool last)
{
    const uint32_t S = g.stride[p], per_sub = g.per_sub[p], count = g.pad * per_sub;
    const float2 *tw = g.twg + g.tw_off[p];                               // plane q-1, entry lo: W_(S R)^(lo q)
    const uint32_t inv = g.inv_stride[p];
    for (uint32_t u = threadIdx.x; u < count; u += GENERIC_THREADS) {
        uint32_t uu;
        const uint32_t sub = run_of(u, per_sub, uu);
        const uint32_t hi = S == 1 ? uu : mulhi_u32(uu, inv), lo = uu - hi * S;
        float2 *b = z + (size_t)sub * g.n + (size_t)hi * S * R + lo;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = b[(size_t)r * S];
        SmallDft<R>::run(v, g.twm, g.m);
        b[0] = v[0];
#pragma unroll
        for (int q = 1; q < R; ++q)
            b[(size_t)q * S] = last ? v[q] : cmul(v[q], __ldg(tw + (size_t)(q - 1) * S + lo));
    }
}

// pass 0 fused with the load: the first radix pass has stride S = N / R_0 over the WHOLE sub-transform, so butterfly lo
// takes the samples n = lo + r S straight from the PCM (window, and W_M^(n q0) for the sub-transforms q0 > 0, applied
// on the way) and nothing is staged in shared memory first -- one trip through the buffer and one barrier fewer.  The
// P sub-transforms of a butterfly re-read its samples (L1 hits), each as a work item of its own: N / R_0 butterflies
// alone would leave most of the CTA idle (150 at N = 2400).
template <int R, int CH>
__device__ __forceinline__ void generic_first_pass(float2 *z, const GenericArgs &g, unsigned long long frame0, bool second_valid)
{
    const KernelArgs &a = g.k;
    const float *pa = a.pcm + frame0 * (unsigned long long)a.hop * CH;
    const float *pb = pa + a.hop;
    const uint32_t S = g.stride[0];                                       // = N / R
    const bool last = g.npass == 1;
    const float2 *tw = g.twg + g.tw_off[0];
    for (uint32_t u = threadIdx.x; u < g.pad * S; u += GENERIC_THREADS) {   // one butterfly of one sub-transform per item
        uint32_t lo;
        const uint32_t q0 = run_of(u, S, lo);
        {
            float2 v[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t n = lo + (uint32_t)r * S;
                float2 s;
                if constexpr (CH == 2) {
                    s = __ldg(reinterpret_cast<const float2 *>(pa) + n);
                } else {
                    s.x = __ldg(pa + n);
                    s.y = second_valid ? __ldg(pb + n) : 0.0f;
                }
                const float wn = __ldg(a.window + n);
                s.x *= wn;  s.y *= wn;                                    // fft.rs:57-63
                v[r] = q0 == 0 ? s : cmul(s, __ldg(g.twm + n * q0));      // n q0 < N P = M
            }
            SmallDft<R>::run(v, g.twm, g.m);
            float2 *b = z + (size_t)q0 * g.n + lo;
            b[0] = v[0];
#pragma unroll
            for (int q = 1; q < R; ++q)
                b[(size_t)q * S] = last ? v[q] : cmul(v[q], __ldg(tw + (size_t)(q - 1) * S + lo));
        }
    }
}

template <int CH>
__device__ __forceinline__ void generic_first_pass_dispatch(float2 *z, const GenericArgs &g, unsigned long long frame0, bool second_valid)
{
    switch (g.radix[0]) {
        case 2: generic_first_pass<2, CH>(z, g, frame0, second_valid); break;
        case 3: generic_first_pass<3, CH>(z, g, frame0, second_valid); break;
        case 4: generic_first_pass<4, CH>(z, g, frame0, second_valid); break;
        case 5: generic_first_pass<5, CH>(z, g, frame0, second_valid); break;
        case 6: generic_first_pass<6, CH>(z, g, frame0, second_valid); break;
        case 7: generic_first_pass<7, CH>(z, g, frame0, second_valid); break;
        case 8: generic_first_pass<8, CH>(z, g, frame0, second_valid); break;
        case 9: generic_first_pass<9, CH>(z, g, frame0, second_valid); break;
        case 10: generic_first_pass<10, CH>(z, g, frame0, second_valid); break;
        case 12: generic_first_pass<12, CH>(z, g, frame0, second_valid); break;
        case 15: generic_first_pass<15, CH>(z, g, frame0, second_valid); break;
        case 16: generic_first_pass<16, CH>(z, g, frame0, second_valid); break;
        default: break;
    }
}

__device__ __forceinline__ void generic_pass_dispatch(float2 *z, const GenericArgs &g, uint32_t p)
{
    const bool last = p + 1 == g.npass;
    switch (g.radix[p]) {
        case 2: generic_pass<2>(z, g, p, last); break;
        case 3: generic_pass<3>(z, g, p, last); break;
        case 4: generic_pass<4>(z, g, p, last); break;
        case 5: generic_pass<5>(z, g, p, last); break;
        case 6: generic_pass<6>(z, g, p, last); break;
        case 7: generic_pass<7>(z, g, p, last); break;
        case 8: generic_pass<8>(z, g, p, last); break;
        case 9: generic_pass<9>(z, g, p, last); break;
        case 10: generic_pass<10>(z, g, p, last); break;
        case 12: generic_pass<12>(z, g, p, last); break;
        case 15: generic_pass<15>(z, g, p, last); break;
        case 16: generic_pass<16>(z, g, p, last); break;
        default: break;
    }
}

// ---- Bluestein (chirp-z) for windows the radix list cannot factor: N = 1102 = 2 * 19 * 29 (22.05 kHz), 551, primes ----
// W_M^(n k) = conj(b[n]) conj(b[k]) b[k - n] with b[j] = exp(+i pi j^2 / M), so the M-point DFT of the N windowed samples
// is X[k] = conj(b[k]) * sum_n (z[n] conj(b[n])) b[k - n]: a convolution, done as a circular one of power-of-two
// length L >= N + M - 1 with the passes above: forward DIF (natural order in, digit-reversed out), a pointwise product
// with the transformed chirp stored in that same digit-reversed order, and the inverse as the exact reverse of the
// forward passes (decimation in time: digit-reversed in, natural order out) -- no reordering pass at all.

// one in-place inverse pass: undoes generic_pass<R>(p): conjugate twiddles on the way in, then the inverse R-point DFT
template <int R>
__device__ __forceinline__ void generic_pass_inverse(float2 *z, const GenericArgs &g, uint32_t p, bool last)
{
    const uint32_t S = g.stride[p], per_sub = g.per_sub[p];
    const float2 *tw = g.twg + g.tw_off[p];
    const uint32_t inv = g.inv_stride[p];
    for (uint32_t u = threadIdx.x; u < per_sub; u += GENERIC_THREADS) {
        const uint32_t hi = S == 1 ? u : mulhi_u32(u, inv), lo = u - hi * S;
        float2 *b = z + (size_t)hi * S * R + lo;
        float2 v[R];
        v[0] = b[0];
        v[0].y = -v[0].y;
#pragma unroll
        for (int q = 1; q < R; ++q) {
            float2 x = b[(size_t)q * S];
            if (!last) {
                const float2 w = __ldg(tw + (size_t)(q - 1) * S + lo);
                x = cmul(x, make_float2(w.x, -w.y));
            }
            v[q] = make_float2(x.x, -x.y);                   // conj in, forward butterfly, conj out = inverse DFT (unscaled)
        }
        Butterfly<R>::run(v);
#pragma unroll
        for (int r = 0; r < R; ++r) b[(size_t)r * S] = make_float2(v[r].x, -v[r].y);
    }
}

__device__ __forceinline__ void generic_pass_inverse_dispatch(float2 *z, const GenericArgs &g, uint32_t p)
{
    const bool last = p + 1 == g.npass;
    switch (g.radix[p]) {                                    // L is a power of two
        case 2: generic_pass_inverse<2>(z, g, p, last); break;
        case 4: generic_pass_inverse<4>(z, g, p, last); break;
        case 8: generic_pass_inverse<8>(z, g, p, last); break;
        case 16: generic_pass_inverse<16>(z, g, p, last); break;
        default: break;
    }
}

// phase 0: PCM -> window * conj(chirp) -> shared memory, zero beyond N
template <int CH>
__device__ __forceinline__ void bluestein_load(float2 *z, const GenericArgs &g, unsigned long long frame0, bool second_valid)
{
    const KernelArgs &a = g.k;
    const float *pa = a.pcm + frame0 * (unsigned long long)a.hop * CH;
    const float *pb = pa + a.hop;
    for (uint32_t n = threadIdx.x; n < g.m; n += GENERIC_THREADS) {
        float2 v = make_float2(0.0f, 0.0f);
        if (n < g.bl_n) {
            float2 s;
            if constexpr (CH == 2) {
                s = __ldg(reinterpret_cast<const float2 *>(pa) + n);
            } else {
                s.x = __ldg(pa + n);
                s.y = second_valid ? __ldg(pb + n) : 0.0f;
            }
            v = cmul(s, __ldg(g.bl_a + n));                  // fft.rs:57-63 (window) folded into the chirp factor
        }
        z[n] = v;
    }
}

// pointwise product with the transformed chirp (both in position order)
__device__ __forceinline__ void bluestein_multiply(float2 *z, const GenericArgs &g)
{
    for (uint32_t i = threadIdx.x; i < g.m; i += GENERIC_THREADS) z[i] = cmul(z[i], __ldg(g.bl_bf + i));
}

// X[k] = conj(b[k]) y[k], then the L/R split of fft.rs:81-95 into the magnitude array indexed by bin kk
__device__ __forceinline__ void bluestein_split(const float2 *z, float2 *mags, const GenericArgs &g)
{
    for (uint32_t kk = 1 + threadIdx.x; kk < g.bl_m / 2; kk += GENERIC_THREADS) {
        const float2 x1 = cmul(z[kk], __ldg(g.bl_c + kk));
        const float2 x2 = cmul(z[g.bl_m - kk], __ldg(g.bl_c + (g.bl_m - kk)));
        mags[kk] = split_pair(x1, x2, g.k.scale);
    }
}

// L/R split (fft.rs:81-95) into a separate magnitude array indexed by bin kk
__device__ __forceinline__ void generic_split(const float2 *z, float2 *mags, const GenericArgs &g)
{
    for (uint32_t kk = 1 + threadIdx.x; kk < g.m / 2; kk += GENERIC_THREADS)
        mags[kk] = split_pair(z[__ldg(g.pos_tab + kk)], z[__ldg(g.pos_tab + (g.m - kk))], g.k.scale);
}

#ifndef SBX_EMU
template <int CH, bool BLUESTEIN = false>
__global__ void __launch_bounds__(GENERIC_THREADS, 2) generic_frames_kernel(const GenericArgs g)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *z = reinterpret_cast<float2 *>(smem_raw);                    // [M] ([L] for Bluestein); reused as resample scratch
    float2 *mags = z + g.m;                                               // [M/2]
    const KernelArgs &a = g.k;
    const unsigned long long frame0 = (unsigned long long)blockIdx.x * (CH == 1 ? 2 : 1);
    const bool second_valid = CH == 1 && frame0 + 1 < a.n_frames;
    // A live single frame (done_flag set: the zero-copy graph) reads its window from PINNED HOST memory: every sample
    // is fetched once and staged (generic_load); batches read device memory and fuse the load into pass 0.
    // Measured on a B200: batch N = 2400 24.1 -> 25.4, N = 2205 19.4 -> 20.2, N = 4800 11.0 -> 11.7 M frames/s; live p50
    // 24.4 us staged vs 25.1 us fused.
    const bool fused_first = !BLUESTEIN && SBX_GENERIC_FUSED_FIRST_PASS && a.done_flag == nullptr;
    if constexpr (BLUESTEIN) {
        bluestein_load<CH>(z, g, frame0, second_valid);
    } else {
        if (fused_first) generic_first_pass_dispatch<CH>(z, g, frame0, second_valid);
        else generic_load<CH>(z, g, frame0, second_valid);
    }
    for (uint32_t p = fused_first ? 1u : 0u; p < g.npass; ++p) {
        __syncthreads();
        generic_pass_dispatch(z, g, p);
    }
    __syncthreads();
    if constexpr (BLUESTEIN) {
        bluestein_multiply(z, g);
        for (uint32_t p = g.npass; p-- > 0;) {
            __syncthreads();
            generic_pass_inverse_dispatch(z, g, p);
        }
        __syncthreads();
        bluestein_split(z, mags, g);
    } else {
        generic_split(z, mags, g);
    }
    __syncthreads();
    if (a.mags_out != nullptr) {
        float2 *out = reinterpret_cast<float2 *>(a.mags_out) + frame0 * a.k_out;
        for (uint32_t j = threadIdx.x; j < a.k_out; j += GENERIC_THREADS) {
            const float2 m = mags[j + 1];
            if constexpr (CH == 1) {
                out[j] = make_float2(m.x, m.x);
                if (second_valid) out[a.k_out + j] = make_float2(m.y, m.y);
            } else {
                out[j] = m;
            }
        }
    }
    if (a.rgba != nullptr) {
        for (uint32_t c = 0; c < a.n_chunks; ++c) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(a.chunks) + c);
            const RowChunk ck = {raw.x, raw.y, raw.z, raw.w};
            if (c != 0) __syncthreads();
#if SBX_GENERIC_PK_EPILOGUE
            // the packed (L, R) epilogue of the q2 kernels: same operations, same roundings (stft_kernels.cuh)
            pk_resample_samples<GENERIC_THREADS, 4>(mags, z, a, ck);
            __syncthreads();
            pk_colour_rows<CH == 1, GENERIC_THREADS, 2>(z, a, ck, frame0, second_valid);
#else
            resample_samples<GENERIC_THREADS>(mags, z, a, ck);
            __syncthreads();
            colour_rows<CH == 1, GENERIC_THREADS>(z, a.thr, a.lut, a, ck, frame0, second_valid);
#endif
        }
    }
    if (a.done_flag != nullptr) {                                         // live single frame, see fused_frames_kernel
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned int v = *a.done_count + 1u;
            *a.done_count = v;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned int *>(a.done_flag) = v;
        }
    }
}

// colourise-only for generic plans: magnitudes from global memory, scratch in shared memory
__global__ void __launch_bounds__(GENERIC_THREADS) static generic_colour_kernel(const KernelArgs a, uint32_t scratch_slots)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *scratch = reinterpret_cast<float2 *>(smem_raw);
    (void)scratch_slots;
    const unsigned long long frame = blockIdx.x;
    const float2 *mags = reinterpret_cast<const float2 *>(a.mags_in) + frame * a.k_out - 1;   // bin kk = j + 1
    for (uint32_t c = 0; c < a.n_chunks; ++c) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(a.chunks) + c);
        const RowChunk ck = {raw.x, raw.y, raw.z, raw.w};
        if (c != 0) __syncthreads();
        resample_samples<GENERIC_THREADS>(mags, scratch, a, ck);
        __syncthreads();
        colour_rows<false, GENERIC_THREADS>(scratch, a.thr, a.lut, a, ck, frame, false);
    }
}
#endif  // SBX_EMU

}  // namespace sbx
