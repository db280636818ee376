// stft_generic.cuh -- windows that are not a power of two (the live app's N = (0.05 s * sr) as usize:
// 2400 @ 48 kHz, 2205 @ 44.1 kHz; src/widgets/gpu_spectrogram.rs:323, src/fourier/fft.rs:19-21).
//
// Same structure as the fused kernel -- P premultiplied sub-transforms of length N, in-place DIF,
// digit-reversed output, conjugate-symmetry split, tabulated resample, threshold dB step -- with the
// radix list chosen at run time from {16, 8, 4, 2, 3, 5, 7}: N = 2400 = 16*2*3*5*5, 2205 = 3*3*5*7*7.
// This is the low-rate live path (819 frames/s in the reference), so it trades the fused kernel's
// swizzles and compile-time index math for generality: bin -> position comes from a host table, twiddles
// from per-pass planes (contiguous in the butterfly index), divisions by host-computed reciprocals.
// Measured on a B200, N = 2400 stereo: 18.3 M frames/s in batch, 27 us per live frame (12.0 M / 41 us with
// strided loads from the full W_M table and run-time divisions).  Composite radices evaluated in registers
// (20, 21, 15, ...: 2400 = 20*20*6, three passes instead of five) were measured and rejected: 128 registers halve
// the occupancy and the batch rate drops to 12.8 M frames/s, with no gain for the single live frame.  One CTA per frame (stereo) or frame pair (mono); the transform (M float2)
// and the magnitudes (M/2 float2) live in shared memory: M <= 16384.
#pragma once
#include "stft_kernels.cuh"

namespace sbx {

#ifndef SBX_GENERIC_THREADS
#define SBX_GENERIC_THREADS 512     // measured on a B200, N = 2400: live p50 43.2 / 40.9 / 40.9 us and batch 12.0 / 12.1 / 8.8 M frames/s at 256 / 512 / 1024
#endif
constexpr int GENERIC_THREADS = SBX_GENERIC_THREADS;
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

// R-point DFT, natural order in and out.  Powers of two reuse the fused kernel's butterflies; 3, 5 and 7
// are evaluated directly from the table (w[j] = W_R^j).
template <int R> struct SmallDft {
    static __device__ __forceinline__ void run(float2 (&v)[R], const float2 *twm, uint32_t m)
    {
        if constexpr (R == 2 || R == 4 || R == 8 || R == 16) {
            (void)twm; (void)m;
            Butterfly<R>::run(v);
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

__device__ __forceinline__ void generic_pass_dispatch(float2 *z, const GenericArgs &g, uint32_t p)
{
    const bool last = p + 1 == g.npass;
    switch (g.radix[p]) {
        case 2: generic_pass<2>(z, g, p, last); break;
        case 3: generic_pass<3>(z, g, p, last); break;
        case 4: generic_pass<4>(z, g, p, last); break;
        case 5: generic_pass<5>(z, g, p, last); break;
        case 7: generic_pass<7>(z, g, p, last); break;
        case 8: generic_pass<8>(z, g, p, last); break;
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
__global__ void __launch_bounds__(GENERIC_THREADS) generic_frames_kernel(const GenericArgs g)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *z = reinterpret_cast<float2 *>(smem_raw);                    // [M] ([L] for Bluestein); reused as resample scratch
    float2 *mags = z + g.m;                                               // [M/2]
    const KernelArgs &a = g.k;
    const unsigned long long frame0 = (unsigned long long)blockIdx.x * (CH == 1 ? 2 : 1);
    const bool second_valid = CH == 1 && frame0 + 1 < a.n_frames;
    if constexpr (BLUESTEIN) bluestein_load<CH>(z, g, frame0, second_valid);
    else generic_load<CH>(z, g, frame0, second_valid);
    for (uint32_t p = 0; p < g.npass; ++p) {
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
            resample_samples<GENERIC_THREADS>(mags, z, a, ck);
            __syncthreads();
            colour_rows<CH == 1, GENERIC_THREADS>(z, a.thr, a.lut, a, ck, frame0, second_valid);
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
