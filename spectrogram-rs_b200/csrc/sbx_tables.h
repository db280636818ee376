// sbx_tables.h -- host-side computation of every table a plan uploads (pure C++, no CUDA calls).
// Shared by sbx_plan.cu (the product) and tests/emu (host emulation of the kernel phases).
// Each function cites the reference lines it restates (paths relative to the reference root).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "../../include/sbx.h"
#include "stft_generic.cuh"
#include "stft_large.cuh"

namespace sbx {
namespace tables {

// src/fourier/fft.rs:59-63 -- periodic Hann, all operations in f32
inline float hann_f32(uint32_t i, uint32_t n)
{
    const float tau = 6.28318530717958647692f;
    return 0.5f * (1.0f - cosf((tau * (float)i) / (float)n));
}

// src/log_scaling.rs:114-119,160-191 via plotters' RangedCoordf64::unmap; f64, then `as f32`
// (src/widgets/simple_spectrogram.rs:142-145)
inline float row_edge_hz(double f_min, double f_max, uint32_t py, uint32_t rows)
{
    const double lo = std::log(f_min), hi = std::log(f_max);
    const double off = (double)(int32_t)py / (double)(int32_t)rows;
    const double lin = (hi - lo) * off + lo;
    return (float)(std::exp(lin) + 0.0);
}

// src/fourier/interpolated_frequency_sample.rs:24-31
inline float index_of(float f, float period, uint32_t k_out)
{
    float idx = f * period;
    const float hi = (float)(k_out - 1);
    if (idx < 0.0f) idx = 0.0f;
    if (idx > hi) idx = hi;
    return idx;
}

// src/colorscheme.rs:59-61
inline float normalised_db(float p, const sbx_config &c)
{
    const float db = 10.0f * log10f(p + c.db_eps);
    return (db - c.db_min) / (c.db_max - c.db_min);
}
// gradient.eval_continuous(t as f64) for 256-entry ramps: clamp(floor(t * 256), 0, 255)
inline uint32_t colour_step(float p, const sbx_config &c)
{
    const double v = std::floor((double)normalised_db(p, c) * 256.0);
    if (!(v >= 0.0)) return 0;
    return v > 255.0 ? 255u : (uint32_t)v;
}
// (alpha * 255.0) as u8 -- src/widgets/simple_spectrogram.rs:159 (saturating cast, NaN -> 0)
inline uint32_t alpha_step(float p, const sbx_config &c)
{
    const float a = normalised_db(p, c) * 255.0f;
    if (!(a >= 0.0f)) return 0;
    return a >= 255.0f ? 255u : (uint32_t)a;
}

// Both step functions are monotone non-decreasing in p (p + eps, log10f, affine map, floor), so
// each is fully described by 255 thresholds: step(p) = #{ i : p >= thr[i] }.  thr[i] = the
// smallest non-negative f32 whose step is >= i+1, found by bisection over the f32 bit patterns.
template <class F> void step_thresholds(F step, float *thr /*256*/)
{
    auto from_bits = [](uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; };
    const uint32_t inf_bits = 0x7f800000u;
    for (uint32_t level = 1; level <= 255; ++level) {
        if (step(from_bits(inf_bits)) < level) { thr[level - 1] = std::numeric_limits<float>::quiet_NaN(); continue; }
        uint32_t lo = 0, hi = inf_bits;                 // invariant: step(hi) >= level
        while (lo < hi) {
            const uint32_t mid = lo + (hi - lo) / 2;
            if (step(from_bits(mid)) >= level) hi = mid; else lo = mid + 1;
        }
        thr[level - 1] = from_bits(hi);
    }
    thr[255] = std::numeric_limits<float>::infinity();
}


struct HostTables {
    uint32_t n = 0, pad = 0, m = 0, k_out = 0, rows = 0;
    float period = 0.0f, scale = 0.0f;
    std::vector<float> window;          // [N]
    std::vector<float> wtw;             // [(P-1) N] x (re, im): window[n] * W_M^(n q0)
    std::vector<float> twp;             // per-pass twiddle planes x (re, im), see KernelArgs::twp
    std::vector<float> wq, wq_real;     // q2 kernels: [S_0] x (re, im) W_M^lo, lo < N / R_0 (complex transform / real-sequence transform)
    std::vector<float> twp_block;       // large path: planes of the 4096-point block transform
    // REAL geometry (one mono frame per CTA as the even/odd packed sequence of N/2 pairs, M in {16384, 32768})
    bool real_ok = false;               // the tables below exist
    bool real_auto = false;             // ... and mono batches take the geometry by default (see build)
    std::vector<float> twp_real;        // per-pass planes of the N/2-point sub-transform
    std::vector<float> wtw4;            // [(P-1) N/2] x {w[2n] c, w[2n+1] s, w[2n] s, w[2n+1] c}, (c, s) = W_(M/2)^(n q0)
    std::vector<float> rtw;             // [M/32] x (re, im): W_M^t, recombination twiddle of bin t
    bool large = false;                 // M > 16384: the multi-kernel path of stft_large.cuh
    bool generic = false;               // N not a power of two: the run-time mixed-radix path of stft_generic.cuh
    std::vector<float> twm;             // generic: [M] x (re, im) full twiddle table W_M^t
    std::vector<uint32_t> pos_tab;      // generic: [M] position of bin kk
    std::vector<uint32_t> radix, stride;// generic: per-pass radix and butterfly stride
    std::vector<float> twg;             // generic: per-pass twiddle planes, see GenericArgs::twg
    std::vector<uint32_t> tw_off;       // generic: first entry of each pass in twg
    // Bluestein (generic windows with a prime factor > 7): radix / stride / twg / tw_off / pos_tab describe the
    // power-of-two convolution length bl_len instead of N
    bool bluestein = false;
    uint32_t bl_len = 0;
    std::vector<float> bl_a;            // [N] x (re, im): window[n] * conj(b[n]), b[j] = exp(+i pi j^2 / M)
    std::vector<float> bl_bf;           // [L] x (re, im): FFT_L(b wrapped mod L) / L, stored at the position of each bin
    std::vector<float> bl_c;            // [M] x (re, im): conj(b[k])
    std::vector<RowMap> rowmap;         // [H]
    std::vector<uint32_t> rowinfo;      // [H] n | first scratch slot << 16 (what the kernels read)
    std::vector<RowChunk> chunks;       // rows grouped so that a chunk's samples fit the kernel's scratch area
    std::vector<uint32_t> samp;         // per (row, sample): {floor(x), weight bits}, see KernelArgs::samp
    uint32_t bank_group = 16;           // lanes / rows served by one shared-memory request (see the sample table in build)
    bool taps_clamped = false;          // some cubic tap falls outside [0, K-1]
    std::vector<float> thr;             // [2][256] pairs (ts[i], ts[i+1]), see step_index
    std::vector<float> thr_raw;         // [2][256] the bare thresholds (introspection)
    float guess_a[2], guess_b[2];
    float step_margin = 1.0f;           // KernelArgs::step_margin (1: the guessed colour step is always checked)
    uint32_t lut[256];
    uint32_t ctab[1024];                // [256] {ts[i], ts[i+1], lut[i], 0} (colour step bracket + colour)
    bool ok = true;                     // false: a row has more samples than the scratch area holds
};

// radix plan of the length-N sub-transform: first pass radix 2^(log2n % 4) when that is not 1, then 16s
struct RadixPlan {
    int log2n, rem, npass;
    explicit RadixPlan(int l) : log2n(l), rem(l & 3), npass(l / 4 + ((l & 3) ? 1 : 0)) {}
    int log2r(int p) const { return (rem != 0 && p == 0) ? rem : 4; }
    int log2stride(int p) const { int s = log2n; for (int i = 0; i <= p; ++i) s -= log2r(i); return s; }
};

inline void build(const sbx_config &cfg, HostTables &t)
{
    t.ok = true;
    t.n = cfg.window_n;
    t.pad = cfg.pad_factor;
    t.m = t.n * t.pad;
    t.k_out = t.m / 2 - 1;              // every bin strictly between DC and Nyquist; N-1 at P = 2 (fft.rs:33)
    t.rows = cfg.rows_h;
    t.scale = 2.0f / (float)t.n;                                                       // fft.rs:92
    // interpolated_frequency_sample.rs:52-54 with SampleRate(sample_rate as u32), simple_spectrogram.rs:138
    t.period = (2.0f * (float)t.k_out) / (float)(uint32_t)cfg.sample_rate;

    t.window.resize(t.n);
    for (uint32_t i = 0; i < t.n; ++i) t.window[i] = hann_f32(i, t.n);

    t.generic = (t.n & (t.n - 1)) != 0 || t.n < 256;
    t.large = !t.generic && t.m > 16384;

    // sub-transform premultiply: window[n] * W_M^(n q0), q0 = 1..P-1 (f64 product of the f32 window);
    // the generic path multiplies by W_M^(n q0) from its full table instead
    t.wtw.clear();
    if (!t.generic) t.wtw.assign((size_t)(t.pad - 1) * t.n * 2, 0.0f);
    for (uint32_t q0 = 1; q0 < t.pad && !t.generic; ++q0)
        for (uint32_t i = 0; i < t.n; ++i) {
            const double ang = -2.0 * M_PI * (double)(((uint64_t)i * q0) % t.m) / (double)t.m;
            const size_t o = ((size_t)(q0 - 1) * t.n + i) * 2;
            t.wtw[o] = (float)((double)t.window[i] * std::cos(ang));
            t.wtw[o + 1] = (float)((double)t.window[i] * std::sin(ang));
        }

    // ---- transform tables ----------------------------------------------------------------------
    t.twm.clear(); t.pos_tab.clear(); t.radix.clear(); t.stride.clear(); t.twg.clear(); t.tw_off.clear();
    if (t.generic) {
        // radix list from {16, 15, 12, 10, 9, 8, 7, 6, 5, 4, 3, 2}: the factorisation with the FEWEST passes (every pass
        // is a barrier-separated trip through shared memory), larger radices first; a left-over factor means a
        // prime > 7.  SBX_GENERIC_PRIME_RADICES=1 restores round 1's {16, 8, 4, 2, 3, 5, 7} list (A/B measurements).
        uint32_t rest = t.n;
        for (uint32_t f : {2u, 3u, 5u, 7u})
            while (rest % f == 0) rest /= f;
        const char *env_prime = std::getenv("SBX_GENERIC_PRIME_RADICES");
        if (rest == 1 && !(env_prime && env_prime[0] == '1')) {
            const uint32_t cand[] = {16, 15, 12, 10, 9, 8, 7, 6, 5, 4, 3, 2};
            std::vector<uint32_t> cur, best;
            // depth-limited search, non-increasing factors
            struct Search {
                static bool go(uint32_t n, size_t max_r, int left, const uint32_t *cand, std::vector<uint32_t> &cur)
                {
                    if (n == 1) return true;
                    if (left == 0) return false;
                    for (size_t i = max_r; i < 12; ++i) {
                        if (n % cand[i] != 0) continue;
                        cur.push_back(cand[i]);
                        if (go(n / cand[i], i, left - 1, cand, cur)) return true;
                        cur.pop_back();
                    }
                    return false;
                }
            };
            for (int depth = 1; depth <= GENERIC_MAX_PASSES && best.empty(); ++depth) {
                cur.clear();
                if (Search::go(t.n, 0, depth, cand, cur)) best = cur;
            }
            // The LAST pass has stride 1: a thread owns R consecutive elements, so the 16 lanes of a shared-memory
            // request sit R elements apart -- conflict-free when R is odd, gcd(R, 16)-way conflicted otherwise
            // (16-way for R = 16).  Put the radix with the smallest gcd(R, 16) last (ties: the larger one).
            if (best.size() > 1) {
                auto g16 = [](uint32_t r) { uint32_t g = 1; while (g < 16 && r % (2 * g) == 0) g *= 2; return g; };
                size_t pick = 0;
                for (size_t i = 1; i < best.size(); ++i)
                    if (g16(best[i]) < g16(best[pick]) || (g16(best[i]) == g16(best[pick]) && best[i] > best[pick])) pick = i;
                const uint32_t last = best[pick];
                best.erase(best.begin() + (long)pick);
                best.push_back(last);
            }
            t.radix = best;
            rest = best.empty() ? t.n : 1;
        } else {
            rest = t.n;
            const uint32_t cand[] = {16, 8, 4, 2, 3, 5, 7};
            for (uint32_t r : cand)
                while (rest % r == 0 && rest > 1) { t.radix.push_back(r); rest /= r; }
        }
        if ((t.m & 1u) || t.m < 8) { t.radix.clear(); t.ok = false; return; }
        uint32_t fft_len = t.n;                                  // length the passes transform (P sub-transforms of N)
        if (rest != 1 || t.radix.size() > (size_t)GENERIC_MAX_PASSES) {
            // a prime factor > 7 (N = 1102 = 2 * 19 * 29 at 22.05 kHz, 551, ...): Bluestein over a power-of-two
            // circular convolution of length L >= N + M - 1, which must fit one SM's shared memory beside the magnitudes
            uint32_t len = 8;
            while (len < t.n + t.m - 1) len <<= 1;
            if (len > BLUESTEIN_MAX_LEN) { t.radix.clear(); t.ok = false; return; }
            t.bluestein = true;
            t.bl_len = fft_len = len;
            t.radix.clear();
            for (uint32_t rest2 = len; rest2 > 1;) {
                const uint32_t r = rest2 % 16 == 0 ? 16u : (rest2 % 8 == 0 ? 8u : (rest2 % 4 == 0 ? 4u : 2u));
                t.radix.push_back(r);
                rest2 /= r;
            }
        }
        uint32_t s = fft_len;
        for (uint32_t r : t.radix) { s /= r; t.stride.push_back(s); }
        t.twm.resize((size_t)t.m * 2);
        for (uint32_t k = 0; k < t.m; ++k) {
            const double ang = -2.0 * M_PI * (double)k / (double)t.m;
            t.twm[2 * (size_t)k] = (float)std::cos(ang);
            t.twm[2 * (size_t)k + 1] = (float)std::sin(ang);
        }
        // per-pass twiddle planes (all passes but the last): plane q-1, entry lo = W_(S R)^(lo q), contiguous in lo
        t.twg.clear(); t.tw_off.clear();
        for (size_t p = 0; p < t.radix.size(); ++p) {
            t.tw_off.push_back((uint32_t)(t.twg.size() / 2));
            if (p + 1 == t.radix.size()) break;
            const uint32_t R = t.radix[p], S = t.stride[p];
            for (uint32_t q = 1; q < R; ++q)
                for (uint32_t lo = 0; lo < S; ++lo) {
                    const double ang = -2.0 * M_PI * (double)(((uint64_t)lo * q) % ((uint64_t)S * R)) / (double)((uint64_t)S * R);
                    t.twg.push_back((float)std::cos(ang));
                    t.twg.push_back((float)std::sin(ang));
                }
        }
        // bin kk = P k + q0, k = sum d_p (R_0 .. R_(p-1))  ->  position q0 N + sum d_p S_p
        const uint32_t tab_len = t.bluestein ? t.bl_len : t.m, tab_pad = t.bluestein ? 1u : t.pad, tab_n = t.bluestein ? t.bl_len : t.n;
        t.pos_tab.resize(tab_len);
        for (uint32_t kk = 0; kk < tab_len; ++kk) {
            uint32_t k = kk / tab_pad, pos = (kk % tab_pad) * tab_n;
            for (size_t p = 0; p < t.radix.size(); ++p) { pos += (k % t.radix[p]) * t.stride[p]; k /= t.radix[p]; }
            t.pos_tab[kk] = pos;
        }
        if (t.bluestein) {
            const uint32_t len = t.bl_len;
            // b[j] = exp(+i pi j^2 / M): the angle from j^2 mod 2M, exactly (integers), then f64 cos / sin
            auto chirp = [&](uint64_t j, double &re, double &im) {
                const uint64_t r = (j * j) % (2ull * t.m);
                const double ang = M_PI * (double)r / (double)t.m;
                re = std::cos(ang); im = std::sin(ang);
            };
            t.bl_a.resize((size_t)t.n * 2);
            for (uint32_t i = 0; i < t.n; ++i) {
                double re, im;
                chirp(i, re, im);
                t.bl_a[2 * (size_t)i] = (float)((double)t.window[i] * re);
                t.bl_a[2 * (size_t)i + 1] = (float)(-(double)t.window[i] * im);
            }
            t.bl_c.resize((size_t)t.m * 2);
            for (uint32_t k = 0; k < t.m; ++k) {
                double re, im;
                chirp(k, re, im);
                t.bl_c[2 * (size_t)k] = (float)re;
                t.bl_c[2 * (size_t)k + 1] = (float)(-im);
            }
            // h[j mod L] = b[j] for j in [-(N-1), M-1]; its L-point DFT in f64 (iterative radix 2), scaled by 1 / L
            std::vector<double> hr(len, 0.0), hi(len, 0.0);
            for (uint32_t j = 0; j < t.m; ++j) chirp(j, hr[j], hi[j]);
            for (uint32_t j = 1; j < t.n; ++j) chirp(j, hr[len - j], hi[len - j]);
            for (uint32_t i = 1, j = 0; i < len; ++i) {                      // bit reversal
                uint32_t bit = len >> 1;
                for (; j & bit; bit >>= 1) j ^= bit;
                j ^= bit;
                if (i < j) { std::swap(hr[i], hr[j]); std::swap(hi[i], hi[j]); }
            }
            for (uint32_t span = 2; span <= len; span <<= 1) {
                const uint32_t half = span >> 1;
                for (uint32_t base = 0; base < len; base += span)
                    for (uint32_t j = 0; j < half; ++j) {
                        const double ang = -2.0 * M_PI * (double)j / (double)span, wr = std::cos(ang), wi = std::sin(ang);
                        const double xr = hr[base + j + half] * wr - hi[base + j + half] * wi;
                        const double xi = hr[base + j + half] * wi + hi[base + j + half] * wr;
                        hr[base + j + half] = hr[base + j] - xr; hi[base + j + half] = hi[base + j] - xi;
                        hr[base + j] += xr; hi[base + j] += xi;
                    }
            }
            t.bl_bf.resize((size_t)len * 2);
            for (uint32_t k = 0; k < len; ++k) {
                t.bl_bf[2 * (size_t)t.pos_tab[k]] = (float)(hr[k] / (double)len);
                t.bl_bf[2 * (size_t)t.pos_tab[k] + 1] = (float)(hi[k] / (double)len);
            }
        }
    }
    // per-pass twiddle planes: plane j of pass p holds W_(S_p R_p)^(2^j lo), lo < S_p
    int log2n = 0;
    while ((1u << log2n) < t.n) ++log2n;
    auto planes = [](const RadixPlan &rp, int first_pass, int last_pass, std::vector<float> &dst) {
        dst.clear();
        for (int p = first_pass; p < last_pass; ++p) {
            const uint32_t s = 1u << rp.log2stride(p), span = s << rp.log2r(p);
            for (int j = 0; j < TW_PLANES; ++j)
                for (uint32_t lo = 0; lo < s; ++lo) {
                    const double ang = -2.0 * M_PI * (double)(((uint64_t)lo << j) % span) / (double)span;
                    dst.push_back((float)std::cos(ang));
                    dst.push_back((float)std::sin(ang));
                }
        }
    };
    const RadixPlan rp(log2n);
    if (t.generic) {
        t.twp.clear();
        t.twp_block.clear();
    } else if (t.large) {
        planes(rp, 0, 1, t.twp);                               // the global radix-R0 pass: W_N^(2^j lo), lo < 4096
        const RadixPlan rb(LARGE_LOG2_BLOCK);
        planes(rb, 0, rb.npass - 1, t.twp_block);              // the 4096-point block transform
    } else {
        planes(rp, 0, rp.npass - 1, t.twp);
        t.twp_block.clear();
    }

    // q2 kernels (pad factor 2): W_M^lo for lo below the first pass's stride; lane B's pass-0 factor is W_M^lo times a constant
    auto wq_table = [](int log2n_sub, uint32_t m_sub, std::vector<float> &dst) {
        const RadixPlan r(log2n_sub);
        const uint32_t s0 = 1u << r.log2stride(0);
        dst.resize((size_t)s0 * 2);
        for (uint32_t lo = 0; lo < s0; ++lo) {
            const double ang = -2.0 * M_PI * (double)lo / (double)m_sub;
            dst[2 * (size_t)lo] = (float)std::cos(ang);
            dst[2 * (size_t)lo + 1] = (float)std::sin(ang);
        }
    };
    t.wq.clear(); t.wq_real.clear();
    if (!t.generic && !t.large && t.pad == 2) wq_table(log2n, t.m, t.wq);

    // REAL geometry: transform of the packed sequence z[n] = x[2n] + i x[2n+1], length M/2 with N/2 non-zero pairs
    t.real_ok = !t.generic && t.m >= REAL_MIN_M && t.m <= REAL_MAX_M && t.n >= 512;
    t.twp_real.clear(); t.wtw4.clear(); t.rtw.clear();
    if (t.real_ok) {
        const RadixPlan rr(log2n - 1);
        planes(rr, 0, rr.npass - 1, t.twp_real);
        const uint32_t nc = t.n / 2, mc = t.m / 2;
        t.wtw4.assign((size_t)(t.pad - 1) * nc * 4, 0.0f);
        for (uint32_t q0 = 1; q0 < t.pad; ++q0)
            for (uint32_t i = 0; i < nc; ++i) {
                const double ang = -2.0 * M_PI * (double)(((uint64_t)i * q0) % mc) / (double)mc;
                const double c = std::cos(ang), sn = std::sin(ang);
                const double w0 = (double)t.window[2 * i], w1 = (double)t.window[2 * i + 1];
                float *o = t.wtw4.data() + ((size_t)(q0 - 1) * nc + i) * 4;
                o[0] = (float)(w0 * c); o[1] = (float)(w1 * sn); o[2] = (float)(w0 * sn); o[3] = (float)(w1 * c);
            }
        if (t.pad == 2) wq_table(log2n - 1, mc, t.wq_real);
        t.rtw.resize((size_t)(t.m / 32) * 2);
        for (uint32_t k = 0; k < t.m / 32; ++k) {
            const double ang = -2.0 * M_PI * (double)k / (double)t.m;
            t.rtw[2 * (size_t)k] = (float)std::cos(ang);
            t.rtw[2 * (size_t)k + 1] = (float)std::sin(ang);
        }
    }

    t.rowmap.resize(t.rows);
    std::vector<uint32_t> samp_first(t.rows);                 // first entry of each row in the per-sample table
    uint64_t total_samples = 0;
    for (uint32_t py = 0; py < t.rows; ++py) {
        const float f0 = row_edge_hz(cfg.f_min, cfg.f_max, py, t.rows);
        const float f1 = row_edge_hz(cfg.f_min, cfg.f_max, py + 1, t.rows);
        const float i0 = index_of(f0, t.period, t.k_out), i1 = index_of(f1, t.period, t.k_out);
        const float d = std::floor(i1 - i0);
        const uint32_t n = d >= 1.0f ? (uint32_t)d : 1u;                                // ifs.rs:64
        RowMap rm;
        rm.f0 = f0;
        rm.step = (f1 - f0) / (float)n;                                                // lin_space step
        rm.n = n;
        rm.aux = 0;                                                                    // scratch slot, set below
        samp_first[py] = (uint32_t)total_samples;
        total_samples += n;
        t.rowmap[py] = rm;
    }
    // Row chunks and scratch slots.  A chunk is a run of rows whose interpolated samples fit the kernel's scratch
    // area; row py's samples occupy slots aux .. aux+n-1 of its chunk.  The threads of colour_rows (one per row) read
    // slot aux + s at step s of their sequential sums, so the rows a shared-memory request serves together -- an
    // aligned group of BANK_GROUP rows: 16 for the 8-byte (L,R) elements of the pair / stereo kernels (a half-warp per
    // wavefront), 32 for the 4-byte elements of the real-sequence geometry -- must start in different residues modulo
    // BANK_GROUP.  Three slot policies, the first that needs the fewest chunks wins:
    //   GREEDY  a row starts at the next slot whose residue no earlier row of its group has taken (no padding at
    //           all where n is odd or 1; one slot per 16/n rows where it is even);
    //   ODD     rows with an even n >= 4 get one slot of padding (odd per-row stride; the round-1 rule);
    //   DENSE   no padding.
    const uint32_t cap = t.large ? (uint32_t)LARGE_BLOCK : t.m / 2;
    const bool real_likely = t.real_ok && !t.large && RadixPlan(log2n - 1).npass < RadixPlan(log2n).npass;
    const uint32_t bank_group = real_likely ? 32u : 16u;
    t.bank_group = bank_group;
    if (cap >= 0x7fffu) t.ok = false;                                     // slots travel in 15 bits, 0x7fff = "no store"
    enum SlotPolicy { GREEDY = 0, ODD = 1, DENSE = 2 };
    auto assign_slots = [&](SlotPolicy policy, std::vector<RowChunk> &chunks, std::vector<uint32_t> &aux, bool &fits) {
        chunks.clear();
        aux.assign(t.rows, 0u);
        fits = true;
        RowChunk ck = {0, 0, 0, 0};
        uint32_t slot = 0, used = 0;                                      // used: residues taken in the current row group
        for (uint32_t py = 0; py < t.rows; ++py) {
            const uint32_t n = t.rowmap[py].n;
            if (n > cap) {
                // a row too long for the scratch area (large path only: few rows over a very long transform) gets a
                // chunk of its own, recognisable by samp_end - samp_begin > cap; the kernel resamples it in parts
                if (!t.large) fits = false;
                if (ck.row_end > ck.row_begin) chunks.push_back(ck);
                chunks.push_back(RowChunk{py, py + 1, samp_first[py], samp_first[py] + n});
                ck = RowChunk{py + 1, py + 1, samp_first[py] + n, samp_first[py] + n};
                slot = 0;
                continue;
            }
            for (int attempt = 0; attempt < 2; ++attempt) {
                if (ck.row_end == ck.row_begin) { ck.row_begin = py; ck.samp_begin = samp_first[py]; slot = 0; }
                if ((py - ck.row_begin) % bank_group == 0) used = 0;
                uint32_t first = slot, need = n;
                if (policy == GREEDY)
                    while (used & (1u << (first % bank_group))) ++first;
                else if (policy == ODD && (n & 1u) == 0 && n >= 4)
                    need = n + 1;
                if (first + need > cap && ck.row_end > ck.row_begin) {    // close the chunk and place the row in a new one
                    chunks.push_back(ck);
                    ck = RowChunk{py, py, samp_first[py], samp_first[py]};
                    continue;
                }
                aux[py] = first;
                used |= 1u << (first % bank_group);
                slot = first + need;
                ck.row_end = py + 1;
                ck.samp_end = samp_first[py] + n;
                break;
            }
        }
        if (ck.row_end > ck.row_begin) chunks.push_back(ck);
    };
    {
        std::vector<RowChunk> best_chunks;
        std::vector<uint32_t> best_aux;
        bool have = false;
        const char *env_slots = std::getenv("SBX_TABLE_SLOTS");               // experiment knob: first policy tried
        const int first_policy = env_slots ? std::atoi(env_slots) : 0;
        for (int pi = 0; pi < 3; ++pi) {
            const SlotPolicy policy = (SlotPolicy)((first_policy + pi) % 3);
            std::vector<RowChunk> chunks;
            std::vector<uint32_t> aux;
            bool fits = true;
            assign_slots(policy, chunks, aux, fits);
            if (!fits) t.ok = false;
            if (!have || chunks.size() < best_chunks.size()) { best_chunks.swap(chunks); best_aux.swap(aux); have = true; }
        }
        t.chunks.swap(best_chunks);
        for (uint32_t py = 0; py < t.rows; ++py) t.rowmap[py].aux = best_aux[py];
    }
    for (const RowChunk &c : t.chunks)
        if (c.samp_end - c.samp_begin > cap) t.real_ok = false;            // rows resampled in parts: large path only
    // Default choice between the frame-pair kernel and the real-sequence geometry for mono batches.  Measured on a
    // B200 (M frames/s, pair vs real): cfg3 N 8192 P 2: 22.6 vs 29.3; N 8192 P 4 (M 32768, pair = the four-kernel
    // large path): 5.7 vs 9.0; but cfg1 N 4096 P 2: 59.0 vs 52.7; N 4096 P 4: 25.2 vs 23.6; N 2048 P 4: 49.2 vs 45.1.
    // The half-length transform wins exactly where it needs one shared-memory pass fewer (N/2 = 16^k) or where the
    // pair no longer fits one SM at all.
    t.real_auto = t.real_ok && (t.large || RadixPlan(log2n - 1).npass < RadixPlan(log2n).npass);
    t.rowinfo.resize(t.rows);
    for (uint32_t py = 0; py < t.rows; ++py) t.rowinfo[py] = t.rowmap[py].n | (t.rowmap[py].aux << 16);

    // per-sample table: x = index_of(lin_space(f0..f1, n)[s])  (ifs.rs:67-69, :24-31), stored as
    //   word 0: floor(x) in bits 0..16, the sample's scratch slot in bits 17..31 (0x7fff: filler entry, nothing stored)
    //   word 1: the interpolation weight: mu = x - floor(x) (cubic, ifs.rs:90) or
    //           (1 - cos((x - low) * PI)) / 2 (cosine, ifs.rs:82-83; f32, host libm)
    // One lane handles one entry, and the lanes of one shared-memory request -- an aligned group of BANK_GROUP entries
    // of a chunk -- read the taps at floor(x) + {-1, 0, 1, 2} and store to their slots: both are conflict-free when the
    // group's bins and slots each span fewer than BANK_GROUP positions.  Rows average samples a little MORE than one
    // bin apart, so a full group would span up to 1.5 BANK_GROUP bins; the table therefore closes a group early and
    // fills it with entries that repeat the previous bin (a broadcast) and store nothing.  (Fused and mixed-radix
    // paths; the four-kernel path reads its taps through L1 and keeps the dense table.)
    t.samp.clear();
    t.samp.reserve((size_t)total_samples * 2 + 64);
    t.taps_clamped = false;
    if (t.k_out >= (1u << 17)) t.ok = false;
    const float pi = 3.14159265358979323846f;
    const char *env_fill = std::getenv("SBX_TABLE_FILL");                     // experiment knob
    const int fill_mode = env_fill ? std::atoi(env_fill) : 0;                 // 0 none, 1 every group, 2 within the round budget
    const bool fill_groups = !t.large && fill_mode != 0;
    for (RowChunk &ck : t.chunks) {
        const uint32_t begin = (uint32_t)(t.samp.size() / 2);
        uint32_t group_x = 0, group_slot = 0, prev_x = 0;
        // the kernels take the table in rounds of 2048 entries (threads x loads in flight): fillers must not add a round
        const uint32_t real_entries = ck.samp_end - ck.samp_begin;
        uint32_t filler_budget = fill_mode == 2 ? (real_entries + 2047u) / 2048u * 2048u - real_entries : 0xffffffffu;
        for (uint32_t py = ck.row_begin; py < ck.row_end; ++py) {
            const RowMap &rm = t.rowmap[py];
            for (uint32_t s = 0; s < rm.n; ++s) {
                const float fq = rm.f0 + (float)s * rm.step;
                const float x = index_of(fq, t.period, t.k_out);
                const float fl = std::floor(x);
                const float off = x - fl;
                const float w = cfg.interp == SBX_INTERP_COSINE ? (1.0f - cosf(off * pi)) / 2.0f : off;
                const uint32_t x1 = (uint32_t)fl;
                if (x1 < 1 || x1 + 2 > t.k_out - 1) t.taps_clamped = true;
                uint32_t wbits;
                std::memcpy(&wbits, &w, 4);
                const uint32_t slot = (rm.aux + s) & 0x7fffu;                           // big rows ignore the slot
                uint32_t pos = (uint32_t)(t.samp.size() / 2) - begin;
                if (fill_groups && pos % bank_group != 0 && bank_group - pos % bank_group <= filler_budget &&
                    (x1 < group_x || x1 - group_x >= bank_group || slot < group_slot || slot - group_slot >= bank_group)) {
                    filler_budget -= bank_group - pos % bank_group;
                    for (; pos % bank_group != 0; ++pos) { t.samp.push_back(prev_x | (0x7fffu << 17)); t.samp.push_back(0u); }
                }
                if (pos % bank_group == 0) { group_x = x1; group_slot = slot; }
                t.samp.push_back(x1 | (slot << 17));
                t.samp.push_back(wbits);
                prev_x = x1;
            }
        }
        ck.samp_begin = begin;
        ck.samp_end = (uint32_t)(t.samp.size() / 2);
    }

    t.thr_raw.resize(512);
    const sbx_config c = cfg;
    step_thresholds([&c](float pw) { return colour_step(pw, c); }, t.thr_raw.data());
    step_thresholds([&c](float pw) { return alpha_step(pw, c); }, t.thr_raw.data() + 256);
    // device layout per set: 256 pairs (ts[i], ts[i+1]) of ts = [-inf, thr[0..254], +inf]
    t.thr.assign(2 * THR_STRIDE, std::numeric_limits<float>::infinity());
    for (int set = 0; set < 2; ++set) {
        float ts[257];
        ts[0] = -std::numeric_limits<float>::infinity();
        for (int i = 0; i < 255; ++i) ts[1 + i] = t.thr_raw[set * 256 + i];
        ts[256] = std::numeric_limits<float>::infinity();
        float *dst = t.thr.data() + set * THR_STRIDE;
        for (int i = 0; i < 256; ++i) { dst[2 * i] = ts[i]; dst[2 * i + 1] = ts[i + 1]; }
    }
    // guess of the step: t = (10 log10(x) - db_min) / (db_max - db_min) * levels = log2(x) * a + b
    const double range = (double)cfg.db_max - (double)cfg.db_min;
    for (int set = 0; set < 2; ++set) {
        const double levels = set == 0 ? 256.0 : 255.0;
        t.guess_a[set] = (float)(10.0 * std::log10(2.0) * levels / range);
        t.guess_b[set] = (float)(-(double)cfg.db_min * levels / range);
    }

    // How far the device's guess  g(p) = log2(p + eps) * a + b  can be from the point where the reference's step
    // function actually changes, in steps.  Both follow T(p) = (10 log10(p + eps) - db_min) / range * 256:
    //   * the reference chain (f32 log10f, subtract, divide, floor): measured here -- T at every exact threshold and
    //     at the float just below it must straddle the integer it belongs to, to within err_ref;
    //   * the device guess: p + eps rounded to f32 (|d log2| <= 1.8e-7), MUFU.LG2 (__log2f: 2 ulp of |log2| <= 128,
    //     i.e. <= 3.1e-5 absolute), the FMA and the f32 roundings of a and b.
    // A guess whose fractional part lies in (margin, 1 - margin) cannot be on the wrong side of a threshold.
    {
        const double range_d = (double)cfg.db_max - (double)cfg.db_min;
        auto T = [&](float pw) {
            const float sum = pw + cfg.db_eps;                                  // the f32 sum both sides start from
            return (10.0 * std::log10((double)sum) - (double)cfg.db_min) / range_d * 256.0;
        };
        double err_ref = 0.0;
        bool valid = true;
        for (int i = 0; i < 255; ++i) {
            const float th = t.thr_raw[i];
            if (std::isnan(th)) continue;                                       // level never reached
            if (!(th > 0.0f) || std::isinf(th)) { if (th != 0.0f) valid = false; continue; }
            const float below = std::nextafterf(th, 0.0f);
            const double hi = T(th) - (double)(i + 1), lo = T(below) - (double)(i + 1);
            if (!std::isfinite(hi) || !std::isfinite(lo)) { valid = false; continue; }
            err_ref = std::max(err_ref, std::max(std::fabs(hi), std::fabs(lo)));
        }
        const double a_abs = std::fabs((double)t.guess_a[0]);
        const double err_dev = a_abs * (1.8e-7 + 3.1e-5) + 2.0 * 256.0 * 1.2e-7 + (a_abs * 128.0 + std::fabs((double)t.guess_b[0])) * 1.2e-7;
        const double margin = 2.0 * (err_ref + err_dev);
        t.step_margin = (valid && margin < 0.25) ? (float)std::max(margin, 1e-3) : 1.0f;
    }

    for (int i = 0; i < 256; ++i)
        t.lut[i] = (uint32_t)cfg.lut_rgb[3 * i] | ((uint32_t)cfg.lut_rgb[3 * i + 1] << 8) |
                   ((uint32_t)cfg.lut_rgb[3 * i + 2] << 16) | (255u << 24);
    for (int i = 0; i < 256; ++i) {
        std::memcpy(&t.ctab[4 * i], &t.thr[2 * i], 8);                 // (ts[i], ts[i+1]) of the colour set
        t.ctab[4 * i + 2] = t.lut[i];
        t.ctab[4 * i + 3] = 0;
    }
}

// run-time pass description of the mixed-radix path (device pointers are set by the caller)
inline void fill_generic_passes(const HostTables &t, GenericArgs &g)
{
    g.n = t.n; g.pad = t.pad; g.m = t.m;
    g.bl_n = t.n; g.bl_m = t.m;
    if (t.bluestein) { g.n = g.m = t.bl_len; g.pad = 1; }     // the passes transform ONE sequence of length L
    g.npass = (uint32_t)t.radix.size();
    for (size_t i = 0; i < t.radix.size(); ++i) {
        g.radix[i] = t.radix[i];
        g.stride[i] = t.stride[i];
        g.per_sub[i] = g.n / t.radix[i];
        g.tw_off[i] = t.tw_off[i];
        g.inv_stride[i] = (uint32_t)((1ull << 32) / t.stride[i]) + 1u;      // unused (and wraps) for S = 1
    }
}

}  // namespace tables
}  // namespace sbx
