// sbx_tables.h -- host-side computation of every table a plan uploads (pure C++, no CUDA calls).
// Shared by sbx_plan.cu (the product) and tests/emu (host emulation of the kernel phases).
// Each function cites the reference lines it restates (paths relative to the reference root).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../../include/sbx.h"
#include "stft_kernels.cuh"

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
    std::vector<float> tw;              // [M] x (re, im)
    std::vector<RowMap> rowmap;         // [H]
    std::vector<float> cos_w;           // COSINE weights per (row, sample)
    std::vector<float> thr;             // [512]
    uint32_t lut[256];
};

inline void build(const sbx_config &cfg, HostTables &t)
{
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

    t.tw.resize((size_t)t.m * 2);
    for (uint32_t k = 0; k < t.m; ++k) {
        const double ang = -2.0 * M_PI * (double)k / (double)t.m;
        t.tw[2 * (size_t)k] = (float)std::cos(ang);
        t.tw[2 * (size_t)k + 1] = (float)std::sin(ang);
    }

    t.rowmap.resize(t.rows);
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
        rm.aux = (uint32_t)total_samples;
        total_samples += n;
        t.rowmap[py] = rm;
    }
    t.cos_w.clear();
    if (cfg.interp == SBX_INTERP_COSINE) {
        // (1 - cos(frac * PI)) / 2 per (row, sample), f32, host libm -- ifs.rs:82-83
        const float pi = 3.14159265358979323846f;
        t.cos_w.resize(total_samples);
        for (uint32_t py = 0; py < t.rows; ++py) {
            const RowMap &rm = t.rowmap[py];
            for (uint32_t s = 0; s < rm.n; ++s) {
                const float fq = rm.f0 + (float)s * rm.step;
                const float x = index_of(fq, t.period, t.k_out);
                const float off = x - std::floor(x);
                t.cos_w[rm.aux + s] = (1.0f - cosf(off * pi)) / 2.0f;
            }
        }
    }

    t.thr.resize(512);
    const sbx_config c = cfg;
    step_thresholds([&c](float pw) { return colour_step(pw, c); }, t.thr.data());
    step_thresholds([&c](float pw) { return alpha_step(pw, c); }, t.thr.data() + 256);

    for (int i = 0; i < 256; ++i)
        t.lut[i] = (uint32_t)cfg.lut_rgb[3 * i] | ((uint32_t)cfg.lut_rgb[3 * i + 1] << 8) |
                   ((uint32_t)cfg.lut_rgb[3 * i + 2] << 16) | (255u << 24);
}

}  // namespace tables
}  // namespace sbx
