/*
 * stft_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the spectrogram-rs frame path
 *     PCM window -> Hann -> zero-pad -> C2C FFT -> L/R split -> 2/N scale
 *     -> log-frequency resample (cubic | cosine) -> dB -> gradient -> RGBA column
 * following the reference operation by operation, in the reference's own mix of
 * f32 / f64 arithmetic.  Citations are relative to /root/reference.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the timed CPU baseline.
 * The product (spectrogram-rs_b200/csrc) never includes, links or calls it.
 *
 * PARITY STATUS: "parity unpinned" by the reference -- the reference ships no tests,
 * golden vectors or fixtures, and neither Rust nor FFTW
 * exist in this image, so the reference cannot be run here.  The oracle is pinned
 * instead by analytic known-answer tests (tests/test_oracle_kat.py), an independent
 * f64 numpy FFT referee, and committed golden vectors generated from this file
 * (tests/golden/, tests/golden/make_golden.py).
 *
 * Third-party arithmetic that lives in crates absent from /root/reference is
 * isolated in single functions marked [unvendored]:
 *   fftw 0.8.0 (FFTW3f)        -> orc_fft_c2c_f32   (any correct f32 FFT; FFTW's result is
 *                                                    plan-dependent to ~1e-7 of peak)
 *   iter_num_tools 0.7.1       -> orc_lin_space_at
 *   plotters 0.3.5             -> orc_ranged_unmap
 *   colorous 1.0.12            -> orc_gradient_index
 *   num-complex 0.4.4 / num-traits pow -> orc_norm, orc_pow2/orc_pow3
 *
 * Build (see oracle/Makefile):  gcc -O2 -ffp-contract=off -fno-fast-math -pthread -shared -fPIC
 * -ffp-contract=off is REQUIRED: Rust never fuses a*b+c, GNU C does by default.
 * liboracle.so (-O2, host-independent) is the checker; liboracle_fast.so (-O3 -march=native, same source, same
 * results -- tests/test_oracle_kat.py compares the two) is what bench.py TIMES as the CPU baseline.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define ORC_API __attribute__((visibility("default")))

typedef struct { float re, im; } c32;

enum { ORC_INTERP_CUBIC = 0, ORC_INTERP_COSINE = 1 };

/* ------------------------------------------------------------------------- */
/* [unvendored] FFT: stand-in for fftw C2CPlan32 forward, unnormalised        */
/* call site: src/fourier/fft.rs:20-24 (plan), :77 (execute)                  */
/* Iterative radix-2 DIT in f32 with twiddles computed in f64 and rounded once */
/* ------------------------------------------------------------------------- */
static int orc_is_pow2(uint32_t m) { return m && !(m & (m - 1)); }

/* O(M^2) f64 DFT (any length): used for non-power-of-two lengths and as referee */
ORC_API void orc_dft_c2c_f64(const double *in_re, const double *in_im,
                             double *out_re, double *out_im, uint32_t m)
{
    for (uint32_t k = 0; k < m; ++k) {
        double sr = 0.0, si = 0.0;
        for (uint32_t n = 0; n < m; ++n) {
            uint64_t kn = ((uint64_t)k * n) % m;
            double a = -2.0 * M_PI * (double)kn / (double)m;
            double c = cos(a), s = sin(a);
            sr += in_re[n] * c - in_im[n] * s;
            si += in_re[n] * s + in_im[n] * c;
        }
        out_re[k] = sr; out_im[k] = si;
    }
}

/* per-size tables (twiddles computed in f64, rounded once; bit-reversal permutation),
 * built once and cached so the timed CPU baseline does not pay for them per frame */
typedef struct { c32 *tw; uint32_t *rev; } orc_fft_tables;
static orc_fft_tables orc_tables[33];

static pthread_mutex_t orc_tables_lock = PTHREAD_MUTEX_INITIALIZER;

static const orc_fft_tables *orc_get_tables(uint32_t m, uint32_t lg)
{
    orc_fft_tables *t = &orc_tables[lg];
    c32 *ready = __atomic_load_n(&t->tw, __ATOMIC_ACQUIRE);
    if (ready) return t;
    pthread_mutex_lock(&orc_tables_lock);
    if (!t->tw) {
        c32 *tw = (c32 *)malloc(sizeof(c32) * (m / 2 + 1));
        uint32_t *rev = (uint32_t *)malloc(sizeof(uint32_t) * m);
        if (tw && rev) {
            for (uint32_t k = 0; k < m / 2; ++k) {
                double a = -2.0 * M_PI * (double)k / (double)m;
                tw[k].re = (float)cos(a); tw[k].im = (float)sin(a);
            }
            for (uint32_t i = 0; i < m; ++i) {
                uint32_t r = 0;
                for (uint32_t b = 0; b < lg; ++b) r |= ((i >> b) & 1u) << (lg - 1 - b);
                rev[i] = r;
            }
            t->rev = rev;
            __atomic_store_n(&t->tw, tw, __ATOMIC_RELEASE);
        }
    }
    pthread_mutex_unlock(&orc_tables_lock);
    return t->tw ? t : NULL;
}

ORC_API int orc_fft_c2c_f32(const c32 *in, c32 *out, uint32_t m)
{
    if (!orc_is_pow2(m)) {
        /* mixed sizes (the live app's 4800/4410): f64 DFT rounded to f32 */
        double *b = (double *)malloc(sizeof(double) * 4 * (size_t)m);
        if (!b) return -1;
        for (uint32_t i = 0; i < m; ++i) { b[i] = in[i].re; b[m + i] = in[i].im; }
        orc_dft_c2c_f64(b, b + m, b + 2 * (size_t)m, b + 3 * (size_t)m, m);
        for (uint32_t i = 0; i < m; ++i) { out[i].re = (float)b[2 * (size_t)m + i]; out[i].im = (float)b[3 * (size_t)m + i]; }
        free(b);
        return 0;
    }
    if (m == 1) { out[0] = in[0]; return 0; }
    uint32_t lg = 0; while ((1u << lg) < m) ++lg;
    const orc_fft_tables *t = orc_get_tables(m, lg);
    if (!t) return -1;
    const c32 *tw = t->tw;
    for (uint32_t i = 0; i < m; ++i) out[t->rev[i]] = in[i];
    for (uint32_t len = 2; len <= m; len <<= 1) {
        uint32_t half = len >> 1, stride = m / len;
        for (uint32_t base = 0; base < m; base += len) {
            for (uint32_t j = 0; j < half; ++j) {
                c32 w = tw[j * stride];
                c32 u = out[base + j], v = out[base + j + half];
                float tr = v.re * w.re - v.im * w.im;
                float ti = v.re * w.im + v.im * w.re;
                out[base + j].re = u.re + tr;        out[base + j].im = u.im + ti;
                out[base + j + half].re = u.re - tr; out[base + j + half].im = u.im - ti;
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* window + pad + FFT + stereo split: FastFourierTransform::process           */
/* src/fourier/fft.rs:43-99                                                   */
/* ------------------------------------------------------------------------- */

/* src/fourier/fft.rs:59-63  periodic Hann, every operation in f32 */
ORC_API float orc_hann(uint32_t i, uint32_t n)
{
    const float tau = 6.28318530717958647692f;          /* f32::TAU */
    return 0.5f * (1.0f - cosf((tau * (float)i) / (float)n));
}

/* [unvendored] num-complex norm(): re.hypot(im) in f32 (src/fourier/fft.rs:87-88) */
static float orc_norm(float re, float im) { return hypotf(re, im); }

ORC_API uint32_t orc_num_output_frequencies(uint32_t n, uint32_t pad)
{
    /* src/fourier/fft.rs:33  N-1 for the reference's hard-coded pad factor 2.  Generalised
     * as "every bin strictly between DC and Nyquist of the padded transform" = M/2-1, which
     * is N-1 at P=2 and keeps period()=2K/sr (ifs.rs:53) one-bin-consistent for P=4. */
    return (n * pad) / 2u - 1u;
}

/* fft.rs:48-69  take N, pack (l + i r), Hann, zero-pad to M = pad*N.  z has M entries. */
ORC_API void orc_window_pad(const float *lr, uint32_t n, uint32_t pad, c32 *z)
{
    uint32_t m = n * pad;
    for (uint32_t i = 0; i < n; ++i) {                          /* fft.rs:57-63 */
        float w = orc_hann(i, n);
        z[i].re = lr[2 * i] * w;  z[i].im = lr[2 * i + 1] * w;   /* c32 * f32 = componentwise */
    }
    for (uint32_t i = n; i < m; ++i) { z[i].re = 0.0f; z[i].im = 0.0f; }  /* fft.rs:65 */
}

/* fft.rs:81-98  pair Z[k] with Z[M-k], split into L/R magnitudes, scale by 2/N */
ORC_API void orc_split_scale(const c32 *zf, uint32_t n, uint32_t pad, float *out_lr)
{
    uint32_t m = n * pad, k_out = orc_num_output_frequencies(n, pad);
    float scale = 2.0f / (float)n;                              /* fft.rs:92 */
    for (uint32_t j = 0; j < k_out; ++j) {                      /* fft.rs:81-95 */
        c32 a = zf[j + 1], b = zf[m - 1 - j];
        float sl = orc_norm(a.re + b.re, a.im - b.im) / 2.0f;   /* |a + conj b| / 2 */
        float sr = orc_norm(a.re - b.re, a.im + b.im) / 2.0f;   /* |a - conj b| / 2 */
        out_lr[2 * j] = sl * scale;  out_lr[2 * j + 1] = sr * scale;
    }
}

/*
 * One frame.  lr = N interleaved (l,r) samples.  out = K interleaved (L,R) magnitudes.
 * Returns 0, or 1 when n_avail < N (the reference's `None`, fft.rs:72).
 */
ORC_API int orc_frame_magnitudes(const float *lr, uint32_t n_avail, uint32_t n, uint32_t pad, float *out_lr)
{
    if (n_avail < n) return 1;                                  /* fft.rs:72 */
    uint32_t m = n * pad;
    c32 *z = (c32 *)malloc((size_t)m * 2 * sizeof(c32));        /* fft.rs:68,76: two buffers per frame */
    if (!z) return -1;
    c32 *zf = z + m;
    orc_window_pad(lr, n, pad, z);
    if (orc_fft_c2c_f32(z, zf, m) != 0) { free(z); return -1; } /* fft.rs:77 */
    orc_split_scale(zf, n, pad, out_lr);
    free(z);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* row -> Hz map: LogCoordf64::unmap  (src/log_scaling.rs:114-119,160-191)    */
/* ------------------------------------------------------------------------- */

/* [unvendored] plotters 0.3.5 RangedCoordf64::unmap: (end-start)*((p-min)/(max-min)) + start */
static double orc_ranged_unmap(double start, double end, int32_t p, int32_t lo, int32_t hi)
{
    double off = (double)(p - lo) / (double)(hi - lo);
    return (end - start) * off + start;
}

/* edge frequency of pixel row py in [0,H]; f64 then cast f32 (simple_spectrogram.rs:142-145) */
ORC_API float orc_row_edge(double f_min, double f_max, uint32_t py, uint32_t rows)
{
    /* log_scaling.rs:160-191: zero_point 0, range positive => linear = ln(f_min)..ln(f_max) */
    double lin = orc_ranged_unmap(log(f_min), log(f_max), (int32_t)py, 0, (int32_t)rows);
    double v = exp(lin) + 0.0;                                   /* f64_to_value: fv + zero_point */
    return (float)v;
}

/* ------------------------------------------------------------------------- */
/* InterpolatedFrequencySample (src/fourier/interpolated_frequency_sample.rs) */
/* ------------------------------------------------------------------------- */

/* ifs.rs:52-54  period = 2.0 * len as f32 / sr.0 as f32, sr truncated to u32 (simple_spectrogram.rs:138) */
ORC_API float orc_period(uint32_t k_out, float sample_rate)
{
    uint32_t sr = (uint32_t)sample_rate;
    return (2.0f * (float)k_out) / (float)sr;
}

/* ifs.rs:24-31 */
static float orc_index_of(float f, float period, uint32_t k_out)
{
    float idx = f * period, hi = (float)(k_out - 1);
    if (idx < 0.0f) idx = 0.0f;
    if (idx > hi) idx = hi;
    return idx;
}

/* [unvendored] iter_num_tools 0.7.1 lin_space over a half-open Range: start + i*((end-start)/n) */
static float orc_lin_space_at(float f0, float step, uint32_t i) { return f0 + (float)i * step; }

/* [unvendored] num_traits::pow(mu,2) / pow(mu,3) by repeated multiplication */
static float orc_pow2(float x) { return x * x; }
static float orc_pow3(float x) { float b = x * x; return x * b; }

/* ifs.rs:88-105.  x0 = (floor(x) as usize - 1).max(0) underflows (panics) in the reference
 * when floor(x)==0; the oracle (and the product) saturate at 0, the evident intent. */
static c32 orc_cubic(const c32 *d, uint32_t len, float x)
{
    float fl = floorf(x), mu = x - fl;
    uint32_t x1 = (uint32_t)fl;
    uint32_t x0 = x1 > 0 ? x1 - 1 : 0;
    uint32_t x2 = x1 + 1 < len - 1 ? x1 + 1 : len - 1;
    uint32_t x3 = x1 + 2 < len - 1 ? x1 + 2 : len - 1;
    c32 y0 = d[x0], y1 = d[x1], y2 = d[x2], y3 = d[x3], o;
    float mu2 = orc_pow2(mu), mu3 = orc_pow3(mu);
#define CUBIC1(F) do { \
        float a0 = ((y3.F - y2.F) - y0.F) + y1.F; \
        float a1 = (y0.F - y1.F) - a0; \
        float a2 = y2.F - y0.F; \
        float a3 = y1.F; \
        o.F = ((a0 * mu3) + (a1 * mu2)) + ((a2 * mu) + a3); } while (0)
    CUBIC1(re); CUBIC1(im);
#undef CUBIC1
    return o;
}

/* ifs.rs:78-86 (dead code in the reference; README/north-star behaviour) */
static c32 orc_cosine(const c32 *d, uint32_t len, float x)
{
    const float pi = 3.14159265358979323846f;
    uint32_t low = (uint32_t)floorf(x);
    uint32_t high = (uint32_t)ceilf(x);
    uint32_t lo_b = low + 1, hi_b = len - 1;
    /* usize::clamp(min,max) panics if min>max (low == len-1); saturate instead */
    if (lo_b > hi_b) lo_b = hi_b;
    if (high < lo_b) high = lo_b;
    if (high > hi_b) high = hi_b;
    float off = x - (float)low;
    off = (1.0f - cosf(off * pi)) / 2.0f;
    c32 o;
    o.re = (d[low].re * (1.0f - off)) + (d[high].re * off);
    o.im = (d[low].im * (1.0f - off)) + (d[high].im * off);
    return o;
}

/* One row of the bin map: the sampling plan of magnitude_in(f0..f1)  (ifs.rs:60-67) */
ORC_API void orc_bin_map_row(double f_min, double f_max, uint32_t py, uint32_t rows,
                             float period, uint32_t k_out,
                             float *f0_out, float *step_out, uint32_t *n_out)
{
    float f0 = orc_row_edge(f_min, f_max, py, rows);
    float f1 = orc_row_edge(f_min, f_max, py + 1, rows);
    float i0 = orc_index_of(f0, period, k_out), i1 = orc_index_of(f1, period, k_out);
    float d = floorf(i1 - i0);
    uint32_t n = d >= 1.0f ? (uint32_t)d : 1u;                   /* (floor as usize).max(1); as-cast saturates <0 -> 0 */
    *f0_out = f0; *step_out = (f1 - f0) / (float)n; *n_out = n;
}

/* ifs.rs:60-75 */
static c32 orc_magnitude_in(const c32 *d, uint32_t len, float period, float f0, float step, uint32_t n, int interp)
{
    c32 acc = {0.0f, 0.0f};
    for (uint32_t s = 0; s < n; ++s) {
        float fq = orc_lin_space_at(f0, step, s);
        float x = orc_index_of(fq, period, len);
        c32 v = interp == ORC_INTERP_COSINE ? orc_cosine(d, len, x) : orc_cubic(d, len, x);
        acc.re += v.re; acc.im += v.im;
    }
    acc.re /= (float)n; acc.im /= (float)n;
    return acc;
}

/* ------------------------------------------------------------------------- */
/* ColorScheme::color_for, mono schemes (src/colorscheme.rs:16-17,55-71)      */
/* ------------------------------------------------------------------------- */

/* colorscheme.rs:60-61 on a power value p = l*l + r*r */
ORC_API float orc_db_chain(float p, float db_min, float db_max, float db_eps)
{
    float db = 10.0f * log10f(p + db_eps);
    return (db - db_min) / (db_max - db_min);
}

/* colorscheme.rs:59-61 */
ORC_API float orc_normalised_db(float l, float r, float db_min, float db_max, float db_eps)
{
    float p = l * l + r * r;                                     /* norm_sqr */
    return orc_db_chain(p, db_min, db_max, db_eps);
}

/* [unvendored] colorous 1.0.12 256-entry ramp gradients: idx = clamp(floor(t*256),0,255) */
ORC_API uint32_t orc_gradient_index(float t)
{
    double v = floor((double)t * 256.0);
    if (!(v >= 0.0)) return 0;                                   /* negatives and NaN */
    if (v > 255.0) return 255;
    return (uint32_t)v;
}

ORC_API uint32_t orc_color_index(float l, float r, float db_min, float db_max, float db_eps)
{
    return orc_gradient_index(orc_normalised_db(l, r, db_min, db_max, db_eps));
}

/* Stereo schemes (colorscheme.rs:63-66): colour = gradient(l/(|l|+|r|)) evaluated in f64,
 * alpha = normalised dB; put_pixel alpha = (alpha*255.0) as u8 (simple_spectrogram.rs:159). */
ORC_API void orc_stereo_color_levels(float l, float r, float db_min, float db_max, float db_eps, uint32_t levels,
                                     uint32_t *lut_index, uint8_t *alpha);
ORC_API void orc_stereo_color(float l, float r, float db_min, float db_max, float db_eps,
                              uint32_t *lut_index, uint8_t *alpha)
{
    orc_stereo_color_levels(l, r, db_min, db_max, db_eps, 256u, lut_index, alpha);
}
/* `levels`: steps of the table standing in for the continuous diverging gradient (256 = the lut_rgb table; more
 * with the product's sbx_plan_set_pan_gradient): index = clamp(floor(pan * levels), 0, levels - 1) [unvendored] */
ORC_API void orc_stereo_color_levels(float l, float r, float db_min, float db_max, float db_eps, uint32_t levels,
                                     uint32_t *lut_index, uint8_t *alpha)
{
    float t = orc_normalised_db(l, r, db_min, db_max, db_eps);
    float l1 = fabsf(l) + fabsf(r);
    double pan = (double)l / (double)l1;
    double v = floor(pan * (double)levels);
    *lut_index = !(v >= 0.0) ? 0u : (v > (double)(levels - 1u) ? levels - 1u : (uint32_t)v);
    float a = t * 255.0f;                                        /* Rust `as u8` saturates, NaN -> 0 */
    *alpha = !(a >= 0.0f) ? 0 : (a >= 255.0f ? 255 : (uint8_t)a);
}

/* ------------------------------------------------------------------------- */
/* column loop: SimpleSpectrogram::snapshot (simple_spectrogram.rs:136-164)   */
/* ------------------------------------------------------------------------- */
typedef struct {
    float sample_rate;
    uint32_t window_n, pad_factor, hop, rows_h;
    double f_min, f_max;
    float db_min, db_max, db_eps;
    uint32_t interp;      /* ORC_INTERP_* */
    uint32_t stereo_scheme; /* 0 mono gradient (alpha 255), 1 stereo pan gradient */
    uint32_t pan_levels;    /* stereo schemes: entries of lut_rgb, indexed by pan (0 = 256) */
} orc_params;

/* magnitudes (K x (L,R)) -> RGBA column of H rows; row 0 = top = highest frequency */
ORC_API void orc_column_from_magnitudes(const orc_params *p, const float *mags_lr, const uint8_t *lut_rgb, uint8_t *rgba)
{
    uint32_t k_out = orc_num_output_frequencies(p->window_n, p->pad_factor);
    float period = orc_period(k_out, p->sample_rate);
    const c32 *d = (const c32 *)mags_lr;                         /* ifs.rs:19 */
    for (uint32_t py = 0; py < p->rows_h; ++py) {                /* simple_spectrogram.rs:141 */
        float f0, step; uint32_t n;
        orc_bin_map_row(p->f_min, p->f_max, py, p->rows_h, period, k_out, &f0, &step, &n);
        c32 m = orc_magnitude_in(d, k_out, period, f0, step, n, (int)p->interp);
        uint32_t row = p->rows_h - py - 1;                       /* :150 */
        uint32_t idx; uint8_t alpha = 255;
        if (p->stereo_scheme) orc_stereo_color_levels(m.re, m.im, p->db_min, p->db_max, p->db_eps,
                                                      p->pan_levels ? p->pan_levels : 256u, &idx, &alpha);
        else idx = orc_color_index(m.re, m.im, p->db_min, p->db_max, p->db_eps);
        rgba[4 * row + 0] = lut_rgb[3 * idx + 0];
        rgba[4 * row + 1] = lut_rgb[3 * idx + 1];
        rgba[4 * row + 2] = lut_rgb[3 * idx + 2];
        rgba[4 * row + 3] = alpha;
    }
}

/* resampled (l,r) per row before colouring, py order (row py = low->high frequency) */
ORC_API void orc_rows_from_magnitudes(const orc_params *p, const float *mags_lr, float *rows_lr)
{
    uint32_t k_out = orc_num_output_frequencies(p->window_n, p->pad_factor);
    float period = orc_period(k_out, p->sample_rate);
    for (uint32_t py = 0; py < p->rows_h; ++py) {
        float f0, step; uint32_t n;
        orc_bin_map_row(p->f_min, p->f_max, py, p->rows_h, period, k_out, &f0, &step, &n);
        c32 m = orc_magnitude_in((const c32 *)mags_lr, k_out, period, f0, step, n, (int)p->interp);
        rows_lr[2 * py] = m.re; rows_lr[2 * py + 1] = m.im;
    }
}

ORC_API int orc_column(const orc_params *p, const float *lr, uint32_t n_avail, const uint8_t *lut_rgb, uint8_t *rgba)
{
    uint32_t k_out = orc_num_output_frequencies(p->window_n, p->pad_factor);
    float *m = (float *)malloc(sizeof(float) * 2 * (size_t)k_out);
    if (!m) return -1;
    int st = orc_frame_magnitudes(lr, n_avail, p->window_n, p->pad_factor, m);
    if (st == 0) orc_column_from_magnitudes(p, m, lut_rgb, rgba);
    free(m);
    return st;
}

/* ------------------------------------------------------------------------- */
/* framing: AudioStreamTransform::process (src/fourier/audio_transform.rs:34-42)
 * frame f covers samples [f*hop, f*hop+N); frames while >= N samples remain.  */
/* ------------------------------------------------------------------------- */
ORC_API uint64_t orc_num_frames(uint64_t n_samples, uint32_t n, uint32_t hop)
{
    if (n_samples < n || hop == 0) return 0;
    return (n_samples - n) / hop + 1;
}

/* hop in samples from a stride in seconds: (stride * sr) as usize, f32 multiply (audio_transform.rs:35) */
ORC_API uint32_t orc_hop_from_stride(float stride_s, float sample_rate)
{
    float h = stride_s * sample_rate;
    return h >= 0.0f ? (uint32_t)h : 0u;
}

/* window length from a period in seconds: (period * sr) as usize (fft.rs:19,41) */
ORC_API uint32_t orc_window_from_period(float period_s, float sample_rate)
{
    float w = period_s * sample_rate;
    return w >= 0.0f ? (uint32_t)w : 0u;
}

/*
 * Offline batch: pcm holds n_samples frames of `channels` (1: mono duplicated to (s,s) as
 * src/devices/audio_input_list_model.rs:66-69 does; 2: interleaved l,r).  Writes
 * frames x H x 4 RGBA (frame-major, each frame one column, row 0 = top) and/or
 * frames x K x 2 magnitudes.  Frames [frame_lo, frame_hi) only (for the sharding tests).
 * threads<=1 -> serial (what the reference does); else frames split over pthreads.
 */
typedef struct {
    const orc_params *p; const float *pcm; uint32_t channels;
    uint64_t frame_lo, lo, hi; const uint8_t *lut; uint8_t *rgba; float *mags; int err;
} orc_job;

static void *orc_batch_worker(void *arg)
{
    orc_job *j = (orc_job *)arg;
    const orc_params *p = j->p;
    uint32_t n = p->window_n, k_out = orc_num_output_frequencies(n, p->pad_factor);
    float *lr = (float *)malloc(sizeof(float) * 2 * (size_t)n);
    float *m = (float *)malloc(sizeof(float) * 2 * (size_t)k_out);
    if (!lr || !m) { j->err = 1; free(lr); free(m); return NULL; }
    for (uint64_t f = j->lo; f < j->hi; ++f) {
        const float *src = j->pcm + (size_t)f * p->hop * j->channels;
        if (j->channels == 1) for (uint32_t i = 0; i < n; ++i) { lr[2 * i] = src[i]; lr[2 * i + 1] = src[i]; }
        else memcpy(lr, src, sizeof(float) * 2 * (size_t)n);
        if (orc_frame_magnitudes(lr, n, n, p->pad_factor, m) != 0) { j->err = 1; break; }
        size_t o = (size_t)(f - j->frame_lo);
        if (j->mags) memcpy(j->mags + o * 2 * k_out, m, sizeof(float) * 2 * (size_t)k_out);
        if (j->rgba) orc_column_from_magnitudes(p, m, j->lut, j->rgba + o * 4 * p->rows_h);
    }
    free(lr); free(m);
    return NULL;
}

ORC_API int64_t orc_batch(const orc_params *p, const float *pcm, uint64_t n_samples, uint32_t channels,
                          uint64_t frame_lo, uint64_t frame_hi, const uint8_t *lut_rgb,
                          uint8_t *rgba, float *mags_lr, int threads)
{
    uint64_t total = orc_num_frames(n_samples, p->window_n, p->hop);
    if (frame_hi > total) frame_hi = total;
    if (frame_lo >= frame_hi) return 0;
    uint64_t count = frame_hi - frame_lo;
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > count) threads = (int)count;
    orc_job *jobs = (orc_job *)calloc((size_t)threads, sizeof(orc_job));
    pthread_t *tid = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    if (!jobs || !tid) { free(jobs); free(tid); return -1; }
    for (int t = 0; t < threads; ++t) {
        orc_job *j = &jobs[t];
        j->p = p; j->pcm = pcm; j->channels = channels; j->frame_lo = frame_lo;
        j->lo = frame_lo + count * (uint64_t)t / (uint64_t)threads;
        j->hi = frame_lo + count * (uint64_t)(t + 1) / (uint64_t)threads;
        j->lut = lut_rgb; j->rgba = rgba; j->mags = mags_lr; j->err = 0;
    }
    int err = 0;
    if (threads == 1) orc_batch_worker(&jobs[0]);
    else {
        char *started = (char *)calloc((size_t)threads, 1);
        for (int t = 0; t < threads; ++t) {
            if (started && pthread_create(&tid[t], NULL, orc_batch_worker, &jobs[t]) == 0) started[t] = 1;
            else orc_batch_worker(&jobs[t]);                    /* could not spawn: run inline */
        }
        for (int t = 0; t < threads; ++t) if (started && started[t]) pthread_join(tid[t], NULL);
        free(started);
    }
    for (int t = 0; t < threads; ++t) err |= jobs[t].err;
    free(jobs); free(tid);
    return err ? -1 : (int64_t)count;
}

/* resample + colour stage alone over a batch of spectra (threads over frames): lets the timed CPU baseline
 * use a faster FFT library for the FFT stage (pocketfft via scipy; FFTW itself is not in this image) */
typedef struct { const orc_params *p; const float *mags; const uint8_t *lut; uint8_t *rgba; uint64_t lo, hi; } orc_cjob;
static void *orc_colour_worker(void *arg)
{
    orc_cjob *j = (orc_cjob *)arg;
    uint32_t k_out = orc_num_output_frequencies(j->p->window_n, j->p->pad_factor);
    for (uint64_t f = j->lo; f < j->hi; ++f)
        orc_column_from_magnitudes(j->p, j->mags + (size_t)f * 2 * k_out, j->lut, j->rgba + (size_t)f * 4 * j->p->rows_h);
    return NULL;
}
ORC_API int orc_colour_batch(const orc_params *p, const float *mags_lr, uint64_t n_frames, const uint8_t *lut_rgb,
                             uint8_t *rgba, int threads)
{
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n_frames) threads = (int)(n_frames ? n_frames : 1);
    orc_cjob *jobs = (orc_cjob *)calloc((size_t)threads, sizeof(orc_cjob));
    pthread_t *tid = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    char *started = (char *)calloc((size_t)threads, 1);
    if (!jobs || !tid || !started) { free(jobs); free(tid); free(started); return -1; }
    for (int t = 0; t < threads; ++t) {
        jobs[t].p = p; jobs[t].mags = mags_lr; jobs[t].lut = lut_rgb; jobs[t].rgba = rgba;
        jobs[t].lo = n_frames * (uint64_t)t / (uint64_t)threads;
        jobs[t].hi = n_frames * (uint64_t)(t + 1) / (uint64_t)threads;
        if (threads > 1 && pthread_create(&tid[t], NULL, orc_colour_worker, &jobs[t]) == 0) started[t] = 1;
        else orc_colour_worker(&jobs[t]);
    }
    for (int t = 0; t < threads; ++t) if (started[t]) pthread_join(tid[t], NULL);
    free(jobs); free(tid); free(started);
    return 0;
}

/* The FFT stage handed to a library (timed CPU baseline only): everything of fft.rs:48-69 before the transform
 * (take N, pack, Hann -- recomputed per frame as the reference does --, zero-pad) for a run of frames into
 * z[frames][M], and everything after it (fft.rs:81-98 split + scale, then the column) from zf[frames][M].
 * bench.py runs scipy.fft (pocketfft, complex64) between the two: FFTW itself is not in this image. */
typedef struct { const orc_params *p; const float *pcm; uint32_t channels; uint64_t frame_lo, lo, hi;
                 c32 *z; const uint8_t *lut; uint8_t *rgba; int err; } orc_sjob;
static void *orc_window_worker(void *arg)
{
    orc_sjob *j = (orc_sjob *)arg;
    uint32_t n = j->p->window_n, m = n * j->p->pad_factor;
    float *lr = (float *)malloc(sizeof(float) * 2 * (size_t)n);
    if (!lr) { j->err = 1; return NULL; }
    for (uint64_t f = j->lo; f < j->hi; ++f) {
        const float *src = j->pcm + (size_t)(j->frame_lo + f) * j->p->hop * j->channels;
        if (j->channels == 1) for (uint32_t i = 0; i < n; ++i) { lr[2 * i] = src[i]; lr[2 * i + 1] = src[i]; }
        else memcpy(lr, src, sizeof(float) * 2 * (size_t)n);
        orc_window_pad(lr, n, j->p->pad_factor, j->z + (size_t)f * m);
    }
    free(lr);
    return NULL;
}
static void *orc_split_colour_worker(void *arg)
{
    orc_sjob *j = (orc_sjob *)arg;
    uint32_t n = j->p->window_n, m = n * j->p->pad_factor, k_out = orc_num_output_frequencies(n, j->p->pad_factor);
    float *mags = (float *)malloc(sizeof(float) * 2 * (size_t)k_out);
    if (!mags) { j->err = 1; return NULL; }
    for (uint64_t f = j->lo; f < j->hi; ++f) {
        orc_split_scale(j->z + (size_t)f * m, n, j->p->pad_factor, mags);
        orc_column_from_magnitudes(j->p, mags, j->lut, j->rgba + (size_t)f * 4 * j->p->rows_h);
    }
    free(mags);
    return NULL;
}
static int orc_run_sjobs(orc_sjob proto, uint64_t n_frames, int threads, void *(*worker)(void *))
{
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n_frames) threads = (int)(n_frames ? n_frames : 1);
    orc_sjob *jobs = (orc_sjob *)calloc((size_t)threads, sizeof(orc_sjob));
    pthread_t *tid = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    char *started = (char *)calloc((size_t)threads, 1);
    if (!jobs || !tid || !started) { free(jobs); free(tid); free(started); return -1; }
    for (int t = 0; t < threads; ++t) {
        jobs[t] = proto;
        jobs[t].lo = n_frames * (uint64_t)t / (uint64_t)threads;
        jobs[t].hi = n_frames * (uint64_t)(t + 1) / (uint64_t)threads;
        if (threads > 1 && pthread_create(&tid[t], NULL, worker, &jobs[t]) == 0) started[t] = 1;
        else worker(&jobs[t]);
    }
    int err = 0;
    for (int t = 0; t < threads; ++t) { if (started[t]) pthread_join(tid[t], NULL); err |= jobs[t].err; }
    free(jobs); free(tid); free(started);
    return err ? -1 : 0;
}
ORC_API int orc_window_pad_batch(const orc_params *p, const float *pcm, uint32_t channels, uint64_t frame_lo,
                                 uint64_t n_frames, float *z, int threads)
{
    orc_sjob proto; memset(&proto, 0, sizeof proto);
    proto.p = p; proto.pcm = pcm; proto.channels = channels; proto.frame_lo = frame_lo; proto.z = (c32 *)z;
    return orc_run_sjobs(proto, n_frames, threads, orc_window_worker);
}
ORC_API int orc_split_colour_batch(const orc_params *p, const float *zf, uint64_t n_frames, const uint8_t *lut_rgb,
                                   uint8_t *rgba, int threads)
{
    orc_sjob proto; memset(&proto, 0, sizeof proto);
    proto.p = p; proto.z = (c32 *)zf; proto.lut = lut_rgb; proto.rgba = rgba;
    return orc_run_sjobs(proto, n_frames, threads, orc_split_colour_worker);
}

ORC_API int orc_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
