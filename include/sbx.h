/*
 * sbx.h -- C ABI of the B200 (sm_100a) spectrogram frame transform.
 *
 * Drop-in boundary for spectrogram-rs's frame path.  Every entry point names the reference
 * interface it replaces (paths relative to the reference repository root).  Plain pointers and
 * sizes only; no C++/torch types; never throws or aborts across the boundary -- every failure is
 * an sbx_status plus a message retrievable with sbx_last_error().
 *
 * Symbols:  N = window length in samples, P = pad factor, M = P*N = FFT length,
 *           K = M/2 - 1 output bins (N-1 at the reference's P = 2), H = output rows.
 *
 * Threading: a plan is NOT thread-safe (one plan <-> one CUDA stream <-> one caller at a time,
 * like the reference's `&mut self`); distinct plans are independent.
 */
#ifndef SBX_H
#define SBX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SBX_API __declspec(dllexport)
#else
#define SBX_API __attribute__((visibility("default")))
#endif

#define SBX_ABI_VERSION 1u

typedef struct sbx_plan sbx_plan;   /* opaque; owns device tables, scratch and a CUDA stream */
typedef struct sbx_live sbx_live;   /* opaque; pinned ring + CUDA graph for one live stream */
typedef struct sbx_multi sbx_multi; /* opaque; one plan per device of one box, for sharded host-pointer batches */
typedef int32_t sbx_status;

enum {
    SBX_OK = 0,
    SBX_NOT_ENOUGH_SAMPLES = 1,   /* the reference's `None` (src/fourier/fft.rs:72) */
    SBX_ERR_INVALID_ARGUMENT = -1,
    SBX_ERR_UNSUPPORTED = -2,     /* configuration outside what the kernels cover */
    SBX_ERR_CUDA = -3,            /* CUDA runtime / driver failure, or no usable GPU */
    SBX_ERR_OUT_OF_MEMORY = -4,
    SBX_ERR_BUFFER_TOO_SMALL = -5
};

enum { SBX_INTERP_CUBIC = 0,      /* cubic_interpolate, the interpolator the reference wires in
                                     (src/fourier/interpolated_frequency_sample.rs:46-48,88-105) */
       SBX_INTERP_COSINE = 1 };   /* cosine_interpolate (ibid. :78-86; README behaviour) */

enum { SBX_SCHEME_MONO = 0,       /* ColorScheme::new_mono: colour = gradient(dB), alpha = 255
                                     (src/colorscheme.rs:67-70) */
       SBX_SCHEME_STEREO = 1 };   /* ColorScheme::new_stereo: colour = gradient(l/(|l|+|r|)),
                                     alpha = normalised dB (src/colorscheme.rs:63-66).  The gradient is read as
                                     lut_rgb[clamp(floor(pan * 256))]: a 256-step quantisation of the reference's
                                     continuous stereo gradients (see lut_rgb below), or from the finer table of
                                     sbx_plan_set_pan_gradient() */

/*
 * Everything the reference hard-codes as literals (SURVEY.md section 5 "Config"), as one struct.
 * Fill with sbx_config_default() first, then override.
 */
typedef struct sbx_config {
    uint32_t struct_size;   /* = sizeof(sbx_config); ABI guard */
    uint32_t abi_version;   /* = SBX_ABI_VERSION */
    float    sample_rate;   /* Hz.  FastFourierTransform::new(sample_rate, ..) src/fourier/fft.rs:18 */
    uint32_t window_n;      /* N = (period * sample_rate) as usize, src/fourier/fft.rs:19,41 */
    uint32_t pad_factor;    /* P; the reference hard-codes 2 (src/fourier/fft.rs:21,44) */
    uint32_t hop;           /* (stride * sample_rate) as usize, src/fourier/audio_transform.rs:35 */
    uint32_t rows_h;        /* H; TEXTURE_HEIGHT = 1024, src/widgets/simple_spectrogram.rs:35 */
    uint32_t channels;      /* PCM layout of the BATCH/LIVE entry points: 1 = mono (treated as the
                               (s,s) duplication of src/devices/audio_input_list_model.rs:66-69),
                               2 = interleaved (l,r).  The single-frame entry points always take
                               interleaved (l,r) like AudioTransform::process. */
    double   f_min, f_max;  /* 32.0 .. 22030.0, src/widgets/simple_spectrogram.rs:107 */
    float    db_min, db_max;/* -70, -10, src/colorscheme.rs:16-17 */
    float    db_eps;        /* 1e-7, src/colorscheme.rs:60 */
    uint32_t interp;        /* SBX_INTERP_* */
    uint32_t scheme;        /* SBX_SCHEME_* */
    int32_t  device;        /* CUDA device ordinal */
    uint32_t reserved;
    uint8_t  lut_rgb[768];  /* 256 x (r,g,b): gradient.eval_continuous((i+0.5)/256), i = 0..255
                               (src/colorscheme.rs:69).  See sbx_builtin_lut().
                               EXACT only for colorous' 256-entry ramps (MAGMA, VIRIDIS, INFERNO, PLASMA), where
                               eval_continuous(t) itself is table[clamp(floor(t*256))].  Every other gradient the
                               reference lists (the stereo schemes RED_YELLOW_BLUE -- the app default --, RED_BLUE,
                               SPECTRAL, ...; CIVIDIS, TURBO, REDS, ...) is a continuous spline/polynomial in
                               colorous: the library evaluates it at 256 steps, so an r/g/b value can differ from
                               eval_continuous(t) by the gradient's change over half a step (a few LSB where the
                               gradient is steep).  Alpha (stereo schemes) and the step index are exact either way. */
} sbx_config;

/* Reference defaults: 48 kHz, N 4096, P 2, hop 512, H 1024, 32..22030 Hz, -70..-10 dB, eps 1e-7,
 * CUBIC, mono scheme, Magma, stereo PCM, device 0. */
SBX_API void sbx_config_default(sbx_config *cfg);

/* Built-in gradients: all 19 of default_color_schemes() (src/colorscheme.rs:129-149), by lower-case name of the
 * colorous constant: "red_yellow_blue", "magma", "viridis", "red_blue", "spectral", "red_yellow_green", "pink_green",
 * "purple_orange", "inferno", "plasma", "cividis", "cubehelix", "turbo", "cool", "reds", "blues", "greens", "greys",
 * "oranges".  sbx_builtin_lut fills the 256-entry table of sbx_config::lut_rgb: table[i] = gradient((i + 0.5) / 256) --
 * for the four ramps (magma, viridis, inferno, plasma) exactly colorous' table.  The definitions restate
 * d3-scale-chromatic (colorous is its port and is not vendored with the reference; csrc/sbx_gradients.h); a Rust
 * caller fills the tables from colorous itself.  SBX_ERR_INVALID_ARGUMENT for other names. */
SBX_API sbx_status sbx_builtin_lut(const char *name, uint8_t lut_rgb[768]);
/* The same gradient as a table of `entries` steps (1 .. 65536): rgb[3 i ..] = gradient((i + 0.5) / entries). */
SBX_API sbx_status sbx_builtin_gradient(const char *name, uint32_t entries, uint8_t *rgb);
/* Enumeration in the reference's list order: name of gradient `index` < sbx_builtin_gradient_count() (NULL beyond),
 * *is_stereo_scheme (may be NULL) = 1 where the reference builds it with ColorScheme::new_stereo. */
SBX_API uint32_t    sbx_builtin_gradient_count(void);
SBX_API const char *sbx_builtin_gradient_name(uint32_t index, int32_t *is_stereo_scheme);

/* (period * sample_rate) as usize -- src/fourier/fft.rs:19,41 -- and
 * (stride * sample_rate) as usize -- src/fourier/audio_transform.rs:35 (f32 multiply, truncation). */
SBX_API uint32_t sbx_samples_from_seconds(float seconds, float sample_rate);

/* ---- plan lifetime: replaces FastFourierTransform::new (src/fourier/fft.rs:18-31), which
 * re-plans FFTW on every sample-rate change (src/widgets/simple_spectrogram.rs:214-219). ----
 * Supported geometry (anything else: SBX_ERR_UNSUPPORTED, never a silent fallback):
 *   - N a power of two, 256 .. 65536, P in {1, 2, 4}, M = P*N <= 262144 (M > 16384 needs N >= 8192);
 *   - N = 2^a 3^b 5^c 7^d otherwise (the application's 2400 / 2205 / 4800 / 9600, the widgets' N = 100),
 *     P in {1, 2, 4} with M even and 12*M bytes within one SM's shared memory (M <= 19370);
 *   - any other N (a prime factor > 7: 1102 = 2*19*29 at 22.05 kHz, 551 at 11.025 kHz, primes) by Bluestein's
 *     chirp-z over a power-of-two convolution, as long as N*(P+1) - 1 <= 16384 (N <= 5461 at P = 2), M even;
 *   - any hop >= 1, H in [1, 2^20], channels 1 or 2, both interpolators, both scheme kinds.
 * Plan creation computes every table on the host (window, twiddles, bin map, per-sample positions and weights,
 * dB-step thresholds) and uploads them; the kernels never evaluate exp / log / cos. */
SBX_API sbx_status sbx_plan_create(const sbx_config *cfg, sbx_plan **out_plan);
SBX_API void       sbx_plan_destroy(sbx_plan *plan);

/* Stereo schemes (SBX_SCHEME_STEREO) colour a pixel with gradient.eval_continuous(pan), pan = l / (|l| + |r|)
 * (src/colorscheme.rs:63-66), and every stereo gradient of the reference is a CONTINUOUS spline (the diverging
 * ColorBrewer schemes), which the 256 steps of lut_rgb only approximate.  This installs a finer table for the pan
 * lookup: rgb = entries x (r, g, b), entry i = gradient((i + 0.5) / entries), read as
 * rgb[clamp(floor(pan * entries), 0, entries - 1)] (pan evaluated in f64 as the reference does).  With entries = 4096
 * a channel differs from eval_continuous(pan) by at most 1 (tests/test_oracle_kat.py states the measured maximum per
 * gradient and table size).  Alpha (the normalised dB) is unaffected.  Call before sbx_live_create: a live stream
 * captures the table that is installed when it is created.  The table is copied. */
SBX_API sbx_status sbx_plan_set_pan_gradient(sbx_plan *plan, const uint8_t *rgb, uint32_t entries);

/* Message for the last failing call on this plan (plan may be NULL: last error of a failed
 * sbx_plan_create on this thread).  Never NULL. */
SBX_API const char *sbx_last_error(const sbx_plan *plan);

/* AudioTransform::sample_rate / num_input_samples (src/fourier/audio_transform.rs:7-8),
 * FastFourierTransform::num_output_frequencies (src/fourier/fft.rs:33). */
SBX_API float    sbx_sample_rate(const sbx_plan *plan);
SBX_API uint32_t sbx_num_input_samples(const sbx_plan *plan);       /* N */
SBX_API uint32_t sbx_num_output_frequencies(const sbx_plan *plan);  /* K */
SBX_API uint32_t sbx_fft_length(const sbx_plan *plan);              /* M */
SBX_API uint32_t sbx_rows(const sbx_plan *plan);                    /* H */
SBX_API uint32_t sbx_hop(const sbx_plan *plan);
SBX_API uint32_t sbx_channels(const sbx_plan *plan);

/* Frames the sliding window yields over n_samples samples: frame f covers [f*hop, f*hop+N)
 * (src/fourier/audio_transform.rs:37-41). */
SBX_API uint64_t sbx_num_frames(const sbx_plan *plan, uint64_t n_samples);

/* ---- single frame, HOST buffers ---------------------------------------------------------- */

/* Exact drop-in for FastFourierTransform::process (src/fourier/fft.rs:43-99).
 * lr: n_avail interleaved (l,r) samples (only the first N are used); out_lr: K interleaved (L,R).
 * Returns SBX_NOT_ENOUGH_SAMPLES when n_avail < N (the reference's None). */
SBX_API sbx_status sbx_process_magnitudes(sbx_plan *plan, const float *lr, uint32_t n_avail, float *out_lr);

/* Replaces the body of SimpleSpectrogram::snapshot's frame loop
 * (src/widgets/simple_spectrogram.rs:136-161): one frame -> one RGBA8 column of H rows,
 * rgba[4*row], row 0 = top = highest frequency (the `height - py - 1` flip at :150). */
SBX_API sbx_status sbx_process_column(sbx_plan *plan, const float *lr, uint32_t n_avail, uint8_t *rgba);

/* InterpolatedFrequencySample::magnitude_in + ColorScheme::color_for on caller-supplied spectra
 * (src/fourier/interpolated_frequency_sample.rs:60-75, src/colorscheme.rs:55-71):
 * mags_lr = n_frames x K x (L,R)  ->  rgba = n_frames x H x 4.  This is the "identical
 * normalised inputs" entry point of the bit-exactness contract. */
SBX_API sbx_status sbx_colorize_magnitudes(sbx_plan *plan, const float *mags_lr, uint64_t n_frames, uint8_t *rgba);

/* ---- offline batch ------------------------------------------------------------------------ */

/* AudioStreamTransform::process (src/fourier/audio_transform.rs:34-42) + the column loop over a
 * whole recording.  pcm: n_samples x channels f32 (HOST memory, pinned or pageable);
 * rgba: frames x H x 4, frame-major (frame f's column at rgba + f*H*4).  Host<->device copies are
 * chunked and overlapped inside the call.  *frames_out = frames written. */
SBX_API sbx_status sbx_process_batch(sbx_plan *plan, const float *pcm, uint64_t n_samples,
                                     uint8_t *rgba, uint64_t rgba_capacity_frames, uint64_t *frames_out);

/* Same, DEVICE pointers, asynchronous on the plan's stream (sbx_synchronize() to wait).
 * d_pcm must be 16-byte aligned.
 * Mono plans (channels = 1) transform frames two at a time -- frame 2c of the call as the real part and frame 2c+1
 * as the imaginary part of one complex sequence, separated afterwards by the reference's own conjugate-symmetry split
 * (src/fourier/fft.rs:81-89).  A frame's magnitudes therefore depend, in the last bit (about 1e-7 of the pair's
 * peak), on which slot it takes and on its partner: two calls return byte-identical columns for a frame only if it is
 * paired identically, i.e. if d_pcm of the second call starts an EVEN number of frames (a multiple of 2*hop samples)
 * after the first.  sharding.time_blocks / sbx_process_batch chunk on even frame counts for this reason.  (The
 * real-sequence geometry, sbx_kernel_path() == "fused-real", has no such dependence.) */
SBX_API sbx_status sbx_process_batch_device(sbx_plan *plan, const float *d_pcm, uint64_t n_samples,
                                            uint8_t *d_rgba, uint64_t rgba_capacity_frames, uint64_t *frames_out);

/* Batched FastFourierTransform::process: DEVICE pointers, d_mags_lr = frames x K x (L,R).
 * (Also the feed format of GPUSpectrogram's texture upload, src/widgets/gpu_spectrogram.rs:255-275.) */
SBX_API sbx_status sbx_magnitudes_batch_device(sbx_plan *plan, const float *d_pcm, uint64_t n_samples,
                                               float *d_mags_lr, uint64_t capacity_frames, uint64_t *frames_out);

/* ---- feed formats of the GL widget (GPUSpectrogram) -------------------------------------------- */

/* Several equally long recordings in one call (BASELINE cfg4: 64 planar channels; the reference runs one
 * AudioStreamTransform per stream, src/fourier/audio_transform.rs:14-42): sequence s starts
 * s * sequence_stride_samples samples into d_pcm (each sample = `channels` floats) and its frames land at
 * d_rgba + s * frames * H * 4, frames = sbx_num_frames(n_samples_per_sequence), returned through
 * *frames_per_sequence_out.  rgba_capacity_frames counts all sequences.  The shared-memory-resident kernels take
 * all sequences in ONE grid (no per-sequence tail); the M > 16384 and mixed-radix paths launch per sequence.
 * DEVICE pointers, asynchronous on the plan's stream. */
SBX_API sbx_status sbx_process_batch_device_planar(sbx_plan *plan, const float *d_pcm, uint32_t n_sequences,
                                                   uint64_t sequence_stride_samples, uint64_t n_samples_per_sequence,
                                                   uint8_t *d_rgba, uint64_t rgba_capacity_frames,
                                                   uint64_t *frames_per_sequence_out);

/* Batched FastFourierTransform::process with the rows packed as IEEE binary16 pairs: the texel format of
 * GPUSpectrogram's F16F16 ring texture (src/widgets/gpu_spectrogram.rs:218-227,255-275), so the rows can be
 * uploaded without a CPU-side Vec<(f32,f32)>.  d_mags_f16 = frames x K x (L,R) uint16 (binary16 bits,
 * round-to-nearest-even).  DEVICE pointers, asynchronous on the plan's stream. */
SBX_API sbx_status sbx_magnitudes_batch_device_f16(sbx_plan *plan, const float *d_pcm, uint64_t n_samples,
                                                   uint16_t *d_mags_f16, uint64_t capacity_frames, uint64_t *frames_out);

/* ColorScheme::lookup_table(resolution) (src/colorscheme.rs:73-92): the resolution x resolution x (r,g,b,a) f32
 * palette GPUSpectrogram uploads as a texture (gpu_spectrogram.rs:232-239).  table[i][j]: magnitude =
 * i/(res-1), pan = 1 - j/(res-1); mono schemes: colour = gradient(magnitude)/256, alpha 1; stereo schemes:
 * colour = gradient(pan)/256, alpha = magnitude.  Host memory, out has resolution*resolution*4 floats. */
SBX_API sbx_status sbx_palette_table(const sbx_plan *plan, uint32_t resolution, float *out_rgba);

/* The column store of SimpleSpectrogram (src/widgets/simple_spectrogram.rs:89-94 pixbuf, :141-164 column loop):
 * a width x rows RGBA8 image, row-major with rowstride_bytes per row (the GdkPixbuf layout, in which a column is
 * strided).  Column f of d_columns (n_frames x rows x 4, row 0 = top, as the batch entry points write them) lands
 * in image column (first_column + f) mod width -- `self.offset = (self.offset + 1) % width`, :164 -- so the store
 * scrolls by wrapping around; when n_frames > width only the last `width` columns are drawn.  *next_column_out
 * (may be NULL) receives (first_column + n_frames) mod width, the cursor for the next call.  DEVICE pointers,
 * asynchronous on the plan's stream; a tiled transpose, both sides coalesced.  rowstride_bytes must be a multiple
 * of 4 and >= 4 * width, d_image 4-byte aligned. */
SBX_API sbx_status sbx_blit_columns_device(sbx_plan *plan, const uint8_t *d_columns, uint64_t n_frames, uint8_t *d_image,
                                           uint32_t width, uint32_t rowstride_bytes, uint32_t first_column,
                                           uint32_t *next_column_out);

/* ---- streams ------------------------------------------------------------------------------ */
/* Use the caller's CUDA stream (a cudaStream_t passed as void*; NULL restores the plan's own). */
SBX_API sbx_status sbx_set_stream(sbx_plan *plan, void *cuda_stream);
SBX_API sbx_status sbx_synchronize(sbx_plan *plan);
/* Kernel launches issued by this plan since creation (bench.py's gpu_launches evidence). */
SBX_API uint64_t   sbx_launch_count(const sbx_plan *plan);
/* Which kernel family the plan's batch entry points run for its channel layout (evidence for bench.py and the
 * tests; static string): "fused-pair" (two mono frames per CTA), "fused-real" (one mono frame per CTA as the
 * even/odd packed half-length sequence; needs an even hop, else "fused-pair" / "large" runs), "fused-stereo",
 * "fused-pair-q2" / "fused-stereo-q2" (pad factor 2: both sub-transforms of a unit in the two lanes of packed FP32
 * instructions, stft_q2.cuh), "large" (four kernels, M > 16384) or "generic" (window not a power of two). */
SBX_API const char *sbx_kernel_path(const sbx_plan *plan);

/* ---- introspection for the bit-exactness tests ------------------------------------------- */
/* The log-frequency bin map the plan computed on the host: per row py (0 = lowest frequency)
 * the sampling plan of magnitude_in(f0..f1): first sample frequency f0, step (f1-f0)/n, count n
 * (src/log_scaling.rs:114-119, src/fourier/interpolated_frequency_sample.rs:60-67).
 * Any output pointer may be NULL.  Arrays have H entries. */
SBX_API sbx_status sbx_get_bin_map(const sbx_plan *plan, float *f0, float *step, uint32_t *n, float *period);
/* 255 power thresholds: LUT index of power p is #{i : p >= thr[i]} -- the exact step function
 * of color_for's dB/normalise/floor(t*256) chain (src/colorscheme.rs:59-61,69). */
SBX_API sbx_status sbx_get_color_thresholds(const sbx_plan *plan, float thr[255]);
/* The f32 periodic Hann window the plan uploaded (src/fourier/fft.rs:59-63); N entries. */
SBX_API sbx_status sbx_get_window(const sbx_plan *plan, float *w);

/* ---- live path: the SPSC ring + hop cursor of L1/L2 --------------------------------------- */
/* Pinned-host ring of (l,r) samples (capacity as HeapRb::new(4096),
 * src/devices/audio_input_list_model.rs:30) with AudioStreamTransform's cursor semantics
 * (src/fourier/audio_transform.rs:34-42) and one captured CUDA graph per frame: a single kernel node that reads
 * the pinned window and writes the pinned column itself and signals completion through a pinned flag the consumer
 * polls (M <= 16384 and the mixed-radix path), or H2D copy -> kernels -> D2H copy + stream synchronize (M > 16384,
 * SBX_LIVE_COPY=1).  One producer thread (push) and one consumer thread (pop) may run concurrently. */
/* The live stream borrows the plan: destroy it (sbx_live_destroy) before sbx_plan_destroy(plan). */
SBX_API sbx_status sbx_live_create(sbx_plan *plan, uint32_t ring_capacity_samples, sbx_live **out_live);
SBX_API void       sbx_live_destroy(sbx_live *live);
/* Producer side (the cpal callback, src/devices/audio_input_list_model.rs:66-72): pushes up to
 * n_samples frames of `channels` floats (mono is duplicated to (s,s)); returns how many fit. */
SBX_API uint32_t   sbx_live_push(sbx_live *live, const float *samples, uint32_t n_samples, uint32_t channels);
/* Consumer side: one iteration of AudioStreamTransform::process's closure -- transform the first
 * N buffered samples without consuming them, then skip(hop) (also when the frame failed, as the
 * reference does).  SBX_OK + one RGBA column, or SBX_NOT_ENOUGH_SAMPLES. */
SBX_API sbx_status sbx_live_pop_column(sbx_live *live, uint8_t *rgba);
SBX_API sbx_status sbx_live_pop_magnitudes(sbx_live *live, float *out_lr);
SBX_API uint32_t   sbx_live_buffered(const sbx_live *live);
/* Measurement helper for the live path (no reference counterpart): `iters` times, push `n_samples` frames of
 * `channels` floats from `samples` and pop one column, timing each push+pop pair on the host with
 * CLOCK_MONOTONIC; latencies_us[i] receives the i-th latency in microseconds.  Returns SBX_OK, or the first
 * failing status (SBX_NOT_ENOUGH_SAMPLES if a pop found fewer than N buffered samples). */
SBX_API sbx_status sbx_live_measure(sbx_live *live, const float *samples, uint32_t n_samples, uint32_t channels,
                                    uint32_t iters, float *latencies_us);

/* ---- several devices: one HOST recording sharded over the GPUs of one box --------------------- */
/* The reference transforms one stream on one thread (src/fourier/audio_transform.rs:34-42); frame f depends only on
 * samples [f*hop, f*hop+N) (ibid. :37-41, src/fourier/fft.rs:48-50), so an offline recording shards without any
 * exchange between devices: by time block -- device g of G transforms frames [F*g/G, F*(g+1)/G) and therefore reads
 * its own samples plus a trailing halo of N - hop samples it shares with its neighbour -- or, for planar
 * multi-channel input, by whole channels.  An sbx_multi owns one plan per device (identical tables) and drives each
 * device from its own host thread with the chunk pipeline of sbx_process_batch; there is no collective and no
 * device-to-device traffic.  Block boundaries of mono plans are rounded down to even frame numbers (frames 2c, 2c+1
 * share one transform, see sbx_process_batch_device), so the sharded result equals the one-device result byte for
 * byte.  An sbx_multi is not thread-safe (one caller at a time). */
/* devices: n_devices distinct CUDA ordinals, or NULL for 0 .. n_devices-1.  cfg->device is ignored. */
SBX_API sbx_status sbx_multi_create(const sbx_config *cfg, const int32_t *devices, uint32_t n_devices, sbx_multi **out_multi);
SBX_API void       sbx_multi_destroy(sbx_multi *multi);
/* Message of the last failing sbx_multi_* call (multi may be NULL: a failed sbx_multi_create on this thread). */
SBX_API const char *sbx_multi_last_error(const sbx_multi *multi);
SBX_API uint32_t   sbx_multi_device_count(const sbx_multi *multi);
/* The plan of device `index` (owned by the sbx_multi; for the accessors, sbx_kernel_path, the device entry points). */
SBX_API sbx_plan  *sbx_multi_plan(sbx_multi *multi, uint32_t index);
/* The time block of device `index` for a recording of n_samples: frames [frame_lo, frame_hi), samples
 * [sample_lo, sample_hi) including the halo.  Any output pointer may be NULL. */
SBX_API sbx_status sbx_multi_time_block(const sbx_multi *multi, uint64_t n_samples, uint32_t index, uint64_t *frame_lo,
                                        uint64_t *frame_hi, uint64_t *sample_lo, uint64_t *sample_hi);
/* The same arithmetic without a handle or a GPU, for callers that run one process per device (MPI, torchrun) and hand
 * shard `index` of `n_shards` its own block: frames [F*index/n, F*(index+1)/n) of the F frames of the recording,
 * boundaries rounded down to even frames when channels == 1 (mono plans pair frames 2c, 2c+1), samples
 * [frame_lo*hop, (frame_hi-1)*hop + N).  Any output pointer may be NULL. */
SBX_API sbx_status sbx_shard_time_block(uint64_t n_samples, uint32_t window_n, uint32_t hop, uint32_t channels, uint32_t index,
                                        uint32_t n_shards, uint64_t *frame_lo, uint64_t *frame_hi, uint64_t *sample_lo,
                                        uint64_t *sample_hi);
/* sbx_process_batch over all devices, time-block sharded.  pcm / rgba: HOST memory (pinned -- sbx_host_alloc,
 * sbx_host_register -- for copies that overlap; pageable works, slowly). */
SBX_API sbx_status sbx_multi_process_batch(sbx_multi *multi, const float *pcm, uint64_t n_samples, uint8_t *rgba,
                                           uint64_t rgba_capacity_frames, uint64_t *frames_out);
/* n_sequences equally long recordings (BASELINE cfg4: planar channels), whole sequences per device; arguments as
 * sbx_process_batch_device_planar but HOST pointers: sequence s starts s * sequence_stride_samples samples into pcm,
 * its columns land at rgba + s * frames * H * 4. */
SBX_API sbx_status sbx_multi_process_batch_planar(sbx_multi *multi, const float *pcm, uint32_t n_sequences,
                                                  uint64_t sequence_stride_samples, uint64_t n_samples_per_sequence,
                                                  uint8_t *rgba, uint64_t rgba_capacity_frames,
                                                  uint64_t *frames_per_sequence_out);
/* Kernel launches of all its plans (see sbx_launch_count). */
SBX_API uint64_t   sbx_multi_launch_count(const sbx_multi *multi);

/* ---- pinned host memory helpers (for callers without a CUDA binding) --------------------- */
SBX_API void *sbx_host_alloc(size_t bytes);
SBX_API void  sbx_host_free(void *p);
/* Page-locks / releases memory the caller already owns (a Rust Vec<f32>, an mmap), usable from every device. */
SBX_API sbx_status sbx_host_register(void *ptr, size_t bytes);
SBX_API sbx_status sbx_host_unregister(void *ptr);

/* Library/ABI version and the SM architecture the kernels were compiled for ("sm_100a"). */
SBX_API uint32_t    sbx_abi_version(void);
SBX_API const char *sbx_build_arch(void);

#ifdef __cplusplus
}
#endif
#endif /* SBX_H */
