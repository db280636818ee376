// sbx_plan.cu -- plan construction (host-computed tables), path selection, the colourise / pack / blit kernels and the C ABI of
// include/sbx.h.  No CPU compute fallback exists: every transform entry point launches the sm_100a
// kernels of stft_kernels.cuh or fails with SBX_ERR_CUDA / SBX_ERR_UNSUPPORTED.
#include <cuda_fp16.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>

#include "luts.inc"
#include "sbx_gradients.h"
#include "sbx_launch.h"
#include "sbx_tables.h"

namespace {

thread_local std::string g_create_error;

int ilog2_exact(uint32_t v)
{
    if (v == 0 || (v & (v - 1)) != 0) return -1;
    int l = 0;
    while ((1u << l) < v) ++l;
    return l;
}

template <class T> cudaError_t upload(T **dst, const void *src, size_t bytes)
{
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(dst), bytes ? bytes : 16);
    if (e != cudaSuccess) return e;
    return bytes ? cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) : cudaSuccess;
}

}  // namespace

namespace sbx {

sbx_status fail(const sbx_plan *p, sbx_status st, const std::string &msg)
{
    if (p) p->last_error = msg; else g_create_error = msg;
    return st;
}
sbx_status cuda_fail(const sbx_plan *p, cudaError_t e, const char *what)
{
    return fail(p, SBX_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

bool real_path_wanted(const sbx_plan *p, int channels, const float *d_pcm, Sequences seqs)
{
    return channels == 1 && p->use_real && (p->hop & 1u) == 0 && (reinterpret_cast<uintptr_t>(d_pcm) & 7u) == 0 &&
           (seqs.stride_floats & 1u) == 0;
}

// SBX_REAL=0 / 1 overrides the plan-time default (tables::HostTables::real_auto): 1 = wherever the tables exist.
bool choose_real(bool real_ok, bool real_auto)
{
    const char *env = std::getenv("SBX_REAL");
    if (env && env[0] == '0') return false;
    if (env && env[0] == '1') return real_ok;
    return real_auto;
}

// Plan-time choice between fused_frames_kernel and fused_q2_kernel (stft_q2.cuh) where both exist, per channel layout.
// SBX_Q2=0 / 1 overrides the measured default (read when the plan is created).
static bool choose_q2(const sbx_plan *p, int channels)
{
    if (!q2_available(p)) return false;
    const char *env = std::getenv("SBX_Q2");
    if (env && env[0] == '0') return false;
    if (env && env[0] == '1') return true;
    // measured on a B200, M frames/s (fused_frames_kernel -> fused_q2_kernel): N = 4096 mono 58.6 -> 63.8, stereo 29.7 -> 32.7;
    // N = 2048 stereo 56.5 -> 57.6; N = 8192 stereo (one 512-thread CTA per SM either way) 12.6 -> 12.5
    // mono N = 2048 (hop 256, H 2048): 85.1 -> 82.6, so the 128-thread CTAs of N = 2048 are used for stereo PCM only
    return p->log2n == 12 || (p->log2n == 11 && channels == 2);
}

static bool choose_q2_real(const sbx_plan *p)
{
    if (!q2_real_available(p)) return false;
    const char *env = std::getenv("SBX_Q2");
    if (env && env[0] == '0') return false;
    if (env && env[0] == '1') return true;
    // measured on a B200, M frames/s (classic real-sequence kernel -> q2): N = 8192 (cfg3) 29.9 -> 33.0; N = 4096 44.5 -> 42.4;
    // N = 16384 (one 512-thread CTA per SM either way) 7.95 -> 7.96
    return p->log2n == 13;
}

sbx_status launch_fused(sbx_plan *p, int channels, const float *d_pcm, uint64_t n_frames, uint8_t *d_rgba,
                        float *d_mags, cudaStream_t stream, Sequences seqs)
{
    if (real_path_wanted(p, channels, d_pcm, seqs)) return launch_real(p, d_pcm, n_frames, d_rgba, d_mags, stream, seqs);
    if (p->generic || p->large) {
        // the mixed-radix and multi-kernel paths take one sequence per launch
        for (uint32_t s = 0; s < seqs.count; ++s) {
            const float *pcm = d_pcm + (size_t)s * seqs.stride_floats;
            uint8_t *rgba = d_rgba ? d_rgba + (size_t)s * n_frames * p->rows * 4 : nullptr;
            float *mags = d_mags ? d_mags + (size_t)s * n_frames * p->k_out * 2 : nullptr;
            const sbx_status st = p->generic ? launch_generic(p, channels, pcm, n_frames, rgba, mags, stream)
                                             : launch_large(p, channels, pcm, n_frames, rgba, mags, stream);
            if (st != SBX_OK) return st;
        }
        return SBX_OK;
    }
    KernelArgs a = p->base_args();
    a.pcm = d_pcm;
    a.rgba = d_rgba;
    a.mags_out = d_mags;
    a.n_frames = n_frames;
    if (n_frames == 1 && seqs.count == 1) { a.done_flag = p->live_done_flag; a.done_count = p->live_done_count; }
    // a lone stereo frame (single-frame entry points, live graphs) is latency-bound: WIDE geometry; SBX_WIDE=0 disables
    static const bool no_wide = [] { const char *e = std::getenv("SBX_WIDE"); return e && e[0] == '0'; }();
    const bool wide = n_frames == 1 && seqs.count == 1 && channels == 2 && !no_wide && wide_available(p);
    // pad factor 2, N = 2048 .. 8192: both sub-transforms in packed FP32 lanes (stft_q2.cuh)
    const bool q2 = !wide && p->use_q2[channels - 1];
    const cudaError_t e = wide ? launch_fused_wide(p, a, stream)
                        : q2 ? launch_fused_q2(p, channels, a, n_frames, stream, seqs)
                               : (p->log2n <= 10 ? launch_fused_small(p, channels, a, n_frames, stream, seqs)
                                                 : launch_fused_big(p, channels, a, n_frames, stream, seqs));
    if (e != cudaSuccess) return cuda_fail(p, e, "fused_frames_kernel launch");
    if (n_frames) p->launches += 1;
    return SBX_OK;
}

template <int LOG2M> cudaError_t launch_colourise_one(const KernelArgs &a, uint64_t n_frames, cudaStream_t stream)
{
    static std::atomic<bool> configured[kMaxDevices];
    const int smem = 8 << LOG2M;
    cudaError_t e = allow_smem(colourise_kernel<LOG2M>, smem, configured);
    if (e != cudaSuccess) return e;
    if (n_frames == 0) return cudaSuccess;
    if (n_frames > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    colourise_kernel<LOG2M><<<(unsigned)n_frames, 256, smem, stream>>>(a);
    return cudaGetLastError();
}

sbx_status launch_colourise(sbx_plan *p, const float *d_mags, uint64_t n_frames, uint8_t *d_rgba, cudaStream_t stream)
{
    if (p->generic) {
        if (n_frames == 0) return SBX_OK;
        if (n_frames > 0x7fffffffull) return fail(p, SBX_ERR_INVALID_ARGUMENT, "too many frames");
        KernelArgs a = p->base_args();
        a.mags_in = d_mags;
        a.rgba = d_rgba;
        a.n_frames = n_frames;
        const uint32_t slots = p->m / 2;
        const size_t smem = (size_t)slots * sizeof(float2);
        static std::atomic<bool> configured[kMaxDevices];
        cudaError_t eg = allow_smem(generic_colour_kernel, kMaxSmemOptIn, configured);
        if (eg == cudaSuccess) { generic_colour_kernel<<<(unsigned)n_frames, GENERIC_THREADS, smem, stream>>>(a, slots); eg = cudaGetLastError(); }
        if (eg != cudaSuccess) return cuda_fail(p, eg, "generic_colour_kernel launch");
        p->launches += 1;
        return SBX_OK;
    }
    if (p->large) {
        if (n_frames == 0) return SBX_OK;
        if (n_frames > 0x7fffffffull) return fail(p, SBX_ERR_INVALID_ARGUMENT, "too many frames");
        LargeArgs la{};
        la.k = p->base_args();
        la.k.rgba = d_rgba;
        la.k.n_frames = n_frames;
        la.units = (uint32_t)n_frames;
        // magnitudes are indexed by bin kk = j + 1: shift the caller's K x (L,R) rows by one entry
        if (la.k.n_chunks > 65535u) return fail(p, SBX_ERR_UNSUPPORTED, "too many row chunks");
        large_colour_kernel<false><<<dim3((unsigned)n_frames, la.k.n_chunks), LARGE_THREADS, 0, stream>>>(
            la, reinterpret_cast<const float2 *>(d_mags) - 1, (size_t)p->k_out);
        cudaError_t el = cudaGetLastError();
        if (el != cudaSuccess) return cuda_fail(p, el, "large_colour_kernel launch");
        p->launches += 1;
        return SBX_OK;
    }
    KernelArgs a = p->base_args();
    a.mags_in = d_mags;
    a.rgba = d_rgba;
    a.n_frames = n_frames;
    cudaError_t e;
    switch (p->log2n + p->log2pad) {
        case 8: e = launch_colourise_one<8>(a, n_frames, stream); break;
        case 9: e = launch_colourise_one<9>(a, n_frames, stream); break;
        case 10: e = launch_colourise_one<10>(a, n_frames, stream); break;
        case 11: e = launch_colourise_one<11>(a, n_frames, stream); break;
        case 12: e = launch_colourise_one<12>(a, n_frames, stream); break;
        case 13: e = launch_colourise_one<13>(a, n_frames, stream); break;
        case 14: e = launch_colourise_one<14>(a, n_frames, stream); break;
        default: e = cudaErrorInvalidConfiguration; break;
    }
    if (e != cudaSuccess) return cuda_fail(p, e, "colourise_kernel launch");
    if (n_frames) p->launches += 1;
    return SBX_OK;
}

}  // namespace sbx

namespace sbx {
// f32 (L,R) rows -> binary16 pairs (round-to-nearest-even), the texel format of the GL widget's F16F16 texture
__global__ void __launch_bounds__(256) pack_f16_kernel(const float2 *__restrict__ in, __half2 *__restrict__ out, size_t count)
{
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < count; i += (size_t)gridDim.x * 256)
        out[i] = __float22half2_rn(in[i]);
}
}  // namespace sbx

namespace sbx {
// columns (frames x rows, one packed pixel each) -> row-major image: a 32 x 32 tiled transpose through shared memory.
// Reads run along a column (consecutive rows), writes along an image row (consecutive frames); the image column
// wraps modulo `width`.  frames <= width (the host drops columns that would be overwritten).
__global__ void __launch_bounds__(256) blit_columns_kernel(const uint32_t *__restrict__ cols, uint32_t frames, uint32_t rows,
                                                            uint32_t *__restrict__ image, uint32_t width, uint32_t stride_px,
                                                            uint32_t first_column)
{
    __shared__ uint32_t tile[32][33];
    const uint32_t f0 = blockIdx.x * 32u, r0 = blockIdx.y * 32u;
    const uint32_t tx = threadIdx.x & 31u, ty = threadIdx.x >> 5;
#pragma unroll
    for (uint32_t i = ty; i < 32u; i += 8u) {
        const uint32_t f = f0 + i, r = r0 + tx;
        if (f < frames && r < rows) tile[i][tx] = cols[(size_t)f * rows + r];
    }
    __syncthreads();
#pragma unroll
    for (uint32_t i = ty; i < 32u; i += 8u) {
        const uint32_t f = f0 + tx, r = r0 + i;
        uint32_t x = first_column + f;                      // both < width
        if (x >= width) x -= width;
        if (f < frames && r < rows) image[(size_t)r * stride_px + x] = tile[tx][i];
    }
}
}  // namespace sbx

sbx::KernelArgs sbx_plan::base_args() const
{
    sbx::KernelArgs a{};
    a.window = d_window;
    a.wtw = d_wtw;
    a.twp = d_twp;
    a.wq = d_wq;
    a.wtw4 = d_wtw4;
    a.rtw = d_rtw;
    a.rowmap = d_rowmap;
    a.rowinfo = d_rowinfo;
    a.chunks = d_chunks;
    a.n_chunks = n_chunks;
    a.chunk0 = chunk0;
    a.samp = d_samp;
    a.taps_clamped = taps_clamped;
    a.thr = d_thr;
    a.lut = d_lut;
    a.ctab = d_ctab;
    a.pan_lut = d_pan_lut ? d_pan_lut : d_lut;
    a.pan_levels = d_pan_lut ? pan_levels : 256u;
    a.hop = hop;
    a.rows = rows;
    a.k_out = k_out;
    a.scale = scale;
    a.db_eps = cfg.db_eps;
    a.neg_zero = -0.0f;
    a.step_margin = step_margin;
    for (int i = 0; i < 2; ++i) { a.guess_a[i] = guess_a[i]; a.guess_b[i] = guess_b[i]; }
    a.interp = (int)cfg.interp;
    a.scheme = (int)cfg.scheme;
    return a;
}

using sbx::cuda_fail;
using sbx::fail;
using sbx::kMaxLog2M;
using sbx::kMaxLog2MLarge;
using sbx::kMaxLog2NLarge;
using sbx::kMinLog2N;

#define SBX_CUDA(plan, call)                                              \
    do {                                                                  \
        cudaError_t e__ = (call);                                         \
        if (e__ != cudaSuccess) return cuda_fail((plan), e__, #call);     \
    } while (0)

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

SBX_API uint32_t sbx_abi_version(void) { return SBX_ABI_VERSION; }
SBX_API const char *sbx_build_arch(void) { return "sm_100a"; }

SBX_API sbx_status sbx_builtin_gradient(const char *name, uint32_t entries, uint8_t *rgb)
{
    if (!name || !rgb || entries == 0 || entries > 65536u) return SBX_ERR_INVALID_ARGUMENT;
    const sbx::gradients::Def *d = sbx::gradients::find(name);
    if (!d) return SBX_ERR_INVALID_ARGUMENT;
    sbx::gradients::table(*d, entries, rgb);
    return SBX_OK;
}

SBX_API sbx_status sbx_builtin_lut(const char *name, uint8_t lut_rgb[768])
{
    return sbx_builtin_gradient(name, 256, lut_rgb);
}

SBX_API uint32_t sbx_builtin_gradient_count(void) { return 19; }

SBX_API const char *sbx_builtin_gradient_name(uint32_t index, int32_t *is_stereo_scheme)
{
    if (index >= 19) return nullptr;
    if (is_stereo_scheme) *is_stereo_scheme = sbx::gradients::kDefs[index].stereo ? 1 : 0;
    return sbx::gradients::kDefs[index].name;
}

SBX_API void sbx_config_default(sbx_config *cfg)
{
    if (!cfg) return;
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = (uint32_t)sizeof(sbx_config);
    cfg->abi_version = SBX_ABI_VERSION;
    cfg->sample_rate = 48000.0f;
    cfg->window_n = 4096;
    cfg->pad_factor = 2;
    cfg->hop = 512;
    cfg->rows_h = 1024;
    cfg->channels = 2;
    cfg->f_min = 32.0;
    cfg->f_max = 22030.0;
    cfg->db_min = -70.0f;
    cfg->db_max = -10.0f;
    cfg->db_eps = 1e-7f;
    cfg->interp = SBX_INTERP_CUBIC;
    cfg->scheme = SBX_SCHEME_MONO;
    cfg->device = 0;
    std::memcpy(cfg->lut_rgb, SBX_LUT_MAGMA, 768);
}

SBX_API uint32_t sbx_samples_from_seconds(float seconds, float sample_rate)
{
    const float v = seconds * sample_rate;           // f32 multiply, then `as usize` (saturating)
    if (!(v >= 0.0f)) return 0;
    return v >= 4294967296.0f ? 0xffffffffu : (uint32_t)v;
}

SBX_API const char *sbx_last_error(const sbx_plan *plan)
{
    return plan ? plan->last_error.c_str() : g_create_error.c_str();
}

SBX_API void sbx_plan_destroy(sbx_plan *p)
{
    if (!p) return;
    cudaSetDevice(p->cfg.device);
    if (p->own_stream) cudaStreamSynchronize(p->own_stream);
    for (auto &s : p->slots) {
        if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
        cudaFree(s.d_pcm);
        cudaFree(s.d_rgba);
    }
    cudaFree(p->d_window); cudaFree(p->d_wtw); cudaFree(p->d_twp); cudaFree(p->d_wq); cudaFree(p->d_wq_real); cudaFree(p->d_rowmap); cudaFree(p->d_rowinfo); cudaFree(p->d_chunks);
    cudaFree(p->d_samp); cudaFree(p->d_f16_scratch); cudaFree(p->d_twm); cudaFree(p->d_bl_a); cudaFree(p->d_bl_bf); cudaFree(p->d_bl_c); cudaFree(p->d_twg); cudaFree(p->d_pos_tab); cudaFree(p->d_twp_block); cudaFree(p->d_zbuf); cudaFree(p->d_magbuf);
    cudaFree(p->d_thr); cudaFree(p->d_lut); cudaFree(p->d_pan_lut); cudaFree(p->d_ctab); cudaFree(p->d_twp_real); cudaFree(p->d_wtw4); cudaFree(p->d_rtw); cudaFree(p->d_cz_in); cudaFree(p->d_cz_out); cudaFree(p->d_frame); cudaFree(p->d_mags1); cudaFree(p->d_col);
    for (uint32_t *old_table : p->retired_pan_luts) cudaFree(old_table);
    if (p->h_frame) cudaFreeHost(p->h_frame);
    if (p->h_col) cudaFreeHost(p->h_col);
    if (p->large_scratch_free) cudaEventDestroy(p->large_scratch_free);
    if (p->own_stream) cudaStreamDestroy(p->own_stream);
    delete p;
}

SBX_API sbx_status sbx_plan_create(const sbx_config *cfg, sbx_plan **out_plan)
{
    if (out_plan) *out_plan = nullptr;
    if (!cfg || !out_plan) return fail(nullptr, SBX_ERR_INVALID_ARGUMENT, "cfg and out_plan must not be NULL");
    if (cfg->struct_size != sizeof(sbx_config) || cfg->abi_version != SBX_ABI_VERSION)
        return fail(nullptr, SBX_ERR_INVALID_ARGUMENT, "sbx_config size/version mismatch (call sbx_config_default first)");
    int log2n = ilog2_exact(cfg->window_n);
    const int log2pad = ilog2_exact(cfg->pad_factor);
    if (cfg->window_n < 4) return fail(nullptr, SBX_ERR_INVALID_ARGUMENT, "window_n must be >= 4");
    if (log2pad < 0 || log2pad > 2) return fail(nullptr, SBX_ERR_UNSUPPORTED, "pad_factor must be 1, 2 or 4");
    const bool generic = log2n < (int)kMinLog2N;                  // not a power of two, or a tiny one
    if (generic) {
        log2n = -1;
        if (((cfg->window_n * cfg->pad_factor) & 1u) != 0)
            return fail(nullptr, SBX_ERR_UNSUPPORTED, "window_n * pad_factor must be even");
        if (cfg->window_n * cfg->pad_factor < 8)
            return fail(nullptr, SBX_ERR_UNSUPPORTED, "window_n * pad_factor must be >= 8");
        uint32_t rest = cfg->window_n;
        for (uint32_t f : {2u, 3u, 5u, 7u})
            while (rest % f == 0) rest /= f;
        if (rest != 1) {
            // a prime factor > 7: Bluestein, a circular convolution of power-of-two length >= N + M - 1 in shared memory
            if ((uint64_t)cfg->window_n * (cfg->pad_factor + 1) - 1 > sbx::BLUESTEIN_MAX_LEN)
                return fail(nullptr, SBX_ERR_UNSUPPORTED, "window_n has a prime factor larger than 7 and window_n * (pad_factor + 1) - 1 exceeds 16384 "
                                                          "(the Bluestein path's convolution length)");
        } else if ((uint64_t)cfg->window_n * cfg->pad_factor * 12 > sbx::kMaxSmemOptIn) {
            return fail(nullptr, SBX_ERR_UNSUPPORTED, "non-power-of-two windows need window_n * pad_factor <= 19370 (one SM's shared memory)");
        }
    } else {
        if (log2n > (int)kMaxLog2NLarge) return fail(nullptr, SBX_ERR_UNSUPPORTED, "window_n must be <= 65536");
        if (log2n + log2pad > (int)kMaxLog2MLarge)
            return fail(nullptr, SBX_ERR_UNSUPPORTED, "window_n * pad_factor must be <= 262144");
        if (log2n + log2pad > (int)kMaxLog2M && log2n < 13)
            return fail(nullptr, SBX_ERR_UNSUPPORTED, "transforms longer than 16384 need window_n >= 8192");
    }
    if (cfg->channels != 1 && cfg->channels != 2) return fail(nullptr, SBX_ERR_INVALID_ARGUMENT, "channels must be 1 or 2");
    if (cfg->hop == 0) return fail(nullptr, SBX_ERR_INVALID_ARGUMENT, "hop must be >= 1");
    if (cfg->rows_h == 0 || cfg->rows_h > (1u << 20)) return fail(nullptr, SBX_ERR_INVALID_ARGUMENT, "rows_h must be in [1, 2^20]");
    if (!(cfg->f_min > 0.0) || !(cfg->f_max > cfg->f_min)) return fail(nullptr, SBX_ERR_INVALID_ARGUMENT, "need 0 < f_min < f_max");
    if (!(cfg->sample_rate >= 1.0f) || !(cfg->sample_rate < 4.0e9f)) return fail(nullptr, SBX_ERR_INVALID_ARGUMENT, "sample_rate out of range");
    if (!(cfg->db_max > cfg->db_min)) return fail(nullptr, SBX_ERR_INVALID_ARGUMENT, "need db_min < db_max");
    if (!(cfg->db_eps >= 0.0f)) return fail(nullptr, SBX_ERR_INVALID_ARGUMENT, "db_eps must be >= 0");
    if (cfg->interp > SBX_INTERP_COSINE) return fail(nullptr, SBX_ERR_INVALID_ARGUMENT, "unknown interp");
    if (cfg->scheme > SBX_SCHEME_STEREO) return fail(nullptr, SBX_ERR_INVALID_ARGUMENT, "unknown scheme");

    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(nullptr, SBX_ERR_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= n_dev) return fail(nullptr, SBX_ERR_INVALID_ARGUMENT, "device ordinal out of range");
    SBX_CUDA(nullptr, cudaSetDevice(cfg->device));
    cudaDeviceProp prop{};
    SBX_CUDA(nullptr, cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10)
        return fail(nullptr, SBX_ERR_CUDA, "kernels are compiled for sm_100a only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor));

    sbx_plan *p = new (std::nothrow) sbx_plan();
    if (!p) return fail(nullptr, SBX_ERR_OUT_OF_MEMORY, "out of host memory");
    p->cfg = *cfg;
    p->n = cfg->window_n;
    p->pad = cfg->pad_factor;
    p->m = p->n * p->pad;
    p->k_out = p->m / 2 - 1;
    p->rows = cfg->rows_h;
    p->hop = cfg->hop;
    p->channels = cfg->channels;
    p->log2n = log2n;
    p->log2pad = log2pad;

    // ---- tables (host-computed once; the kernels never evaluate exp/log/cos) ---------------------
    sbx::tables::HostTables t;
    sbx::tables::build(*cfg, t);
    p->period = t.period;
    p->scale = t.scale;
    if (!t.ok) {
        delete p;
        // (prime factors > 7 and transforms shorter than 8 were rejected above with their own messages)
        return fail(nullptr, SBX_ERR_UNSUPPORTED, t.generic && t.radix.empty()
                        ? "window_n needs more radix passes than the mixed-radix kernel runs (GENERIC_MAX_PASSES)"
                        : "a row of the bin map averages more samples than the shared-memory scratch holds, or the bin count exceeds the sample table's index width");
    }
    p->generic = t.generic;
    p->radix = t.radix;
    p->stride = t.stride;
    if (t.generic) sbx::tables::fill_generic_passes(t, p->generic_passes);
    p->h_window = t.window;
    p->h_rowmap = t.rowmap;
    p->h_thr_raw = t.thr_raw;
    p->large = t.large;
    p->real_ok = t.real_ok;
    p->use_real = sbx::choose_real(t.real_ok, t.real_auto);
    for (int c = 1; c <= 2; ++c) p->use_q2[c - 1] = sbx::choose_q2(p, c);
    p->use_q2_real = sbx::choose_q2_real(p);
    p->n_chunks = (uint32_t)t.chunks.size();
    if (!t.chunks.empty()) p->chunk0 = t.chunks[0];
    p->taps_clamped = t.taps_clamped ? 1 : 0;
    for (int i = 0; i < 2; ++i) { p->guess_a[i] = t.guess_a[i]; p->guess_b[i] = t.guess_b[i]; }
    p->step_margin = t.step_margin;

#define SBX_TRY(call)                                                                   \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            sbx_status st__ = e__ == cudaErrorMemoryAllocation ? SBX_ERR_OUT_OF_MEMORY : SBX_ERR_CUDA; \
            fail(nullptr, st__, std::string(#call) + ": " + cudaGetErrorString(e__));   \
            sbx_plan_destroy(p);                                                        \
            return st__;                                                                \
        }                                                                               \
    } while (0)

    SBX_TRY(upload(&p->d_window, p->h_window.data(), p->n * sizeof(float)));
    SBX_TRY(upload(&p->d_wtw, t.wtw.data(), t.wtw.size() * sizeof(float)));
    SBX_TRY(upload(&p->d_twp, t.twp.data(), t.twp.size() * sizeof(float)));
    SBX_TRY(upload(&p->d_wq, t.wq.data(), t.wq.size() * sizeof(float)));
    SBX_TRY(upload(&p->d_wq_real, t.wq_real.data(), t.wq_real.size() * sizeof(float)));
    SBX_TRY(upload(&p->d_twp_real, t.twp_real.data(), t.twp_real.size() * sizeof(float)));
    SBX_TRY(upload(&p->d_wtw4, t.wtw4.data(), t.wtw4.size() * sizeof(float)));
    SBX_TRY(upload(&p->d_rtw, t.rtw.data(), t.rtw.size() * sizeof(float)));
    SBX_TRY(upload(&p->d_twm, t.twm.data(), t.twm.size() * sizeof(float)));
    p->bluestein = t.bluestein;
    p->bl_len = t.bl_len;
    SBX_TRY(upload(&p->d_bl_a, t.bl_a.data(), t.bl_a.size() * sizeof(float)));
    SBX_TRY(upload(&p->d_bl_bf, t.bl_bf.data(), t.bl_bf.size() * sizeof(float)));
    SBX_TRY(upload(&p->d_bl_c, t.bl_c.data(), t.bl_c.size() * sizeof(float)));
    SBX_TRY(upload(&p->d_twg, t.twg.data(), t.twg.size() * sizeof(float)));
    SBX_TRY(upload(&p->d_pos_tab, t.pos_tab.data(), t.pos_tab.size() * sizeof(uint32_t)));
    SBX_TRY(upload(&p->d_twp_block, t.twp_block.data(), t.twp_block.size() * sizeof(float)));
    SBX_TRY(upload(&p->d_rowmap, p->h_rowmap.data(), p->rows * sizeof(sbx::RowMap)));
    SBX_TRY(upload(&p->d_rowinfo, t.rowinfo.data(), t.rowinfo.size() * sizeof(uint32_t)));
    SBX_TRY(upload(&p->d_chunks, t.chunks.data(), t.chunks.size() * sizeof(sbx::RowChunk)));
    SBX_TRY(upload(&p->d_samp, t.samp.data(), t.samp.size() * sizeof(uint32_t)));
    SBX_TRY(upload(&p->d_thr, t.thr.data(), t.thr.size() * sizeof(float)));
    SBX_TRY(upload(&p->d_lut, t.lut, sizeof(t.lut)));
    SBX_TRY(upload(&p->d_ctab, t.ctab, sizeof(t.ctab)));
    SBX_TRY(cudaMalloc(&p->d_frame, (size_t)p->n * 2 * sizeof(float)));
    SBX_TRY(cudaMalloc(&p->d_mags1, (size_t)p->k_out * 2 * sizeof(float)));
    SBX_TRY(cudaMalloc(&p->d_col, (size_t)p->rows * 4));
    SBX_TRY(cudaMallocHost(&p->h_frame, (size_t)(p->n > p->k_out ? p->n : p->k_out) * 2 * sizeof(float)));
    SBX_TRY(cudaMallocHost(&p->h_col, (size_t)p->rows * 4));
    SBX_TRY(cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking));
#undef SBX_TRY
    p->stream = p->own_stream;
    *out_plan = p;
    return SBX_OK;
}

SBX_API float sbx_sample_rate(const sbx_plan *p) { return p ? p->cfg.sample_rate : 0.0f; }
SBX_API uint32_t sbx_num_input_samples(const sbx_plan *p) { return p ? p->n : 0; }
SBX_API uint32_t sbx_num_output_frequencies(const sbx_plan *p) { return p ? p->k_out : 0; }
SBX_API uint32_t sbx_fft_length(const sbx_plan *p) { return p ? p->m : 0; }
SBX_API uint32_t sbx_rows(const sbx_plan *p) { return p ? p->rows : 0; }
SBX_API uint32_t sbx_hop(const sbx_plan *p) { return p ? p->hop : 0; }
SBX_API uint32_t sbx_channels(const sbx_plan *p) { return p ? p->channels : 0; }
SBX_API uint64_t sbx_launch_count(const sbx_plan *p) { return p ? p->launches : 0; }

SBX_API const char *sbx_kernel_path(const sbx_plan *p)
{
    if (!p) return "";
    if (p->generic) return p->bluestein ? "bluestein" : "generic";
    if (p->channels == 1 && p->use_real && (p->hop & 1u) == 0) return p->use_q2_real ? "fused-real-q2" : "fused-real";
    if (p->large) return "large";
    if (p->use_q2[p->channels - 1]) return p->channels == 1 ? "fused-pair-q2" : "fused-stereo-q2";
    return p->channels == 1 ? "fused-pair" : "fused-stereo";
}

SBX_API uint64_t sbx_num_frames(const sbx_plan *p, uint64_t n_samples)
{
    if (!p || n_samples < p->n) return 0;
    return (n_samples - p->n) / p->hop + 1;
}

SBX_API sbx_status sbx_set_stream(sbx_plan *p, void *cuda_stream)
{
    if (!p) return SBX_ERR_INVALID_ARGUMENT;
    p->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : p->own_stream;
    return SBX_OK;
}

SBX_API sbx_status sbx_synchronize(sbx_plan *p)
{
    if (!p) return SBX_ERR_INVALID_ARGUMENT;
    SBX_CUDA(p, cudaSetDevice(p->cfg.device));
    SBX_CUDA(p, cudaStreamSynchronize(p->stream));
    return SBX_OK;
}

// ---- single frame, host buffers ----------------------------------------------------------------
static sbx_status single_frame(sbx_plan *p, const float *lr, uint32_t n_avail, float *out_lr, uint8_t *rgba)
{
    if (!p) return SBX_ERR_INVALID_ARGUMENT;
    if (!lr && n_avail) return fail(p, SBX_ERR_INVALID_ARGUMENT, "lr is NULL");
    if (n_avail < p->n) return SBX_NOT_ENOUGH_SAMPLES;                 // fft.rs:72
    SBX_CUDA(p, cudaSetDevice(p->cfg.device));
    cudaStream_t s = p->stream;
    std::memcpy(p->h_frame, lr, (size_t)p->n * 2 * sizeof(float));
    SBX_CUDA(p, cudaMemcpyAsync(p->d_frame, p->h_frame, (size_t)p->n * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
    sbx_status st = sbx::launch_fused(p, 2, p->d_frame, 1, rgba ? p->d_col : nullptr, out_lr ? p->d_mags1 : nullptr, s);
    if (st != SBX_OK) return st;
    if (out_lr) SBX_CUDA(p, cudaMemcpyAsync(p->h_frame, p->d_mags1, (size_t)p->k_out * 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (rgba) SBX_CUDA(p, cudaMemcpyAsync(p->h_col, p->d_col, (size_t)p->rows * 4, cudaMemcpyDeviceToHost, s));
    SBX_CUDA(p, cudaStreamSynchronize(s));
    if (out_lr) std::memcpy(out_lr, p->h_frame, (size_t)p->k_out * 2 * sizeof(float));
    if (rgba) std::memcpy(rgba, p->h_col, (size_t)p->rows * 4);
    return SBX_OK;
}

SBX_API sbx_status sbx_process_magnitudes(sbx_plan *p, const float *lr, uint32_t n_avail, float *out_lr)
{
    if (p && !out_lr) return fail(p, SBX_ERR_INVALID_ARGUMENT, "out_lr is NULL");
    return single_frame(p, lr, n_avail, out_lr, nullptr);
}

SBX_API sbx_status sbx_process_column(sbx_plan *p, const float *lr, uint32_t n_avail, uint8_t *rgba)
{
    if (p && !rgba) return fail(p, SBX_ERR_INVALID_ARGUMENT, "rgba is NULL");
    return single_frame(p, lr, n_avail, nullptr, rgba);
}

SBX_API sbx_status sbx_colorize_magnitudes(sbx_plan *p, const float *mags_lr, uint64_t n_frames, uint8_t *rgba)
{
    if (!p) return SBX_ERR_INVALID_ARGUMENT;
    if (n_frames == 0) return SBX_OK;
    if (!mags_lr || !rgba) return fail(p, SBX_ERR_INVALID_ARGUMENT, "NULL buffer");
    SBX_CUDA(p, cudaSetDevice(p->cfg.device));
    cudaStream_t s = p->stream;
    const size_t in_bytes = (size_t)n_frames * p->k_out * 2 * sizeof(float), out_bytes = (size_t)n_frames * p->rows * 4;
    // plan-owned staging, grown on demand and kept: no cudaMalloc / cudaFree (a device-wide synchronisation) per call
    if (in_bytes > p->cz_in_bytes) {
        cudaFree(p->d_cz_in);
        p->d_cz_in = nullptr;
        p->cz_in_bytes = 0;
        const cudaError_t e = cudaMalloc(&p->d_cz_in, in_bytes);
        if (e != cudaSuccess) return cuda_fail(p, e, "cudaMalloc(colourise input)");
        p->cz_in_bytes = in_bytes;
    }
    if (out_bytes > p->cz_out_bytes) {
        cudaFree(p->d_cz_out);
        p->d_cz_out = nullptr;
        p->cz_out_bytes = 0;
        const cudaError_t e = cudaMalloc(&p->d_cz_out, out_bytes);
        if (e != cudaSuccess) return cuda_fail(p, e, "cudaMalloc(colourise output)");
        p->cz_out_bytes = out_bytes;
    }
    sbx_status st = SBX_OK;
    cudaError_t e = cudaMemcpyAsync(p->d_cz_in, mags_lr, in_bytes, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) st = sbx::launch_colourise(p, p->d_cz_in, n_frames, p->d_cz_out, s);
    if (e == cudaSuccess && st == SBX_OK) e = cudaMemcpyAsync(rgba, p->d_cz_out, out_bytes, cudaMemcpyDeviceToHost, s);
    const cudaError_t e2 = cudaStreamSynchronize(s);
    if (st != SBX_OK) return st;
    if (e != cudaSuccess) return cuda_fail(p, e, "colourise copy");
    if (e2 != cudaSuccess) return cuda_fail(p, e2, "colourise sync");
    return SBX_OK;
}

// ---- batch, device buffers ---------------------------------------------------------------------
static sbx_status batch_device(sbx_plan *p, const float *d_pcm, uint64_t n_samples, uint8_t *d_rgba, float *d_mags,
                               uint64_t capacity_frames, uint64_t *frames_out)
{
    if (frames_out) *frames_out = 0;
    if (!p) return SBX_ERR_INVALID_ARGUMENT;
    const uint64_t frames = sbx_num_frames(p, n_samples);
    if (frames == 0) return SBX_OK;
    if (!d_pcm || (!d_rgba && !d_mags)) return fail(p, SBX_ERR_INVALID_ARGUMENT, "NULL device buffer");
    if (frames > capacity_frames) return fail(p, SBX_ERR_BUFFER_TOO_SMALL, "output capacity is smaller than the frame count");
    if ((reinterpret_cast<uintptr_t>(d_pcm) & 15u) != 0) return fail(p, SBX_ERR_INVALID_ARGUMENT, "d_pcm must be 16-byte aligned");
    SBX_CUDA(p, cudaSetDevice(p->cfg.device));
    sbx_status st = sbx::launch_fused(p, (int)p->channels, d_pcm, frames, d_rgba, d_mags, p->stream);
    if (st == SBX_OK && frames_out) *frames_out = frames;
    return st;
}

SBX_API sbx_status sbx_process_batch_device(sbx_plan *p, const float *d_pcm, uint64_t n_samples, uint8_t *d_rgba,
                                            uint64_t rgba_capacity_frames, uint64_t *frames_out)
{
    if (p && !d_rgba && sbx_num_frames(p, n_samples)) return fail(p, SBX_ERR_INVALID_ARGUMENT, "d_rgba is NULL");
    return batch_device(p, d_pcm, n_samples, d_rgba, nullptr, rgba_capacity_frames, frames_out);
}

SBX_API sbx_status sbx_process_batch_device_planar(sbx_plan *p, const float *d_pcm, uint32_t n_sequences, uint64_t sequence_stride_samples,
                                                   uint64_t n_samples_per_sequence, uint8_t *d_rgba, uint64_t rgba_capacity_frames,
                                                   uint64_t *frames_per_sequence_out)
{
    if (frames_per_sequence_out) *frames_per_sequence_out = 0;
    if (!p) return SBX_ERR_INVALID_ARGUMENT;
    const uint64_t frames = sbx_num_frames(p, n_samples_per_sequence);
    if (frames == 0 || n_sequences == 0) return SBX_OK;
    if (!d_pcm || !d_rgba) return fail(p, SBX_ERR_INVALID_ARGUMENT, "NULL device buffer");
    if (n_sequences > 1 && sequence_stride_samples < n_samples_per_sequence)
        return fail(p, SBX_ERR_INVALID_ARGUMENT, "sequence_stride_samples is smaller than n_samples_per_sequence");
    if (frames > rgba_capacity_frames / n_sequences)
        return fail(p, SBX_ERR_BUFFER_TOO_SMALL, "output capacity is smaller than n_sequences * frames per sequence");
    if ((reinterpret_cast<uintptr_t>(d_pcm) & 15u) != 0) return fail(p, SBX_ERR_INVALID_ARGUMENT, "d_pcm must be 16-byte aligned");
    SBX_CUDA(p, cudaSetDevice(p->cfg.device));
    sbx::Sequences seqs;
    seqs.count = n_sequences;
    seqs.stride_floats = sequence_stride_samples * p->channels;
    const sbx_status st = sbx::launch_fused(p, (int)p->channels, d_pcm, frames, d_rgba, nullptr, p->stream, seqs);
    if (st == SBX_OK && frames_per_sequence_out) *frames_per_sequence_out = frames;
    return st;
}

SBX_API sbx_status sbx_magnitudes_batch_device(sbx_plan *p, const float *d_pcm, uint64_t n_samples, float *d_mags_lr,
                                               uint64_t capacity_frames, uint64_t *frames_out)
{
    if (p && !d_mags_lr && sbx_num_frames(p, n_samples)) return fail(p, SBX_ERR_INVALID_ARGUMENT, "d_mags_lr is NULL");
    return batch_device(p, d_pcm, n_samples, nullptr, d_mags_lr, capacity_frames, frames_out);
}

SBX_API sbx_status sbx_magnitudes_batch_device_f16(sbx_plan *p, const float *d_pcm, uint64_t n_samples, uint16_t *d_mags_f16,
                                                   uint64_t capacity_frames, uint64_t *frames_out)
{
    if (frames_out) *frames_out = 0;
    if (!p) return SBX_ERR_INVALID_ARGUMENT;
    const uint64_t frames = sbx_num_frames(p, n_samples);
    if (frames == 0) return SBX_OK;
    if (!d_pcm || !d_mags_f16) return fail(p, SBX_ERR_INVALID_ARGUMENT, "NULL device buffer");
    if (frames > capacity_frames) return fail(p, SBX_ERR_BUFFER_TOO_SMALL, "output capacity is smaller than the frame count");
    if ((reinterpret_cast<uintptr_t>(d_pcm) & 15u) != 0) return fail(p, SBX_ERR_INVALID_ARGUMENT, "d_pcm must be 16-byte aligned");
    SBX_CUDA(p, cudaSetDevice(p->cfg.device));
    // f32 rows go through a chunk-sized scratch (even frame count: mono pairs), then are packed
    uint64_t chunk = ((uint64_t)64 << 20) / ((uint64_t)p->k_out * 8);
    chunk = chunk < 2 ? 2 : (chunk & ~1ull);
    if (!p->d_f16_scratch) {
        SBX_CUDA(p, cudaMalloc(&p->d_f16_scratch, (size_t)chunk * p->k_out * 8));
    }
    for (uint64_t f0 = 0; f0 < frames; f0 += chunk) {
        const uint64_t nf = frames - f0 < chunk ? frames - f0 : chunk;
        sbx_status st = sbx::launch_fused(p, (int)p->channels, d_pcm + (size_t)f0 * p->hop * p->channels, nf, nullptr,
                                          p->d_f16_scratch, p->stream);
        if (st != SBX_OK) return st;
        const size_t count = (size_t)nf * p->k_out;
        const unsigned blocks = (unsigned)((count + 255) / 256 > 148 * 16 ? 148 * 16 : (count + 255) / 256);
        sbx::pack_f16_kernel<<<blocks, 256, 0, p->stream>>>(reinterpret_cast<const float2 *>(p->d_f16_scratch),
                                                           reinterpret_cast<__half2 *>(d_mags_f16) + (size_t)f0 * p->k_out, count);
        SBX_CUDA(p, cudaGetLastError());
        p->launches += 1;
    }
    if (frames_out) *frames_out = frames;
    return SBX_OK;
}

SBX_API sbx_status sbx_plan_set_pan_gradient(sbx_plan *p, const uint8_t *rgb, uint32_t entries)
{
    if (!p) return SBX_ERR_INVALID_ARGUMENT;
    if (!rgb || entries < 2 || entries > 65536u) return fail(p, SBX_ERR_INVALID_ARGUMENT, "pan gradient: need rgb and 2 <= entries <= 65536");
    SBX_CUDA(p, cudaSetDevice(p->cfg.device));
    std::vector<uint32_t> packed(entries);
    for (uint32_t i = 0; i < entries; ++i)
        packed[i] = (uint32_t)rgb[3 * i] | ((uint32_t)rgb[3 * i + 1] << 8) | ((uint32_t)rgb[3 * i + 2] << 16) | (255u << 24);
    SBX_CUDA(p, cudaStreamSynchronize(p->stream));           // no launch may still read the table being replaced
    uint32_t *d_new = nullptr;
    SBX_CUDA(p, cudaMalloc(&d_new, (size_t)entries * sizeof(uint32_t)));
    const cudaError_t e = cudaMemcpy(d_new, packed.data(), (size_t)entries * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(d_new); return cuda_fail(p, e, "pan gradient upload"); }
    // a live stream created earlier has captured the previous table's address in its graphs: keep that table alive
    // until the plan goes (a few KB each), so a late call can make such a stream stale but never dangling
    if (p->d_pan_lut) p->retired_pan_luts.push_back(p->d_pan_lut);
    p->d_pan_lut = d_new;
    p->pan_levels = entries;
    p->h_pan_rgb.assign(rgb, rgb + (size_t)entries * 3);
    return SBX_OK;
}

SBX_API sbx_status sbx_palette_table(const sbx_plan *p, uint32_t resolution, float *out)
{
    if (!p || !out || resolution < 2) return SBX_ERR_INVALID_ARGUMENT;
    // gradient.eval_continuous(t) as a table of `levels` steps: table[clamp(floor(t * levels), 0, levels - 1)]
    // (exact for the 256-entry ramps; the stereo schemes use the pan gradient where one is installed)
    const bool pan_table = p->cfg.scheme == SBX_SCHEME_STEREO && !p->h_pan_rgb.empty();
    const uint8_t *table = pan_table ? p->h_pan_rgb.data() : p->cfg.lut_rgb;
    const double levels = pan_table ? (double)p->pan_levels : 256.0;
    auto eval = [table, levels](double t, float rgb[3]) {
        const double v = std::floor(t * levels);
        const int idx = !(v >= 0.0) ? 0 : (v > levels - 1.0 ? (int)levels - 1 : (int)v);
        for (int c = 0; c < 3; ++c) rgb[c] = (float)table[3 * idx + c] / 256.0f;               // colorscheme.rs:83,86
    };
    for (uint32_t i = 0; i < resolution; ++i)
        for (uint32_t j = 0; j < resolution; ++j) {
            const float magnitude = (float)i / (float)(resolution - 1);                          // :79
            const float pan = 1.0f - ((float)j / (float)(resolution - 1));                       // :80
            float *o = out + ((size_t)i * resolution + j) * 4;
            if (p->cfg.scheme == SBX_SCHEME_STEREO) { eval((double)pan, o); o[3] = magnitude; }
            else { eval((double)magnitude, o); o[3] = 1.0f; }
        }
    return SBX_OK;
}

SBX_API sbx_status sbx_blit_columns_device(sbx_plan *p, const uint8_t *d_columns, uint64_t n_frames, uint8_t *d_image,
                                           uint32_t width, uint32_t rowstride_bytes, uint32_t first_column, uint32_t *next_column_out)
{
    if (!p) return SBX_ERR_INVALID_ARGUMENT;
    if (width == 0 || first_column >= width) return fail(p, SBX_ERR_INVALID_ARGUMENT, "need width >= 1 and first_column < width");
    if ((rowstride_bytes & 3u) != 0 || (uint64_t)rowstride_bytes < 4ull * width)
        return fail(p, SBX_ERR_INVALID_ARGUMENT, "rowstride_bytes must be a multiple of 4 and >= 4 * width");
    if (next_column_out) *next_column_out = (uint32_t)((first_column + n_frames) % width);
    if (n_frames == 0) return SBX_OK;
    if (!d_columns || !d_image) return fail(p, SBX_ERR_INVALID_ARGUMENT, "NULL device buffer");
    if (((reinterpret_cast<uintptr_t>(d_columns) | reinterpret_cast<uintptr_t>(d_image)) & 3u) != 0)
        return fail(p, SBX_ERR_INVALID_ARGUMENT, "buffers must be 4-byte aligned");
    // columns older than one image width are overwritten by later ones: draw only the last `width`
    const uint64_t skip = n_frames > width ? n_frames - width : 0;
    const uint32_t frames = (uint32_t)(n_frames - skip);
    const uint32_t first = (uint32_t)((first_column + skip) % width);
    SBX_CUDA(p, cudaSetDevice(p->cfg.device));
    const dim3 grid((frames + 31) / 32, (p->rows + 31) / 32);
    if (grid.y > 65535u) return fail(p, SBX_ERR_UNSUPPORTED, "rows_h too large for the blit grid");
    sbx::blit_columns_kernel<<<grid, 256, 0, p->stream>>>(reinterpret_cast<const uint32_t *>(d_columns) + (size_t)skip * p->rows, frames,
                                                         p->rows, reinterpret_cast<uint32_t *>(d_image), width, rowstride_bytes / 4,
                                                         first);
    SBX_CUDA(p, cudaGetLastError());
    p->launches += 1;
    return SBX_OK;
}

// ---- batch, host buffers: chunked, H2D / kernel / D2H overlapped over kSlots streams -------------
}  // extern "C"

namespace sbx {
// The chunk pipeline behind sbx_process_batch and the sbx_multi_* calls: every segment is a run of frames of one
// recording (HOST pointers); the chunks of all segments go round-robin over the plan's slot streams and only the
// end of the call drains them, so the copies of one segment overlap the kernels of the next.
sbx_status batch_host_segments(sbx_plan *p, const HostSegment *segs, size_t n_segs)
{
    SBX_CUDA(p, cudaSetDevice(p->cfg.device));
    if (p->chunk_frames == 0) {
        // at most ~32 MiB of RGBA (SBX_BATCH_CHUNK_MB) and ~64 MiB of PCM per chunk, even frame count (mono pairs),
        // at least 2 frames.  Measured end to end on a B200 (cfg1, 131072 frames per call): 8 / 32 / 128 MiB chunks
        // -> 12.2 / 12.1 / 11.0 M frames/s; the D2H link (~50 GB/s) is the limit, not the chunking.
        const char *env = std::getenv("SBX_BATCH_CHUNK_MB");
        const uint64_t chunk_mb = env && std::atoi(env) > 0 ? (uint64_t)std::atoi(env) : 32u;
        uint64_t cf = (chunk_mb << 20) / ((uint64_t)p->rows * 4);
        const uint64_t cf_pcm = ((uint64_t)64 << 20) / ((uint64_t)p->hop * p->channels * sizeof(float));
        cf = cf < cf_pcm ? cf : cf_pcm;
        cf = cf < 2 ? 2 : (cf & ~1ull);
        const size_t pcm_bytes = ((size_t)(cf - 1) * p->hop + p->n) * p->channels * sizeof(float);
        cudaError_t e = cudaSuccess;
        for (auto &s : p->slots) {
            if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaMalloc(&s.d_pcm, pcm_bytes);
            if (e == cudaSuccess) e = cudaMalloc(&s.d_rgba, (size_t)cf * p->rows * 4);
        }
        if (e != cudaSuccess) {                                      // leave no half-built pipeline behind
            for (auto &s : p->slots) {
                if (s.stream) cudaStreamDestroy(s.stream);
                cudaFree(s.d_pcm);
                cudaFree(s.d_rgba);
                s = sbx_plan::Slot{};
            }
            return cuda_fail(p, e, "batch pipeline allocation");
        }
        p->chunk_frames = cf;
    }
    // order after whatever the caller queued on the plan stream
    SBX_CUDA(p, cudaStreamSynchronize(p->stream));
    const uint64_t cf = p->chunk_frames;
    uint64_t chunk = 0;
    sbx_status st = SBX_OK;
    for (size_t g = 0; g < n_segs && st == SBX_OK; ++g) {
        const float *pcm = segs[g].pcm;
        uint8_t *rgba = segs[g].rgba;
        const uint64_t frames = segs[g].frames;
        for (uint64_t f0 = 0; f0 < frames && st == SBX_OK; f0 += cf, ++chunk) {
            sbx_plan::Slot &s = p->slots[chunk % sbx_plan::kSlots];
            const uint64_t nf = frames - f0 < cf ? frames - f0 : cf;
            const size_t samples = (size_t)(nf - 1) * p->hop + p->n;
            cudaError_t e = cudaMemcpyAsync(s.d_pcm, pcm + (size_t)f0 * p->hop * p->channels,
                                            samples * p->channels * sizeof(float), cudaMemcpyHostToDevice, s.stream);
            if (e != cudaSuccess) { st = cuda_fail(p, e, "batch H2D copy"); break; }
            // The multi-kernel path works in ONE plan-wide scratch (d_zbuf, d_magbuf): chunk k+1's kernels on the next
            // slot stream must not start before chunk k's have finished with it.  Only the kernels are chained -- the
            // H2D copy above and the D2H copy below still overlap with the neighbouring chunks' kernels.
            if (p->large) {
                if (!p->large_scratch_free) {
                    e = cudaEventCreateWithFlags(&p->large_scratch_free, cudaEventDisableTiming);
                    if (e != cudaSuccess) { st = cuda_fail(p, e, "cudaEventCreate(large scratch)"); break; }
                } else {
                    e = cudaStreamWaitEvent(s.stream, p->large_scratch_free, 0);
                    if (e != cudaSuccess) { st = cuda_fail(p, e, "cudaStreamWaitEvent(large scratch)"); break; }
                }
            }
            st = sbx::launch_fused(p, (int)p->channels, s.d_pcm, nf, s.d_rgba, nullptr, s.stream);
            if (st != SBX_OK) break;
            if (p->large) {
                e = cudaEventRecord(p->large_scratch_free, s.stream);
                if (e != cudaSuccess) { st = cuda_fail(p, e, "cudaEventRecord(large scratch)"); break; }
            }
            e = cudaMemcpyAsync(rgba + (size_t)f0 * p->rows * 4, s.d_rgba, (size_t)nf * p->rows * 4,
                                cudaMemcpyDeviceToHost, s.stream);
            if (e != cudaSuccess) st = cuda_fail(p, e, "batch D2H copy");
        }
    }
    // always drain: no copy into the caller's buffers may still be in flight when this returns, error or not
    for (auto &s : p->slots) {
        const cudaError_t e = cudaStreamSynchronize(s.stream);
        if (e != cudaSuccess && st == SBX_OK) st = cuda_fail(p, e, "batch stream synchronize");
    }
    return st;
}
}  // namespace sbx

extern "C" {

SBX_API sbx_status sbx_process_batch(sbx_plan *p, const float *pcm, uint64_t n_samples, uint8_t *rgba,
                                     uint64_t rgba_capacity_frames, uint64_t *frames_out)
{
    if (frames_out) *frames_out = 0;
    if (!p) return SBX_ERR_INVALID_ARGUMENT;
    const uint64_t frames = sbx_num_frames(p, n_samples);
    if (frames == 0) return SBX_OK;
    if (!pcm || !rgba) return fail(p, SBX_ERR_INVALID_ARGUMENT, "NULL host buffer");
    if (frames > rgba_capacity_frames) return fail(p, SBX_ERR_BUFFER_TOO_SMALL, "rgba capacity is smaller than the frame count");
    const sbx::HostSegment seg{pcm, frames, rgba};
    const sbx_status st = sbx::batch_host_segments(p, &seg, 1);
    if (st != SBX_OK) return st;
    if (frames_out) *frames_out = frames;
    return SBX_OK;
}

// ---- introspection -----------------------------------------------------------------------------
SBX_API sbx_status sbx_get_bin_map(const sbx_plan *p, float *f0, float *step, uint32_t *n, float *period)
{
    if (!p) return SBX_ERR_INVALID_ARGUMENT;
    for (uint32_t py = 0; py < p->rows; ++py) {
        if (f0) f0[py] = p->h_rowmap[py].f0;
        if (step) step[py] = p->h_rowmap[py].step;
        if (n) n[py] = p->h_rowmap[py].n;
    }
    if (period) *period = p->period;
    return SBX_OK;
}

SBX_API sbx_status sbx_get_color_thresholds(const sbx_plan *p, float thr[255])
{
    if (!p || !thr) return SBX_ERR_INVALID_ARGUMENT;
    std::memcpy(thr, p->h_thr_raw.data(), 255 * sizeof(float));
    return SBX_OK;
}

SBX_API sbx_status sbx_get_window(const sbx_plan *p, float *w)
{
    if (!p || !w) return SBX_ERR_INVALID_ARGUMENT;
    std::memcpy(w, p->h_window.data(), p->n * sizeof(float));
    return SBX_OK;
}

SBX_API void *sbx_host_alloc(size_t bytes)
{
    void *ptr = nullptr;
    if (cudaMallocHost(&ptr, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return ptr;
}
SBX_API void sbx_host_free(void *ptr)
{
    if (ptr) cudaFreeHost(ptr);
}

}  // extern "C"
