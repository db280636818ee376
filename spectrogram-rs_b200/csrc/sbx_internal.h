// sbx_internal.h -- host-side plan object shared by sbx_plan.cu and sbx_live.cu (not installed).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/sbx.h"
#include "stft_generic.cuh"
#include "stft_large.cuh"

struct sbx_plan {
    sbx_config cfg;
    uint32_t n, pad, m, k_out, rows, hop, channels;
    int log2n, log2pad;
    float period, scale;

    // host copies of the tables (for introspection and the bit-exactness tests)
    std::vector<float> h_window;
    std::vector<sbx::RowMap> h_rowmap;
    std::vector<float> h_thr_raw;       // [2][256] bare thresholds
    float guess_a[2] = {0, 0}, guess_b[2] = {0, 0};
    float step_margin = 1.0f;
    uint32_t n_chunks = 0;
    sbx::RowChunk chunk0{};             // chunks[0], passed by value in the kernel arguments

    // device tables
    float *d_window = nullptr;
    float2 *d_wtw = nullptr;
    float2 *d_twp = nullptr;
    float2 *d_wq = nullptr, *d_wq_real = nullptr;   // q2 kernels: W_M^lo (pass 0, lane B)
    sbx::RowMap *d_rowmap = nullptr;
    uint32_t *d_rowinfo = nullptr;
    sbx::RowChunk *d_chunks = nullptr;
    uint2 *d_samp = nullptr;
    int taps_clamped = 0;

    // generic path (window not a power of two): run-time radix list, full twiddle table, bin -> position table
    bool generic = false;
    float2 *d_twm = nullptr;
    uint32_t *d_pos_tab = nullptr;
    std::vector<uint32_t> radix, stride;
    sbx::GenericArgs generic_passes{};  // pass description (radices, strides, table offsets), filled at plan creation
    float2 *d_twg = nullptr;            // per-pass twiddle planes
    // Bluestein (generic windows with a prime factor > 7): chirp tables, convolution length
    bool bluestein = false;
    uint32_t bl_len = 0;
    float2 *d_bl_a = nullptr, *d_bl_bf = nullptr, *d_bl_c = nullptr;

    // REAL geometry (mono batches of long windows, one frame per CTA): tables of the half-length packed transform
    bool real_ok = false;               // tables exist
    bool use_real = false;              // mono batches take it (plan-time choice, see choose_real)
    bool use_q2_real = false;           // ... and for the real-sequence geometry
    bool use_q2[2] = {false, false};    // P = 2: the packed-lane kernel of stft_q2.cuh for mono / stereo PCM (choose_q2)
    float2 *d_twp_real = nullptr;
    float4 *d_wtw4 = nullptr;
    float2 *d_rtw = nullptr;

    // large path (M > 16384): block-transform twiddles and per-launch scratch
    bool large = false;
    float2 *d_twp_block = nullptr;
    float2 *d_zbuf = nullptr;           // [large_units][M]
    float2 *d_magbuf = nullptr;         // [large_units][M/2]
    uint32_t large_units = 0;           // scratch capacity in units (frame pairs / frames)
    // d_zbuf / d_magbuf are plan-wide: launches of the large path on DIFFERENT streams (the chunk pipeline of
    // sbx_process_batch) must not overlap.  Every such launch waits for this event and re-records it.
    cudaEvent_t large_scratch_free = nullptr;
    float *d_thr = nullptr;
    uint32_t *d_lut = nullptr;
    uint4 *d_ctab = nullptr;
    uint32_t *d_pan_lut = nullptr;      // stereo schemes: finer gradient table indexed by pan (sbx_plan_set_pan_gradient), else d_lut
    uint32_t pan_levels = 256;
    std::vector<uint8_t> h_pan_rgb;
    std::vector<uint32_t *> retired_pan_luts;   // replaced tables, freed with the plan (live graphs may still hold them)

    // single-frame staging (host pinned + device)
    float *h_frame = nullptr;           // max(N, K) x 2 floats: N x 2 in, reused for K x 2 out
    uint8_t *h_col = nullptr;           // H x 4
    float *d_frame = nullptr;           // N x 2
    float *d_mags1 = nullptr;           // K x 2
    uint8_t *d_col = nullptr;           // H x 4

    float *d_cz_in = nullptr;           // sbx_colorize_magnitudes staging (grown on demand, kept)
    uint8_t *d_cz_out = nullptr;
    size_t cz_in_bytes = 0, cz_out_bytes = 0;

    float *d_f16_scratch = nullptr;     // f32 rows of one chunk, packed to binary16 by sbx_magnitudes_batch_device_f16

    // batch pipeline slots (host-pointer batch entry point)
    static constexpr int kSlots = 3;
    struct Slot {
        cudaStream_t stream = nullptr;
        float *d_pcm = nullptr;
        uint8_t *d_rgba = nullptr;
    } slots[kSlots];
    uint64_t chunk_frames = 0;

    // launch geometry of the persistent kernel for this plan's hop, per channel layout (index CH - 1; 2 = REAL)
    struct PersistentGeom { bool ready = false; int ctas = 0, occ = 0; size_t smem = 0; } pgeom[3];

    // set by sbx_live while it captures its single-frame graphs (KernelArgs::done_flag / done_count), else NULL
    unsigned int *live_done_flag = nullptr, *live_done_count = nullptr;

    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;      // current stream (own or caller's)
    uint64_t launches = 0;
    mutable std::string last_error;

    sbx::KernelArgs base_args() const;
};

namespace sbx {
// equally long PCM sequences (planar channels) handled by one call: sequence s starts s * stride_floats into the PCM
// buffer and writes frames [s * n_frames, (s + 1) * n_frames) of the output
struct Sequences { uint32_t count = 1; unsigned long long stride_floats = 0; };
// a run of `frames` frames of one recording in HOST memory and where its columns go (sbx_process_batch, sbx_multi_*)
struct HostSegment { const float *pcm; uint64_t frames; uint8_t *rgba; };
sbx_status batch_host_segments(sbx_plan *p, const HostSegment *segs, size_t n_segs);
// launches the plan's kernel family for `channels`-layout PCM on `stream`
sbx_status launch_fused(sbx_plan *p, int channels, const float *d_pcm, uint64_t n_frames, uint8_t *d_rgba,
                        float *d_mags, cudaStream_t stream, Sequences seqs = {});
sbx_status launch_generic(sbx_plan *p, int channels, const float *d_pcm, uint64_t n_frames, uint8_t *d_rgba,
                          float *d_mags, cudaStream_t stream);
sbx_status launch_large(sbx_plan *p, int channels, const float *d_pcm, uint64_t n_frames, uint8_t *d_rgba,
                        float *d_mags, cudaStream_t stream);
sbx_status launch_real(sbx_plan *p, const float *d_pcm, uint64_t n_frames, uint8_t *d_rgba, float *d_mags, cudaStream_t stream,
                       Sequences seqs);
cudaError_t launch_fused_small(sbx_plan *p, int channels, const KernelArgs &a, uint64_t n_frames, cudaStream_t stream, Sequences seqs);   // N <= 1024
cudaError_t launch_fused_big(sbx_plan *p, int channels, const KernelArgs &a, uint64_t n_frames, cudaStream_t stream, Sequences seqs);     // N >= 2048
bool q2_available(const sbx_plan *p);                                                                                   // sbx_dispatch_q2.cu
cudaError_t launch_fused_q2(sbx_plan *p, int channels, const KernelArgs &a, uint64_t n_frames, cudaStream_t stream, Sequences seqs);      // P = 2, packed lanes
bool q2_real_available(const sbx_plan *p);
cudaError_t launch_real_q2(sbx_plan *p, const KernelArgs &a, uint64_t n_frames, cudaStream_t stream, Sequences seqs);                     // REAL geometry, P = 2
bool wide_available(const sbx_plan *p);                                                                                 // sbx_dispatch_wide.cu
cudaError_t launch_fused_wide(sbx_plan *p, const KernelArgs &a, cudaStream_t stream);                                   // one stereo frame
sbx_status launch_colourise(sbx_plan *p, const float *d_mags, uint64_t n_frames, uint8_t *d_rgba, cudaStream_t stream);
sbx_status fail(const sbx_plan *p, sbx_status st, const std::string &msg);
sbx_status cuda_fail(const sbx_plan *p, cudaError_t e, const char *what);
}  // namespace sbx
