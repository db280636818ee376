// sbx_dispatch_large.cu -- launches of the multi-kernel path (stft_large.cuh, M > 16384) and of the mixed-radix
// path (stft_generic.cuh, windows that are not a power of two).
#include "sbx_launch.h"

namespace sbx {

// ---- generic path (window not a power of two) -----------------------------------------------------
sbx_status launch_generic(sbx_plan *p, int channels, const float *d_pcm, uint64_t n_frames, uint8_t *d_rgba,
                          float *d_mags, cudaStream_t stream)
{
    if (n_frames == 0) return SBX_OK;
    GenericArgs g = p->generic_passes;                          // radices, strides, table offsets
    g.k = p->base_args();
    g.k.pcm = d_pcm;
    g.k.rgba = d_rgba;
    g.k.mags_out = d_mags;
    g.k.n_frames = n_frames;
    if (n_frames == 1) { g.k.done_flag = p->live_done_flag; g.k.done_count = p->live_done_count; }
    g.twm = p->d_twm;
    g.twg = p->d_twg;
    g.pos_tab = p->d_pos_tab;
    g.bl_a = p->d_bl_a;
    g.bl_bf = p->d_bl_bf;
    g.bl_c = p->d_bl_c;
    // transform [M] (Bluestein: the convolution buffer [L]) + magnitudes [M/2], float2
    const size_t smem = p->bluestein ? (size_t)p->bl_len * 8 + (size_t)p->m * 4 : (size_t)p->m * 12;
    const uint64_t ctas = channels == 1 ? (n_frames + 1) / 2 : n_frames;
    if (ctas > 0x7fffffffull) return fail(p, SBX_ERR_INVALID_ARGUMENT, "too many frames in one launch");
    static std::atomic<bool> configured1[kMaxDevices], configured2[kMaxDevices], configured1b[kMaxDevices], configured2b[kMaxDevices];
    cudaError_t e;
    if (p->bluestein && channels == 1) {
        e = allow_smem(generic_frames_kernel<1, true>, kMaxSmemOptIn, configured1b);
        if (e == cudaSuccess) { generic_frames_kernel<1, true><<<(unsigned)ctas, GENERIC_THREADS, smem, stream>>>(g); e = cudaGetLastError(); }
    } else if (p->bluestein) {
        e = allow_smem(generic_frames_kernel<2, true>, kMaxSmemOptIn, configured2b);
        if (e == cudaSuccess) { generic_frames_kernel<2, true><<<(unsigned)ctas, GENERIC_THREADS, smem, stream>>>(g); e = cudaGetLastError(); }
    } else if (channels == 1) {
        e = allow_smem(generic_frames_kernel<1>, kMaxSmemOptIn, configured1);
        if (e == cudaSuccess) { generic_frames_kernel<1><<<(unsigned)ctas, GENERIC_THREADS, smem, stream>>>(g); e = cudaGetLastError(); }
    } else {
        e = allow_smem(generic_frames_kernel<2>, kMaxSmemOptIn, configured2);
        if (e == cudaSuccess) { generic_frames_kernel<2><<<(unsigned)ctas, GENERIC_THREADS, smem, stream>>>(g); e = cudaGetLastError(); }
    }
    if (e != cudaSuccess) return cuda_fail(p, e, "generic_frames_kernel launch");
    p->launches += 1;
    return SBX_OK;
}

// ---- large path (M > 16384): the cluster kernel + colour, or four kernels per chunk of units ----------
// Measured on a B200 (cfg5 geometries, M frames/s, four kernels vs cluster kernel + colour; profiles/README.md):
//   2-CTA clusters (M = 32768, stereo N 16384 P 2): 3.31 vs 3.55-3.67  -> cluster path by default
//   4-CTA clusters (M = 65536, mono N 16384 P 4):   3.36 vs 2.76;  stereo N 16384 P 4: 1.72 vs 1.34
//   8-CTA clusters (M = 131072, mono N 32768 P 4):  1.73 vs 1.28
// One 1024-thread CTA owns an SM (128 KB of shared memory), so nothing overlaps its load, barrier and store phases, and
// clusters of 4 / 8 leave 16 / 20 of the 148 SMs without work; the four-kernel path stays the default there.
// SBX_LARGE_CLUSTER=1 / 0 forces the cluster kernel wherever it exists (M <= 131072) / nowhere.
static bool cluster_path_enabled(int ctas)
{
    const char *e = std::getenv("SBX_LARGE_CLUSTER");           // read per launch: the tests compare both paths in one process
    if (e && e[0] == '0') return false;
    if (e && e[0] == '1') return true;
    return ctas == 2;
}

template <int LOG2N, int LOG2PAD, int CH>
cudaError_t launch_large_cluster(const LargeArgs &la, bool want_rgba, cudaStream_t stream)
{
    using GC = GeomC<LOG2N, LOG2PAD>;
    static std::atomic<bool> configured[kMaxDevices];
    cudaError_t e = allow_smem(large_cluster_kernel<LOG2N, LOG2PAD, CH>, GC::SMEM_BYTES, configured);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(la.units * GC::CTAS);
    cfg.blockDim = dim3(CLUSTER_THREADS);
    cfg.dynamicSmemBytes = GC::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = GC::CTAS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, large_cluster_kernel<LOG2N, LOG2PAD, CH>, la);
    if (e != cudaSuccess) return e;
    if (want_rgba)
        large_colour_kernel<CH == 1><<<dim3(la.units, la.k.n_chunks), LARGE_THREADS, 0, stream>>>(la, la.magbuf, (size_t)GeomL<LOG2N, LOG2PAD>::M / 2);
    return cudaGetLastError();
}

template <int LOG2N, int LOG2PAD, int CH>
cudaError_t launch_large_one(const LargeArgs &la, bool want_rgba, cudaStream_t stream)
{
    using GL = GeomL<LOG2N, LOG2PAD>;
    if constexpr (GL::Q / 4 <= 8) {
        if (cluster_path_enabled(GL::Q / 4)) return launch_large_cluster<LOG2N, LOG2PAD, CH>(la, want_rgba, stream);
    }
    const unsigned units = la.units;
    large_first_kernel<LOG2N, LOG2PAD, CH><<<units * GL::PAD * (LARGE_BLOCK / LARGE_THREADS), LARGE_THREADS, 0, stream>>>(la);
    large_block_kernel<<<units * (GL::M / LARGE_BLOCK), LARGE_THREADS, 0, stream>>>(la);
    static std::atomic<bool> split_configured[kMaxDevices];
    cudaError_t e = allow_smem(large_split_kernel<LOG2N, LOG2PAD, CH>, GL::SPLIT_SMEM, split_configured);
    if (e != cudaSuccess) return e;
    large_split_kernel<LOG2N, LOG2PAD, CH><<<units * (GL::Q / 2), LARGE_THREADS, GL::SPLIT_SMEM, stream>>>(la);
    if (want_rgba)
        large_colour_kernel<CH == 1><<<dim3(units, la.k.n_chunks), LARGE_THREADS, 0, stream>>>(la, la.magbuf, (size_t)GL::M / 2);
    return cudaGetLastError();
}

template <int LOG2N, int LOG2PAD>
cudaError_t launch_large_ch(int channels, const LargeArgs &la, bool want_rgba, cudaStream_t stream)
{
    if constexpr (LOG2N + LOG2PAD <= (int)kMaxLog2M || LOG2N + LOG2PAD > (int)kMaxLog2MLarge) {
        return cudaErrorInvalidConfiguration;
    } else {
        return channels == 1 ? launch_large_one<LOG2N, LOG2PAD, 1>(la, want_rgba, stream)
                             : launch_large_one<LOG2N, LOG2PAD, 2>(la, want_rgba, stream);
    }
}

template <int LOG2N>
cudaError_t launch_large_pad(int log2pad, int channels, const LargeArgs &la, bool want_rgba, cudaStream_t stream)
{
    switch (log2pad) {
        case 0: return launch_large_ch<LOG2N, 0>(channels, la, want_rgba, stream);
        case 1: return launch_large_ch<LOG2N, 1>(channels, la, want_rgba, stream);
        case 2: return launch_large_ch<LOG2N, 2>(channels, la, want_rgba, stream);
        default: return cudaErrorInvalidConfiguration;
    }
}

sbx_status launch_large(sbx_plan *p, int channels, const float *d_pcm, uint64_t n_frames, uint8_t *d_rgba,
                        float *d_mags, cudaStream_t stream)
{
    if (n_frames == 0) return SBX_OK;
    if (p->large_units == 0) {
        // scratch: (M + M/2) float2 per unit, SBX_LARGE_SCRATCH_MB (default 512) MiB per chunk of units.  Measured
        // on a B200 (cfg5, N = 65536): 512 MiB 0.68 M frames/s, 96 MiB (L2-resident) 0.57, 48 MiB 0.53 -- bigger
        // grids beat L2 residency of the hand-off between the four kernels.
        const size_t per_unit = (size_t)p->m * 12;
        const char *env = std::getenv("SBX_LARGE_SCRATCH_MB");
        const uint64_t budget_mb = env && std::atoi(env) > 0 ? (uint64_t)std::atoi(env) : 512u;
        uint64_t units = (budget_mb << 20) / per_unit;
        units = units < 1 ? 1 : (units > 8192 ? 8192 : units);
        cudaError_t e = cudaMalloc(&p->d_zbuf, units * (size_t)p->m * sizeof(float2));
        if (e == cudaSuccess) e = cudaMalloc(&p->d_magbuf, units * (size_t)(p->m / 2) * sizeof(float2));
        if (e != cudaSuccess) { cudaFree(p->d_zbuf); p->d_zbuf = nullptr; return cuda_fail(p, e, "cudaMalloc(large-transform scratch)"); }
        p->large_units = (uint32_t)units;
    }
    const uint32_t per = channels == 1 ? 2 : 1;
    const uint64_t total_units = (n_frames + per - 1) / per;
    LargeArgs la{};
    la.k = p->base_args();
    la.k.pcm = d_pcm;
    la.k.rgba = d_rgba;
    la.k.mags_out = d_mags;
    la.k.n_frames = n_frames;
    la.twp_block = p->d_twp_block;
    la.zbuf = p->d_zbuf;
    la.magbuf = p->d_magbuf;
    la.log2n = (uint32_t)p->log2n;
    la.log2pad = (uint32_t)p->log2pad;
    for (uint64_t u0 = 0; u0 < total_units; u0 += p->large_units) {
        la.frame_base = u0 * per;
        la.units = (uint32_t)(total_units - u0 < p->large_units ? total_units - u0 : p->large_units);
        cudaError_t e;
        switch (p->log2n) {
            case 13: e = launch_large_pad<13>(p->log2pad, channels, la, d_rgba != nullptr, stream); break;
            case 14: e = launch_large_pad<14>(p->log2pad, channels, la, d_rgba != nullptr, stream); break;
            case 15: e = launch_large_pad<15>(p->log2pad, channels, la, d_rgba != nullptr, stream); break;
            case 16: e = launch_large_pad<16>(p->log2pad, channels, la, d_rgba != nullptr, stream); break;
            default: e = cudaErrorInvalidConfiguration; break;
        }
        if (e != cudaSuccess) return cuda_fail(p, e, "large-transform kernels launch");
        p->launches += d_rgba != nullptr ? 4 : 3;
    }
    return SBX_OK;
}

}  // namespace sbx
