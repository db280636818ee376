// sbx_dispatch_q2.cu -- launches of fused_q2_kernel (stft_q2.cuh): pad factor 2, both sub-transforms of a unit in the
// two lanes of packed FP32 instructions, one CTA of N/16 threads per unit.
#include "sbx_launch.h"
#include "stft_q2.cuh"

namespace sbx {

template <class G>
static cudaError_t launch_q2_geom(KernelArgs args, uint64_t n_frames, cudaStream_t stream, Sequences seqs)
{
    static std::atomic<bool> configured[kMaxDevices];
    cudaError_t e = allow_smem(fused_q2_kernel<G>, G::SMEM_BYTES, configured);
    if (e != cudaSuccess) return e;
    const uint64_t units = (n_frames + G::FRAMES_PER_CTA - 1) / G::FRAMES_PER_CTA;
    const uint64_t ctas = units * seqs.count;
    if (ctas == 0) return cudaSuccess;
    if (ctas > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    if (seqs.count > 1) {
        args.units_per_seq = (uint32_t)units;
        args.seq_stride = seqs.stride_floats;
    }
    fused_q2_kernel<G><<<(unsigned)ctas, G::THREADS, G::SMEM_BYTES, stream>>>(args);
    return cudaGetLastError();
}

bool q2_available(const sbx_plan *p)
{
    return !p->generic && !p->large && p->log2pad == 1 && p->log2n >= 11 && p->log2n <= 13;
}

cudaError_t launch_fused_q2(sbx_plan *p, int channels, const KernelArgs &a, uint64_t n_frames, cudaStream_t stream, Sequences seqs)
{
    switch (p->log2n * 2 + (channels == 1 ? 0 : 1)) {
        case 11 * 2 + 0: return launch_q2_geom<Q2Geom<11, 1>>(a, n_frames, stream, seqs);
        case 11 * 2 + 1: return launch_q2_geom<Q2Geom<11, 2>>(a, n_frames, stream, seqs);
        case 12 * 2 + 0: return launch_q2_geom<Q2Geom<12, 1>>(a, n_frames, stream, seqs);
        case 12 * 2 + 1: return launch_q2_geom<Q2Geom<12, 2>>(a, n_frames, stream, seqs);
        case 13 * 2 + 0: return launch_q2_geom<Q2Geom<13, 1>>(a, n_frames, stream, seqs);
        case 13 * 2 + 1: return launch_q2_geom<Q2Geom<13, 2>>(a, n_frames, stream, seqs);
        default: return cudaErrorInvalidConfiguration;
    }
}

// real-sequence geometry (one mono frame per CTA as N/2 sample pairs) at pad factor 2: args as launch_real prepares them
bool q2_real_available(const sbx_plan *p)
{
    return p->real_ok && p->log2pad == 1 && p->log2n >= 12 && p->log2n <= 14;
}

cudaError_t launch_real_q2(sbx_plan *p, const KernelArgs &a, uint64_t n_frames, cudaStream_t stream, Sequences seqs)
{
    switch (p->log2n) {
        case 12: return launch_q2_geom<Q2Geom<11, 2, true>>(a, n_frames, stream, seqs);
        case 13: return launch_q2_geom<Q2Geom<12, 2, true>>(a, n_frames, stream, seqs);
        case 14: return launch_q2_geom<Q2Geom<13, 2, true>>(a, n_frames, stream, seqs);
        default: return cudaErrorInvalidConfiguration;
    }
}

}  // namespace sbx

#ifdef SBX_PHASE_CLOCK
// variant builds only (tools/phase_split.py): read and clear the per-phase cycle counters of fused_q2_kernel
extern "C" __attribute__((visibility("default"))) int sbx_debug_phase_cycles(unsigned long long out[8], int clear)
{
    if (cudaMemcpyFromSymbol(out, sbx::g_phase_cycles, 8 * sizeof(unsigned long long)) != cudaSuccess) return -1;
    if (clear) {
        const unsigned long long zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (cudaMemcpyToSymbol(sbx::g_phase_cycles, zero, sizeof(zero)) != cudaSuccess) return -1;
    }
    return 0;
}
#endif
