// sbx_dispatch_real.cu -- launches of fused_frames_kernel<Geom<.., REAL>>.
#include "sbx_launch.h"

namespace sbx {

// REAL geometry: one mono frame per CTA, transformed as the packed sequence of N/2 sample pairs (M/2 points).
// Wins where a frame PAIR no longer fits two CTAs per SM (M >= 16384) and makes M = 32768 a single-kernel path.
template <int LOG2NC, int LOG2PAD>
cudaError_t launch_real_one(sbx_plan *p, const KernelArgs &a, uint64_t n_frames, cudaStream_t stream, Sequences seqs)
{
    return launch_geom<Geom<LOG2NC, LOG2PAD, 2, true>>(p, a, n_frames, stream, seqs);
}

sbx_status launch_real(sbx_plan *p, const float *d_pcm, uint64_t n_frames, uint8_t *d_rgba, float *d_mags, cudaStream_t stream,
                       Sequences seqs)
{
    KernelArgs a = p->base_args();
    a.pcm = d_pcm;
    a.rgba = d_rgba;
    a.mags_out = d_mags;
    a.n_frames = n_frames;
    a.twp = p->d_twp_real;
    a.wq = p->d_wq_real;
    a.hop = p->hop / 2;                                    // in sample pairs
    cudaError_t e = cudaErrorInvalidConfiguration;
    if (p->use_q2_real) {
        e = launch_real_q2(p, a, n_frames, stream, seqs);
        if (e != cudaSuccess) return cuda_fail(p, e, "fused_q2_kernel (real-sequence geometry) launch");
        if (n_frames) p->launches += 1;
        return SBX_OK;
    }
    switch (p->log2n + p->log2pad) {
        case 13:
            if (p->log2pad == 0) e = launch_real_one<12, 0>(p, a, n_frames, stream, seqs);
            else if (p->log2pad == 1) e = launch_real_one<11, 1>(p, a, n_frames, stream, seqs);
            else if (p->log2pad == 2) e = launch_real_one<10, 2>(p, a, n_frames, stream, seqs);
            break;
        case 14:
            if (p->log2pad == 0) e = launch_real_one<13, 0>(p, a, n_frames, stream, seqs);
            else if (p->log2pad == 1) e = launch_real_one<12, 1>(p, a, n_frames, stream, seqs);
            else if (p->log2pad == 2) e = launch_real_one<11, 2>(p, a, n_frames, stream, seqs);
            break;
        case 15:
            if (p->log2pad == 0) e = launch_real_one<14, 0>(p, a, n_frames, stream, seqs);
            else if (p->log2pad == 1) e = launch_real_one<13, 1>(p, a, n_frames, stream, seqs);
            else if (p->log2pad == 2) e = launch_real_one<12, 2>(p, a, n_frames, stream, seqs);
            break;
        default: break;
    }
    if (e != cudaSuccess) return cuda_fail(p, e, "fused_frames_kernel (real-sequence geometry) launch");
    if (n_frames) p->launches += 1;
    return SBX_OK;
}

}  // namespace sbx
