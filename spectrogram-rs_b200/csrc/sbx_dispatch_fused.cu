// sbx_dispatch_fused.cu -- launches of fused_frames_kernel for the frame-pair (mono) and stereo geometries.
// SBX_FUSED_PART selects the window lengths this translation unit instantiates (0: N = 256 .. 1024, 1: N = 2048 ..
// 16384); build.py compiles the file once per part, in parallel.
#include "sbx_launch.h"

#ifndef SBX_FUSED_PART
#error "compile with -DSBX_FUSED_PART=0 or 1"
#endif

namespace sbx {

template <int LOG2N, int LOG2PAD>
cudaError_t launch_ch(sbx_plan *p, int channels, const KernelArgs &args, uint64_t n_frames, cudaStream_t stream, Sequences seqs)
{
    if constexpr (LOG2N + LOG2PAD > (int)kMaxLog2M) {
        return cudaErrorInvalidConfiguration;
    } else {
        return channels == 1 ? launch_geom<Geom<LOG2N, LOG2PAD, 1>>(p, args, n_frames, stream, seqs)
                             : launch_geom<Geom<LOG2N, LOG2PAD, 2>>(p, args, n_frames, stream, seqs);
    }
}

template <int LOG2N>
cudaError_t launch_pad(sbx_plan *p, int channels, const KernelArgs &args, uint64_t n_frames, cudaStream_t stream, Sequences seqs)
{
    switch (p->log2pad) {
        case 0: return launch_ch<LOG2N, 0>(p, channels, args, n_frames, stream, seqs);
        case 1: return launch_ch<LOG2N, 1>(p, channels, args, n_frames, stream, seqs);
        case 2: return launch_ch<LOG2N, 2>(p, channels, args, n_frames, stream, seqs);
        default: return cudaErrorInvalidConfiguration;
    }
}


#if SBX_FUSED_PART == 0
cudaError_t launch_fused_small(sbx_plan *p, int channels, const KernelArgs &a, uint64_t n_frames, cudaStream_t stream, Sequences seqs)
{
    switch (p->log2n) {
        case 8: return launch_pad<8>(p, channels, a, n_frames, stream, seqs);
        case 9: return launch_pad<9>(p, channels, a, n_frames, stream, seqs);
        case 10: return launch_pad<10>(p, channels, a, n_frames, stream, seqs);
        default: return cudaErrorInvalidConfiguration;
    }
}
#else
cudaError_t launch_fused_big(sbx_plan *p, int channels, const KernelArgs &a, uint64_t n_frames, cudaStream_t stream, Sequences seqs)
{
    switch (p->log2n) {
        case 11: return launch_pad<11>(p, channels, a, n_frames, stream, seqs);
        case 12: return launch_pad<12>(p, channels, a, n_frames, stream, seqs);
        case 13: return launch_pad<13>(p, channels, a, n_frames, stream, seqs);
        case 14: return launch_pad<14>(p, channels, a, n_frames, stream, seqs);      // P = 1 only (M <= 16384)
        default: return cudaErrorInvalidConfiguration;
    }
}
#endif

}  // namespace sbx
