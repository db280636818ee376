// sbx_dispatch_wide.cu -- single stereo frames (sbx_process_magnitudes / sbx_process_column / the live graphs) on the
// WIDE geometries: one CTA with up to four times the usual threads, because a lone frame is latency-bound and the
// split, resample and colour phases parallelise over rows and samples.  Measured on a B200, live p50 per frame:
// cfg2 (M = 4096, 256 -> 1024 threads) 23.7 -> 21.2 us; M = 8192 (512 -> 1024 threads) 18.7 -> 19.3 us, so the
// geometries stop at M = 4096.
#include "sbx_launch.h"

namespace sbx {

template <int LOG2N, int LOG2PAD>
cudaError_t launch_wide_one(sbx_plan *p, const KernelArgs &a, cudaStream_t stream)
{
    if constexpr (LOG2N + LOG2PAD < 9 || LOG2N + LOG2PAD > 12) {
        return cudaErrorInvalidConfiguration;           // M < 512: nothing to widen; M >= 8192: no gain (see above)
    } else {
        return launch_geom<Geom<LOG2N, LOG2PAD, 2, false, true>>(p, a, 1, stream);
    }
}

template <int LOG2N> cudaError_t launch_wide_pad(sbx_plan *p, const KernelArgs &a, cudaStream_t stream)
{
    switch (p->log2pad) {
        case 0: return launch_wide_one<LOG2N, 0>(p, a, stream);
        case 1: return launch_wide_one<LOG2N, 1>(p, a, stream);
        case 2: return launch_wide_one<LOG2N, 2>(p, a, stream);
        default: return cudaErrorInvalidConfiguration;
    }
}

bool wide_available(const sbx_plan *p)
{
    const int log2m = p->log2n + p->log2pad;
    return !p->generic && !p->large && p->log2n >= 8 && log2m >= 9 && log2m <= 12;
}

cudaError_t launch_fused_wide(sbx_plan *p, const KernelArgs &a, cudaStream_t stream)
{
    switch (p->log2n) {
        case 8: return launch_wide_pad<8>(p, a, stream);
        case 9: return launch_wide_pad<9>(p, a, stream);
        case 10: return launch_wide_pad<10>(p, a, stream);
        case 11: return launch_wide_pad<11>(p, a, stream);
        case 12: return launch_wide_pad<12>(p, a, stream);
        default: return cudaErrorInvalidConfiguration;
    }
}

}  // namespace sbx
