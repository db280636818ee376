// sbx_launch.h -- launch helpers shared by the dispatch translation units (sbx_dispatch_*.cu, sbx_plan.cu).  The
// kernels are split over several .cu files only so that nvcc can compile their template instances in parallel.
#pragma once
#include <atomic>
#include <cstdlib>

#include "sbx_internal.h"

namespace sbx {

constexpr uint32_t kMinLog2N = 8, kMaxLog2M = 14;                      // fused (single-SM) kernel: M <= 16384
constexpr uint32_t kMaxLog2NLarge = 16, kMaxLog2MLarge = 18;             // multi-kernel path of stft_large.cuh
constexpr size_t kMaxSmemOptIn = 227 * 1024;           // opt-in dynamic shared memory per CTA on sm_100

// Opt a kernel in to `bytes` of dynamic shared memory, once per device.  `done` is the kernel's own flag array
// (one static per template instance); the attribute call is idempotent, so two host threads racing on the first
// launch only repeat it.
constexpr int kMaxDevices = 64;
template <class Kernel> cudaError_t allow_smem(Kernel kernel, size_t bytes, std::atomic<bool> (&done)[kMaxDevices])
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
    if (done[dev].load(std::memory_order_acquire)) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) done[dev].store(true, std::memory_order_release);
    return e;
}

// seqs.count > 1: that many equally long sequences of n_frames frames each in one grid (KernelArgs::units_per_seq)
template <class G>
cudaError_t launch_geom(sbx_plan *p, KernelArgs args, uint64_t n_frames, cudaStream_t stream, Sequences seqs = {})
{
    constexpr int CH = G::CH;
    static std::atomic<bool> configured[kMaxDevices], configured_p[kMaxDevices];
    cudaError_t e = allow_smem(fused_frames_kernel<G>, G::SMEM_BYTES, configured);
    if (e != cudaSuccess) return e;
    const uint64_t units = (n_frames + G::FRAMES_PER_CTA - 1) / G::FRAMES_PER_CTA;
    const uint64_t ctas = units * seqs.count;
    if (ctas == 0) return cudaSuccess;
    if (ctas > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    if (seqs.count > 1) {
        args.units_per_seq = (uint32_t)units;
        args.seq_stride = seqs.stride_floats;
        fused_frames_kernel<G><<<(unsigned)ctas, G::THREADS, G::SMEM_BYTES, stream>>>(args);
        return cudaGetLastError();
    }

    // Persistent, TMA-staged variant: grid = SMs x resident CTAs, the next unit's PCM span prefetched into shared
    // memory by cp.async.bulk while the current unit is transformed.  Used where only ONE CTA fits an SM
    // (M = 16384: 128 KB of transform buffer), i.e. where no second CTA hides the load latency of pass 0:
    // measured on a B200, cfg3 19.9 -> 21.8 M frames/s.  With two or more resident CTAs the one-CTA-per-unit
    // grid is faster (cfg1: 51.3 vs 47.7 M frames/s; staggering the start of co-resident persistent CTAs by
    // 3-12 k cycles changed nothing, so it is not a lock-step effect -- the extra 18 KB of staging per CTA
    // leaves less L1 for the plan tables, and there is one more barrier per unit).
    // Needs 16-byte aligned sample spans and enough units to fill the machine.  (Never for the WIDE single-frame
    // geometries.)
    if constexpr (!G::WIDE) {
    constexpr uint32_t FPC = G::FRAMES_PER_CTA;
    const uint32_t unit_stride_bytes = FPC * args.hop * CH * 4u;
    const uint32_t span_bytes = (G::N + (FPC - 1) * args.hop) * CH * 4u;
    static const bool disabled = std::getenv("SBX_NO_PERSISTENT") != nullptr;
    const bool aligned = (unit_stride_bytes % 16u) == 0 && (span_bytes % 16u) == 0 &&
                         (reinterpret_cast<uintptr_t>(args.pcm) % 16u) == 0;
    const size_t smem_p = (size_t)G::SMEM_BYTES + span_bytes;
    constexpr size_t kPersistentDynMax = kMaxSmemOptIn - 64;     // the kernel also has a static mbarrier
    if (!disabled && aligned && smem_p <= kPersistentDynMax) {
        // the opt-in is the device maximum (the staging span depends on the plan's hop); occupancy and grid are
        // per plan and channel layout
        e = allow_smem(fused_frames_persistent_kernel<G>, kPersistentDynMax, configured_p);
        if (e != cudaSuccess) return e;
        sbx_plan::PersistentGeom &pg = p->pgeom[G::REAL ? 2 : CH - 1];
        if (!pg.ready) {
            int occ = 0, sms = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fused_frames_persistent_kernel<G>, G::THREADS, smem_p);
            if (e != cudaSuccess) return e;
            e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->cfg.device);
            if (e != cudaSuccess) return e;
            pg.ctas = occ * sms;
            pg.occ = occ;
            pg.smem = smem_p;
            pg.ready = true;
        }
        static const bool forced = std::getenv("SBX_FORCE_PERSISTENT") != nullptr;
        if (pg.ctas > 0 && ctas >= (uint64_t)pg.ctas * 2 && (pg.occ == 1 || forced)) {
            fused_frames_persistent_kernel<G><<<(unsigned)pg.ctas, G::THREADS, pg.smem, stream>>>(args, (uint32_t)ctas);
            return cudaGetLastError();
        }
    }
    }  // !G::WIDE
    fused_frames_kernel<G><<<(unsigned)ctas, G::THREADS, G::SMEM_BYTES, stream>>>(args);
    return cudaGetLastError();
}

}  // namespace sbx
