// sbx_live.cu -- the live path: a pinned-host SPSC ring of (l,r) samples with the cursor semantics
// of AudioStreamTransform::process (src/fourier/audio_transform.rs:34-42) and one captured CUDA
// graph per frame: the fused kernel reading the pinned window and writing the pinned column directly
// (zero-copy), or H2D copy -> kernel -> D2H copy where staging through device memory is faster.
//
// Producer = the audio callback thread (src/devices/audio_input_list_model.rs:63-79), consumer =
// the draw thread; exactly one of each, as in the reference (HeapRb::new(4096).split()).
#include <time.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <new>

#include "sbx_internal.h"

struct sbx_live {
    sbx_plan *plan = nullptr;
    int device = 0;                        // copied from the plan: sbx_live_destroy must not touch the plan (it may be gone)
    uint32_t capacity = 0;                 // samples
    float *ring = nullptr;                 // pinned, capacity x (l,r)
    std::atomic<uint64_t> head{0};         // total samples consumed
    std::atomic<uint64_t> tail{0};         // total samples produced
    float *h_in = nullptr;                 // pinned linearised window, N x 2
    float *h_mags = nullptr;               // pinned K x 2
    uint8_t *h_col = nullptr;              // pinned H x 4
    float *d_in = nullptr;
    float *d_mags = nullptr;
    uint8_t *d_col = nullptr;
    cudaStream_t stream = nullptr;
    cudaGraphExec_t graph_col = nullptr, graph_mags = nullptr;
    bool zero_copy = true;                 // kernel reads/writes the pinned host buffers directly (one graph node)
    // completion by polling (zero-copy graphs only): the kernel stores a launch counter to *h_flag when the column is
    // visible; the consumer spins on it instead of paying cudaStreamSynchronize's completion path
    bool poll = false;
    unsigned int *h_flag = nullptr;        // pinned
    unsigned int *d_count = nullptr;       // device
    unsigned int seq = 0;                  // flagged launches issued so far
};

using sbx::cuda_fail;
using sbx::fail;

#define SBX_CUDA(plan, call)                                              \
    do {                                                                  \
        cudaError_t e__ = (call);                                         \
        if (e__ != cudaSuccess) return cuda_fail((plan), e__, #call);     \
    } while (0)

static sbx_status build_graph(sbx_live *lv, bool column, cudaGraphExec_t *out)
{
    sbx_plan *p = lv->plan;
    cudaGraph_t graph = nullptr;
    SBX_CUDA(p, cudaStreamBeginCapture(lv->stream, cudaStreamCaptureModeThreadLocal));
    cudaError_t e = cudaSuccess;
    sbx_status st = SBX_OK;
    if (lv->zero_copy) {
        // pinned host memory is device-addressable (UVA): the 16 KB window is read and the 4 KB column written
        // over PCIe by the kernel itself -- one graph node, no copy-engine hand-offs
        if (lv->poll) { p->live_done_flag = lv->h_flag; p->live_done_count = lv->d_count; }
        st = sbx::launch_fused(p, 2, lv->h_in, 1, column ? lv->h_col : nullptr, column ? nullptr : lv->h_mags, lv->stream);
        p->live_done_flag = p->live_done_count = nullptr;
    } else {
        e = cudaMemcpyAsync(lv->d_in, lv->h_in, (size_t)p->n * 2 * sizeof(float), cudaMemcpyHostToDevice, lv->stream);
        if (e == cudaSuccess)
            st = sbx::launch_fused(p, 2, lv->d_in, 1, column ? lv->d_col : nullptr, column ? nullptr : lv->d_mags, lv->stream);
        if (e == cudaSuccess && st == SBX_OK) {
            e = column ? cudaMemcpyAsync(lv->h_col, lv->d_col, (size_t)p->rows * 4, cudaMemcpyDeviceToHost, lv->stream)
                       : cudaMemcpyAsync(lv->h_mags, lv->d_mags, (size_t)p->k_out * 2 * sizeof(float), cudaMemcpyDeviceToHost, lv->stream);
        }
    }
    cudaError_t e_end = cudaStreamEndCapture(lv->stream, &graph);
    if (st != SBX_OK) { if (graph) cudaGraphDestroy(graph); return st; }
    if (e != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return cuda_fail(p, e, "live graph capture"); }
    if (e_end != cudaSuccess) return cuda_fail(p, e_end, "cudaStreamEndCapture");
    e = cudaGraphInstantiate(out, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return cuda_fail(p, e, "cudaGraphInstantiate");
    p->launches -= 1;                       // the captured launch did not execute
    return SBX_OK;
}

extern "C" {

SBX_API void sbx_live_destroy(sbx_live *lv)
{
    if (!lv) return;
    cudaSetDevice(lv->device);              // never lv->plan: a caller may have destroyed the plan first
    if (lv->stream) cudaStreamSynchronize(lv->stream);
    if (lv->graph_col) cudaGraphExecDestroy(lv->graph_col);
    if (lv->graph_mags) cudaGraphExecDestroy(lv->graph_mags);
    if (lv->ring) cudaFreeHost(lv->ring);
    if (lv->h_in) cudaFreeHost(lv->h_in);
    if (lv->h_mags) cudaFreeHost(lv->h_mags);
    if (lv->h_col) cudaFreeHost(lv->h_col);
    if (lv->h_flag) cudaFreeHost(lv->h_flag);
    cudaFree(lv->d_count);
    cudaFree(lv->d_in); cudaFree(lv->d_mags); cudaFree(lv->d_col);
    if (lv->stream) cudaStreamDestroy(lv->stream);
    delete lv;
}

SBX_API sbx_status sbx_live_create(sbx_plan *p, uint32_t ring_capacity_samples, sbx_live **out_live)
{
    if (out_live) *out_live = nullptr;
    if (!p || !out_live) return SBX_ERR_INVALID_ARGUMENT;
    if (ring_capacity_samples == 0) return fail(p, SBX_ERR_INVALID_ARGUMENT, "ring capacity must be >= 1");
    SBX_CUDA(p, cudaSetDevice(p->cfg.device));
    sbx_live *lv = new (std::nothrow) sbx_live();
    if (!lv) return fail(p, SBX_ERR_OUT_OF_MEMORY, "out of host memory");
    lv->plan = p;
    lv->device = p->cfg.device;
    lv->capacity = ring_capacity_samples;
    // SBX_LIVE_COPY=1 selects the three-node graph (H2D copy -> kernel -> D2H copy) instead of zero-copy access.
    // The multi-kernel large path always stages through device memory.  The mixed-radix path reads every sample
    // once (generic_load), so zero-copy pays there too: N = 2400 p50 26.9 us against 30.8 us with copy nodes.
    // Measured p50 on a B200, cfg2 (N 2048 stereo): 16.2 us zero-copy (WIDE geometry, polled completion; a plain
    // kernel launch instead of the one-node graph: 17.6 us), 31.8 us with copy nodes.
    const char *env = std::getenv("SBX_LIVE_COPY");
    lv->zero_copy = !(env && env[0] == '1') && !p->large;
    // SBX_LIVE_POLL=0: wait with cudaStreamSynchronize instead of polling the completion flag
    const char *env_poll = std::getenv("SBX_LIVE_POLL");
    lv->poll = lv->zero_copy && !(env_poll && env_poll[0] == '0');
    cudaError_t e = cudaMallocHost(&lv->ring, (size_t)lv->capacity * 2 * sizeof(float));
    if (e == cudaSuccess) e = cudaMallocHost(&lv->h_in, (size_t)p->n * 2 * sizeof(float));
    if (e == cudaSuccess) e = cudaMallocHost(&lv->h_mags, (size_t)p->k_out * 2 * sizeof(float));
    if (e == cudaSuccess) e = cudaMallocHost(&lv->h_col, (size_t)p->rows * 4);
    if (e == cudaSuccess) e = cudaMalloc(&lv->d_in, (size_t)p->n * 2 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&lv->d_mags, (size_t)p->k_out * 2 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&lv->d_col, (size_t)p->rows * 4);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&lv->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess && lv->poll) {
        e = cudaMallocHost(&lv->h_flag, sizeof(unsigned int));
        if (e == cudaSuccess) { *lv->h_flag = 0; e = cudaMalloc(&lv->d_count, sizeof(unsigned int)); }
        if (e == cudaSuccess) e = cudaMemset(lv->d_count, 0, sizeof(unsigned int));
    }
    if (e != cudaSuccess) { sbx_live_destroy(lv); return cuda_fail(p, e, "live buffers"); }
    // one eager launch configures the kernel attributes outside of stream capture
    std::memset(lv->h_in, 0, (size_t)p->n * 2 * sizeof(float));
    e = cudaMemcpyAsync(lv->d_in, lv->h_in, (size_t)p->n * 2 * sizeof(float), cudaMemcpyHostToDevice, lv->stream);
    sbx_status st = e == cudaSuccess ? sbx::launch_fused(p, 2, lv->d_in, 1, lv->d_col, lv->d_mags, lv->stream) : SBX_ERR_CUDA;
    if (st == SBX_OK) e = cudaStreamSynchronize(lv->stream);
    if (st == SBX_OK && e == cudaSuccess) st = build_graph(lv, true, &lv->graph_col);
    if (st == SBX_OK && e == cudaSuccess) st = build_graph(lv, false, &lv->graph_mags);
    if (e != cudaSuccess) st = cuda_fail(p, e, "live warm-up");
    if (st != SBX_OK) { sbx_live_destroy(lv); return st; }
    *out_live = lv;
    return SBX_OK;
}

SBX_API uint32_t sbx_live_buffered(const sbx_live *lv)
{
    if (!lv) return 0;
    return (uint32_t)(lv->tail.load(std::memory_order_acquire) - lv->head.load(std::memory_order_acquire));
}

// push_iter: writes as many samples as fit, drops the rest (ringbuf Producer::push_iter)
SBX_API uint32_t sbx_live_push(sbx_live *lv, const float *samples, uint32_t n_samples, uint32_t channels)
{
    if (!lv || !samples || (channels != 1 && channels != 2)) return 0;
    const uint64_t tail = lv->tail.load(std::memory_order_relaxed);
    const uint64_t head = lv->head.load(std::memory_order_acquire);
    const uint32_t free_slots = lv->capacity - (uint32_t)(tail - head);
    const uint32_t n = n_samples < free_slots ? n_samples : free_slots;
    // at most two contiguous runs (the ring wraps once); no per-sample modulo
    const uint32_t start = (uint32_t)(tail % lv->capacity);
    const uint32_t first = lv->capacity - start < n ? lv->capacity - start : n;
    auto put = [&](float *dst, const float *src, uint32_t count) {
        if (channels == 2) {
            std::memcpy(dst, src, (size_t)count * 2 * sizeof(float));           // audio_input_list_model.rs:70
        } else {
            for (uint32_t i = 0; i < count; ++i) { dst[2 * i] = src[i]; dst[2 * i + 1] = src[i]; }   // (s, s), :67
        }
    };
    put(lv->ring + 2 * (size_t)start, samples, first);
    if (first < n) put(lv->ring, samples + (size_t)first * channels, n - first);
    lv->tail.store(tail + n, std::memory_order_release);
    return n;
}

static sbx_status live_pop(sbx_live *lv, uint8_t *rgba, float *out_lr)
{
    if (!lv) return SBX_ERR_INVALID_ARGUMENT;
    sbx_plan *p = lv->plan;
    const uint64_t head = lv->head.load(std::memory_order_relaxed);
    const uint64_t tail = lv->tail.load(std::memory_order_acquire);
    const uint32_t avail = (uint32_t)(tail - head);
    // input_stream.skip(stride_samples) runs whether or not the frame was produced (audio_transform.rs:38-39)
    const uint32_t skip = p->hop < avail ? p->hop : avail;
    if (avail < p->n) {
        lv->head.store(head + skip, std::memory_order_release);
        return SBX_NOT_ENOUGH_SAMPLES;
    }
    // linearise the window (non-consuming iter(), audio_transform.rs:38)
    const uint32_t start = (uint32_t)(head % lv->capacity);
    const uint32_t first = lv->capacity - start < p->n ? lv->capacity - start : p->n;
    std::memcpy(lv->h_in, lv->ring + 2 * (size_t)start, (size_t)first * 2 * sizeof(float));
    if (first < p->n) std::memcpy(lv->h_in + 2 * (size_t)first, lv->ring, (size_t)(p->n - first) * 2 * sizeof(float));
    SBX_CUDA(p, cudaSetDevice(p->cfg.device));
    SBX_CUDA(p, cudaGraphLaunch(rgba ? lv->graph_col : lv->graph_mags, lv->stream));
    p->launches += 1;
    bool done = false;
    if (lv->poll) {
        // the kernel stores the launch counter once the column is visible in pinned memory; a lost or faulted launch
        // falls through to the stream synchronize below, which reports it
        const unsigned int expect = ++lv->seq;
        volatile unsigned int *flag = lv->h_flag;
        timespec t0;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        for (uint32_t spins = 0; !done; ++spins) {
            if (*flag == expect) { done = true; break; }
            if ((spins & 0x3ffu) == 0x3ffu) {
                timespec t1;
                clock_gettime(CLOCK_MONOTONIC, &t1);
                if ((t1.tv_sec - t0.tv_sec) * 1000000000ll + (t1.tv_nsec - t0.tv_nsec) > 2000000ll) break;   // 2 ms
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
    }
    if (!done) {
        SBX_CUDA(p, cudaStreamSynchronize(lv->stream));
        if (lv->poll) lv->seq = *lv->h_flag;          // resynchronise the counter after a timeout
    }
    if (rgba) std::memcpy(rgba, lv->h_col, (size_t)p->rows * 4);
    else std::memcpy(out_lr, lv->h_mags, (size_t)p->k_out * 2 * sizeof(float));
    lv->head.store(head + skip, std::memory_order_release);
    return SBX_OK;
}

SBX_API sbx_status sbx_live_pop_column(sbx_live *lv, uint8_t *rgba)
{
    if (!lv || !rgba) return SBX_ERR_INVALID_ARGUMENT;
    return live_pop(lv, rgba, nullptr);
}

SBX_API sbx_status sbx_live_measure(sbx_live *lv, const float *samples, uint32_t n_samples, uint32_t channels,
                                    uint32_t iters, float *latencies_us)
{
    if (!lv || !samples || !latencies_us) return SBX_ERR_INVALID_ARGUMENT;
    std::vector<uint8_t> col((size_t)lv->plan->rows * 4);
    for (uint32_t i = 0; i < iters; ++i) {
        timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        sbx_live_push(lv, samples, n_samples, channels);
        const sbx_status st = live_pop(lv, col.data(), nullptr);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        if (st != SBX_OK) return st;
        latencies_us[i] = (float)((double)(t1.tv_sec - t0.tv_sec) * 1e6 + (double)(t1.tv_nsec - t0.tv_nsec) * 1e-3);
    }
    return SBX_OK;
}

SBX_API sbx_status sbx_live_pop_magnitudes(sbx_live *lv, float *out_lr)
{
    if (!lv || !out_lr) return SBX_ERR_INVALID_ARGUMENT;
    return live_pop(lv, nullptr, out_lr);
}

}  // extern "C"
