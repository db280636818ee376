// sbx_multi.cu -- one HOST recording over several GPUs of one box (include/sbx.h, "several devices").
//
// Frame f depends only on samples [f*hop, f*hop + N) (src/fourier/audio_transform.rs:37-41, src/fourier/fft.rs:48-50):
// the recording splits into contiguous frame blocks, each with a trailing halo of N - hop samples it shares with its
// neighbour, and planar multi-channel input splits into whole channels.  One plan and one host thread per device,
// every thread runs the same chunk pipeline as sbx_process_batch on its own block; nothing is exchanged between
// devices (no collective).
#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "sbx_internal.h"

struct sbx_multi {
    std::vector<sbx_plan *> plans;
    bool mono = false;
    mutable std::string last_error;
};

namespace {

thread_local std::string g_multi_create_error;

// The plans of an sbx_multi live on different devices and every plan entry point makes its device current: the
// sbx_multi_* calls put the caller's current device back before they return (a host application -- or the CUDA
// runtime user above it, e.g. torch -- must not find its device switched by a library call).
struct DeviceGuard {
    int saved = -1;
    DeviceGuard() { if (cudaGetDevice(&saved) != cudaSuccess) { cudaGetLastError(); saved = -1; } }
    ~DeviceGuard() { if (saved >= 0) cudaSetDevice(saved); }
};

sbx_status mfail(const sbx_multi *m, sbx_status st, const std::string &msg)
{
    if (m) m->last_error = msg; else g_multi_create_error = msg;
    return st;
}

// Block boundaries as sharding.time_blocks(frame_align): device g of G owns frames [F*g/G, F*(g+1)/G), rounded down to
// a multiple of `align` (2 for mono plans: frames 2c and 2c+1 share one transform, so a block that started on an odd
// frame would pair every frame differently and match the unsharded call only up to the last bit).
void frame_block(uint64_t total, uint32_t g, uint32_t G, uint64_t align, uint64_t *lo, uint64_t *hi)
{
    // total * g cannot overflow for any recording that fits a host: total < 2^56, G <= 2^8
    *lo = (total * g / G) / align * align;
    *hi = g + 1 == G ? total : (total * (g + 1) / G) / align * align;
}

// runs fn(i) for every plan on its own host thread and returns the first failing status
template <class F> sbx_status for_each_device_impl(sbx_multi *m, F fn);
template <class F> sbx_status for_each_device(sbx_multi *m, F fn)
{
    try {
        return for_each_device_impl(m, fn);
    } catch (...) {                                             // bookkeeping allocations: report, never throw across the C ABI
        return mfail(m, SBX_ERR_OUT_OF_MEMORY, "out of host memory");
    }
}
template <class F> sbx_status for_each_device_impl(sbx_multi *m, F fn)
{
    const size_t G = m->plans.size();
    std::vector<sbx_status> st(G, SBX_OK);
    std::vector<std::thread> threads;
    threads.reserve(G);
    std::vector<char> spawned(G, 0);
    for (size_t i = 1; i < G; ++i) {
        try {
            threads.emplace_back([&, i] {
                try { st[i] = fn(i); } catch (...) { st[i] = SBX_ERR_OUT_OF_MEMORY; }
            });
            spawned[i] = 1;
        } catch (...) {                                         // no thread to be had: nothing may throw across the C ABI
        }
    }
    try { st[0] = fn(0); } catch (...) { st[0] = SBX_ERR_OUT_OF_MEMORY; }      // the caller's thread drives device 0
    for (auto &t : threads) t.join();
    for (size_t i = 1; i < G; ++i)
        if (!spawned[i]) {                                      // ... that device's block runs on the caller's thread instead
            try { st[i] = fn(i); } catch (...) { st[i] = SBX_ERR_OUT_OF_MEMORY; }
        }
    for (size_t i = 0; i < G; ++i)
        if (st[i] != SBX_OK)
            return mfail(m, st[i], "device " + std::to_string(m->plans[i]->cfg.device) + ": " + sbx_last_error(m->plans[i]));
    return SBX_OK;
}

}  // namespace

extern "C" {

SBX_API sbx_status sbx_multi_create(const sbx_config *cfg, const int32_t *devices, uint32_t n_devices, sbx_multi **out_multi)
{
    if (out_multi) *out_multi = nullptr;
    if (!cfg || !out_multi) return mfail(nullptr, SBX_ERR_INVALID_ARGUMENT, "cfg and out_multi must not be NULL");
    if (n_devices == 0 || n_devices > 256) return mfail(nullptr, SBX_ERR_INVALID_ARGUMENT, "n_devices must be in [1, 256]");
    for (uint32_t i = 0; devices && i < n_devices; ++i)
        for (uint32_t j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return mfail(nullptr, SBX_ERR_INVALID_ARGUMENT, "a device ordinal is listed twice");
    sbx_multi *m = new (std::nothrow) sbx_multi();
    if (!m) return mfail(nullptr, SBX_ERR_OUT_OF_MEMORY, "out of host memory");
    DeviceGuard guard;
    try {
        m->plans.reserve(n_devices);
    } catch (...) {
        delete m;
        return mfail(nullptr, SBX_ERR_OUT_OF_MEMORY, "out of host memory");
    }
    m->mono = cfg->channels == 1;
    for (uint32_t i = 0; i < n_devices; ++i) {
        sbx_config c = *cfg;
        c.device = devices ? devices[i] : (int32_t)i;          // NULL: devices 0 .. n_devices-1
        sbx_plan *p = nullptr;
        const sbx_status st = sbx_plan_create(&c, &p);
        if (st != SBX_OK) {
            const std::string why = sbx_last_error(nullptr);
            sbx_multi_destroy(m);
            return mfail(nullptr, st, "device " + std::to_string(c.device) + ": " + why);
        }
        m->plans.push_back(p);
    }
    *out_multi = m;
    return SBX_OK;
}

SBX_API void sbx_multi_destroy(sbx_multi *m)
{
    if (!m) return;
    DeviceGuard guard;
    for (sbx_plan *p : m->plans) sbx_plan_destroy(p);
    delete m;
}

SBX_API const char *sbx_multi_last_error(const sbx_multi *m)
{
    return m ? m->last_error.c_str() : g_multi_create_error.c_str();
}

SBX_API uint32_t sbx_multi_device_count(const sbx_multi *m) { return m ? (uint32_t)m->plans.size() : 0; }

SBX_API sbx_plan *sbx_multi_plan(sbx_multi *m, uint32_t index)
{
    return m && index < m->plans.size() ? m->plans[index] : nullptr;
}

SBX_API sbx_status sbx_shard_time_block(uint64_t n_samples, uint32_t window_n, uint32_t hop, uint32_t channels, uint32_t index,
                                        uint32_t n_shards, uint64_t *frame_lo, uint64_t *frame_hi, uint64_t *sample_lo,
                                        uint64_t *sample_hi)
{
    if (window_n == 0 || hop == 0 || n_shards == 0 || index >= n_shards || (channels != 1 && channels != 2))
        return SBX_ERR_INVALID_ARGUMENT;
    const uint64_t total = n_samples < window_n ? 0 : (n_samples - window_n) / hop + 1;    // audio_transform.rs:37-41
    if (total >> 56) return SBX_ERR_INVALID_ARGUMENT;
    uint64_t lo, hi;
    frame_block(total, index, n_shards, channels == 1 ? 2 : 1, &lo, &hi);
    if (frame_lo) *frame_lo = lo;
    if (frame_hi) *frame_hi = hi;
    if (sample_lo) *sample_lo = lo * hop;
    if (sample_hi) *sample_hi = hi > lo ? (hi - 1) * hop + window_n : lo * hop;            // includes the N - hop halo
    return SBX_OK;
}

SBX_API sbx_status sbx_multi_time_block(const sbx_multi *m, uint64_t n_samples, uint32_t index, uint64_t *frame_lo, uint64_t *frame_hi,
                                        uint64_t *sample_lo, uint64_t *sample_hi)
{
    if (!m || index >= m->plans.size()) return SBX_ERR_INVALID_ARGUMENT;
    const sbx_plan *p = m->plans[0];
    return sbx_shard_time_block(n_samples, p->n, p->hop, p->channels, index, (uint32_t)m->plans.size(), frame_lo, frame_hi,
                                sample_lo, sample_hi);
}

SBX_API sbx_status sbx_multi_process_batch(sbx_multi *m, const float *pcm, uint64_t n_samples, uint8_t *rgba,
                                           uint64_t rgba_capacity_frames, uint64_t *frames_out)
{
    if (frames_out) *frames_out = 0;
    if (!m) return SBX_ERR_INVALID_ARGUMENT;
    const sbx_plan *p0 = m->plans[0];
    const uint64_t frames = sbx_num_frames(p0, n_samples);
    if (frames == 0) return SBX_OK;
    if (!pcm || !rgba) return mfail(m, SBX_ERR_INVALID_ARGUMENT, "NULL host buffer");
    if (frames > rgba_capacity_frames) return mfail(m, SBX_ERR_BUFFER_TOO_SMALL, "rgba capacity is smaller than the frame count");
    const uint32_t G = (uint32_t)m->plans.size();
    DeviceGuard guard;
    const sbx_status st = for_each_device(m, [&](size_t i) -> sbx_status {
        sbx_plan *p = m->plans[i];
        uint64_t lo, hi;
        frame_block(frames, (uint32_t)i, G, m->mono ? 2 : 1, &lo, &hi);
        if (hi <= lo) return SBX_OK;
        const sbx::HostSegment seg{pcm + (size_t)lo * p->hop * p->channels, hi - lo, rgba + (size_t)lo * p->rows * 4};
        return sbx::batch_host_segments(p, &seg, 1);
    });
    if (st == SBX_OK && frames_out) *frames_out = frames;
    return st;
}

SBX_API sbx_status sbx_multi_process_batch_planar(sbx_multi *m, const float *pcm, uint32_t n_sequences, uint64_t sequence_stride_samples,
                                                  uint64_t n_samples_per_sequence, uint8_t *rgba, uint64_t rgba_capacity_frames,
                                                  uint64_t *frames_per_sequence_out)
{
    if (frames_per_sequence_out) *frames_per_sequence_out = 0;
    if (!m) return SBX_ERR_INVALID_ARGUMENT;
    const sbx_plan *p0 = m->plans[0];
    const uint64_t frames = sbx_num_frames(p0, n_samples_per_sequence);
    if (frames == 0 || n_sequences == 0) return SBX_OK;
    if (!pcm || !rgba) return mfail(m, SBX_ERR_INVALID_ARGUMENT, "NULL host buffer");
    if (n_sequences > 1 && sequence_stride_samples < n_samples_per_sequence)
        return mfail(m, SBX_ERR_INVALID_ARGUMENT, "sequence_stride_samples is smaller than n_samples_per_sequence");
    if (frames > rgba_capacity_frames / n_sequences)
        return mfail(m, SBX_ERR_BUFFER_TOO_SMALL, "output capacity is smaller than n_sequences * frames per sequence");
    const uint32_t G = (uint32_t)m->plans.size();
    DeviceGuard guard;
    const sbx_status st = for_each_device(m, [&](size_t i) -> sbx_status {
        sbx_plan *p = m->plans[i];
        // whole sequences per device (sharding.channel_blocks), no halo
        const uint32_t s_lo = (uint32_t)((uint64_t)n_sequences * i / G), s_hi = (uint32_t)((uint64_t)n_sequences * (i + 1) / G);
        std::vector<sbx::HostSegment> segs;
        segs.reserve(s_hi - s_lo);
        for (uint32_t s = s_lo; s < s_hi; ++s)
            segs.push_back({pcm + (size_t)s * sequence_stride_samples * p->channels, frames, rgba + (size_t)s * frames * p->rows * 4});
        return segs.empty() ? (sbx_status)SBX_OK : sbx::batch_host_segments(p, segs.data(), segs.size());
    });
    if (st == SBX_OK && frames_per_sequence_out) *frames_per_sequence_out = frames;
    return st;
}

SBX_API uint64_t sbx_multi_launch_count(const sbx_multi *m)
{
    uint64_t n = 0;
    if (m)
        for (const sbx_plan *p : m->plans) n += sbx_launch_count(p);
    return n;
}

SBX_API sbx_status sbx_host_register(void *ptr, size_t bytes)
{
    if (!ptr || !bytes) return SBX_ERR_INVALID_ARGUMENT;
    const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return e == cudaErrorMemoryAllocation ? SBX_ERR_OUT_OF_MEMORY : SBX_ERR_CUDA; }
    return SBX_OK;
}

SBX_API sbx_status sbx_host_unregister(void *ptr)
{
    if (!ptr) return SBX_ERR_INVALID_ARGUMENT;
    if (cudaHostUnregister(ptr) != cudaSuccess) { cudaGetLastError(); return SBX_ERR_CUDA; }
    return SBX_OK;
}

}  // extern "C"
