// audio_transform.hpp -- C++17 host-side mirror of the reference's frame-transform interface over the C ABI
// of include/sbx.h (the reference is Rust; no Rust toolchain exists in the build image -- INTEGRATION.md has
// the Rust shim).  Same names, argument meaning and error behaviour as
//   trait AudioTransform / struct AudioStreamTransform<T>   src/fourier/audio_transform.rs:4-42
//   FastFourierTransform                                    src/fourier/fft.rs:11-99
//   HeapRb::new(cap).split()                                src/devices/audio_input_list_model.rs:30
// Header-only; link with -lsbx_b200.  All arithmetic happens in the library, on the GPU.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/sbx.h"

namespace sbx {

using StereoMagnitude = std::pair<float, float>;   // src/fourier/mod.rs:13
using Period = float;                              // :14
using Frequency = float;                           // :15
static_assert(sizeof(StereoMagnitude) == 2 * sizeof(float), "(f32, f32) must be two packed floats");

struct Error : std::runtime_error {
    sbx_status status;
    Error(sbx_status st, const std::string &msg) : std::runtime_error(msg), status(st) {}
};

// Single-producer single-consumer ring of (l, r) samples -- ringbuf::HeapRb as the reference uses it.
class HeapRb {
   public:
    explicit HeapRb(size_t capacity) : buf_(capacity), cap_(capacity) {}
    // Producer::push_iter: pushes while there is room, returns the number pushed (the rest is dropped)
    template <class It> size_t push_iter(It first, It last)
    {
        const uint64_t tail = tail_.load(std::memory_order_relaxed), head = head_.load(std::memory_order_acquire);
        size_t room = cap_ - static_cast<size_t>(tail - head), n = 0;
        size_t at = static_cast<size_t>(tail % cap_);                       // running index: no per-sample modulo
        for (; first != last && n < room; ++first, ++n) {
            buf_[at] = *first;
            if (++at == cap_) at = 0;
        }
        tail_.store(tail + n, std::memory_order_release);
        return n;
    }
    size_t push_mono(const float *data, size_t n)            // audio_input_list_model.rs:66-69
    {
        std::vector<StereoMagnitude> tmp(n);
        for (size_t i = 0; i < n; ++i) tmp[i] = {data[i], data[i]};
        return push_iter(tmp.begin(), tmp.end());
    }
    size_t push_interleaved(const float *data, size_t n_frames)   // :70-72
    {
        const StereoMagnitude *p = reinterpret_cast<const StereoMagnitude *>(data);
        return push_iter(p, p + n_frames);
    }
    size_t len() const { return static_cast<size_t>(tail_.load(std::memory_order_acquire) - head_.load(std::memory_order_acquire)); }
    // Consumer::iter(): copies up to `max` samples WITHOUT consuming them
    size_t peek(StereoMagnitude *out, size_t max) const
    {
        const uint64_t head = head_.load(std::memory_order_relaxed);
        const size_t n = std::min(max, len());
        const size_t start = static_cast<size_t>(head % cap_), first = std::min(n, cap_ - start);
        std::copy(buf_.begin() + start, buf_.begin() + start + first, out);
        std::copy(buf_.begin(), buf_.begin() + (n - first), out + first);
        return n;
    }
    size_t skip(size_t n)                                     // Consumer::skip
    {
        n = std::min(n, len());
        head_.store(head_.load(std::memory_order_relaxed) + n, std::memory_order_release);
        return n;
    }

   private:
    std::vector<StereoMagnitude> buf_;
    size_t cap_;
    std::atomic<uint64_t> head_{0}, tail_{0};
};

// trait AudioTransform (src/fourier/audio_transform.rs:4-11) as a C++ "concept by convention":
//   using Output; Frequency sample_rate() const; size_t num_input_samples() const;
//   std::optional<Output> process(const StereoMagnitude* samples, size_t n_available);

// FastFourierTransform::new(sample_rate, period) drop-in (src/fourier/fft.rs:18-31)
class GpuFrameTransform {
   public:
    using Output = std::vector<StereoMagnitude>;

    GpuFrameTransform(Frequency sample_rate, Period period, const sbx_config *overrides = nullptr)
    {
        sbx_config cfg;
        if (overrides) cfg = *overrides; else sbx_config_default(&cfg);
        cfg.sample_rate = sample_rate;
        cfg.window_n = sbx_samples_from_seconds(period, sample_rate);      // (period * sample_rate) as usize
        const sbx_status st = sbx_plan_create(&cfg, &plan_);
        if (st != SBX_OK) throw Error(st, sbx_last_error(nullptr));
        sample_rate_ = sample_rate;
    }
    GpuFrameTransform(GpuFrameTransform &&o) noexcept : plan_(o.plan_), sample_rate_(o.sample_rate_) { o.plan_ = nullptr; }
    GpuFrameTransform &operator=(GpuFrameTransform &&o) noexcept
    {
        if (this != &o) { sbx_plan_destroy(plan_); plan_ = o.plan_; sample_rate_ = o.sample_rate_; o.plan_ = nullptr; }
        return *this;
    }
    GpuFrameTransform(const GpuFrameTransform &) = delete;
    GpuFrameTransform &operator=(const GpuFrameTransform &) = delete;
    ~GpuFrameTransform() { sbx_plan_destroy(plan_); }

    Frequency sample_rate() const { return sample_rate_; }
    size_t num_input_samples() const { return sbx_num_input_samples(plan_); }
    size_t num_output_frequencies() const { return sbx_num_output_frequencies(plan_); }   // fft.rs:33
    size_t rows() const { return sbx_rows(plan_); }
    sbx_plan *plan() const { return plan_; }

    // AudioTransform::process (fft.rs:43-99): nullopt when fewer than N samples are available (fft.rs:72)
    std::optional<Output> process(const StereoMagnitude *samples, size_t n_available)
    {
        Output out(num_output_frequencies());
        const sbx_status st = sbx_process_magnitudes(plan_, reinterpret_cast<const float *>(samples),
                                                     static_cast<uint32_t>(n_available), reinterpret_cast<float *>(out.data()));
        if (st == SBX_NOT_ENOUGH_SAMPLES) return std::nullopt;
        if (st != SBX_OK) throw Error(st, sbx_last_error(plan_));
        return out;
    }
    // the loop body of SimpleSpectrogram::snapshot (simple_spectrogram.rs:136-161): one RGBA8 column, row 0 = top
    bool process_column(const StereoMagnitude *samples, size_t n_available, uint8_t *rgba)
    {
        const sbx_status st = sbx_process_column(plan_, reinterpret_cast<const float *>(samples),
                                                 static_cast<uint32_t>(n_available), rgba);
        if (st == SBX_NOT_ENOUGH_SAMPLES) return false;
        if (st != SBX_OK) throw Error(st, sbx_last_error(plan_));
        return true;
    }

   private:
    sbx_plan *plan_ = nullptr;
    Frequency sample_rate_ = 0;
};

// struct AudioStreamTransform<T> (src/fourier/audio_transform.rs:14-42); public fields, as in the reference
template <class T> struct AudioStreamTransform {
    HeapRb *input_stream;
    T transform;
    Period stride;

    AudioStreamTransform(HeapRb *stream, T t, Period s) : input_stream(stream), transform(std::move(t)), stride(s) {}

    // process(): calls `sink(output)` once per frame while a whole window is buffered.  As in the reference,
    // skip(stride_samples) also runs after the failing attempt (:38-41).  Returns the number of frames.
    template <class Sink> size_t process(Sink &&sink)
    {
        const size_t stride_samples = sbx_samples_from_seconds(stride, transform.sample_rate());
        std::vector<StereoMagnitude> window(transform.num_input_samples());
        size_t frames = 0;
        for (;;) {
            const size_t got = input_stream->peek(window.data(), window.size());
            auto out = transform.process(window.data(), got);
            input_stream->skip(stride_samples);
            if (!out) return frames;
            sink(std::move(*out));
            ++frames;
        }
    }
};

// The pixbuf of SimpleSpectrogram (simple_spectrogram.rs:89-94) and its column cursor (:150-164): a row-major
// RGBA8 image whose column `offset` receives each new frame, then advances modulo the width.  Host side, for the
// live path's one-column-at-a-time flow; batches blit on the device with sbx_blit_columns_device.
struct ColumnStore {
    size_t width, height, rowstride, offset = 0;
    std::vector<uint8_t> pixels;

    ColumnStore(size_t w, size_t h) : width(w), height(h), rowstride(w * 4), pixels(w * h * 4, 0) {}

    // column: height x (r,g,b,a), row 0 = top (what process_column / sbx_live_pop_column produce)
    void put_column(const uint8_t *column)
    {
        for (size_t y = 0; y < height; ++y) std::memcpy(&pixels[y * rowstride + offset * 4], column + y * 4, 4);
        offset = (offset + 1) % width;                                              // :164
    }
};

// A whole recording on every GPU of the box through ONE call (no reference counterpart: the application is live-only;
// frames are independent, src/fourier/audio_transform.rs:37-41): time blocks with an N - hop halo, or whole planar
// channels, one plan and one host thread per device inside the library (sbx_multi_*).
class OfflineSpectrogram {
public:
    // devices: CUDA ordinals; empty = 0 .. n_devices-1
    OfflineSpectrogram(const sbx_config &cfg, uint32_t n_devices, const std::vector<int32_t> &devices = {})
    {
        const sbx_status st = sbx_multi_create(&cfg, devices.empty() ? nullptr : devices.data(), n_devices, &multi_);
        if (st != SBX_OK) throw Error(st, std::string("sbx_multi_create: ") + sbx_multi_last_error(nullptr));
        rows_ = sbx_rows(sbx_multi_plan(multi_, 0));
    }
    ~OfflineSpectrogram() { sbx_multi_destroy(multi_); }
    OfflineSpectrogram(const OfflineSpectrogram &) = delete;
    OfflineSpectrogram &operator=(const OfflineSpectrogram &) = delete;

    uint64_t num_frames(uint64_t n_samples) const { return sbx_num_frames(sbx_multi_plan(multi_, 0), n_samples); }

    // pcm: n_samples x channels (host; pinned or sbx_host_register'ed for overlapping copies) -> frames x H x 4 RGBA
    std::vector<uint8_t> process(const float *pcm, uint64_t n_samples)
    {
        std::vector<uint8_t> rgba((size_t)num_frames(n_samples) * rows_ * 4);
        uint64_t frames = 0;
        const sbx_status st = sbx_multi_process_batch(multi_, pcm, n_samples, rgba.data(), num_frames(n_samples), &frames);
        if (st != SBX_OK) throw Error(st, std::string("sbx_multi_process_batch: ") + sbx_multi_last_error(multi_));
        rgba.resize((size_t)frames * rows_ * 4);
        return rgba;
    }

private:
    sbx_multi *multi_ = nullptr;
    uint32_t rows_ = 0;
};

}  // namespace sbx
