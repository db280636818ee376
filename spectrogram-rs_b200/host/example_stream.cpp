// example_stream.cpp -- the reference's live loop written against the C++ mirror:
// ring -> AudioStreamTransform<GpuFrameTransform>::process() -> RGBA columns.
//   g++ -std=c++17 example_stream.cpp -L.. -lsbx_b200 -Wl,-rpath,'$ORIGIN/..' -o example_stream
// Prints "frames=<n> k=<K> peak_bin=<j> peak=<L>" for a 1 kHz tone; exit code 0 on success.
#include <cmath>
#include <cstdio>

#include "audio_transform.hpp"

int main()
{
    try {
        const float sr = 48000.0f, period = 2048.0f / 48000.0f, stride = 512.0f / 48000.0f;
        sbx::HeapRb ring(4096);                                           // audio_input_list_model.rs:30
        std::vector<float> mono(4096);
        for (size_t i = 0; i < mono.size(); ++i) mono[i] = 0.5f * std::sin(6.283185307179586 * 1000.0 * i / sr);
        ring.push_mono(mono.data(), mono.size());
        sbx::AudioStreamTransform<sbx::GpuFrameTransform> fft(&ring, sbx::GpuFrameTransform(sr, period), stride);
        size_t peak_bin = 0;
        float peak = 0.0f;
        const size_t frames = fft.process([&](sbx::GpuFrameTransform::Output spectrum) {
            for (size_t j = 0; j < spectrum.size(); ++j)
                if (spectrum[j].first > peak) { peak = spectrum[j].first; peak_bin = j; }
        });
        std::vector<uint8_t> column(fft.transform.rows() * 4);
        std::vector<sbx::StereoMagnitude> win(2048, {0.25f, -0.25f});
        const bool ok = fft.transform.process_column(win.data(), win.size(), column.data());
        std::printf("frames=%zu k=%zu peak_bin=%zu peak=%.4f column=%d alpha=%u\n", frames,
                    fft.transform.num_output_frequencies(), peak_bin, peak, ok ? 1 : 0, column[3]);
        return frames == 5 && ok ? 0 : 1;
    } catch (const sbx::Error &e) {
        std::fprintf(stderr, "sbx error %d: %s\n", e.status, e.what());
        return 2;
    }
}
