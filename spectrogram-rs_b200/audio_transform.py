"""Host-side mirror of the reference's frame-transform interface, with the same names, argument
meaning and error behaviour, so parity tests read like tests of the reference itself:

    trait AudioTransform, struct AudioStreamTransform<T>   src/fourier/audio_transform.rs:4-42
    FastFourierTransform                                   src/fourier/fft.rs:11-99
    ColorScheme / default_color_schemes                    src/colorscheme.rs:22-151
    the column loop of SimpleSpectrogram::snapshot         src/widgets/simple_spectrogram.rs:136-164
    HeapRb::new(cap).split()                               src/devices/audio_input_list_model.rs:30

All arithmetic happens in libsbx_b200.so on the GPU; these classes only move buffers and cursors.
"""
import collections

import numpy as np

from . import binding
from .plan import Plan, builtin_gradient, builtin_lut, samples_from_seconds


class AudioTransform:
    """trait AudioTransform (src/fourier/audio_transform.rs:4-11)."""

    def sample_rate(self):
        raise NotImplementedError

    def num_input_samples(self):
        raise NotImplementedError

    def process(self, samples):
        raise NotImplementedError


class FastFourierTransform(AudioTransform):
    """FastFourierTransform::new(sample_rate, period) (src/fourier/fft.rs:18-31): window of
    (period * sample_rate) as usize samples, zero-padded 2x, C2C transform, L/R split.
    `pad_factor` generalises the hard-coded 2 (fft.rs:21,44)."""

    def __init__(self, sample_rate, period, pad_factor=2, device=0, **plan_kwargs):
        self._sample_rate = float(np.float32(sample_rate))
        self._period = float(np.float32(period))
        n = samples_from_seconds(self._period, self._sample_rate)
        self.plan = Plan(sample_rate=self._sample_rate, window_n=n, pad_factor=pad_factor, device=device, **plan_kwargs)

    def sample_rate(self):
        return self._sample_rate

    def num_input_samples(self):
        return self.plan.n

    def num_output_frequencies(self):
        return self.plan.k

    def process(self, samples):
        """samples: iterable of (l, r).  Consumes at most N of them; returns a (K, 2) float32 array of
        (L, R) magnitudes, or None if fewer than N samples were supplied (fft.rs:72)."""
        if not isinstance(samples, np.ndarray):
            samples = np.fromiter((v for s in samples for v in s), dtype=np.float32).reshape(-1, 2)
        return self.plan.process_magnitudes(samples[: self.plan.n])


class HeapRb:
    """Bounded SPSC ring of (l, r) samples: ringbuf::HeapRb as the reference uses it
    (src/devices/audio_input_list_model.rs:30,66-72).  push_iter drops what does not fit."""

    def __init__(self, capacity):
        self.capacity = int(capacity)
        self._q = collections.deque()

    def split(self):
        return self, self

    def push_iter(self, samples):
        pushed = 0
        for s in samples:
            if len(self._q) >= self.capacity:
                break
            self._q.append((float(s[0]), float(s[1])))
            pushed += 1
        return pushed

    def push_mono(self, data):
        return self.push_iter((s, s) for s in data)                         # :67-69

    def push_interleaved(self, data):
        it = iter(data)
        return self.push_iter(zip(it, it))                                  # :70-72

    def __len__(self):
        return len(self._q)

    def iter(self):
        return iter(self._q)

    def skip(self, n):
        n = min(n, len(self._q))
        for _ in range(n):
            self._q.popleft()
        return n


class AudioStreamTransform:
    """struct AudioStreamTransform<T> (src/fourier/audio_transform.rs:14-42); the three fields are
    public and mutated directly by the widgets (gpu_spectrogram.rs:34,321)."""

    def __init__(self, input_stream, transform, stride):
        self.input_stream = input_stream
        self.transform = transform
        self.stride = stride

    def process(self):
        """Yields one transform output per hop while a whole window is buffered.  As in the
        reference, skip(stride_samples) also runs after the failing (None) attempt (:38-41)."""
        stride_samples = samples_from_seconds(self.stride, self.transform.sample_rate())
        while True:
            n = self.transform.num_input_samples()
            window = np.fromiter((v for s, _ in zip(self.input_stream.iter(), range(n)) for v in s),
                                 dtype=np.float32).reshape(-1, 2)
            out = self.transform.process(window)
            self.input_stream.skip(stride_samples)
            if out is None:
                return
            yield out


class ColorScheme:
    """ColorScheme (src/colorscheme.rs:22-92).  The gradient is held as a table: 256 entries for sbx_config::lut_rgb
    and, for the stereo schemes (continuous diverging gradients indexed by pan), a finer one for
    sbx_plan_set_pan_gradient."""

    PAN_ENTRIES = 4096

    def __init__(self, lut_rgb, name, background=None, pan_rgb=None):
        self.lut = np.ascontiguousarray(lut_rgb, dtype=np.uint8).reshape(256, 3)
        self.pan_gradient = None if pan_rgb is None else np.ascontiguousarray(pan_rgb, dtype=np.uint8).reshape(-1, 3)
        self.name = name
        self._background = background

    @classmethod
    def new_mono(cls, gradient, name):
        return cls(builtin_lut(gradient) if isinstance(gradient, str) else gradient, name)

    @classmethod
    def new_stereo(cls, gradient, background, name):
        if isinstance(gradient, str):
            return cls(builtin_lut(gradient), name, tuple(background), builtin_gradient(gradient, cls.PAN_ENTRIES))
        return cls(gradient, name, tuple(background))

    @property
    def is_stereo(self):
        return self._background is not None

    def background(self):
        return self._background if self._background is not None else tuple(int(v) for v in self.lut[0])   # :41-44

    def foreground(self):
        return tuple(int(v) for v in (self.lut[255] if self._background is None else self.lut[128]))      # :46-53


def default_color_schemes():
    """default_color_schemes() (src/colorscheme.rs:125-151): all 19, in the reference's order and with its names."""
    black = (0, 0, 0)
    return [ColorScheme.new_stereo("red_yellow_blue", black, "Blue-Yellow-Red (Stereo)"),
            ColorScheme.new_mono("magma", "Magma"), ColorScheme.new_mono("viridis", "Viridis"),
            ColorScheme.new_stereo("red_blue", black, "Blue-Red (Stereo)"),
            ColorScheme.new_stereo("spectral", black, "Spectral (Stereo)"),
            ColorScheme.new_stereo("red_yellow_green", black, "Green-Yellow-Red (Stereo)"),
            ColorScheme.new_stereo("pink_green", black, "Green-Pink (Stereo)"),
            ColorScheme.new_stereo("purple_orange", black, "Orange-Purple (Stereo)"),
            ColorScheme.new_mono("inferno", "Inferno"), ColorScheme.new_mono("plasma", "Plasma"),
            ColorScheme.new_mono("cividis", "Cividis"), ColorScheme.new_mono("cubehelix", "Cube-helix"),
            ColorScheme.new_mono("turbo", "Turbo"), ColorScheme.new_mono("cool", "Cool"),
            ColorScheme.new_mono("reds", "Reds"), ColorScheme.new_mono("blues", "Blues"),
            ColorScheme.new_mono("greens", "Greens"), ColorScheme.new_mono("greys", "Greys"),
            ColorScheme.new_mono("oranges", "Oranges")]


class SpectrogramColumns:
    """The frame loop of SimpleSpectrogram::snapshot (src/widgets/simple_spectrogram.rs:136-164): a
    width x height RGBA8 image whose column `offset` receives each new frame, then advances
    modulo the width.  sample-rate changes re-plan with period 0.05 s (:214-219)."""

    TEXTURE_WIDTH = 1024
    TEXTURE_HEIGHT = 1024

    def __init__(self, sample_stream, sample_rate=48000.0, period=0.05, stride=2.0 / 1024.0, palette=None,
                 width=TEXTURE_WIDTH, height=TEXTURE_HEIGHT, window_n=None, pad_factor=2, device=0):
        self.palette = palette or ColorScheme.new_mono("magma", "magma")
        self.width, self.height = width, height
        self.buffer = np.zeros((height, width, 4), np.uint8)
        self.offset = 0
        self._period, self._pad, self._device = period, pad_factor, device
        n = window_n if window_n is not None else samples_from_seconds(np.float32(period), np.float32(sample_rate))
        self._window_n_override = window_n
        self.stride = stride
        self.input_stream = sample_stream
        self._make_plan(sample_rate, n)

    def _make_plan(self, sample_rate, n):
        hop = max(1, samples_from_seconds(np.float32(self.stride), np.float32(sample_rate)))
        old = getattr(self, "plan", None)
        if old is not None:
            old.close()                     # re-plan (simple_spectrogram.rs:214-219): release the previous plan now
        self.plan = Plan(sample_rate=sample_rate, window_n=n, pad_factor=self._pad, hop=hop, rows_h=self.height,
                         channels=2, lut=self.palette.lut, device=self._device,
                         scheme=binding.SCHEME_STEREO if self.palette.is_stereo else binding.SCHEME_MONO)
        if self.palette.is_stereo and self.palette.pan_gradient is not None:
            self.plan.set_pan_gradient(self.palette.pan_gradient)

    def set_sample_rate(self, sample_rate):
        n = self._window_n_override or samples_from_seconds(np.float32(self._period), np.float32(sample_rate))
        self._make_plan(float(sample_rate), n)

    def snapshot(self):
        """Drains the stream: one RGBA column per available frame.  Returns the number of columns drawn."""
        drawn = 0
        while True:
            window = np.fromiter((v for s, _ in zip(self.input_stream.iter(), range(self.plan.n)) for v in s),
                                 dtype=np.float32).reshape(-1, 2)
            col = self.plan.process_column(window)
            self.input_stream.skip(self.plan.hop)
            if col is None:
                return drawn
            self.buffer[:, self.offset, :] = col
            self.offset = (self.offset + 1) % self.width
            drawn += 1
