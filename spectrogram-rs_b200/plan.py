"""Thin RAII wrapper over the C ABI (include/sbx.h).  Every transform goes through libsbx_b200.so;
nothing here computes spectra on the CPU."""
import ctypes as C
import weakref

import numpy as np

from . import binding
from .binding import SbxConfig


class SbxError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"sbx status {status}: {message}")
        self.status = status


def builtin_gradient(name, entries=256):
    """(entries, 3) uint8 table of one of the reference's 19 gradients (src/colorscheme.rs:129-149), entry i =
    gradient((i + 0.5) / entries): 'magma', 'viridis', 'red_yellow_blue', 'turbo', ... (builtin_gradient_names())."""
    lut = np.empty(entries * 3, np.uint8)
    st = binding.load().sbx_builtin_gradient(name.encode(), entries, lut.ctypes.data)
    if st != binding.SBX_OK:
        raise SbxError(st, f"unknown built-in gradient {name!r} (or entries outside 1..65536)")
    return lut.reshape(entries, 3)


def builtin_lut(name):
    """256x3 uint8 table for sbx_config::lut_rgb."""
    return builtin_gradient(name, 256)


def builtin_gradient_names():
    """[(name, is_stereo_scheme)] in the order of default_color_schemes() (src/colorscheme.rs:129-149)."""
    lib = binding.load()
    out = []
    for i in range(lib.sbx_builtin_gradient_count()):
        stereo = C.c_int32()
        out.append((lib.sbx_builtin_gradient_name(i, C.byref(stereo)).decode(), bool(stereo.value)))
    return out


def samples_from_seconds(seconds, sample_rate):
    """(seconds * sample_rate) as usize in f32 -- fft.rs:19, audio_transform.rs:35."""
    return int(binding.load().sbx_samples_from_seconds(seconds, sample_rate))


def _ptr(x):
    """Device/host pointer of a numpy array, torch tensor, int or None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError(f"cannot take a pointer of {type(x)}")


class Plan:
    """One sbx_plan: replaces FastFourierTransform::new's FFTW plan (src/fourier/fft.rs:18-31) plus the
    constants of the column loop (src/widgets/simple_spectrogram.rs:107,136-161)."""

    def __init__(self, sample_rate=48000.0, window_n=4096, pad_factor=2, hop=512, rows_h=1024, channels=2,
                 f_min=32.0, f_max=22030.0, db_min=-70.0, db_max=-10.0, db_eps=1e-7, interp=binding.INTERP_CUBIC,
                 scheme=binding.SCHEME_MONO, lut="magma", device=0):
        self._lib = binding.load()
        self._h = None
        self._live = weakref.WeakSet()      # live streams hold graphs and buffers of this plan: closed before it
        cfg = SbxConfig()
        self._lib.sbx_config_default(C.byref(cfg))
        cfg.sample_rate, cfg.window_n, cfg.pad_factor, cfg.hop = sample_rate, window_n, pad_factor, hop
        cfg.rows_h, cfg.channels, cfg.f_min, cfg.f_max = rows_h, channels, f_min, f_max
        cfg.db_min, cfg.db_max, cfg.db_eps, cfg.interp, cfg.scheme, cfg.device = db_min, db_max, db_eps, interp, scheme, device
        lut_arr = builtin_lut(lut) if isinstance(lut, str) else np.ascontiguousarray(lut, dtype=np.uint8)
        if lut_arr.size != 768:
            raise ValueError("lut must be 256x3 uint8")
        C.memmove(cfg.lut_rgb, lut_arr.ctypes.data, 768)
        self.lut = lut_arr.reshape(256, 3).copy()
        self.config = cfg
        handle = C.c_void_p()
        st = self._lib.sbx_plan_create(C.byref(cfg), C.byref(handle))
        if st != binding.SBX_OK:
            raise SbxError(st, self._lib.sbx_last_error(None).decode())
        self._h = handle
        self.n = int(self._lib.sbx_num_input_samples(handle))
        self.k = int(self._lib.sbx_num_output_frequencies(handle))
        self.m = int(self._lib.sbx_fft_length(handle))
        self.rows = int(self._lib.sbx_rows(handle))
        self.hop = int(self._lib.sbx_hop(handle))
        self.channels = int(self._lib.sbx_channels(handle))
        self.sample_rate = float(self._lib.sbx_sample_rate(handle))

    # -- lifetime -----------------------------------------------------------------------------
    def close(self):
        if self._h is not None:
            for lv in list(self._live):     # a stream must never outlive its plan (sbx.h: lifetime of sbx_live)
                lv.close()
            self._lib.sbx_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self):
        return self._h

    def _check(self, st):
        if st < 0:
            raise SbxError(st, self._lib.sbx_last_error(self._h).decode())
        return st

    # -- geometry ----------------------------------------------------------------------------
    def num_frames(self, n_samples):
        return int(self._lib.sbx_num_frames(self._h, n_samples))

    # -- single frame (host) -----------------------------------------------------------------
    def process_magnitudes(self, lr):
        """FastFourierTransform::process: (n_avail, 2) f32 -> (K, 2) f32, or None when n_avail < N."""
        lr = np.ascontiguousarray(lr, dtype=np.float32).reshape(-1, 2)
        out = np.empty((self.k, 2), np.float32)
        st = self._check(self._lib.sbx_process_magnitudes(self._h, lr.ctypes.data, lr.shape[0], out.ctypes.data))
        return None if st == binding.SBX_NOT_ENOUGH_SAMPLES else out

    def process_column(self, lr):
        lr = np.ascontiguousarray(lr, dtype=np.float32).reshape(-1, 2)
        out = np.empty((self.rows, 4), np.uint8)
        st = self._check(self._lib.sbx_process_column(self._h, lr.ctypes.data, lr.shape[0], out.ctypes.data))
        return None if st == binding.SBX_NOT_ENOUGH_SAMPLES else out

    def colorize_magnitudes(self, mags_lr):
        mags_lr = np.ascontiguousarray(mags_lr, dtype=np.float32).reshape(-1, self.k, 2)
        out = np.empty((mags_lr.shape[0], self.rows, 4), np.uint8)
        self._check(self._lib.sbx_colorize_magnitudes(self._h, mags_lr.ctypes.data, mags_lr.shape[0], out.ctypes.data))
        return out

    # -- batch -------------------------------------------------------------------------------
    def process_batch(self, pcm, out=None):
        """Host PCM (numpy, or the address of pinned memory via (ptr, n_samples)) -> (frames, H, 4) uint8."""
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        n_samples = pcm.shape[0]
        frames = self.num_frames(n_samples)
        if out is None:
            out = np.empty((frames, self.rows, 4), np.uint8)
        got = C.c_uint64()
        self._check(self._lib.sbx_process_batch(self._h, pcm.ctypes.data, n_samples, out.ctypes.data, out.shape[0], C.byref(got)))
        return out[:got.value]

    def process_batch_ptr(self, pcm_ptr, n_samples, rgba_ptr, capacity_frames):
        got = C.c_uint64()
        self._check(self._lib.sbx_process_batch(self._h, _ptr(pcm_ptr), n_samples, _ptr(rgba_ptr), capacity_frames, C.byref(got)))
        return int(got.value)

    def process_batch_device(self, d_pcm, n_samples, d_rgba, capacity_frames):
        """Device pointers (torch CUDA tensors or ints); asynchronous on the plan's stream."""
        got = C.c_uint64()
        self._check(self._lib.sbx_process_batch_device(self._h, _ptr(d_pcm), n_samples, _ptr(d_rgba), capacity_frames, C.byref(got)))
        return int(got.value)

    def magnitudes_batch_device(self, d_pcm, n_samples, d_mags, capacity_frames):
        got = C.c_uint64()
        self._check(self._lib.sbx_magnitudes_batch_device(self._h, _ptr(d_pcm), n_samples, _ptr(d_mags), capacity_frames, C.byref(got)))
        return int(got.value)

    def magnitudes_batch_device_f16(self, d_pcm, n_samples, d_mags_f16, capacity_frames):
        """Rows as binary16 (L,R) pairs: the GL widget's F16F16 texel format (gpu_spectrogram.rs:218-227)."""
        got = C.c_uint64()
        self._check(self._lib.sbx_magnitudes_batch_device_f16(self._h, _ptr(d_pcm), n_samples, _ptr(d_mags_f16), capacity_frames, C.byref(got)))
        return int(got.value)

    def palette_table(self, resolution=32):
        """ColorScheme::lookup_table(resolution) (colorscheme.rs:73-92): (res, res, 4) float32."""
        out = np.empty((resolution, resolution, 4), np.float32)
        self._check(self._lib.sbx_palette_table(self._h, resolution, out.ctypes.data))
        return out

    def process_batch_device_planar(self, d_pcm, n_sequences, sequence_stride_samples, n_samples_per_sequence, d_rgba,
                                    capacity_frames):
        """n_sequences equally long recordings (planar channels) in one call; returns frames per sequence."""
        got = C.c_uint64(0)
        self._check(self._lib.sbx_process_batch_device_planar(self._h, _ptr(d_pcm), n_sequences, sequence_stride_samples,
                                                              n_samples_per_sequence, _ptr(d_rgba), capacity_frames, C.byref(got)))
        return int(got.value)

    def blit_columns_device(self, d_columns, n_frames, d_image, width, rowstride_bytes, first_column=0):
        """The column store of SimpleSpectrogram (simple_spectrogram.rs:141-164) on the device: column f of
        d_columns (frames x rows x 4) -> image column (first_column + f) % width of a row-major RGBA8 image.
        Returns the next column cursor.  Asynchronous on the plan's stream."""
        nxt = C.c_uint32(0)
        self._check(self._lib.sbx_blit_columns_device(self._h, _ptr(d_columns), n_frames, _ptr(d_image), width,
                                                      rowstride_bytes, first_column, C.byref(nxt)))
        return int(nxt.value)

    def set_pan_gradient(self, rgb):
        """Stereo schemes: a finer table for the pan lookup, (entries, 3) uint8 with entry i = gradient((i + 0.5) /
        entries) (sbx_plan_set_pan_gradient); e.g. builtin_gradient('red_yellow_blue', 4096)."""
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8).reshape(-1, 3)
        self._check(self._lib.sbx_plan_set_pan_gradient(self._h, rgb.ctypes.data, rgb.shape[0]))

    def set_stream(self, cuda_stream):
        self._check(self._lib.sbx_set_stream(self._h, cuda_stream))

    def synchronize(self):
        self._check(self._lib.sbx_synchronize(self._h))

    @property
    def launch_count(self):
        return int(self._lib.sbx_launch_count(self._h))

    @property
    def kernel_path(self):
        """'fused-pair[-q2]' | 'fused-real' | 'fused-stereo[-q2]' | 'large' | 'generic' (sbx_kernel_path)."""
        return self._lib.sbx_kernel_path(self._h).decode()

    # -- introspection -----------------------------------------------------------------------
    def bin_map(self):
        f0, step = np.empty(self.rows, np.float32), np.empty(self.rows, np.float32)
        n = np.empty(self.rows, np.uint32)
        period = C.c_float()
        fp, up = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
        self._check(self._lib.sbx_get_bin_map(self._h, f0.ctypes.data_as(fp), step.ctypes.data_as(fp),
                                              n.ctypes.data_as(up), C.byref(period)))
        return f0, step, n, np.float32(period.value)

    def color_thresholds(self):
        thr = np.empty(255, np.float32)
        self._check(self._lib.sbx_get_color_thresholds(self._h, thr.ctypes.data_as(C.POINTER(C.c_float))))
        return thr

    def window(self):
        w = np.empty(self.n, np.float32)
        self._check(self._lib.sbx_get_window(self._h, w.ctypes.data_as(C.POINTER(C.c_float))))
        return w

    # -- live --------------------------------------------------------------------------------
    def live(self, ring_capacity=4096):
        return LiveStream(self, ring_capacity)


class _PlanView(Plan):
    """A plan owned by an sbx_multi: same accessors and device entry points, no ownership."""

    def __init__(self, lib, handle):                          # noqa: D401 -- deliberately skips Plan.__init__
        self._lib, self._h = lib, C.c_void_p(handle)
        self._live = weakref.WeakSet()
        self.n = int(lib.sbx_num_input_samples(self._h))
        self.k = int(lib.sbx_num_output_frequencies(self._h))
        self.m = int(lib.sbx_fft_length(self._h))
        self.rows = int(lib.sbx_rows(self._h))
        self.hop = int(lib.sbx_hop(self._h))
        self.channels = int(lib.sbx_channels(self._h))
        self.sample_rate = float(lib.sbx_sample_rate(self._h))

    def close(self):
        self._h = None


class MultiPlan:
    """sbx_multi: one plan per device of one box; a HOST recording is sharded by time block (N - hop halo, no
    exchange between devices) or, for planar channels, by whole channels (SURVEY.md 8e;
    src/fourier/audio_transform.rs:37-41)."""

    def __init__(self, devices, **plan_kwargs):
        self._lib = binding.load()
        self._h = None
        lut = plan_kwargs.pop("lut", "magma")
        cfg = SbxConfig()
        self._lib.sbx_config_default(C.byref(cfg))
        for key, val in plan_kwargs.items():
            if not hasattr(cfg, key):
                raise TypeError(f"unknown plan parameter {key!r}")
            setattr(cfg, key, val)
        lut_arr = builtin_lut(lut) if isinstance(lut, str) else np.ascontiguousarray(lut, dtype=np.uint8)
        C.memmove(cfg.lut_rgb, lut_arr.ctypes.data, 768)
        devs = list(range(devices)) if isinstance(devices, int) else [int(d) for d in devices]
        arr = (C.c_int32 * len(devs))(*devs)
        handle = C.c_void_p()
        st = self._lib.sbx_multi_create(C.byref(cfg), arr, len(devs), C.byref(handle))
        if st != binding.SBX_OK:
            raise SbxError(st, self._lib.sbx_multi_last_error(None).decode())
        self._h = handle
        self.devices = devs
        self.plans = [_PlanView(self._lib, self._lib.sbx_multi_plan(handle, i)) for i in range(len(devs))]
        self.rows, self.hop, self.n, self.channels = self.plans[0].rows, self.plans[0].hop, self.plans[0].n, self.plans[0].channels

    def close(self):
        if self._h is not None:
            for p in self.plans:
                p.close()
            self._lib.sbx_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, st):
        if st < 0:
            raise SbxError(st, self._lib.sbx_multi_last_error(self._h).decode())
        return st

    def num_frames(self, n_samples):
        return self.plans[0].num_frames(n_samples)

    def time_block(self, n_samples, index):
        """(frame_lo, frame_hi, sample_lo, sample_hi) of device `index` -- sharding.time_blocks(frame_align=2 if mono)."""
        v = [C.c_uint64() for _ in range(4)]
        self._check(self._lib.sbx_multi_time_block(self._h, n_samples, index, *[C.byref(x) for x in v]))
        return tuple(int(x.value) for x in v)

    def process_batch_ptr(self, pcm_ptr, n_samples, rgba_ptr, capacity_frames):
        got = C.c_uint64()
        self._check(self._lib.sbx_multi_process_batch(self._h, _ptr(pcm_ptr), n_samples, _ptr(rgba_ptr), capacity_frames, C.byref(got)))
        return int(got.value)

    def process_batch(self, pcm, out=None):
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        n_samples = pcm.shape[0]
        if out is None:
            out = np.empty((self.num_frames(n_samples), self.rows, 4), np.uint8)
        return out[:self.process_batch_ptr(pcm.ctypes.data, n_samples, out.ctypes.data, out.shape[0])]

    def process_batch_planar_ptr(self, pcm_ptr, n_sequences, sequence_stride_samples, n_samples_per_sequence, rgba_ptr, capacity_frames):
        got = C.c_uint64()
        self._check(self._lib.sbx_multi_process_batch_planar(self._h, _ptr(pcm_ptr), n_sequences, sequence_stride_samples,
                                                             n_samples_per_sequence, _ptr(rgba_ptr), capacity_frames, C.byref(got)))
        return int(got.value)

    def process_batch_planar(self, pcm):
        """(n_sequences, n_samples[, channels]) host array -> (n_sequences, frames, H, 4) uint8."""
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        n_seq, n_samples = pcm.shape[0], pcm.shape[1]
        frames = self.num_frames(n_samples)
        out = np.empty((n_seq, frames, self.rows, 4), np.uint8)
        self.process_batch_planar_ptr(pcm.ctypes.data, n_seq, n_samples, n_samples, out.ctypes.data, n_seq * frames)
        return out

    @property
    def launch_count(self):
        return int(self._lib.sbx_multi_launch_count(self._h))


class LiveStream:
    """sbx_live: pinned SPSC ring + hop cursor + one CUDA graph per frame."""

    def __init__(self, plan, ring_capacity=4096):
        self.plan = plan
        self._lib = plan._lib
        h = C.c_void_p()
        plan._check(self._lib.sbx_live_create(plan.handle, ring_capacity, C.byref(h)))
        self._h = h
        plan._live.add(self)

    def close(self):
        if self._h is not None:
            self._lib.sbx_live_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def push(self, samples, channels=2):
        samples = np.ascontiguousarray(samples, dtype=np.float32)
        n = samples.size // channels
        return int(self._lib.sbx_live_push(self._h, samples.ctypes.data, n, channels))

    def buffered(self):
        return int(self._lib.sbx_live_buffered(self._h))

    def pop_column(self, out=None):
        if out is None:
            out = np.empty((self.plan.rows, 4), np.uint8)
        st = self.plan._check(self._lib.sbx_live_pop_column(self._h, out.ctypes.data))
        return None if st == binding.SBX_NOT_ENOUGH_SAMPLES else out

    def measure(self, samples, iters, channels=2):
        """Native timing loop: `iters` x (push all of `samples`, pop one column); latencies in microseconds."""
        samples = np.ascontiguousarray(samples, dtype=np.float32)
        lat = np.empty(iters, np.float32)
        self.plan._check(self._lib.sbx_live_measure(self._h, samples.ctypes.data, samples.size // channels, channels,
                                                    iters, lat.ctypes.data))
        return lat

    def pop_magnitudes(self):
        out = np.empty((self.plan.k, 2), np.float32)
        st = self.plan._check(self._lib.sbx_live_pop_magnitudes(self._h, out.ctypes.data))
        return None if st == binding.SBX_NOT_ENOUGH_SAMPLES else out
