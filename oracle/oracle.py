"""ctypes loader for the CPU oracle (oracle/stft_oracle.c).  TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")

INTERP_CUBIC, INTERP_COSINE = 0, 1


class OrcParams(C.Structure):
    _fields_ = [("sample_rate", C.c_float), ("window_n", C.c_uint32), ("pad_factor", C.c_uint32),
                ("hop", C.c_uint32), ("rows_h", C.c_uint32), ("f_min", C.c_double), ("f_max", C.c_double),
                ("db_min", C.c_float), ("db_max", C.c_float), ("db_eps", C.c_float),
                ("interp", C.c_uint32), ("stereo_scheme", C.c_uint32), ("pan_levels", C.c_uint32)]


FAST_LIB_PATH = os.path.join(HERE, "liboracle_fast.so")


def build(force=False):
    src = os.path.join(HERE, "stft_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return LIB_PATH


def build_fast():
    """liboracle_fast.so: -O3 -march=native, for the TIMED CPU baseline only.  Always rebuilt on the machine that
    times it (a library built for this container's CPU may not run on the GPU box's)."""
    subprocess.run(["make", "-C", HERE, "-B", "liboracle_fast.so"], check=True, capture_output=True)
    return FAST_LIB_PATH


_lib = None
_fast = None


def fast_lib():
    """The timed build with the entry points the baseline uses."""
    global _fast
    if _fast is None:
        build_fast()
        L = C.CDLL(FAST_LIB_PATH)
        f32p, u8p = C.POINTER(C.c_float), C.POINTER(C.c_uint8)
        L.orc_batch.restype = C.c_int64
        L.orc_batch.argtypes = [C.POINTER(OrcParams), f32p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64,
                                u8p, u8p, f32p, C.c_int]
        L.orc_window_pad_batch.restype = C.c_int
        L.orc_window_pad_batch.argtypes = [C.POINTER(OrcParams), f32p, C.c_uint32, C.c_uint64, C.c_uint64, f32p, C.c_int]
        L.orc_split_colour_batch.restype = C.c_int
        L.orc_split_colour_batch.argtypes = [C.POINTER(OrcParams), f32p, C.c_uint64, u8p, u8p, C.c_int]
        _fast = L
    return _fast


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        f32p, u8p, u32p = C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)
        L.orc_hann.restype = C.c_float
        L.orc_hann.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_num_output_frequencies.restype = C.c_uint32
        L.orc_num_output_frequencies.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_frame_magnitudes.restype = C.c_int
        L.orc_frame_magnitudes.argtypes = [f32p, C.c_uint32, C.c_uint32, C.c_uint32, f32p]
        L.orc_window_pad.restype = None
        L.orc_window_pad.argtypes = [f32p, C.c_uint32, C.c_uint32, f32p]
        L.orc_split_scale.restype = None
        L.orc_split_scale.argtypes = [f32p, C.c_uint32, C.c_uint32, f32p]
        L.orc_fft_c2c_f32.restype = C.c_int
        L.orc_fft_c2c_f32.argtypes = [f32p, f32p, C.c_uint32]
        L.orc_row_edge.restype = C.c_float
        L.orc_row_edge.argtypes = [C.c_double, C.c_double, C.c_uint32, C.c_uint32]
        L.orc_period.restype = C.c_float
        L.orc_period.argtypes = [C.c_uint32, C.c_float]
        L.orc_bin_map_row.restype = None
        L.orc_bin_map_row.argtypes = [C.c_double, C.c_double, C.c_uint32, C.c_uint32, C.c_float, C.c_uint32,
                                      f32p, f32p, u32p]
        L.orc_normalised_db.restype = C.c_float
        L.orc_normalised_db.argtypes = [C.c_float] * 5
        L.orc_db_chain.restype = C.c_float
        L.orc_db_chain.argtypes = [C.c_float] * 4
        L.orc_gradient_index.restype = C.c_uint32
        L.orc_gradient_index.argtypes = [C.c_float]
        L.orc_color_index.restype = C.c_uint32
        L.orc_color_index.argtypes = [C.c_float] * 5
        L.orc_column_from_magnitudes.restype = None
        L.orc_column_from_magnitudes.argtypes = [C.POINTER(OrcParams), f32p, u8p, u8p]
        L.orc_rows_from_magnitudes.restype = None
        L.orc_rows_from_magnitudes.argtypes = [C.POINTER(OrcParams), f32p, f32p]
        L.orc_column.restype = C.c_int
        L.orc_column.argtypes = [C.POINTER(OrcParams), f32p, C.c_uint32, u8p, u8p]
        L.orc_num_frames.restype = C.c_uint64
        L.orc_num_frames.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
        L.orc_hop_from_stride.restype = C.c_uint32
        L.orc_hop_from_stride.argtypes = [C.c_float, C.c_float]
        L.orc_window_from_period.restype = C.c_uint32
        L.orc_window_from_period.argtypes = [C.c_float, C.c_float]
        L.orc_batch.restype = C.c_int64
        L.orc_batch.argtypes = [C.POINTER(OrcParams), f32p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64,
                                u8p, u8p, f32p, C.c_int]
        L.orc_colour_batch.restype = C.c_int
        L.orc_colour_batch.argtypes = [C.POINTER(OrcParams), f32p, C.c_uint64, u8p, u8p, C.c_int]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _f32(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def params(sample_rate=48000.0, window_n=4096, pad_factor=2, hop=512, rows_h=1024, f_min=32.0, f_max=22030.0,
           db_min=-70.0, db_max=-10.0, db_eps=1e-7, interp=INTERP_CUBIC, stereo_scheme=0, pan_levels=0):
    """pan_levels: stereo schemes only -- entries of the lut passed to batch / column_from_magnitudes (0 = 256)."""
    return OrcParams(sample_rate, window_n, pad_factor, hop, rows_h, f_min, f_max, db_min, db_max, db_eps,
                     interp, stereo_scheme, pan_levels)


def num_output_frequencies(n, pad):
    return int(lib().orc_num_output_frequencies(n, pad))


def hann(n):
    L = lib()
    return np.array([L.orc_hann(i, n) for i in range(n)], dtype=np.float32)


def frame_magnitudes(lr, n, pad):
    """lr: (>=n, 2) float32.  Returns (K, 2) float32 or None (the reference's None)."""
    lr = np.ascontiguousarray(lr, dtype=np.float32)
    out = np.empty((num_output_frequencies(n, pad), 2), dtype=np.float32)
    st = lib().orc_frame_magnitudes(_f32(lr), lr.shape[0], n, pad, _f32(out))
    if st == 1:
        return None
    if st != 0:
        raise RuntimeError("orc_frame_magnitudes failed")
    return out


def window_pad(lr, n, pad):
    lr = np.ascontiguousarray(lr, dtype=np.float32)
    z = np.empty((n * pad, 2), dtype=np.float32)
    lib().orc_window_pad(_f32(lr), n, pad, _f32(z))
    return z


def split_scale(zf, n, pad):
    zf = np.ascontiguousarray(zf, dtype=np.float32)
    out = np.empty((num_output_frequencies(n, pad), 2), dtype=np.float32)
    lib().orc_split_scale(_f32(zf), n, pad, _f32(out))
    return out


def fft_c2c_f32(z):
    z = np.ascontiguousarray(z, dtype=np.float32)
    out = np.empty_like(z)
    if lib().orc_fft_c2c_f32(_f32(z), _f32(out), z.shape[0]) != 0:
        raise RuntimeError("orc_fft_c2c_f32 failed")
    return out


def bin_map(p):
    """Returns (f0, step, n, period) arrays for every row of the params' log-frequency map."""
    L = lib()
    k = num_output_frequencies(p.window_n, p.pad_factor)
    period = L.orc_period(k, p.sample_rate)
    f0 = np.empty(p.rows_h, np.float32)
    step = np.empty(p.rows_h, np.float32)
    n = np.empty(p.rows_h, np.uint32)
    a, b, c = C.c_float(), C.c_float(), C.c_uint32()
    for py in range(p.rows_h):
        L.orc_bin_map_row(p.f_min, p.f_max, py, p.rows_h, period, k, C.byref(a), C.byref(b), C.byref(c))
        f0[py], step[py], n[py] = a.value, b.value, c.value
    return f0, step, n, np.float32(period)


def column_from_magnitudes(p, mags_lr, lut_rgb):
    mags_lr = np.ascontiguousarray(mags_lr, dtype=np.float32)
    lut_rgb = np.ascontiguousarray(lut_rgb, dtype=np.uint8)
    out = np.empty((p.rows_h, 4), dtype=np.uint8)
    lib().orc_column_from_magnitudes(C.byref(p), _f32(mags_lr), _u8(lut_rgb), _u8(out))
    return out


def rows_from_magnitudes(p, mags_lr):
    mags_lr = np.ascontiguousarray(mags_lr, dtype=np.float32)
    out = np.empty((p.rows_h, 2), dtype=np.float32)
    lib().orc_rows_from_magnitudes(C.byref(p), _f32(mags_lr), _f32(out))
    return out


def color_index(l, r, db_min=-70.0, db_max=-10.0, db_eps=1e-7):
    return int(lib().orc_color_index(l, r, db_min, db_max, db_eps))


def num_frames(n_samples, n, hop):
    return int(lib().orc_num_frames(n_samples, n, hop))


def batch(p, pcm, channels, lut_rgb, want_rgba=True, want_mags=False, frame_lo=0, frame_hi=None, threads=1, fast=False):
    """pcm: float32 (n_samples,) mono or (n_samples, 2).  Returns (rgba or None, mags or None).
    fast=True: the -O3 -march=native build (bench.py's timed baseline; same results)."""
    pcm = np.ascontiguousarray(pcm, dtype=np.float32)
    n_samples = pcm.shape[0]
    total = num_frames(n_samples, p.window_n, p.hop)
    if frame_hi is None or frame_hi > total:
        frame_hi = total
    count = max(0, frame_hi - frame_lo)
    k = num_output_frequencies(p.window_n, p.pad_factor)
    rgba = np.empty((count, p.rows_h, 4), np.uint8) if want_rgba else None
    mags = np.empty((count, k, 2), np.float32) if want_mags else None
    lut_rgb = np.ascontiguousarray(lut_rgb, dtype=np.uint8)
    if count:
        L = fast_lib() if fast else lib()
        got = L.orc_batch(C.byref(p), _f32(pcm), n_samples, channels, frame_lo, frame_hi, _u8(lut_rgb),
                          _u8(rgba) if want_rgba else None, _f32(mags) if want_mags else None, threads)
        if got != count:
            raise RuntimeError(f"orc_batch returned {got}, expected {count}")
    return rgba, mags


def max_threads():
    return int(lib().orc_max_threads())


def batch_pocketfft(p, pcm, channels, lut_rgb, threads=1, chunk_frames=4096):
    """The same path with the FFT stage done by pocketfft (scipy.fft, complex64, `workers=threads`) instead of the
    oracle's own radix-2 FFT: the stand-in for the reference's FFTW call in the timed CPU baseline (FFTW is not
    in this image).  Everything around the transform -- take N / pack / Hann / zero-pad (fft.rs:48-69) and split /
    scale / resample / dB / gradient (fft.rs:81-98 onwards) -- is the C oracle (-O3 -march=native build), threaded
    over frames like the transform itself.  Returns rgba (frames, H, 4)."""
    import scipy.fft
    pcm = np.ascontiguousarray(pcm, dtype=np.float32)
    n, pad, hop = p.window_n, p.pad_factor, p.hop
    m = n * pad
    total = num_frames(pcm.shape[0], n, hop)
    lut_rgb = np.ascontiguousarray(lut_rgb, dtype=np.uint8)
    rgba = np.empty((total, p.rows_h, 4), np.uint8)
    L = fast_lib()
    z = np.empty((min(chunk_frames, max(total, 1)), m), np.complex64)
    for f0 in range(0, total, chunk_frames):
        f1 = min(total, f0 + chunk_frames)
        zc = z[:f1 - f0]
        if L.orc_window_pad_batch(C.byref(p), _f32(pcm), channels, f0, f1 - f0, zc.ctypes.data_as(C.POINTER(C.c_float)), threads) != 0:
            raise RuntimeError("orc_window_pad_batch failed")
        zf = scipy.fft.fft(zc, axis=1, workers=threads, overwrite_x=True)
        zf = np.ascontiguousarray(zf, dtype=np.complex64)
        if L.orc_split_colour_batch(C.byref(p), zf.ctypes.data_as(C.POINTER(C.c_float)), f1 - f0, _u8(lut_rgb), _u8(rgba[f0:f1]), threads) != 0:
            raise RuntimeError("orc_split_colour_batch failed")
    return rgba
