"""ctypes loader for tests/emu/libemu.so -- host emulation of the kernel's phase functions
(TEST INFRASTRUCTURE; see emu_shim.h)."""
import ctypes as C
import importlib.util
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB_PATH = os.path.join(HERE, "libemu.so")
CSRC = os.path.join(ROOT, "spectrogram-rs_b200", "csrc")


def _binding():
    spec = importlib.util.spec_from_file_location("sbx_binding", os.path.join(ROOT, "spectrogram-rs_b200", "binding.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


binding = _binding()


def build(force=False):
    deps = [os.path.join(HERE, f) for f in ("emu_harness.cpp", "emu_shim.h")] + \
           [os.path.join(CSRC, f) for f in ("stft_kernels.cuh", "stft_q2.cuh", "sbx_pk.cuh", "stft_large.cuh", "stft_generic.cuh", "sbx_tables.h")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps):
        subprocess.run(["g++", "-O1", "-std=c++17", "-I", HERE, "-ffp-contract=off", "-fPIC", "-shared",
                        "-fvisibility=hidden", "-o", LIB_PATH, os.path.join(HERE, "emu_harness.cpp")],
                       check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.emu_batch.restype = C.c_longlong
        L.emu_batch.argtypes = [C.POINTER(binding.SbxConfig), C.c_void_p, C.c_ulonglong, C.c_void_p, C.c_void_p]
        L.emu_tables.restype = C.c_int
        L.emu_tables.argtypes = [C.POINTER(binding.SbxConfig)] + [C.c_void_p] * 6
        _lib = L
    return _lib


def make_config(lut_rgb, sample_rate=48000.0, window_n=4096, pad_factor=2, hop=512, rows_h=1024, channels=2,
                f_min=32.0, f_max=22030.0, db_min=-70.0, db_max=-10.0, db_eps=1e-7, interp=0, scheme=0, device=0):
    cfg = binding.SbxConfig()
    cfg.struct_size = C.sizeof(binding.SbxConfig)
    cfg.abi_version = binding.ABI_VERSION
    cfg.sample_rate, cfg.window_n, cfg.pad_factor, cfg.hop = sample_rate, window_n, pad_factor, hop
    cfg.rows_h, cfg.channels, cfg.f_min, cfg.f_max = rows_h, channels, f_min, f_max
    cfg.db_min, cfg.db_max, cfg.db_eps, cfg.interp, cfg.scheme, cfg.device = db_min, db_max, db_eps, interp, scheme, device
    lut = np.ascontiguousarray(lut_rgb, dtype=np.uint8).reshape(768)
    C.memmove(cfg.lut_rgb, lut.ctypes.data, 768)
    return cfg


def batch(cfg, pcm, want_mags=True, want_rgba=True, staged=False, real=None, wide=False, q2=False, cluster=False):
    """staged=True emulates the persistent kernel's pass 0 (PCM span in a stage buffer).
    real: mono batches on the real-sequence geometry (one frame per CTA) -- None: the product's default choice,
    False / True: what SBX_REAL=0 / 1 selects in the product."""
    lib().emu_set_staged(1 if staged else 0)
    lib().emu_set_real(2 if real is None else (1 if real else 0))
    lib().emu_set_q2(1 if q2 else 0)                       # pad factor 2: the packed-lane kernel of stft_q2.cuh
    lib().emu_set_wide(1 if wide else 0)                   # stereo: the WIDE single-frame geometries (more threads per CTA)
    lib().emu_set_cluster(1 if cluster else 0)             # M = 32768 .. 131072: the cluster kernel's phases (stft_large.cuh)
    pcm = np.ascontiguousarray(pcm, dtype=np.float32)
    n_samples = pcm.shape[0]
    n, hop = cfg.window_n, cfg.hop
    frames = 0 if n_samples < n else (n_samples - n) // hop + 1
    k = (n * cfg.pad_factor) // 2 - 1
    mags = np.zeros((frames, k, 2), np.float32) if want_mags else None
    rgba = np.zeros((frames, cfg.rows_h, 4), np.uint8) if want_rgba else None
    got = lib().emu_batch(C.byref(cfg), pcm.ctypes.data, n_samples,
                          mags.ctypes.data if want_mags else None, rgba.ctypes.data if want_rgba else None)
    if got != frames:
        raise RuntimeError(f"emu_batch returned {got}, expected {frames}")
    return mags, rgba


def tables(cfg):
    n, h = cfg.window_n, cfg.rows_h
    window = np.empty(n, np.float32)
    f0, step, cnt = np.empty(h, np.float32), np.empty(h, np.float32), np.empty(h, np.uint32)
    period = C.c_float()
    thr = np.empty(512, np.float32)
    lib().emu_tables(C.byref(cfg), window.ctypes.data, f0.ctypes.data, step.ctypes.data, cnt.ctypes.data,
                     C.addressof(period), thr.ctypes.data)
    return dict(window=window, f0=f0, step=step, n=cnt, period=np.float32(period.value), thr=thr)


def sample_table(cfg):
    """The product's per-sample table: dict(samp (entries, 2) uint32, chunks (n, 4) = row_begin, row_end, samp_begin,
    samp_end, rowinfo (H,), bank_group, ok)."""
    cap = int(cfg.rows_h) * 64 + int(cfg.window_n) * int(cfg.pad_factor) * 2 + 4096
    samp = np.empty((cap, 2), np.uint32)
    chunks = np.empty((int(cfg.rows_h) + 8, 4), np.uint32)
    rowinfo = np.empty(int(cfg.rows_h), np.uint32)
    n_chunks, bank_group, ok = C.c_uint32(), C.c_uint32(), C.c_int()
    fn = lib().emu_sample_table
    fn.restype = C.c_longlong
    fn.argtypes = [C.POINTER(binding.SbxConfig), C.c_void_p, C.c_ulonglong, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                   C.c_void_p, C.c_void_p]
    got = fn(C.byref(cfg), samp.ctypes.data, cap, chunks.ctypes.data, chunks.shape[0], C.addressof(n_chunks),
             rowinfo.ctypes.data, C.addressof(bank_group), C.addressof(ok))
    if got < 0:
        raise RuntimeError("emu_sample_table: buffers too small")
    return dict(samp=samp[:got].copy(), chunks=chunks[:n_chunks.value].copy(), rowinfo=rowinfo, bank_group=int(bank_group.value),
                ok=bool(ok.value))


def set_pan_gradient(rgb):
    """Stereo schemes: install (entries, 3) uint8 as the pan table of the emulated kernels; None restores the 256-entry lut."""
    fn = lib().emu_set_pan_gradient
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_uint32]
    if rgb is None:
        fn(None, 0)
    else:
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8).reshape(-1, 3)
        fn(rgb.ctypes.data, rgb.shape[0])
