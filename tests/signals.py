"""Deterministic synthetic inputs shared by tests, smoke() and bench.py."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def lut(name):
    """256x3 uint8 ramp parsed from the product's luts.inc (no GPU or library needed)."""
    text = open(os.path.join(ROOT, "spectrogram-rs_b200", "csrc", "luts.inc")).read()
    m = re.search(r"SBX_LUT_%s\[768\] = \{(.*?)\};" % name.upper(), text, re.S)
    vals = np.array([int(v) for v in re.findall(r"\d+", m.group(1))], dtype=np.uint8)
    assert vals.size == 768
    return vals.reshape(256, 3)


def log_sweep(n_samples, sr=48000.0, f0=20.0, f1=20000.0, amp=0.5, duration=None):
    """cfg1: exponential sine sweep, computed in f64 then cast to f32."""
    t_total = duration if duration is not None else n_samples / sr
    n = np.arange(n_samples, dtype=np.float64)
    k = np.log(f1 / f0)
    phase = 2.0 * np.pi * f0 * t_total / k * (np.power(f1 / f0, n / (sr * t_total)) - 1.0)
    return (amp * np.sin(phase)).astype(np.float32)


def stereo_loopback(n_samples, sr=48000.0, seed=1):
    """cfg2: L = sweep, R = 0.25 N(0,1) + 1 kHz tone."""
    rng = np.random.Generator(np.random.PCG64(seed))
    left = log_sweep(n_samples, sr)
    t = np.arange(n_samples, dtype=np.float64) / sr
    right = (0.25 * rng.standard_normal(n_samples) + 0.25 * np.sin(2 * np.pi * 1000.0 * t)).astype(np.float32)
    return np.stack([left, right], axis=1)


def sweeps_plus_noise(n_samples, sr=96000.0, seed=2):
    """cfg3: sum of 8 slow sweeps + N(0, 0.05^2)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = np.zeros(n_samples, np.float64)
    for i in range(8):
        x += 0.1 * log_sweep(n_samples, sr, 40.0 * (i + 1), 4000.0 * (i + 1), 1.0).astype(np.float64)
    x += 0.05 * rng.standard_normal(n_samples)
    return x.astype(np.float32)


def channel_tone(c, n_samples, sr=48000.0):
    """cfg4: channel c = sine at 100 * 2^(c/8) Hz + noise (seed 100 + c)."""
    rng = np.random.Generator(np.random.PCG64(100 + c))
    t = np.arange(n_samples, dtype=np.float64) / sr
    return (0.5 * np.sin(2 * np.pi * 100.0 * 2.0 ** (c / 8.0) * t) + 0.05 * rng.standard_normal(n_samples)).astype(np.float32)


def white_noise(n_samples, seed=3, amp=0.25):
    rng = np.random.Generator(np.random.PCG64(seed))
    return (amp * rng.standard_normal(n_samples)).astype(np.float32)
