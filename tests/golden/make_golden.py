"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/stft_oracle.c) on seeded inputs.

The reference ships no fixtures, and neither Rust nor FFTW exist in this image, so
these vectors pin the ORACLE's output (parity unpinned by the reference itself); the analytic
known-answer tests in tests/test_oracle_kat.py pin the oracle independently.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from tests import signals  # noqa: E402

CASES = {
    # name: (params kwargs, channels, lut, n_frames, signal)
    "cfg1_sweep_mono": (dict(sample_rate=48000.0, window_n=4096, pad_factor=2, hop=512, rows_h=1024), 1, "magma", 3, "sweep"),
    "cfg2_stereo_live": (dict(sample_rate=48000.0, window_n=2048, pad_factor=2, hop=2048, rows_h=1024), 2, "magma", 2, "stereo"),
    "cfg3_96k_viridis": (dict(sample_rate=96000.0, window_n=8192, pad_factor=2, hop=256, rows_h=1024), 1, "viridis", 2, "sweeps"),
    "cfg5_pad4_n1024": (dict(sample_rate=48000.0, window_n=1024, pad_factor=4, hop=128, rows_h=2048), 1, "magma", 2, "noise"),
    "small_cosine": (dict(sample_rate=44100.0, window_n=256, pad_factor=2, hop=64, rows_h=96, interp=1), 2, "viridis", 4, "stereo"),
}


def make_signal(kind, n_samples, sr):
    if kind == "sweep":
        return signals.log_sweep(n_samples, sr, duration=n_samples / sr)
    if kind == "stereo":
        return signals.stereo_loopback(n_samples, sr, seed=1)
    if kind == "sweeps":
        return signals.sweeps_plus_noise(n_samples, sr, seed=2)
    return signals.white_noise(n_samples, seed=3)


def generate(name):
    kw, ch, lut_name, frames, kind = CASES[name]
    p = oracle.params(**kw)
    ns = kw["window_n"] + kw["hop"] * (frames - 1)
    pcm = make_signal(kind, ns, kw["sample_rate"])
    rgba, mags = oracle.batch(p, pcm, ch, signals.lut(lut_name), want_mags=True)
    return dict(pcm=pcm, mags=mags, rgba=rgba)


def main():
    for name in CASES:
        out = generate(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
