"""Randomised soak of the CUDA path against the CPU oracle (more and wider cases than tests/test_gpu_parity.py's
seeded test; run by hand on a GPU box, not part of the suite):
    python tools/gpu_soak.py [--cases 300] [--seed 1] [--seconds 240]
Every case: magnitudes within 1e-5 of the frame peak, RGBA bit-exact given the GPU's own magnitudes."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import spectrogram_rs_b200 as sbx  # noqa: E402
from oracle import oracle  # noqa: E402
from tests import signals  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cases", type=int, default=300)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--seconds", type=float, default=240.0)
ap.add_argument("--real-bias", action="store_true", help="mostly mono, even hops, transform lengths 8192 .. 32768 (the real-sequence geometry)")
a = ap.parse_args()
rng = np.random.Generator(np.random.PCG64(a.seed))
lut = signals.lut("magma")
pow2 = [256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536]
mixed = [100, 120, 300, 441, 600, 1000, 1470, 2205, 2400, 4800, 9600,
         551, 1102, 1103, 94, 347, 2731, 4099]             # the last row: a prime factor > 7 (Bluestein path)
t0, done, worst, paths = time.time(), 0, 0.0, {}
for case in range(a.cases):
    if time.time() - t0 > a.seconds:
        break
    if rng.random() < 0.25:
        n, pad = int(rng.choice(mixed)), int(rng.choice([2, 4]))
        if n * pad * 12 > 227 * 1024:
            pad = 2
        if n in (551, 1102, 1103, 94, 347, 2731, 4099) and n * (pad + 1) - 1 > 16384:
            pad = 2 if n * 3 - 1 <= 16384 else 1
    else:
        n = int(rng.choice(pow2))
        pads = [p for p in (1, 2, 4) if n * p <= 262144 and (n * p <= 16384 or n >= 8192)]
        pad = int(rng.choice(pads))
    if a.real_bias and rng.random() < 0.8:
        n, pad = [(2048, 4), (4096, 2), (8192, 1), (4096, 4), (8192, 2), (16384, 1), (8192, 4), (16384, 2), (32768, 1)][int(rng.integers(0, 9))]
    m = n * pad
    sr = float(rng.choice([8000.0, 22050.0, 44100.0, 48000.0, 96000.0, 192000.0]))
    hop = max(1, int(rng.choice([1, 2, 3, n // 8, n // 4, n // 2, n, n + 17, int(rng.integers(1, 2 * n))])))
    big = m > 16384
    rows = int(rng.choice([16, 64, 333, 1024] if big else [1, 2, 7, 33, 100, 257, 1024, 2048]))
    f_min = float(rng.choice([1.0, 20.0, 32.0, 500.0]))
    f_max = float(rng.choice([f_min * 1.5, 4000.0, sr / 2, 22030.0, sr]))
    f_max = f_max if f_max > f_min else f_min * 2
    interp, scheme, ch = int(rng.integers(0, 2)), int(rng.integers(0, 2)), int(rng.integers(1, 3))
    if a.real_bias and rng.random() < 0.8:
        ch, hop = 1, max(2, hop & ~1)
    frames = int(rng.integers(1, 6))
    real = str(rng.choice(["", "0", "1"]))
    os.environ.pop("SBX_REAL", None)
    if real:
        os.environ["SBX_REAL"] = real
    ns = n + hop * (frames - 1) + int(rng.integers(0, 3))
    frames = (ns - n) // hop + 1
    amp = 10.0 ** rng.uniform(-4, 0.3)
    x = (amp * (0.6 * np.sin(2 * np.pi * rng.uniform(10, sr / 2) * np.arange(ns) / sr) + 0.4 * rng.standard_normal(ns))).astype(np.float32)
    pcm = x if ch == 1 else np.stack([x, (0.3 * amp * rng.standard_normal(ns)).astype(np.float32)], 1)
    kw = dict(sample_rate=sr, window_n=n, pad_factor=pad, hop=hop, rows_h=rows, f_min=f_min, f_max=f_max, interp=interp)
    try:
        plan = sbx.Plan(channels=ch, scheme=scheme, lut=lut, **kw)
    except sbx.SbxError as e:
        if e.status == -2:                                    # SBX_ERR_UNSUPPORTED (e.g. a row longer than the scratch area)
            continue
        raise
    with plan:
        p = oracle.params(stereo_scheme=scheme, **kw)
        o_rgba, o_mags = oracle.batch(p, pcm, ch, lut, want_mags=True, threads=8)
        d = torch.zeros((frames, plan.k, 2), dtype=torch.float32, device="cuda")
        plan.magnitudes_batch_device(torch.from_numpy(pcm).cuda(), ns, d, frames)
        g_rgba = plan.process_batch(pcm)
        plan.synchronize()
        g_mags = d.cpu().numpy()
        peak = np.maximum(np.abs(o_mags).max(axis=(1, 2), keepdims=True), 1e-30)
        err = float((np.abs(g_mags - o_mags) / peak).max())
        rec = np.stack([oracle.column_from_magnitudes(p, g_mags[f], lut) for f in range(frames)])
        ok = err <= 1e-5 and np.array_equal(g_rgba, rec)
        paths[plan.kernel_path] = paths.get(plan.kernel_path, 0) + 1
        worst = max(worst, err)
        if not ok:
            print("FAIL", case, plan.kernel_path, kw, "ch", ch, "scheme", scheme, "real", real, "frames", frames, "err", err,
                  "pixels differing", int((g_rgba != rec).any(axis=2).sum()))
            sys.exit(1)
    done += 1
print(f"soak ok: {done} cases in {time.time() - t0:.0f} s, worst magnitude error {worst:.2e} of peak, paths {paths}")
