"""Launches the fused kernel a few times on a mid-sized batch; the command ncu wraps.
    python tools/profile_kernel.py [--cfg 1|3|5] [--frames F] [--iters I]
Prints the CUDA-event time per launch when run without ncu."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import spectrogram_rs_b200 as sbx  # noqa: E402
from tests import signals  # noqa: E402

CFG = {1: dict(sample_rate=48000.0, window_n=4096, pad_factor=2, hop=512, rows_h=1024, channels=1, lut="magma"),
       2: dict(sample_rate=48000.0, window_n=2048, pad_factor=2, hop=2048, rows_h=1024, channels=2, lut="magma"),
       3: dict(sample_rate=96000.0, window_n=8192, pad_factor=2, hop=256, rows_h=1024, channels=1, lut="viridis"),
       5: dict(sample_rate=48000.0, window_n=4096, pad_factor=4, hop=512, rows_h=2048, channels=1, lut="magma"),
       7: dict(sample_rate=96000.0, window_n=8192, pad_factor=2, hop=1024, rows_h=1024, channels=2, lut="viridis"),   # stereo M = 16384
       8: dict(sample_rate=48000.0, window_n=4096, pad_factor=2, hop=512, rows_h=1024, channels=2, lut="magma"),     # cfg1 geometry, stereo
       6: dict(sample_rate=48000.0, window_n=2400, pad_factor=2, hop=58, rows_h=1024, channels=2, lut="magma")}   # the app's own geometry

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", type=int, default=1)
ap.add_argument("--frames", type=int, default=32768)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--mags", action="store_true", help="magnitude output instead of RGBA")
ap.add_argument("--n", type=int, default=0, help="cfg5 sweep: window N with P = 4, hop = N/8, H = 2048")
ap.add_argument("--pad", type=int, default=0, help="with --n: pad factor (default 4, the cfg5 sweep)")
ap.add_argument("--hop", type=int, default=0)
ap.add_argument("--ch", type=int, default=0)
a = ap.parse_args()
kw = dict(CFG[a.cfg])
if a.n:
    kw.update(window_n=a.n, pad_factor=a.pad or 4, hop=a.hop or a.n // 8, rows_h=2048, channels=a.ch or 1)
plan = sbx.Plan(**kw)
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
plan.set_stream(s.cuda_stream)
ns = (a.frames - 1) * kw["hop"] + kw["window_n"]
ch = kw["channels"]
x = torch.from_numpy(signals.white_noise(ns * ch)).cuda()
if a.mags:
    out = torch.empty((a.frames, plan.k, 2), dtype=torch.float32, device="cuda")
    run = lambda: plan.magnitudes_batch_device(x, ns, out, a.frames)
else:
    out = torch.empty((a.frames, plan.rows, 4), dtype=torch.uint8, device="cuda")
    run = lambda: plan.process_batch_device(x, ns, out, a.frames)
for _ in range(3):
    run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record(s)
for _ in range(a.iters):
    run()
e1.record(s)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
print(f"cfg{a.cfg} N={kw['window_n']} P={kw['pad_factor']} hop={kw['hop']} frames={a.frames} {ms:.4f} ms/launch  {a.frames / ms / 1e3:.3f} Mframes/s")
