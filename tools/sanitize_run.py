"""Small invocations of every kernel family (fused, persistent, large, generic, colourise, live) for
compute-sanitizer:  compute-sanitizer --tool racecheck|memcheck|synccheck python tools/sanitize_run.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import spectrogram_rs_b200 as sbx  # noqa: E402
from tests import signals  # noqa: E402

CASES = [dict(window_n=4096, pad_factor=2, hop=512, rows_h=1024, channels=1),       # cfg1 fused, mono pairs
         dict(window_n=2048, pad_factor=2, hop=2048, rows_h=512, channels=2),       # cfg2 fused, stereo
         dict(window_n=1024, pad_factor=4, hop=128, rows_h=256, channels=1, interp=1),
         dict(window_n=8192, pad_factor=4, hop=1024, rows_h=512, channels=1),       # large path
         dict(window_n=2400, pad_factor=2, hop=58, rows_h=256, channels=2)]         # mixed radix
frames = int(os.environ.get("SBX_SANITIZE_FRAMES", "5"))
for kw in CASES:
    n, hop, ch = kw["window_n"], kw["hop"], kw["channels"]
    ns = n + hop * (frames - 1)
    pcm = signals.white_noise(ns * ch).reshape(ns, ch) if ch == 2 else signals.white_noise(ns)
    with sbx.Plan(lut="magma", **kw) as plan:
        rgba = plan.process_batch(pcm)
        d = torch.zeros((frames, plan.k, 2), dtype=torch.float32, device="cuda")
        plan.magnitudes_batch_device(torch.from_numpy(pcm).cuda(), ns, d, frames)
        plan.synchronize()
        col = plan.colorize_magnitudes(d.cpu().numpy())
        assert np.array_equal(col, rgba)
        print("ok", kw, rgba.shape)
# persistent TMA-staged kernel (needs >= 2 units per resident CTA)
if os.environ.get("SBX_SANITIZE_PERSISTENT", "1") == "1":
    kw = dict(window_n=8192, pad_factor=2, hop=256, rows_h=128, channels=1)
    fr = 2 * 148 * 2 + 3
    ns = kw["window_n"] + kw["hop"] * (fr - 1)
    with sbx.Plan(lut="viridis", **kw) as plan:
        out = plan.process_batch(signals.white_noise(ns))
        print("ok persistent", out.shape)
with sbx.Plan(window_n=2048, pad_factor=2, hop=2048, rows_h=256, channels=2, lut="magma") as plan:
    live = plan.live(4096)
    live.push(signals.stereo_loopback(2048))
    assert live.pop_column() is not None
    live.close()
    print("ok live")
