"""Per-phase share of a CTA's residency in fused_q2_kernel, from a -DSBX_PHASE_CLOCK variant of the library.

    tools/build_variant.sh phase -DSBX_PHASE_CLOCK          (here, no GPU needed)
    SBX_LIB_PATH=variants/libsbx_phase.so python tools/phase_split.py --cfg 1 [--frames F] [--out profiles/phase_split_cfg1.json]

Thread 0 of every CTA reads clock64() at the barrier-delimited phase boundaries; the counters are summed over all CTAs of
a launch.  With three CTAs resident per SM the phases of different CTAs overlap, so a phase's share is its share of a
CTA's time on the SM (what shortening it would free), not of the SM's issue slots.  The instrumented kernel is a few
per cent slower than the shipped one and is never benchmarked."""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import spectrogram_rs_b200 as sbx  # noqa: E402
from tests import signals  # noqa: E402

CFG = {1: dict(sample_rate=48000.0, window_n=4096, pad_factor=2, hop=512, rows_h=1024, channels=1, lut="magma"),
       2: dict(sample_rate=48000.0, window_n=2048, pad_factor=2, hop=2048, rows_h=1024, channels=2, lut="magma"),
       8: dict(sample_rate=48000.0, window_n=4096, pad_factor=2, hop=512, rows_h=1024, channels=2, lut="magma")}
PHASES = ["pass0 (PCM + window loads, first radix pass)", "shared-memory passes", "last pass + partner exchange + L/R split",
          "resample (lane per sample)", "row sums + dB step + colour + RGBA stores", "magnitude output (unused here)"]

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", type=int, default=1)
ap.add_argument("--frames", type=int, default=262144)
ap.add_argument("--out", default="")
a = ap.parse_args()
lib = sbx.binding.load()
fn = lib.sbx_debug_phase_cycles                      # AttributeError: not a -DSBX_PHASE_CLOCK build
fn.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
kw = CFG[a.cfg]
plan = sbx.Plan(**kw)
assert plan.kernel_path.endswith("-q2"), plan.kernel_path
ns = (a.frames - 1) * kw["hop"] + kw["window_n"]
x = torch.from_numpy(signals.white_noise(ns * kw["channels"])).cuda()
out = torch.empty((a.frames, plan.rows, 4), dtype=torch.uint8, device="cuda")
for _ in range(3):
    plan.process_batch_device(x, ns, out, a.frames)
plan.synchronize()
buf = (C.c_ulonglong * 8)()
assert fn(buf, 1) == 0
plan.process_batch_device(x, ns, out, a.frames)
plan.synchronize()
assert fn(buf, 1) == 0
cycles = [int(v) for v in buf]
ctas, total = cycles[7], sum(cycles[:6])
res = {"kernel": "fused_q2_kernel, " + plan.kernel_path, "config": {k: v for k, v in kw.items()}, "frames": a.frames, "ctas": ctas,
       "cycles_per_cta": total / ctas,
       "phases": [{"phase": PHASES[i], "cycles_per_cta": cycles[i] / ctas, "share": cycles[i] / total} for i in range(6) if cycles[i]],
       "method": "clock64() of thread 0 at the barrier-delimited phase boundaries, summed over all CTAs of one launch (-DSBX_PHASE_CLOCK variant)"}
text = json.dumps(res, indent=1)
print(text)
if a.out:
    with open(a.out, "w") as f:
        f.write(text + "\n")
