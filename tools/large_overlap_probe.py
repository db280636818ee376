"""Does the multi-kernel path gain from running two batches on two streams at once?  (Design probe: if it does,
launch_large could alternate its chunks between two internal streams.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import spectrogram_rs_b200 as sbx
from tests import signals

for n in (16384, 65536):
    hop, F = n // 8, 2048 if n == 65536 else 4096
    ns = (F - 1) * hop + n
    x = torch.from_numpy(signals.white_noise(ns)).cuda()
    plans, streams, outs = [], [], []
    for i in range(2):
        p = sbx.Plan(sample_rate=48000.0, window_n=n, pad_factor=4, hop=hop, rows_h=2048, channels=1, lut="magma")
        s = torch.cuda.Stream()
        p.set_stream(s.cuda_stream)
        plans.append(p); streams.append(s)
        outs.append(torch.empty((F, 2048, 4), dtype=torch.uint8, device="cuda"))
    def run(k):
        for i in range(k):
            plans[i].process_batch_device(x, ns, outs[i], F)
    for k in (1, 2):
        run(k); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for s in streams: s.wait_event(a)
        for _ in range(3): run(k)
        for s in streams: torch.cuda.current_stream().wait_stream(s)
        b.record(); torch.cuda.synchronize()
        print(f"N={n}: {k} concurrent batch(es): {3 * k * F / (a.elapsed_time(b) * 1e-3) / 1e6:.3f} Mframes/s aggregate")
    for p in plans: p.close()
