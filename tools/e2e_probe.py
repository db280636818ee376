"""End-to-end (host pointers, pinned) throughput of sbx_process_batch for a chunk size given by SBX_BATCH_CHUNK_MB."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import spectrogram_rs_b200 as sbx
from tests import signals
F = 1 << 18
plan = sbx.Plan(sample_rate=48000.0, window_n=4096, pad_factor=2, hop=512, rows_h=1024, channels=1, lut="magma")
ns = (F - 1) * 512 + 4096
h_pcm = torch.from_numpy(signals.white_noise(ns)).pin_memory()
h_rgba = torch.empty((F, 1024, 4), dtype=torch.uint8).pin_memory()
for _ in range(2):
    plan.process_batch_ptr(h_pcm.data_ptr(), ns, h_rgba.data_ptr(), F)
t0 = time.perf_counter()
for _ in range(5):
    plan.process_batch_ptr(h_pcm.data_ptr(), ns, h_rgba.data_ptr(), F)
dt = (time.perf_counter() - t0) / 5
print(f"chunk {os.environ.get('SBX_BATCH_CHUNK_MB', '32')} MiB: {F / dt / 1e6:.2f} Mframes/s, D2H {F * 4096 / dt / 1e9:.1f} GB/s")
