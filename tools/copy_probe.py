"""Host<->device copy ceiling of this box with the end-to-end leg's traffic pattern and NO kernels.

    python tools/copy_probe.py --gpus 1,2,4,8 [--mode processes|threads|both] [--out profiles/copy_ceiling_r02.json]

Per GPU and per step: H2D of 2 KB per frame (cfg1: hop * 4 bytes of PCM) and D2H of 4 KB per frame (one RGBA column)
from / to pinned host memory, chunked into 32 MiB of RGBA per chunk over three streams exactly like
sbx_process_batch's pipeline, for 131072 frames per step (bench.py's --e2e-frames).  Reported as frames/s (the unit of
the end-to-end number it bounds) and GB/s per direction, for G GPUs driven by G processes (one per GPU, what
torchrun does) or by G threads of one process (what sbx_multi_process_batch does).  The e2e numbers of bench.py
are to be read against THIS ceiling: whatever the kernels do, a frame cannot leave the box faster than its 4 KB column
crosses PCIe."""
import argparse
import json
import multiprocessing as mp
import os
import sys
import threading
import time

FRAMES, HOP_BYTES, COL_BYTES, CHUNK_FRAMES, STEPS = 131072, 2048, 4096, 8192, 8


def worker_setup(dev, write_combined=False):
    import torch
    torch.cuda.set_device(dev)
    h_in = torch.empty(FRAMES * HOP_BYTES, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(FRAMES * COL_BYTES, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    d_in = [torch.empty(CHUNK_FRAMES * HOP_BYTES, dtype=torch.uint8, device=f"cuda:{dev}") for _ in range(3)]
    d_out = [torch.empty(CHUNK_FRAMES * COL_BYTES, dtype=torch.uint8, device=f"cuda:{dev}") for _ in range(3)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
    return torch, h_in, h_out, d_in, d_out, streams


def one_step(ctx, h2d=True, d2h=True):
    torch, h_in, h_out, d_in, d_out, streams = ctx
    for c, f0 in enumerate(range(0, FRAMES, CHUNK_FRAMES)):
        s = streams[c % 3]
        with torch.cuda.stream(s):
            if h2d:
                d_in[c % 3].copy_(h_in[f0 * HOP_BYTES:(f0 + CHUNK_FRAMES) * HOP_BYTES], non_blocking=True)
            if d2h:
                h_out[f0 * COL_BYTES:(f0 + CHUNK_FRAMES) * COL_BYTES].copy_(d_out[c % 3], non_blocking=True)
    for s in streams:
        s.synchronize()


def proc_main(dev, n, barrier, results, h2d, d2h):
    ctx = worker_setup(dev)
    for _ in range(2):
        one_step(ctx, h2d, d2h)
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(STEPS):
        one_step(ctx, h2d, d2h)
    barrier.wait()                                               # all devices done: the box-wide time
    results[dev] = time.perf_counter() - t0


def run_processes(n, h2d=True, d2h=True):
    ctx = mp.get_context("spawn")
    barrier, results = ctx.Barrier(n), ctx.Manager().dict()
    procs = [ctx.Process(target=proc_main, args=(d, n, barrier, results, h2d, d2h)) for d in range(n)]
    for p in procs:
        p.start()
    for p in procs:
        p.join()
    return max(results.values())


def run_threads(n, h2d=True, d2h=True):
    ctxs = [worker_setup(d) for d in range(n)]
    barrier, times = threading.Barrier(n), [0.0] * n

    def body(d):
        for _ in range(2):
            one_step(ctxs[d], h2d, d2h)
        barrier.wait()
        t0 = time.perf_counter()
        for _ in range(STEPS):
            one_step(ctxs[d], h2d, d2h)
        barrier.wait()
        times[d] = time.perf_counter() - t0

    ts = [threading.Thread(target=body, args=(d,)) for d in range(n)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return max(times)


def describe(n, sec, h2d, d2h):
    fps = n * FRAMES * STEPS / sec
    return {"frames_per_s": fps, "h2d_gbs_total": fps * HOP_BYTES / 1e9 if h2d else 0.0, "d2h_gbs_total": fps * COL_BYTES / 1e9 if d2h else 0.0,
            "d2h_gbs_per_gpu": fps * COL_BYTES / 1e9 / n if d2h else 0.0}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1")
    ap.add_argument("--mode", default="both")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch
    have = torch.cuda.device_count()
    res = {"what": "pinned H2D 2 KB/frame + D2H 4 KB/frame per GPU, chunked as sbx_process_batch does (32 MiB RGBA chunks, 3 streams), no kernels; "
                   f"{FRAMES} frames per GPU per step, {STEPS} steps", "devices_present": have, "cpus": os.cpu_count(), "runs": {}}
    for n in [int(v) for v in a.gpus.split(",")]:
        if n > have:
            continue
        entry = {}
        if a.mode in ("processes", "both"):
            entry["processes_h2d_and_d2h"] = describe(n, run_processes(n), True, True)
            entry["processes_d2h_only"] = describe(n, run_processes(n, h2d=False), False, True)
            entry["processes_h2d_only"] = describe(n, run_processes(n, d2h=False), True, False)
        if a.mode in ("threads", "both"):
            entry["threads_h2d_and_d2h"] = describe(n, run_threads(n), True, True)
        res["runs"][str(n)] = entry
        print(n, json.dumps(entry), file=sys.stderr, flush=True)
    text = json.dumps(res, indent=1)
    print(text)
    if a.out:
        with open(a.out, "w") as f:
            f.write(text + "\n")
