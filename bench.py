#!/usr/bin/env python
"""bench.py -- STFT frames/s of the fused sm_100a frame transform (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input: F frames per GPU of the
cfg1 geometry ("batched 4096-pt frames": 48 kHz mono log sweep, N 4096, zero-padded to M 8192,
hop 512, H 1024 rows, Magma).  Multi-GPU = time-block sharding: rank r owns frames [rF, (r+1)F) of
one long recording, i.e. its own sample block including the N-hop halo, and there is no data-path
collective -- torch.distributed is used only for the barrier and the max-over-ranks
of the device time.  Weak scaling: F per GPU is fixed.

Rank 0 prints ONE JSON line.  `value` is kernel throughput with the PCM resident in HBM (CUDA
events on the launching stream); `e2e` goes through the host-pointer C-ABI call sbx_process_batch
with pinned host buffers, H2D and D2H inside the timed region; `roofline` compares the
algorithmic bytes (hop*4 in + H*4 out per frame) with the measured HBM copy peak;
`cpu_baseline` is the CPU oracle (a port of the reference's path; the Rust+FFTW reference itself
cannot be built in this image) timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# cfg1 geometry
SR, N, PAD, HOP, ROWS = 48000.0, 4096, 2, 512, 1024
B_FRAME = HOP * 1 * 4 + ROWS * 4          # algorithmic bytes per frame (mono in, RGBA out)
FLOPS_FRAME = 5 * (N * PAD) * 13          # 5 M log2 M, the C2C figure SURVEY.md quotes
WORKLOAD = "cfg1-batched: 48 kHz mono log sweep 20 Hz-20 kHz, N=4096 Hann, zero-pad to M=8192, hop 512, H=1024 log rows 32-22030 Hz, cubic resample, Magma"
# the SAME dict in both arms (ours / --impl reference), so the driver can compare them key by key; everything that
# differs between the arms or between runs (frames per step, sharding, sample sizes) is reported under "run"
CONFIG = {"workload": WORKLOAD, "sample_rate": SR, "window_n": N, "pad_factor": PAD, "hop": HOP, "rows_h": ROWS, "channels": 1,
          "interp": "cubic", "lut": "magma", "bytes_per_frame": B_FRAME,
          "l2": "inputs and outputs of one step are far larger than the 126 MB L2 (no flush needed)"}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_REAL_STDOUT = None


def claim_stdout():
    """Libraries print to stdout on their own (NCCL's "NCCL version ..." banner, for one).  The contract is ONE
    JSON line on stdout, so everything written to fd 1 from here on goes to stderr and only emit() writes to
    the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(local_rank):
    """Pins this rank to the CPUs NVML reports as local to its GPU (before CUDA starts its threads), so that the
    pinned host buffers of the end-to-end leg are first-touched on the GPU's own socket and its PCIe traffic does
    not cross the inter-socket link: what `numactl --cpunodebind` does for a per-GPU service process.
    SBX_BENCH_NUMA=0 disables it.  Returns a short description for the JSON line."""
    if os.environ.get("SBX_BENCH_NUMA", "1") == "0":
        return "off"
    try:
        import pynvml
        pynvml.nvmlInit()
        index = local_rank
        visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        if visible:
            ids = [v.strip() for v in visible.split(",") if v.strip()]
            if local_rank < len(ids) and ids[local_rank].isdigit():
                index = int(ids[local_rank])
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (max(os.cpu_count() or 1, max(os.sched_getaffinity(0)) + 1) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        local = {64 * w + b for w in range(words) for b in range(64) if (int(mask[w]) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        bind_to_gpu_numa_node.original = allowed
        use = local & allowed
        if not use or use == allowed:
            return f"no narrower GPU-local cpu set ({len(allowed)} cpus allowed)"
        os.sched_setaffinity(0, use)
        return f"bound to {len(use)} of {len(allowed)} cpus (NVML cpu affinity of GPU {index})"
    except Exception as exc:                                          # no NVML, cpuset restrictions, ...
        return f"unavailable ({type(exc).__name__})"


def _cpu_time_one(oracle, p, x, lut, pocket, threads):
    t0 = time.perf_counter()
    if pocket:
        oracle.batch_pocketfft(p, x, 1, lut, threads=threads)
    else:
        oracle.batch(p, x, 1, lut, want_rgba=True, threads=threads, fast=True)
    return time.perf_counter() - t0


def cpu_rate(p, signal, frames_budget_s, threads, pocket):
    """frames/s of the CPU port (-O3 -march=native build of the oracle) on `threads` threads over frames, on a bounded
    sample of about frames_budget_s seconds of a recording tiled from `signal`."""
    from oracle import oracle
    from tests import signals
    lut = signals.lut("magma")
    n, hop = p.window_n, p.hop

    def run(frames):
        return _cpu_time_one(oracle, p, np.resize(signal, (frames - 1) * hop + n), lut, pocket, threads)

    run(max(8, threads * 4))                                    # warm-up (tables, page faults, thread pool)
    probe = max(16, threads * 16)
    t = run(probe)
    frames = int(max(probe, min(2_000_000, probe * frames_budget_s / max(t, 1e-6))))
    dt = run(frames)
    return frames / dt, frames, dt


def cpu_reference_rate(frames_budget_s=12.0, threads=None):
    """frames/s of the CPU oracle on this box's host cores, on a bounded sample of the workload: the -O3 -march=native
    build (oracle/Makefile: liboracle_fast.so, built on this machine; same results as the checker build).  Two FFT
    stages are timed -- the oracle's own radix-2 f32 FFT and pocketfft (scipy.fft, complex64) between C-side window /
    split stages, neither of which is FFTW (absent from this image) -- each on ONE thread (what the reference uses:
    everything on the GTK main thread) and on all host cores (frames partitioned over threads); the faster all-core
    number is the baseline value."""
    from oracle import oracle
    from tests import signals
    threads = threads or oracle.max_threads()
    p = oracle.params(sample_rate=SR, window_n=N, pad_factor=PAD, hop=HOP, rows_h=ROWS)
    sweep = signals.log_sweep(int(SR * 60))
    results = {}
    for pocket in (False, True):
        name = "pocketfft" if pocket else "radix2"
        try:
            results[name] = {"all_cores": cpu_rate(p, sweep, frames_budget_s / 3, threads, pocket),
                             "one_thread": cpu_rate(p, sweep, frames_budget_s / 6, 1, pocket)}
        except Exception as exc:                                    # scipy missing: keep the radix-2 number
            log(f"cpu baseline ({name}) failed: {exc!r}")
    best = max(results, key=lambda k: results[k]["all_cores"][0])
    rate, frames, dt = results[best]["all_cores"]
    table = {k: {"all_cores": round(v["all_cores"][0], 1), "one_thread": round(v["one_thread"][0], 1)} for k, v in results.items()}
    fft = "pocketfft complex64 (scipy.fft)" if best == "pocketfft" else "oracle radix-2 f32 FFT"
    cpu_reference_rate.last_step_ms = dt * 1e3                      # duration of the timed sample (the arm's "step")
    cpu_reference_rate.table = table
    return rate, threads, (f"{frames} frames of the workload ({dt:.1f} s, {threads} threads over frames, gcc -O3 -march=native, FFT stage: {fft} "
                           f"standing in for FFTW; window/exp/log recomputed per frame as the reference does; frames/s by FFT stage and "
                           f"thread count: {table})")


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port), all host threads."""
    if rank != 0:
        return
    rates, step_ms = [], []
    threads, sample = None, ""
    for i in range(args.warmup + args.steps):
        rate, threads, sample = cpu_reference_rate(frames_budget_s=min(8.0, 120.0 / max(1, args.steps + args.warmup)))
        if i >= args.warmup:
            rates.append(rate)
            step_ms.append(cpu_reference_rate.last_step_ms)
    value = float(np.mean(rates))
    peaks, src = measured_peaks()
    out = {
        "impl": "reference", "metric": "stft_frames_per_sec", "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.mean(step_ms)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": dict(CONFIG),
        "run": {"note": "CPU path; n_gpus is the arm it is paired with; a step is one bounded sample of the workload (cpu_baseline.sample)"},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample,
                         "by_fft_stage_and_threads": cpu_reference_rate.table, "build": "gcc -O3 -march=native -ffp-contract=off"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=1 << 20, help="frames per GPU per step")
    ap.add_argument("--e2e-frames", type=int, default=1 << 17, help="frames per GPU per end-to-end step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    claim_stdout()

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    numa = bind_to_gpu_numa_node(local_rank)
    log(f"[rank {rank}] numa: {numa}")
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the transform has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # a collective that cannot complete (a rank died) aborts after 5 minutes instead of hanging the box
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=300))

    import spectrogram_rs_b200 as sbx
    from tests import signals

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=torch.device("cuda", local_rank))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    F = args.frames
    plan = sbx.Plan(sample_rate=SR, window_n=N, pad_factor=PAD, hop=HOP, rows_h=ROWS, channels=1, lut="magma", device=local_rank)
    # a non-default stream: sbx_set_stream(NULL) means "the plan's own stream", and the CUDA events
    # below must be recorded on the stream the kernels are launched on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    plan.set_stream(stream.cuda_stream)

    # ---- synthetic input: this rank's time block of one long tiled sweep, generated on the device ----
    sweep = torch.from_numpy(signals.log_sweep(int(SR * 60))).cuda()
    L = sweep.numel()
    n_samples = (F - 1) * HOP + N
    offset = (rank * F * HOP) % L
    reps = (offset + n_samples + L - 1) // L
    pcm = sweep.repeat(reps)[offset:offset + n_samples].contiguous()
    rgba = torch.empty((F, ROWS, 4), dtype=torch.uint8, device="cuda")
    in_bytes, out_bytes = pcm.numel() * 4, rgba.numel()
    log(f"[rank {rank}] F={F} frames/GPU/step, input {in_bytes / 1e9:.2f} GB, output {out_bytes / 1e9:.2f} GB (both > L2: no flush needed)")

    def step():
        got = plan.process_batch_device(pcm, n_samples, rgba, F)
        assert got == F

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = plan.launch_count
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record(stream)
    for i in range(args.steps):
        step()
        ev[i + 1].record(stream)
    barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop()
    launches = plan.launch_count - launches0
    total_ms = max_over_ranks(total_ms)
    ms_per_step = total_ms / args.steps
    value = world * F / (ms_per_step * 1e-3)
    kernel_ms = float(np.mean(per_launch_ms))            # one launch per step: the fused kernel

    # ---- end to end through the host-pointer C ABI: pinned host PCM in, pinned host RGBA out ----------
    Fe = args.e2e_frames
    ns_e = (Fe - 1) * HOP + N
    h_pcm = torch.empty(ns_e, dtype=torch.float32).pin_memory()
    h_pcm.copy_(pcm[:ns_e].cpu())
    h_rgba = torch.empty((Fe, ROWS, 4), dtype=torch.uint8).pin_memory()
    plan.set_stream(None)
    for _ in range(2):
        plan.process_batch_ptr(h_pcm.data_ptr(), ns_e, h_rgba.data_ptr(), Fe)
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        got = plan.process_batch_ptr(h_pcm.data_ptr(), ns_e, h_rgba.data_ptr(), Fe)
        checksum = int(h_rgba[0, 0, 0])                  # the step's result is read on the host
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * Fe * e2e_steps / e2e_s
    assert got == Fe and torch.equal(h_rgba[:64].cuda(), rgba[:64])
    plan.set_stream(stream.cuda_stream)

    # ---- the copy ceiling of this run: the SAME pinned buffers, bytes, chunking (32 MiB of RGBA per chunk, three
    # streams) and ranks as the end-to-end leg, but no kernels -- what the host<->device path alone allows -------------
    CH_F = 8192
    cstreams = [torch.cuda.Stream() for _ in range(3)]
    d_cin = [torch.empty(CH_F * HOP, dtype=torch.float32, device="cuda") for _ in range(3)]
    d_cout = [torch.empty((CH_F, ROWS, 4), dtype=torch.uint8, device="cuda") for _ in range(3)]

    def copy_only_step():
        for c, f0 in enumerate(range(0, Fe, CH_F)):
            nf = min(CH_F, Fe - f0)
            with torch.cuda.stream(cstreams[c % 3]):
                d_cin[c % 3][:nf * HOP].copy_(h_pcm[f0 * HOP:(f0 + nf) * HOP], non_blocking=True)
                h_rgba[f0:f0 + nf].copy_(d_cout[c % 3][:nf], non_blocking=True)
        for cs in cstreams:
            cs.synchronize()

    for _ in range(2):
        copy_only_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        copy_only_step()
    barrier()
    copy_s = max_over_ranks(time.perf_counter() - t0)
    copy_ceiling = world * Fe * e2e_steps / copy_s
    del d_cin, d_cout

    # ---- the same call over ALL devices from one process: sbx_multi_process_batch (rank 0 drives every GPU from its
    # own host threads while the other ranks wait at the barrier): one host recording of world * Fe frames ------------
    e2e_multi = None
    if world > 1:
        barrier()
        if rank == 0:
            try:
                Fm = Fe * world
                ns_m = (Fm - 1) * HOP + N
                hm_pcm = torch.empty(ns_m, dtype=torch.float32).pin_memory()
                hm_pcm.copy_(sweep.repeat((ns_m + L - 1) // L)[:ns_m].cpu())
                hm_rgba = torch.empty((Fm, ROWS, 4), dtype=torch.uint8).pin_memory()
                multi = sbx.MultiPlan(world, sample_rate=SR, window_n=N, pad_factor=PAD, hop=HOP, rows_h=ROWS, channels=1, lut="magma")
                for _ in range(2):
                    multi.process_batch_ptr(hm_pcm.data_ptr(), ns_m, hm_rgba.data_ptr(), Fm)
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    got_m = multi.process_batch_ptr(hm_pcm.data_ptr(), ns_m, hm_rgba.data_ptr(), Fm)
                    checksum = int(hm_rgba[0, 0, 0])
                dt = time.perf_counter() - t0
                # the first and the last device's columns against this rank's device-resident result of the same frames
                assert got_m == Fm and torch.equal(hm_rgba[:64].cuda(), rgba[:64])
                e2e_multi = {"value": Fm * e2e_steps / dt, "frames_per_step": Fm, "steps": e2e_steps,
                             "api": "sbx_multi_process_batch: one process, one host thread per device, one pinned recording"}
                multi.close()
                del hm_pcm, hm_rgba
            except Exception as exc:                        # never lose the headline line
                e2e_multi = {"error": repr(exc)}
            torch.cuda.set_device(local_rank)
        barrier()

    peaks, peak_src = measured_peaks()
    achieved = (F * B_FRAME / 1e9) / (kernel_ms * 1e-3)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")   # written from the ncu --set full capture
    if os.path.exists(tpath):
        try:
            t = json.load(open(tpath))
            traffic = t["dram_bytes_per_frame"] * F
        except Exception:
            traffic = None
    # what actually binds the kernel: instruction issue and the L1/shared-memory pipe (ncu), not HBM.  The secondary
    # roofline is issue slots: warp instructions per frame from the committed ncu capture x the frame rate measured
    # here, against 4 schedulers x 1 instruction/clk x SMs x the SM clock sampled during the timed region.
    secondary = None
    cpath = os.path.join(ROOT, "profiles", "kernel_counters_r02.json")
    if os.path.exists(cpath) and plan.kernel_path.endswith("-q2"):
        try:
            kc = json.load(open(cpath))
            props = torch.cuda.get_device_properties(local_rank)
            clk_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0
            inst_rate = F / (kernel_ms * 1e-3) * kc["warp_instructions_per_frame"]
            issue_peak = 4.0 * props.multi_processor_count * clk_mhz * 1e6
            secondary = {"bound": "issue", "achieved": inst_rate / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-instructions/s",
                         "frac": inst_rate / issue_peak, "warp_instructions_per_frame": kc["warp_instructions_per_frame"],
                         "ncu_pct_of_peak": {"issue_slots": kc["issue_active_pct"], "l1_shared_pipe": kc["l1tex_throughput_pct"],
                                             "fp32_fma_pipe": kc["fma_pipe_pct"], "alu_pipe": kc["alu_pipe_pct"]},
                         "source": "profiles/kernel_counters_r02.json (ncu --set full of the same kernel) x this run's frame rate and SM clock"}
        except Exception:
            secondary = None

    out = {
        "metric": "stft_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(CONFIG),
        "run": {"frames_per_gpu_per_step": F, "sharding": "time-block with N-hop halo, no collectives",
                "step_bytes": "input %.2f GB, output %.2f GB per GPU per step" % (in_bytes / 1e9, out_bytes / 1e9)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_src,
                     "kernel": "fused_q2_kernel<Q2Geom<12,1>>" if plan.kernel_path.endswith("-q2") else "fused_frames_kernel<Geom<12,1,1>>",
                     "kernel_path": plan.kernel_path, "kernel_ms": kernel_ms,
                     "fp32_tflops_5MlogM": value / world * FLOPS_FRAME / 1e12,
                     "binding_resource": "instruction issue and the L1/shared-memory pipe, not HBM: see `secondary`",
                     "secondary": secondary,
                     "note": "the HBM fraction is the mandated figure (algorithmic bytes hop*4 + H*4 per frame against the measured copy "
                             "bandwidth); the path does 44-87 flop per compulsory byte against a 12 flop/B ridge (SURVEY.md D6), so what "
                             "binds it is the SM: ncu (profiles/ncu_r02_cfg1_*) L1/shared pipe 75.5 %, issue slots 61 %, FMA pipe 56 %, "
                             "DRAM traffic = 0.98 x the algorithmic bytes"},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": ns_e * 4 * world, "d2h_bytes_per_step": Fe * ROWS * 4 * world,
                "frames_per_gpu_per_step": Fe, "steps": e2e_steps, "api": "sbx_process_batch (host pointers, pinned), one process per GPU", "numa": numa,
                "d2h_gbs_per_gpu": e2e_value / world * ROWS * 4 / 1e9, "h2d_gbs_per_gpu": e2e_value / world * HOP * 4 / 1e9,
                "copy_ceiling": {"frames_per_s": copy_ceiling, "frac": e2e_value / copy_ceiling,
                                 "what": "the same pinned buffers, bytes per frame (2 KB in, 4 KB out), chunking and ranks with NO kernels, "
                                         "timed right after the end-to-end leg: the host<->device path's own limit in this run"},
                "note": "bound by the PCIe D2H of the RGBA output (4 KB per frame): chunked, three streams, H2D/kernel/D2H overlapped"
                        + ("; with several GPUs the ranks share the host's PCIe / memory path (DESIGN.md section 6, profiles/copy_ceiling_r02.json)" if world > 1 else "")},
        "gpu_launches": launches * world,
        "clocks": clocks,
    }

    if e2e_multi is not None:
        out["e2e"]["one_process_multi_device"] = e2e_multi
        out["e2e"]["one_process_per_gpu"] = {"value": e2e_value, "api": out["e2e"]["api"]}
        if e2e_multi.get("value", 0.0) > e2e_value:          # the headline is the better of the two public calls
            out["e2e"]["copy_ceiling"]["frac"] = e2e_multi["value"] / copy_ceiling
            out["e2e"].update(value=e2e_multi["value"], api=e2e_multi["api"], frames_per_gpu_per_step=Fe,
                              d2h_gbs_per_gpu=e2e_multi["value"] / world * ROWS * 4 / 1e9,
                              h2d_gbs_per_gpu=e2e_multi["value"] / world * HOP * 4 / 1e9)
    if not args.no_extras:
        # cfg4 / cfg5 as the multi-GPU configurations they are: every rank takes part (channel / time-block sharding)
        try:
            sharded = sharded_extras(sbx, signals, torch, dist, rank, world, local_rank, barrier, max_over_ranks)
        except Exception as exc:
            sharded = {"error": repr(exc)}
    if world > 1:                                            # the last collective: the rank-0-only legs below run alone
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and not args.no_extras:
        try:
            out["extra"] = extras(sbx, signals, torch)
            out["extra"].update(sharded)
        except Exception as exc:                            # extras must never lose the headline line
            out["extra"] = {"error": repr(exc)}
        if isinstance(out["extra"].get("cfg2_live_latency_us"), dict):
            # the second half of BASELINE.json's metric ("p50 live-frame latency"), surfaced at the top level
            lat = out["extra"]["cfg2_live_latency_us"]
            out["live_latency_us"] = {"p50": lat["p50"], "p99": lat["p99"], "workload": "cfg2: 48 kHz stereo, 2048-pt frames, one at a time from the pinned ring, 1 GPU"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if getattr(bind_to_gpu_numa_node, "original", None):          # the CPU baseline gets every host core back
            os.sched_setaffinity(0, bind_to_gpu_numa_node.original)
        rate, cores, sample = cpu_reference_rate()
        out["cpu_baseline"] = {"value": rate, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample,
                               "by_fft_stage_and_threads": cpu_reference_rate.table, "build": "gcc -O3 -march=native -ffp-contract=off"}
    if rank == 0:
        emit(out)


def _checksum(torch, rgba):
    """Order-sensitive checksum of a block of columns: (sum of pixels, sum of pixel * position), both mod 2^61."""
    v = rgba.reshape(-1).view(torch.int32).to(torch.int64) & 0xffffffff
    w = (torch.arange(v.numel(), device=v.device, dtype=torch.int64) % 65521) + 1
    mod = (1 << 61) - 1
    return [int(v.sum().item() % mod), int((v * w).sum().item() % mod)]


def sharded_extras(sbx, signals, torch, dist, rank, world, local_rank, barrier, max_over_ranks):
    """BASELINE cfg4 and cfg5 as the multi-GPU configurations they are.  Every rank takes part; rank 0 returns the
    results.  cfg4: 64 planar channels, 64 / G whole channels per rank (no halo).  cfg5: for every FFT size ONE shared
    recording split into G time blocks with the N - hop halo (frame_align 2: mono pairs).  No data-path collective:
    torch.distributed carries the barrier, the max of the device times and the gathered checksums only.  Each rank's
    columns are checked against rank 0's own UNSHARDED transform of the same recording (checksums per block)."""
    from oracle import oracle
    dev = torch.cuda.current_device()
    stream = torch.cuda.current_stream()
    peak = measured_peaks()[0]["hbm_gbs"]
    res = {}

    def gather(obj):
        if world == 1:
            return [obj]
        got = [None] * world
        dist.all_gather_object(got, obj)
        return got

    def timed(fn, reps):
        fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        barrier()
        return max_over_ranks(a.elapsed_time(b)) * 1e-3 / reps

    # ---- cfg4 -------------------------------------------------------------------------------------------------
    n4, hop4, ch4, ns4 = 4096, 512, 64, 48000 * 60
    plan = sbx.Plan(sample_rate=48000.0, window_n=n4, pad_factor=2, hop=hop4, rows_h=1024, channels=1, lut="magma", device=dev)
    plan.set_stream(stream.cuda_stream)
    F4 = plan.num_frames(ns4)
    base = np.stack([signals.channel_tone(c, ns4) for c in range(8)])        # channel c = tone c mod 8 (host time)
    c_lo, c_hi = sbx.sharding.channel_blocks(ch4, world)[rank]
    mine = torch.from_numpy(base[[c % 8 for c in range(c_lo, c_hi)]]).cuda().contiguous()
    out4 = torch.empty((c_hi - c_lo, F4, 1024, 4), dtype=torch.uint8, device="cuda")
    sec = timed(lambda: plan.process_batch_device_planar(mine, c_hi - c_lo, ns4, ns4, out4, (c_hi - c_lo) * F4), 3)
    sums = gather([_checksum(torch, out4[i]) for i in range(c_hi - c_lo)])
    ok4 = None
    if rank == 0:
        flat = [c for part in sums for c in part]
        ref = {}
        for c in range(8):                                                   # rank 0's own transform of the 8 distinct channels
            one = torch.empty((F4, 1024, 4), dtype=torch.uint8, device="cuda")
            plan.process_batch_device(torch.from_numpy(base[c]).cuda(), ns4, one, F4)
            plan.synchronize()
            ref[c] = _checksum(torch, one)
        ok4 = len(flat) == ch4 and all(flat[c] == ref[c % 8] for c in range(ch4))
    res["cfg4_frames_per_s"] = ch4 * F4 / sec
    res["cfg4_sharded"] = {"n_gpus": world, "channels_per_gpu": ch4 // world, "frames_per_s": ch4 * F4 / sec,
                           "hbm_frac": ch4 * F4 / sec * (hop4 * 4 + 4096) / 1e9 / (peak * world),
                           "checksums_equal_rank0_unsharded": ok4,
                           "workload": "64 planar channels x 60 s @ 48 kHz (5618 frames each; channel c = tone c mod 8 + noise), N=4096 -> M=8192, "
                                       "hop 512, H=1024, Magma; 64/G whole channels per GPU in one grid (sbx_process_batch_device_planar), no halo"}
    plan.close()
    del mine, out4

    # ---- cfg5 -------------------------------------------------------------------------------------------------
    sweep = {}
    ns5 = 48000 * 600                                                        # 600 s @ 48 kHz, white noise seed 3
    x_host = signals.white_noise(ns5, seed=3)
    x_all = torch.from_numpy(x_host).cuda() if rank == 0 else None
    for n5 in (512, 1024, 2048, 4096, 8192, 16384, 32768, 65536):
        hop5 = n5 // 8
        plan = sbx.Plan(sample_rate=48000.0, window_n=n5, pad_factor=4, hop=hop5, rows_h=2048, channels=1, lut="magma", device=dev)
        plan.set_stream(stream.cuda_stream)
        blocks = sbx.sharding.time_blocks(ns5, n5, hop5, world, frame_align=2)
        b = blocks[rank]
        F5 = blocks[-1].frame_hi
        x5 = torch.from_numpy(x_host[b.sample_lo:b.sample_hi]).cuda()
        out5 = torch.empty((max(b.frames, 1), 2048, 4), dtype=torch.uint8, device="cuda")
        reps = 3 if n5 <= 8192 else 2
        sec = timed(lambda: plan.process_batch_device(x5, b.samples, out5, b.frames), reps)
        sums = gather(_checksum(torch, out5[:b.frames]))
        ok5 = None
        if rank == 0:
            whole = torch.empty((F5, 2048, 4), dtype=torch.uint8, device="cuda")
            plan.process_batch_device(x_all, ns5, whole, F5)
            plan.synchronize()
            ok5 = all(sums[g] == _checksum(torch, whole[blk.frame_lo:blk.frame_hi]) for g, blk in enumerate(blocks))
            del whole
        fps = F5 / sec
        entry = {"frames_per_s": fps, "frames": F5, "hbm_frac": fps * (hop5 * 4 + 2048 * 4) / 1e9 / (peak * world),
                 "kernel_path": plan.kernel_path, "halo_samples": n5 - hop5, "checksums_equal_rank0_unsharded": ok5}
        if rank == 0:                                                        # the CPU port beside every size, bounded samples
            try:
                p5 = oracle.params(sample_rate=48000.0, window_n=n5, pad_factor=4, hop=hop5, rows_h=2048)
                threads = oracle.max_threads()
                entry["cpu_frames_per_s"] = {"one_thread": round(cpu_rate(p5, x_host, 0.6, 1, False)[0], 1),
                                             "all_cores": round(cpu_rate(p5, x_host, 0.6, threads, False)[0], 1), "cores": threads,
                                             "fft": "oracle radix-2 f32 (-O3 -march=native)"}
            except Exception as exc:
                entry["cpu_frames_per_s"] = {"error": repr(exc)}
        sweep[str(n5)] = entry
        plan.close()
        del x5, out5
    res["cfg5_sweep"] = {"n_gpus": world, "sizes": sweep,
                         "workload": "FFT sizes 512-65536, P=4, hop N/8, H=2048, 600 s of 48 kHz white noise (seed 3): ONE recording split into G "
                                     "time blocks with the N-hop halo, block boundaries on even frames; frames/s = all frames / max over ranks"}
    if world == 1:
        res["cfg5_sweep_1gpu"] = {k: {"frames_per_s": v["frames_per_s"], "hbm_frac": v["hbm_frac"]} for k, v in sweep.items()}
    return res if rank == 0 else {}


def extras(sbx, signals, torch):
    """Secondary measurements reported beside the headline: cfg3 throughput, cfg2 live latency."""
    res = {}
    # cfg3 geometry: 96 kHz mono, N 8192 -> M 16384, hop 256, Viridis
    n3, hop3 = 8192, 256
    F3 = (345_600_000 - n3) // hop3 + 1                  # the configured hour at 96 kHz: 1 349 969 frames, 1.38 GB in, 5.53 GB out
    plan = sbx.Plan(sample_rate=96000.0, window_n=n3, pad_factor=2, hop=hop3, rows_h=1024, channels=1, lut="viridis",
                    device=torch.cuda.current_device())
    stream = torch.cuda.current_stream()                 # main() made a side stream current
    assert stream.cuda_stream != 0
    plan.set_stream(stream.cuda_stream)
    ns = (F3 - 1) * hop3 + n3
    x = torch.from_numpy(signals.sweeps_plus_noise(1 << 22, 96000.0)).cuda()
    x = x.repeat((ns + x.numel() - 1) // x.numel())[:ns].contiguous()
    out = torch.empty((F3, 1024, 4), dtype=torch.uint8, device="cuda")
    res["cfg3_seconds_per_hour_of_audio"] = None
    for _ in range(3):
        plan.process_batch_device(x, ns, out, F3)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record(stream)
    for _ in range(5):
        plan.process_batch_device(x, ns, out, F3)
    b.record(stream)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    res["cfg3_frames_per_s"] = F3 / (ms * 1e-3)
    res["cfg3_seconds_per_hour_of_audio"] = ms * 1e-3
    res["cfg3_workload"] = "1 hour of 96 kHz mono (345.6 M samples), N=8192 -> M=16384, hop 256, H=1024, Viridis, %d frames/step (the whole hour in one call)" % F3
    res["cfg3_hbm_frac"] = res["cfg3_frames_per_s"] * (hop3 * 4 + 4096) / 1e9 / measured_peaks()[0]["hbm_gbs"]
    res["cfg3_kernel_path"] = plan.kernel_path
    # the column store: cfg3's columns blitted into a 65536-column row-major image (256 MiB), a pure
    # HBM-streaming transpose: bytes = columns read + image written
    W = 1 << 16
    img = torch.empty((1024, W, 4), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        plan.blit_columns_device(out, W, img, W, W * 4, 0)
    torch.cuda.synchronize()
    a.record(stream)
    for _ in range(5):
        plan.blit_columns_device(out, W, img, W, W * 4, 0)
    b.record(stream)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    res["column_store_blit"] = {"gbs": 2 * W * 4096 / (ms * 1e-3) / 1e9,
                                "hbm_frac": 2 * W * 4096 / (ms * 1e-3) / 1e9 / measured_peaks()[0]["hbm_gbs"],
                                "what": "sbx_blit_columns_device: 65536 columns x 1024 rows -> row-major RGBA image (tiled transpose)"}
    plan.close()
    del x, out, img

    # sbx_process_batch from PAGEABLE host memory (a plain numpy array, what a caller without pinned buffers passes):
    # the runtime stages every copy through its own pinned buffer, so the copies do not overlap the kernels
    plan = sbx.Plan(sample_rate=SR, window_n=N, pad_factor=PAD, hop=HOP, rows_h=ROWS, channels=1, lut="magma",
                    device=torch.cuda.current_device())
    Fp = 1 << 15
    xp = np.resize(signals.log_sweep(int(SR * 60)), (Fp - 1) * HOP + N)
    outp = np.empty((Fp, ROWS, 4), np.uint8)
    plan.process_batch(xp, outp)
    t0 = time.perf_counter()
    for _ in range(3):
        plan.process_batch(xp, outp)
    res["e2e_pageable_frames_per_s"] = {"value": 3 * Fp / (time.perf_counter() - t0), "frames_per_step": Fp,
                                        "what": "sbx_process_batch with pageable (numpy) input and output buffers, cfg1, 1 GPU"}
    plan.close()
    del xp, outp

    # cfg2: live single stream, 48 kHz stereo, 2048-pt frames pushed one at a time through the pinned ring
    plan = sbx.Plan(sample_rate=48000.0, window_n=2048, pad_factor=2, hop=2048, rows_h=1024, channels=2, lut="magma",
                    device=torch.cuda.current_device())
    live = plan.live(4096)
    lr = signals.stereo_loopback(2048 * 64)
    live.measure(lr[:2048], 1000)                                        # warm-up
    lat = np.concatenate([live.measure(lr[(i % 64) * 2048:(i % 64 + 1) * 2048], 100) for i in range(100)])
    res["cfg2_live_latency_us"] = {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99)),
                                   "mean": float(lat.mean()), "samples": int(lat.size),
                                   "what": "host CLOCK_MONOTONIC around: push 2048 stereo samples into the pinned ring + pop one RGBA column (one CUDA graph with a single kernel node that reads the 16 KB pinned window and writes the 4 KB pinned column itself; completion polled on a pinned flag); native loop sbx_live_measure"}
    # the same frame, kernel only (device-resident input and output, CUDA events on the launching stream)
    d_lr = torch.from_numpy(lr[:2048]).cuda()
    d_col = torch.empty((1, 1024, 4), dtype=torch.uint8, device="cuda")
    plan.set_stream(stream.cuda_stream)
    for _ in range(20):
        plan.process_batch_device(d_lr, 2048, d_col, 1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record(stream)
    for _ in range(200):
        plan.process_batch_device(d_lr, 2048, d_col, 1)
    b.record(stream)
    torch.cuda.synchronize()
    res["cfg2_kernel_only_us"] = a.elapsed_time(b) * 1e3 / 200
    # the reference application's own live geometry: N = (0.05 s * 48 kHz) = 2400, hop 58 (mixed-radix path)
    plan_l = sbx.Plan(sample_rate=48000.0, window_n=2400, pad_factor=2, hop=58, rows_h=1024, channels=2, lut="magma",
                      device=torch.cuda.current_device())
    live_l = plan_l.live(4096)
    live_l.push(lr[:2400])
    live_l.measure(lr[:58], 500)
    lat = live_l.measure(lr[58:116], 5000)
    res["app_live_n2400_latency_us"] = {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99)),
                                        "what": "reference app geometry: N 2400, M 4800, hop 58 samples (819 frames/s needed)"}
    live_l.close(); plan_l.close()
    live.close()
    plan.close()

    return res


if __name__ == "__main__":
    main()
