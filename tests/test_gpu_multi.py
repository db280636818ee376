"""-m gpu: the several-device entry points (sbx_multi_*), the two-thread use of the live ring, and the named regression
test for the early abort in sbx_live_push.

A recording sharded over devices must equal the one-device result byte for byte: frame f depends only on samples
[f*hop, f*hop+N) (src/fourier/audio_transform.rs:37-41), blocks carry an N-hop halo, and nothing is exchanged.
The tests that need a second device skip on a one-GPU box; the same code path (host threads, segments, block
arithmetic) runs with one device everywhere.
"""
import threading

import numpy as np
import pytest

from tests import signals

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sbx():
    import spectrogram_rs_b200 as s
    return s


def _n_devices():
    import torch
    return torch.cuda.device_count()


def _recording(channels, n_samples, seed=5):
    rng = np.random.Generator(np.random.PCG64(seed))
    if channels == 1:
        return (signals.log_sweep(n_samples) + 0.05 * rng.standard_normal(n_samples)).astype(np.float32)
    return signals.stereo_loopback(n_samples, seed=seed)


GEOMETRIES = [
    # N, P, channels, hop, H, frames
    (4096, 2, 1, 512, 1024, 1001),     # cfg1 geometry, odd frame count, mono pairs
    (2048, 2, 2, 2048, 256, 130),      # cfg2 geometry, no overlap
    (8192, 2, 1, 256, 128, 333),       # cfg3 geometry (real-sequence kernel)
    (2400, 2, 2, 58, 64, 77),          # the application's window (mixed-radix path)
    (16384, 4, 1, 2048, 64, 41),       # four-kernel path
]


@pytest.mark.parametrize("n,pad,ch,hop,rows,frames", GEOMETRIES)
def test_multi_one_device_equals_the_plan(sbx, magma, monkeypatch, n, pad, ch, hop, rows, frames):
    """sbx_multi with a single device is the same pipeline as sbx_process_batch (several chunks: 1 MiB of RGBA each)."""
    monkeypatch.setenv("SBX_BATCH_CHUNK_MB", "1")
    x = _recording(ch, n + hop * (frames - 1) + 3)
    kw = dict(window_n=n, pad_factor=pad, hop=hop, rows_h=rows, channels=ch, lut=magma)
    with sbx.Plan(**kw) as plan:
        want = plan.process_batch(x)
    with sbx.MultiPlan([0], **kw) as multi:
        got = multi.process_batch(x)
        assert multi.time_block(x.shape[0], 0)[:2] == (0, frames)
        assert multi.launch_count > 0
    assert got.shape == (frames, rows, 4) and np.array_equal(got, want)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_multi_time_blocks_are_sharding_py(sbx, magma, world):
    """sbx_multi_time_block == sharding.time_blocks (frame_align 2 for mono plans): same boundaries, same halo."""
    if _n_devices() < world:
        pytest.skip(f"needs {world} devices")
    for ch, n, hop, ns in ((1, 1024, 256, 1024 + 256 * 1000 + 17), (2, 512, 512, 512 * 37), (1, 1024, 100, 1024)):
        with sbx.MultiPlan(world, window_n=n, hop=hop, rows_h=32, channels=ch, lut=magma) as multi:
            blocks = sbx.sharding.time_blocks(ns, n, hop, world, frame_align=2 if ch == 1 else 1)
            for b in blocks:
                assert multi.time_block(ns, b.rank) == (b.frame_lo, b.frame_hi, b.sample_lo, b.sample_hi)


@pytest.mark.parametrize("n,pad,ch,hop,rows,frames", GEOMETRIES)
def test_multi_devices_equal_one_device(sbx, magma, monkeypatch, n, pad, ch, hop, rows, frames):
    """Every available device takes a time block of ONE host recording; the columns equal the one-device call's
    byte for byte, and each device's own slab equals what that device computes from its block alone."""
    world = min(_n_devices(), 8)
    if world < 2:
        pytest.skip("needs at least 2 devices")
    monkeypatch.setenv("SBX_BATCH_CHUNK_MB", "1")
    x = _recording(ch, n + hop * (frames - 1) + 3)
    kw = dict(window_n=n, pad_factor=pad, hop=hop, rows_h=rows, channels=ch, lut=magma)
    with sbx.Plan(device=0, **kw) as plan:
        want = plan.process_batch(x)
    import torch
    torch.cuda.set_device(0)
    with sbx.MultiPlan(world, **kw) as multi:
        got = multi.process_batch(x)
        assert torch.cuda.current_device() == 0                # the caller's current device is put back
        assert np.array_equal(got, want)
        # the last device alone, on its own block of samples (a real second device, not a shard on device 0)
        lo, hi, s_lo, s_hi = multi.time_block(x.shape[0], world - 1)
        assert hi == frames and s_hi <= x.shape[0]
    with sbx.Plan(device=world - 1, **kw) as plan:
        assert np.array_equal(plan.process_batch(x[s_lo:s_hi]), want[lo:hi])


@pytest.mark.parametrize("ch,n_seq", [(1, 5), (2, 3)])
def test_multi_planar_equals_per_sequence(sbx, magma, monkeypatch, ch, n_seq):
    """cfg4's layout: whole channels per device, several sequences pipelined through one device's slots."""
    monkeypatch.setenv("SBX_BATCH_CHUNK_MB", "1")
    n, hop, rows, frames = 4096, 512, 256, 301
    ns = n + hop * (frames - 1)
    seqs = np.stack([_recording(ch, ns, seed=20 + s) for s in range(n_seq)])
    kw = dict(window_n=n, pad_factor=2, hop=hop, rows_h=rows, channels=ch, lut=magma)
    with sbx.Plan(**kw) as plan:
        want = np.stack([plan.process_batch(seqs[s]) for s in range(n_seq)])
    for world in sorted({1, min(_n_devices(), n_seq)}):
        with sbx.MultiPlan(world, **kw) as multi:
            assert np.array_equal(multi.process_batch_planar(seqs), want), world


def test_multi_errors(sbx, magma):
    with pytest.raises(sbx.SbxError):
        sbx.MultiPlan([0, 0], window_n=1024, lut=magma)                      # a device listed twice
    with pytest.raises(sbx.SbxError):
        sbx.MultiPlan([_n_devices()], window_n=1024, lut=magma)              # ordinal out of range
    with pytest.raises(sbx.SbxError):
        sbx.MultiPlan([0], window_n=5462, lut=magma)                         # unsupported window: plan error surfaces
    with sbx.MultiPlan([0], window_n=1024, hop=256, rows_h=64, lut=magma) as multi:
        x = signals.stereo_loopback(1024 * 8)
        small = np.empty((3, 64, 4), np.uint8)
        with pytest.raises(sbx.SbxError) as e:
            multi.process_batch_ptr(x.ctypes.data, x.shape[0], small.ctypes.data, 3)
        assert e.value.status == -5
        assert multi.process_batch(x[:100]).shape[0] == 0                    # fewer than N samples: no frames


def test_live_ring_two_threads_stress(sbx, magma):
    """One producer thread (the cpal callback, audio_input_list_model.rs:63-72) and one consumer thread (the GTK
    draw callback) on the same sbx_live: 10^6 samples in random chunk sizes.  The consumer only pops when a full
    window is buffered and the producer re-offers what did not fit, so no sample is dropped and the columns must equal
    the batch transform of the same recording, in order."""
    n, hop, rows, total = 1024, 256, 64, 1_000_000
    x = signals.stereo_loopback(total, seed=9)
    with sbx.Plan(window_n=n, hop=hop, rows_h=rows, channels=2, lut=magma) as plan:
        want = plan.process_batch(x)
        live = plan.live(4096)
        cols, errors = [], []
        done = threading.Event()

        def produce():
            try:
                rng = np.random.Generator(np.random.PCG64(3))
                pos = 0
                while pos < total:
                    chunk = int(rng.integers(1, 1500))
                    end = min(total, pos + chunk)
                    while pos < end:                              # push returns how many fitted: offer the rest again
                        pos += live.push(x[pos:end])
            except Exception as exc:                              # pragma: no cover
                errors.append(exc)
            finally:
                done.set()

        def consume():
            try:
                while True:
                    if live.buffered() >= n:
                        col = live.pop_column()
                        assert col is not None
                        cols.append(col)
                    elif done.is_set() and live.buffered() < n:
                        break
            except Exception as exc:                              # pragma: no cover
                errors.append(exc)

        threads = [threading.Thread(target=produce), threading.Thread(target=consume)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=300)
        assert not errors, errors
        assert not any(t.is_alive() for t in threads)
        live.close()
    got = np.stack(cols)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(got, want)


def test_live_push_abort_regression(sbx, magma):
    """Regression test for the early abort recorded in round 1's hang.log (sbx_live_push): pushes that overrun the
    ring, wrap it at odd offsets, arrive empty, mono or NULL must return a count and never write outside the ring or
    take the process down."""
    import ctypes as C
    lib = sbx.binding.load()
    n, hop = 256, 64
    with sbx.Plan(window_n=n, hop=hop, rows_h=32, channels=2, lut=magma) as plan:
        for capacity in (257, 1000, 4096):
            live = plan.live(capacity)
            h = live._h
            x = signals.stereo_loopback(3 * capacity + 11)
            assert lib.sbx_live_push(h, None, 100, 2) == 0                   # NULL samples
            assert lib.sbx_live_push(h, x.ctypes.data, 100, 3) == 0          # channel count the ring does not take
            assert lib.sbx_live_push(None, x.ctypes.data, 100, 2) == 0       # NULL stream
            assert live.push(x[:0]) == 0                                     # empty
            assert live.push(x) == capacity and live.buffered() == capacity  # three times the capacity: fits once
            assert live.push(x[:5]) == 0                                     # full
            model = x[:capacity].copy()
            pos, popped = capacity, 0
            for step in range(200):                                          # wrap the ring many times at odd offsets
                col = live.pop_column()
                assert col is not None and np.array_equal(col, plan.process_column(model[:n]))
                model = model[hop:]
                chunk = x[(pos % capacity):(pos % capacity) + 1 + (step * 37) % (2 * hop)]
                mono = step % 3 == 0
                took = live.push(chunk[:, 0].copy(), channels=1) if mono else live.push(chunk)
                assert took == min(len(chunk), capacity - len(model))
                add = chunk[:took].copy()
                if mono:
                    add[:, 1] = add[:, 0]
                model = np.concatenate([model, add])
                pos += took
                assert live.buffered() == len(model)
                if len(model) < n:
                    break
            live.close()


def test_host_register_makes_caller_memory_usable(sbx, magma):
    """sbx_host_register / sbx_host_unregister: page-lock memory the caller already owns (a Rust Vec<f32>, here numpy
    arrays) and run the host-pointer batch on it; results equal the pageable call."""
    lib = sbx.binding.load()
    n, hop, rows, frames = 2048, 512, 128, 700
    x = _recording(1, n + hop * (frames - 1))
    out = np.zeros((frames, rows, 4), np.uint8)
    with sbx.Plan(window_n=n, hop=hop, rows_h=rows, channels=1, lut=magma) as plan:
        want = plan.process_batch(x)
        assert lib.sbx_host_register(x.ctypes.data, x.nbytes) == 0
        assert lib.sbx_host_register(out.ctypes.data, out.nbytes) == 0
        try:
            assert plan.process_batch_ptr(x.ctypes.data, x.shape[0], out.ctypes.data, frames) == frames
        finally:
            assert lib.sbx_host_unregister(out.ctypes.data) == 0
            assert lib.sbx_host_unregister(x.ctypes.data) == 0
        assert np.array_equal(out, want)
    assert lib.sbx_host_register(None, 16) == -1 and lib.sbx_host_unregister(None) == -1
    assert lib.sbx_host_unregister(x.ctypes.data) == -3                     # not registered (any more)
