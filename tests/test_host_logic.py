"""CPU: host-side mirror of the reference interface (ring buffer, hop cursor, literals) and the
multi-GPU sharding arithmetic, exercised with a fake transform / the oracle -- no GPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle
from tests import signals

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sbx():
    import spectrogram_rs_b200 as s
    return s


class OracleTransform:
    """AudioTransform backed by the CPU oracle: lets the host logic run without a GPU."""

    def __init__(self, sample_rate, n, pad=2):
        self.sr, self.n, self.pad, self.calls = sample_rate, n, pad, 0

    def sample_rate(self):
        return self.sr

    def num_input_samples(self):
        return self.n

    def process(self, samples):
        self.calls += 1
        return oracle.frame_magnitudes(np.asarray(samples, np.float32).reshape(-1, 2), self.n, self.pad)


def test_heaprb_push_iter_drops_overflow(sbx):
    prod, cons = sbx.HeapRb(8).split()
    assert prod.push_mono([0.1, 0.2, 0.3]) == 3                       # audio_input_list_model.rs:67-69
    assert list(cons.iter()) == [(0.1, 0.1), (0.2, 0.2), (0.3, 0.3)]
    assert prod.push_interleaved([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12]) == 5   # only 5 of 6 pairs fit
    assert len(cons) == 8 and cons.skip(3) == 3 and len(cons) == 5
    assert cons.skip(100) == 5 and len(cons) == 0


def test_audio_stream_transform_framing_and_skip_quirk(sbx):
    """Frame f covers [f*hop, f*hop+N); skip(hop) also runs after the failing attempt
    (src/fourier/audio_transform.rs:37-41), discarding up to hop samples per call."""
    sr, n = 1000.0, 256
    hop_s = 0.064                                                     # (0.064 * 1000) as usize = 64
    ring = sbx.HeapRb(4096)
    x = signals.stereo_loopback(n + 64 * 3 + 10, sr)
    ring.push_iter(map(tuple, x))
    t = OracleTransform(sr, n)
    stream = sbx.AudioStreamTransform(ring, t, hop_s)
    frames = list(stream.process())
    assert len(frames) == 4 == oracle.num_frames(len(x), n, 64)
    for f, m in enumerate(frames):
        assert np.array_equal(m, oracle.frame_magnitudes(x[f * 64:f * 64 + n], n, 2))
    assert t.calls == 5                                               # 4 frames + the failing attempt
    # 458 pushed, 4 successful skips (256) + the skip after the failed attempt (64)
    assert len(ring) == len(x) - 5 * 64
    assert list(stream.process()) == []                               # and it keeps eating samples
    assert len(ring) == len(x) - 6 * 64


def test_public_fields_are_mutable_like_the_reference(sbx):
    ring = sbx.HeapRb(16)
    s = sbx.AudioStreamTransform(sbx.HeapRb(1), OracleTransform(100.0, 100), 2.0 / 1024.0)
    s.input_stream = ring                                             # gpu_spectrogram.rs:34
    s.transform = OracleTransform(48000.0, 2400)                      # gpu_spectrogram.rs:321
    assert s.transform.num_input_samples() == 2400 and s.input_stream is ring


def test_literal_conversions_match_oracle(sbx):
    L = oracle.lib()
    for sec, sr in ((0.05, 48000.0), (0.05, 44100.0), (2.0 / 1024.0, 48000.0), (2.5 / 2048.0, 48000.0), (1.0, 100.0)):
        assert sbx.samples_from_seconds(sec, sr) == L.orc_window_from_period(sec, sr)


def test_color_scheme_objects(sbx):
    schemes = sbx.default_color_schemes()
    # default_color_schemes(), src/colorscheme.rs:129-149: all 19, in the reference's order and with its display names
    assert [s.name for s in schemes] == [
        "Blue-Yellow-Red (Stereo)", "Magma", "Viridis", "Blue-Red (Stereo)", "Spectral (Stereo)", "Green-Yellow-Red (Stereo)",
        "Green-Pink (Stereo)", "Orange-Purple (Stereo)", "Inferno", "Plasma", "Cividis", "Cube-helix", "Turbo", "Cool", "Reds",
        "Blues", "Greens", "Greys", "Oranges"]
    assert [s.is_stereo for s in schemes] == [True, False, False] + [True] * 5 + [False] * 11
    assert all(s.background() == (0, 0, 0) for s in schemes if s.is_stereo)            # Color { r: 0, g: 0, b: 0 }
    magma = schemes[1]
    assert not magma.is_stereo and magma.background() == (0, 0, 4) and magma.foreground() == (252, 253, 191)
    st = sbx.ColorScheme.new_stereo("viridis", (0, 0, 0), "x")
    assert st.is_stereo and st.background() == (0, 0, 0) and st.foreground() == tuple(int(v) for v in st.lut[128])


def test_time_blocks_partition_frames_exactly(sbx):
    sh = sbx.sharding
    for (S, n, hop) in ((2880000, 4096, 512), (100000, 8192, 256), (5000, 4096, 512), (4096, 4096, 512), (10, 4096, 1)):
        total = sh.num_frames(S, n, hop)
        assert total == oracle.num_frames(S, n, hop)
        for world in (1, 2, 3, 4, 8):
            blocks = sh.time_blocks(S, n, hop, world)
            assert [b.rank for b in blocks] == list(range(world))
            assert blocks[0].frame_lo == 0 and blocks[-1].frame_hi == total
            for a, b in zip(blocks, blocks[1:]):
                assert a.frame_hi == b.frame_lo
            for b in blocks:
                if b.frames:
                    assert b.sample_lo == b.frame_lo * hop and b.sample_hi == (b.frame_hi - 1) * hop + n <= S
                    assert sh.num_frames(b.samples, n, hop) == b.frames          # halo = N - hop samples
    for world in (2, 3, 8):                                             # mono plans: pair-aligned boundaries
        blocks = sh.time_blocks(2880000, 4096, 512, world, frame_align=2)
        assert all(b.frame_lo % 2 == 0 for b in blocks) and blocks[-1].frame_hi == 5618
        assert sum(b.frames for b in blocks) == 5618
    assert sh.channel_blocks(64, 8) == [(8 * g, 8 * g + 8) for g in range(8)]
    assert sh.channel_blocks(3, 2) == [(0, 1), (1, 3)]


def test_sharded_oracle_equals_unsharded(sbx, magma):
    """shard == unsharded, byte for byte (the only thing there is to test about the multi-GPU path)."""
    n, hop = 1024, 128
    x = signals.log_sweep(n + hop * 37, duration=0.2)
    p = oracle.params(window_n=n, hop=hop, rows_h=128)
    whole, _ = oracle.batch(p, x, 1, magma)
    parts = [oracle.batch(p, x[b.sample_lo:b.sample_hi], 1, magma)[0] for b in sbx.sharding.time_blocks(len(x), n, hop, 3)]
    assert np.array_equal(np.concatenate(parts), whole)


WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
import spectrogram_rs_b200 as sbx
from oracle import oracle
from tests import signals
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n, hop, rows = 1024, 128, 64
x = signals.log_sweep(n + hop * 41, duration=0.25)
lut = signals.lut("magma")
p = oracle.params(window_n=n, hop=hop, rows_h=rows)
blk = sbx.sharding.time_blocks(len(x), n, hop, world)[rank]
mine, _ = oracle.batch(p, x[blk.sample_lo:blk.sample_hi], 1, lut)       # no data-path collective
assert mine.shape[0] == blk.frames
dist.barrier()
t = sbx.sharding.max_over_ranks(10.0 + rank, dist)                       # the timing reduction bench.py uses
assert t == 10.0 + world - 1
gathered = [None] * world
dist.all_gather_object(gathered, mine)                                   # test-only gather to compare
if rank == 0:
    whole, _ = oracle.batch(p, x, 1, lut)
    assert np.array_equal(np.concatenate(gathered), whole)
    print("SHARD_OK", world, whole.shape[0])
dist.destroy_process_group()
'''


def test_world_size_2_gloo_time_block_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "SHARD_OK 2 42" in out.stdout


def test_native_time_blocks_equal_sharding_py(sbx):
    """sbx_shard_time_block (C ABI, no GPU needed) == sharding.time_blocks, over a randomised sweep incl. recordings
    shorter than a window, one-frame recordings and more shards than frames."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=300, deadline=None)
    @given(st.integers(0, 10 ** 7), st.integers(1, 70000), st.integers(1, 70000), st.sampled_from([1, 2]), st.integers(1, 16))
    def check(n_samples, window_n, hop, channels, world):
        blocks = sbx.sharding.time_blocks(n_samples, window_n, hop, world, frame_align=2 if channels == 1 else 1)
        for b in blocks:
            assert sbx.sharding.time_block_native(n_samples, window_n, hop, channels, b.rank, world) == b
    check()
    with pytest.raises(ValueError):
        sbx.sharding.time_block_native(1000, 256, 64, 1, 3, 3)              # index out of range
    with pytest.raises(ValueError):
        sbx.sharding.time_block_native(1000, 256, 0, 1, 0, 2)               # hop 0
