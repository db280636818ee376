"""-m gpu: the BASELINE.json configurations at (or near) their full sizes, checked through size-independent
properties -- exact linearity, pair-slot/time-shift invariance, shard == unsharded, silence, tone placement --
plus the oracle on a sparse sample of frames (the oracle is far too slow to run the full sizes)."""
import numpy as np
import pytest

from oracle import oracle
from tests import signals

pytestmark = pytest.mark.gpu
MAG_TOL = 1e-5


@pytest.fixture(scope="module")
def sbx():
    import spectrogram_rs_b200 as s
    return s


def _lut_index(rgba_t, lut):
    """LUT index image of a uint8 (..., 4) CUDA tensor (ramps with unique entries)."""
    import torch
    key = (rgba_t[..., 0].int() << 16) | (rgba_t[..., 1].int() << 8) | rgba_t[..., 2].int()
    lk = (lut[:, 0].astype(np.int64) << 16) | (lut[:, 1].astype(np.int64) << 8) | lut[:, 2].astype(np.int64)
    order = np.argsort(lk)
    sorted_keys = torch.from_numpy(lk[order]).to(key.device).int()
    pos = torch.searchsorted(sorted_keys, key.contiguous())
    return torch.from_numpy(order).to(key.device)[pos.clamp(max=255)]


def test_cfg1_full_60s_sweep(sbx, magma):
    """cfg1: 60 s of a 48 kHz log sweep, N 4096 -> M 8192, hop 512, H 1024, Magma: 5618 frames."""
    import torch
    sr, n, hop, rows = 48000.0, 4096, 512, 1024
    x = signals.log_sweep(int(sr * 60))
    frames = oracle.num_frames(len(x), n, hop)
    assert frames == 5618
    with sbx.Plan(sample_rate=sr, window_n=n, hop=hop, rows_h=rows, channels=1, lut=magma) as plan:
        d_x = torch.from_numpy(x).cuda()
        rgba = torch.empty((frames, rows, 4), dtype=torch.uint8, device="cuda")
        mags = torch.empty((frames, plan.k, 2), dtype=torch.float32, device="cuda")
        assert plan.process_batch_device(d_x, len(x), rgba, frames) == frames
        assert plan.magnitudes_batch_device(d_x, len(x), mags, frames) == frames
        # 1. exact linearity under a power-of-two gain (every operation up to the magnitudes is exact in the scaling)
        mags2 = torch.empty_like(mags)
        plan.magnitudes_batch_device((d_x * 2.0).contiguous(), len(x), mags2, frames)
        plan.synchronize()
        assert torch.equal(mags2, mags * 2.0)
        assert torch.equal(mags[..., 0], mags[..., 1])                       # mono == (s, s): L == R
        # 2. time shift by one hop moves every frame to the other slot of its mono pair (real <-> imaginary part of
        #    the packed transform): same columns up to the last bit of the FFT -> at most one LUT step, rarely
        shifted = torch.empty((frames - 1, rows, 4), dtype=torch.uint8, device="cuda")
        plan.process_batch_device(d_x[hop:].contiguous(), len(x) - hop, shifted, frames - 1)
        plan.synchronize()
        di = (_lut_index(shifted, magma) - _lut_index(rgba[1:], magma)).abs()
        assert int(di.max()) <= 1 and float((di != 0).float().mean()) <= 1e-3
        # 3. eight time blocks with an N - hop halo == unsharded, byte for byte
        parts = []
        for b in sbx.sharding.time_blocks(len(x), n, hop, 8, frame_align=2):
            out = torch.empty((b.frames, rows, 4), dtype=torch.uint8, device="cuda")
            plan.process_batch_device(d_x[b.sample_lo:b.sample_hi].contiguous(), b.samples, out, b.frames)
            parts.append(out)
        plan.synchronize()
        assert torch.equal(torch.cat(parts), rgba)
        # 4. the sweep is where it should be: the brightest row climbs monotonically with time
        idx = _lut_index(rgba, magma)
        peak_row = idx[200:5400:100].float().argmax(dim=1)
        assert bool((peak_row[1:] <= peak_row[:-1]).all())                   # row 0 = top = highest frequency
        # 5. the oracle on a sparse sample of frames
        p = oracle.params(sample_rate=sr, window_n=n, hop=hop, rows_h=rows)
        g_m, g_c = mags.cpu().numpy(), rgba.cpu().numpy()
        for f in range(0, frames, 401):
            s = x[f * hop:f * hop + n]
            o = oracle.frame_magnitudes(np.stack([s, s], 1), n, 2)
            assert np.abs(g_m[f] - o).max() <= MAG_TOL * np.abs(o).max()
            assert np.array_equal(oracle.column_from_magnitudes(p, g_m[f], magma), g_c[f])
    # 6. silence -> the first LUT entry everywhere (t <= 0 clamps to index 0), alpha 255
    with sbx.Plan(sample_rate=sr, window_n=n, hop=hop, rows_h=rows, channels=1, lut=magma) as plan:
        z = plan.process_batch(np.zeros(n + hop * 20, np.float32))
        assert np.all(z[..., :3] == magma[0]) and np.all(z[..., 3] == 255)


def test_cfg3_geometry_persistent_kernel(sbx, viridis):
    """cfg3 geometry (96 kHz mono, N 8192 -> M 16384, hop 256, Viridis) on 60 s: 22 469 frames through the
    persistent TMA-staged kernel; shard == unsharded; oracle on a sparse sample."""
    import torch
    sr, n, hop, rows = 96000.0, 8192, 256, 1024
    x = signals.sweeps_plus_noise(int(sr * 60), sr)
    frames = oracle.num_frames(len(x), n, hop)
    with sbx.Plan(sample_rate=sr, window_n=n, hop=hop, rows_h=rows, channels=1, lut=viridis) as plan:
        d_x = torch.from_numpy(x).cuda()
        rgba = torch.empty((frames, rows, 4), dtype=torch.uint8, device="cuda")
        assert plan.process_batch_device(d_x, len(x), rgba, frames) == frames
        parts = []
        for b in sbx.sharding.time_blocks(len(x), n, hop, 4, frame_align=2):
            out = torch.empty((b.frames, rows, 4), dtype=torch.uint8, device="cuda")
            plan.process_batch_device(d_x[b.sample_lo:b.sample_hi].contiguous(), b.samples, out, b.frames)
            parts.append(out)
        plan.synchronize()
        assert torch.equal(torch.cat(parts), rgba)
        p = oracle.params(sample_rate=sr, window_n=n, hop=hop, rows_h=rows)
        g = rgba.cpu().numpy()
        differing = 0
        picks = list(range(0, frames, 2503))
        for f in picks:
            s = x[f * hop:f * hop + n]
            o = oracle.column_from_magnitudes(p, oracle.frame_magnitudes(np.stack([s, s], 1), n, 2), viridis)
            differing += int((o != g[f]).any(axis=1).sum())
        assert differing <= max(1, int(1e-3 * len(picks) * rows))


def test_cfg4_channel_array(sbx, magma):
    """cfg4: independent channels, channel c = tone at 100 * 2^(c/8) Hz; every channel's brightest row is the
    row of its tone; processing a channel alone or after the others gives the same bytes."""
    sr, n, hop, rows, seconds = 48000.0, 4096, 512, 1024, 4
    p = oracle.params(sample_rate=sr, window_n=n, hop=hop, rows_h=rows)
    f0, _, _, _ = oracle.bin_map(p)
    with sbx.Plan(sample_rate=sr, window_n=n, hop=hop, rows_h=rows, channels=1, lut=magma) as plan:
        first = None
        for c in range(0, 62, 6):                                           # tones up to 18.1 kHz (the top rows end at 22.03 kHz)
            x = signals.channel_tone(c, int(sr * seconds))
            col = plan.process_batch(x)
            if first is None:
                first = (x, col.copy())
            freq = 100.0 * 2.0 ** (c / 8.0)
            want_py = int(np.searchsorted(f0, freq, side="right") - 1)
            bright = col[:, :, :3].astype(int).sum(axis=2)                 # frames x rows
            got_py = rows - 1 - int(np.median(bright.argmax(axis=1)))
            assert abs(got_py - want_py) <= 2, (c, freq, want_py, got_py)
        assert np.array_equal(plan.process_batch(first[0]), first[1])


@pytest.mark.parametrize("n", [512, 1024, 2048, 4096, 8192, 16384, 32768, 65536])
def test_cfg5_size_sweep(sbx, magma, n):
    """cfg5: N = 512 .. 65536 at P = 4, hop = N/8, H = 2048 on white noise; three frames against the oracle."""
    hop, frames = n // 8, 24
    x = signals.white_noise(n + hop * (frames - 1))
    p = oracle.params(window_n=n, pad_factor=4, hop=hop, rows_h=2048)
    with sbx.Plan(window_n=n, pad_factor=4, hop=hop, rows_h=2048, channels=1, lut=magma) as plan:
        import torch
        d_x = torch.from_numpy(x).cuda()
        mags = torch.empty((frames, plan.k, 2), dtype=torch.float32, device="cuda")
        rgba = torch.empty((frames, 2048, 4), dtype=torch.uint8, device="cuda")
        plan.magnitudes_batch_device(d_x, len(x), mags, frames)
        plan.process_batch_device(d_x, len(x), rgba, frames)
        plan.synchronize()
        g_m, g_c = mags.cpu().numpy(), rgba.cpu().numpy()
        for f in (0, 11, 23):
            s = x[f * hop:f * hop + n]
            o = oracle.frame_magnitudes(np.stack([s, s], 1), n, 4)
            assert np.abs(g_m[f] - o).max() <= MAG_TOL * np.abs(o).max()
            assert np.array_equal(oracle.column_from_magnitudes(p, g_m[f], magma), g_c[f])


def test_cfg2_live_stream_equals_batch(sbx, magma):
    """cfg2: 48 kHz stereo, 2048-pt frames pushed one at a time through the pinned ring: the live columns are
    the batch columns, byte for byte, and the hop cursor never drifts."""
    n = hop = 2048
    frames = 300
    lr = signals.stereo_loopback(n * frames)
    with sbx.Plan(sample_rate=48000.0, window_n=n, hop=hop, rows_h=1024, channels=2, lut=magma) as plan:
        batch = plan.process_batch(lr)
        live = plan.live(4096)
        for f in range(frames):
            assert live.push(lr[f * n:(f + 1) * n]) == n
            col = live.pop_column()
            assert col is not None and np.array_equal(col, batch[f]), f
            assert live.buffered() == 0
        live.close()
