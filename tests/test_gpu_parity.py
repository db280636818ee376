"""-m gpu: the CUDA path, called through the C ABI (libsbx_b200.so), against the CPU oracle.

Tolerances (BASELINE.json north_star):
  * magnitudes: max |gpu - oracle| <= 1e-5 * frame peak   (f32 FFT vs the oracle's f32 FFT; the
    literal per-bin relative error is meaningless on leakage-floor bins)
  * bin map / thresholds / window: bit-exact
  * RGBA from identical magnitudes: bit-exact
  * RGBA end to end: every pixel within one LUT step, and at most 0.1 % of pixels differing at all
"""
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


def _pcm(channels, n_samples, seed=0):
    rng = np.random.Generator(np.random.PCG64(seed))
    if channels == 1:
        return (signals.log_sweep(n_samples) + 0.02 * rng.standard_normal(n_samples)).astype(np.float32)
    return signals.stereo_loopback(n_samples, seed=seed + 1)


def _lut_index_image(rgba, lut):
    """Map RGBA pixels back to LUT indices (ramps used here have unique entries)."""
    key = {tuple(int(v) for v in c): i for i, c in enumerate(lut)}
    flat = rgba.reshape(-1, 4)
    return np.array([key[tuple(int(v) for v in px[:3])] for px in flat]).reshape(rgba.shape[:-1])


CASES = [
    # (N, P, channels, hop, H, interp)
    (256, 2, 2, 64, 64, 0),
    (512, 2, 1, 128, 128, 0),
    (1024, 4, 2, 256, 256, 1),
    (2048, 2, 2, 2048, 1024, 0),       # cfg2 geometry
    (4096, 2, 1, 512, 1024, 0),        # cfg1 / cfg4 geometry
    (4096, 2, 2, 512, 1024, 1),
    (8192, 2, 1, 256, 1024, 0),        # cfg3 geometry
    (4096, 4, 1, 512, 2048, 0),        # cfg5 geometry (N = 4096)
    (2048, 1, 2, 512, 512, 0),
    (16384, 1, 2, 2048, 512, 0),       # the longest single-SM transform without padding
    (256, 1, 1, 100, 77, 0),           # ragged hop / odd H
    # windows that are not a power of two: the live app's (0.05 s * sr) as usize, run-time mixed radix
    (2400, 2, 2, 58, 1024, 0),         # 48 kHz, hop = (2.5/2048 * 48000) as usize  (gpu_spectrogram.rs:21-23,323)
    (2205, 2, 2, 86, 1024, 0),         # 44.1 kHz
    (2400, 2, 1, 93, 512, 1),
    (1600, 4, 1, 200, 256, 0),
    (100, 2, 2, 10, 64, 0),            # the dummy transform the widgets start with (gpu_spectrogram.rs:80)
    (4800, 2, 2, 512, 1024, 0),        # 96 kHz
    # windows with a prime factor > 7: Bluestein over a power-of-two convolution (22.05 kHz: N = 1102 = 2 * 19 * 29)
    (1102, 2, 2, 43, 1024, 0),
    (551, 2, 1, 21, 512, 1),           # 11.025 kHz, mono pairs
    (1103, 2, 2, 100, 256, 0),         # a prime window
    (347, 4, 1, 50, 128, 0),
    (5461, 2, 2, 1000, 256, 0),        # the longest: N + M - 1 = 16382 <= 16384
    # M > 16384: the multi-kernel path (cfg5's larger sizes)
    (8192, 4, 1, 1024, 2048, 0),
    (16384, 2, 2, 2048, 512, 1),
    (32768, 1, 1, 4096, 256, 0),
    (16384, 4, 1, 2048, 2048, 0),
    (32768, 4, 2, 4096, 1024, 0),
    (65536, 4, 1, 8192, 2048, 0),      # cfg5's largest: M = 262144
    (8192, 4, 1, 4096, 16, 0),         # 16 rows over 16383 bins: rows longer than the resample scratch area
]


@pytest.mark.parametrize("n,pad,ch,hop,rows,interp", CASES)
def test_batch_matches_oracle(sbx, magma, n, pad, ch, hop, rows, interp):
    import torch
    frames = 7                                        # odd: exercises the half-filled mono pair
    n_samples = n + hop * (frames - 1) + 5
    pcm = _pcm(ch, n_samples)
    p = oracle.params(window_n=n, pad_factor=pad, hop=hop, rows_h=rows, interp=interp)
    o_rgba, o_mags = oracle.batch(p, pcm, ch, magma, want_mags=True)

    with sbx.Plan(window_n=n, pad_factor=pad, hop=hop, rows_h=rows, channels=ch, interp=interp, lut=magma) as plan:
        assert plan.num_frames(n_samples) == frames == oracle.num_frames(n_samples, n, hop)
        d_pcm = torch.from_numpy(pcm).cuda()
        d_mags = torch.zeros((frames, plan.k, 2), dtype=torch.float32, device="cuda")
        d_rgba = torch.zeros((frames, rows, 4), dtype=torch.uint8, device="cuda")
        assert plan.magnitudes_batch_device(d_pcm, n_samples, d_mags, frames) == frames
        assert plan.process_batch_device(d_pcm, n_samples, d_rgba, frames) == frames
        plan.synchronize()
        g_mags, g_rgba = d_mags.cpu().numpy(), d_rgba.cpu().numpy()
        # host-pointer batch entry point gives the same bytes
        h_rgba = plan.process_batch(pcm)
        assert np.array_equal(h_rgba, g_rgba)
        # magnitudes
        peak = np.abs(o_mags).max(axis=(1, 2), keepdims=True)
        assert (np.abs(g_mags - o_mags) / peak).max() <= MAG_TOL
        # identical normalised inputs -> bit-exact RGBA
        assert np.array_equal(plan.colorize_magnitudes(g_mags), np.stack(
            [oracle.column_from_magnitudes(p, g_mags[f], magma) for f in range(frames)]))
        assert np.array_equal(plan.colorize_magnitudes(o_mags), o_rgba)
        # fused kernel == its own magnitudes run through the colour stage
        assert np.array_equal(plan.colorize_magnitudes(g_mags), g_rgba)
        # end to end
        assert np.array_equal(g_rgba[..., 3], o_rgba[..., 3])
        di = np.abs(_lut_index_image(g_rgba, magma) - _lut_index_image(o_rgba, magma))
        assert di.max() <= 1
        assert (di != 0).mean() <= 1e-3


@pytest.mark.parametrize("n,pad", [(256, 2), (2048, 2), (4096, 2), (1024, 4)])
def test_single_frame_entry_points(sbx, magma, n, pad):
    lr = signals.stereo_loopback(n + 10)
    p = oracle.params(window_n=n, pad_factor=pad, rows_h=256)
    with sbx.Plan(window_n=n, pad_factor=pad, rows_h=256, lut=magma) as plan:
        assert plan.process_magnitudes(lr[: n - 1]) is None          # the reference's None
        assert plan.process_column(lr[: n - 1]) is None
        g = plan.process_magnitudes(lr)
        o = oracle.frame_magnitudes(lr, n, pad)
        assert g.shape == o.shape == (n * pad // 2 - 1, 2)
        assert np.abs(g - o).max() <= MAG_TOL * np.abs(o).max()
        col = plan.process_column(lr)
        assert np.array_equal(col, plan.colorize_magnitudes(g)[0])
        ocol = oracle.column_from_magnitudes(p, g, magma)
        assert np.array_equal(col, ocol)


def test_tables_bit_exact(sbx, magma):
    for kw in (dict(window_n=4096, pad_factor=2, rows_h=1024, sample_rate=48000.0),
               dict(window_n=8192, pad_factor=2, rows_h=1024, sample_rate=96000.0),
               dict(window_n=1024, pad_factor=4, rows_h=2048, sample_rate=48000.0),
               dict(window_n=2048, pad_factor=2, rows_h=333, sample_rate=44100.0, f_min=20.0, f_max=20000.0)):
        with sbx.Plan(lut=magma, **kw) as plan:
            p = oracle.params(**kw)
            f0, step, n, period = plan.bin_map()
            of0, ostep, on, operiod = oracle.bin_map(p)
            assert np.array_equal(f0.view(np.uint32), of0.view(np.uint32))
            assert np.array_equal(step.view(np.uint32), ostep.view(np.uint32))
            assert np.array_equal(n, on)
            assert np.float32(period).view(np.uint32) == np.float32(operiod).view(np.uint32)
            assert np.array_equal(plan.window().view(np.uint32), oracle.hann(kw["window_n"]).view(np.uint32))


def test_known_answer_sine(sbx, magma):
    """K1/K2: unit sine at exact padded bin k0 -> L[k0-1] = 0.5; L-only input -> R ~ 0."""
    n, pad, k0 = 4096, 2, 400
    t = np.arange(n, dtype=np.float64)
    s = np.sin(2 * np.pi * k0 * t / (n * pad)).astype(np.float32)
    with sbx.Plan(window_n=n, pad_factor=pad, lut=magma) as plan:
        m = plan.process_magnitudes(np.stack([s, s], 1))
        assert abs(m[k0 - 1, 0] - 0.5) < 2e-6 and abs(m[k0 - 1, 1] - 0.5) < 2e-6
        assert abs(m[k0 + 1, 0] - 0.25) < 2e-6                      # +-2 padded bins = +-1 Hann bin
        assert abs(m[k0, 0] - 0.42441318) < 2e-6
        assert m[k0 + 7, 0] < 1e-6
        m = plan.process_magnitudes(np.stack([s, np.zeros_like(s)], 1))
        assert abs(m[k0 - 1, 0] - 0.5) < 2e-6 and m[:, 1].max() < 1e-6


def test_mono_equals_duplicated_stereo(sbx, magma):
    """channels=1 is defined as the (s,s) duplication of audio_input_list_model.rs:66-69."""
    import torch
    n, hop, frames = 1024, 256, 6
    ns = n + hop * (frames - 1)
    x = signals.white_noise(ns)
    with sbx.Plan(window_n=n, hop=hop, rows_h=256, channels=1, lut=magma) as pm, \
            sbx.Plan(window_n=n, hop=hop, rows_h=256, channels=2, lut=magma) as ps:
        a = torch.zeros((frames, 256, 4), dtype=torch.uint8, device="cuda")
        b = torch.zeros_like(a)
        pm.process_batch_device(torch.from_numpy(x).cuda(), ns, a, frames)
        ps.process_batch_device(torch.from_numpy(np.stack([x, x], 1)).cuda(), ns, b, frames)
        pm.synchronize(); ps.synchronize()
        ia, ib = _lut_index_image(a.cpu().numpy(), magma), _lut_index_image(b.cpu().numpy(), magma)
        assert np.abs(ia - ib).max() <= 1 and (ia != ib).mean() <= 1e-3


def test_time_block_sharding_is_exact(sbx, magma):
    """frames [lo, hi) computed from the sample block [lo*hop, (hi-1)*hop + N) equal the
    same frames of the unsharded run, byte for byte."""
    n, hop = 2048, 256
    ns = n + hop * 40
    x = signals.log_sweep(ns)
    with sbx.Plan(window_n=n, hop=hop, rows_h=512, channels=1, lut=magma) as plan:
        whole = plan.process_batch(x)
        total = whole.shape[0]
        parts = []
        for b in sbx.sharding.time_blocks(ns, n, hop, 3, frame_align=2):      # mono: pair-aligned boundaries
            parts.append(plan.process_batch(x[b.sample_lo:b.sample_hi]))
        assert np.array_equal(np.concatenate(parts), whole)


def test_live_ring_semantics(sbx, magma):
    """sbx_live_*: non-consuming window, skip(hop) even on the failing attempt, push drops overflow."""
    n, hop = 2048, 2048
    lr = signals.stereo_loopback(n * 3)
    with sbx.Plan(window_n=n, hop=hop, rows_h=1024, lut=magma) as plan:
        live = plan.live(4096)
        assert live.pop_column() is None
        assert live.push(lr[:4000]) == 4000
        assert live.push(lr[4000:4200]) == 96                       # capacity 4096
        c0 = live.pop_column()
        assert np.array_equal(c0, plan.process_column(lr[:n]))
        assert live.buffered() == 4096 - hop
        c1 = live.pop_column()
        assert np.array_equal(c1, plan.process_column(lr[hop:hop + n]))
        assert live.buffered() == 0
        live.push(lr[:100])
        assert live.pop_column() is None and live.buffered() == 0   # skip also on failure
        m = None
        live.push(lr[:n])
        m = live.pop_magnitudes()
        assert np.array_equal(m, plan.process_magnitudes(lr[:n]))
        live.close()


def test_errors_do_not_abort(sbx, magma):
    with pytest.raises(sbx.SbxError) as e:
        sbx.Plan(window_n=5462, lut=magma)                          # 2 * 2731: prime factor > 7 and too long for Bluestein
    assert e.value.status == -2
    with pytest.raises(sbx.SbxError):
        sbx.Plan(window_n=131072, pad_factor=2, lut=magma)
    with pytest.raises(sbx.SbxError):
        sbx.Plan(window_n=4096, pad_factor=8, lut=magma)
    with pytest.raises(sbx.SbxError):
        sbx.Plan(window_n=1024, hop=0, lut=magma)
    with sbx.Plan(window_n=1024, lut=magma) as plan:
        import torch
        small = torch.zeros((1, 1024, 4), dtype=torch.uint8, device="cuda")
        pcm = torch.zeros(1024 * 2 * 4, device="cuda")
        with pytest.raises(sbx.SbxError) as e:
            plan.process_batch_device(pcm, 1024 * 4, small, 1)
        assert e.value.status == -5


def test_cpp_host_mirror_runs(tmp_path):
    """The C++ mirror of AudioStreamTransform<T>: 4096 buffered samples, N 2048, hop 512 -> 5 frames; 1 kHz tone."""
    import re
    from tests.test_cabi import _build_cpp_example
    import subprocess
    out = subprocess.run([_build_cpp_example(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    m = re.search(r"frames=(\d+) k=(\d+) peak_bin=(\d+) peak=([\d.]+) column=1 alpha=255", out.stdout)
    assert m, out.stdout
    assert int(m.group(1)) == 5 and int(m.group(2)) == 2047
    assert abs(int(m.group(3)) + 1 - 1000.0 * 4096 / 48000.0) <= 1 and 0.2 < float(m.group(4)) <= 0.25   # A/2 for a sine of amplitude A = 0.5


@pytest.mark.parametrize("n,pad,ch,hop,rows", [(1024, 2, 2, 256, 256), (4096, 2, 1, 512, 1024), (16384, 2, 2, 2048, 512)])
def test_stereo_pan_scheme(sbx, viridis, n, pad, ch, hop, rows):
    """SBX_SCHEME_STEREO (colorscheme.rs:63-66): bit-exact with the oracle given identical magnitudes."""
    ns = n + hop * 4
    pcm = signals.stereo_loopback(ns) if ch == 2 else signals.white_noise(ns)
    if ch == 2:
        pcm[:, 1] *= np.linspace(0, 2, ns).astype(np.float32)
    p = oracle.params(window_n=n, pad_factor=pad, hop=hop, rows_h=rows, stereo_scheme=1)
    with sbx.Plan(window_n=n, pad_factor=pad, hop=hop, rows_h=rows, channels=ch, lut=viridis, scheme=1) as plan:
        g = plan.process_batch(pcm)
        o_rgba, o_mags = oracle.batch(p, pcm, ch, viridis, want_mags=True)
        assert np.array_equal(plan.colorize_magnitudes(o_mags), o_rgba)
        assert (g != o_rgba).any(axis=2).mean() <= 2e-3 and np.abs(g.astype(int) - o_rgba.astype(int))[..., 3].max() <= 1
        assert g[..., 3].min() < 64 and g[..., 3].max() > 128


@pytest.mark.parametrize("n,pad,ch,hop,frames", [(8192, 2, 1, 256, 701), (4096, 4, 2, 512, 330)])
def test_persistent_tma_staged_kernel(sbx, viridis, monkeypatch, n, pad, ch, hop, frames):
    """Enough units to take the persistent kernel (grid = SMs x resident CTAs, PCM staged into shared memory by
    cp.async.bulk): one CTA per SM at M = 16384.  Odd frame count: the trailing half-filled mono pair copies a
    shorter span.  Compared with the oracle and with the one-CTA-per-unit kernel (a short batch of the same data).
    SBX_REAL=0 keeps the mono case on the frame-pair kernel (the default there is the real-sequence geometry,
    test_real_sequence_geometry*)."""
    monkeypatch.setenv("SBX_REAL", "0")
    ns = n + hop * (frames - 1)
    pcm = signals.sweeps_plus_noise(ns, 96000.0) if ch == 1 else signals.stereo_loopback(ns)
    p = oracle.params(sample_rate=96000.0, window_n=n, pad_factor=pad, hop=hop, rows_h=512)
    o_rgba, o_mags = oracle.batch(p, pcm, ch, viridis, want_mags=True, threads=8)
    with sbx.Plan(sample_rate=96000.0, window_n=n, pad_factor=pad, hop=hop, rows_h=512, channels=ch, lut=viridis) as plan:
        import torch
        d_pcm = torch.from_numpy(pcm).cuda()
        d_mags = torch.zeros((frames, plan.k, 2), dtype=torch.float32, device="cuda")
        launches = plan.launch_count
        assert plan.magnitudes_batch_device(d_pcm, ns, d_mags, frames) == frames
        g_rgba = plan.process_batch(pcm)
        plan.synchronize()
        g_mags = d_mags.cpu().numpy()
        assert plan.launch_count > launches
        peak = np.abs(o_mags).max(axis=(1, 2), keepdims=True)
        assert (np.abs(g_mags - o_mags) / peak).max() <= MAG_TOL
        assert (g_rgba != o_rgba).any(axis=2).mean() <= 1e-3
        assert np.array_equal(plan.colorize_magnitudes(g_mags[:8]), g_rgba[:8])
        # the same leading frames through a short batch (one CTA per unit, global loads)
        short = plan.process_batch(pcm[: n + hop * 9])
        assert np.array_equal(short, g_rgba[:10])


@pytest.mark.parametrize("n,pad,hop,rows,interp,scheme", [
    (4096, 2, 512, 1024, 0, 0), (2048, 4, 256, 512, 0, 0), (8192, 1, 1024, 256, 1, 0), (8192, 2, 256, 1024, 0, 0),
    (4096, 4, 512, 2048, 0, 1), (16384, 1, 2048, 512, 0, 0), (8192, 4, 1024, 2048, 0, 0), (16384, 2, 2048, 1024, 1, 0),
    (32768, 1, 4096, 512, 0, 1)])
def test_real_sequence_geometry(sbx, magma, monkeypatch, n, pad, hop, rows, interp, scheme):
    """Mono batches: the frame-pair kernel (SBX_REAL=0; the four-kernel path beyond M = 16384) and the one-frame
    real-sequence geometry (SBX_REAL=1) against the oracle and against each other, on an odd frame count."""
    import torch
    frames = 9
    ns = n + hop * (frames - 1) + 3
    pcm = _pcm(1, ns, seed=n + pad)
    p = oracle.params(window_n=n, pad_factor=pad, hop=hop, rows_h=rows, interp=interp, stereo_scheme=scheme)
    o_rgba, o_mags = oracle.batch(p, pcm, 1, magma, want_mags=True, threads=4)
    peak = np.abs(o_mags).max(axis=(1, 2), keepdims=True)
    got = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("SBX_REAL", mode)
        with sbx.Plan(window_n=n, pad_factor=pad, hop=hop, rows_h=rows, channels=1, interp=interp, scheme=scheme, lut=magma) as plan:
            assert plan.kernel_path.startswith("fused-real") == (mode == "1")
            d_pcm = torch.from_numpy(pcm).cuda()
            d_mags = torch.zeros((frames, plan.k, 2), dtype=torch.float32, device="cuda")
            assert plan.magnitudes_batch_device(d_pcm, ns, d_mags, frames) == frames
            g_rgba = plan.process_batch(pcm)
            plan.synchronize()
            g_mags = d_mags.cpu().numpy()
            assert (np.abs(g_mags - o_mags) / peak).max() <= MAG_TOL
            assert np.array_equal(g_mags[..., 0], g_mags[..., 1])
            rec = np.stack([oracle.column_from_magnitudes(p, g_mags[f], magma) for f in range(frames)])
            assert np.array_equal(g_rgba, rec)
            assert (g_rgba != o_rgba).any(axis=2).mean() <= 2e-3
            got[mode] = g_rgba
    assert (got["0"] != got["1"]).any(axis=2).mean() <= 2e-3
    monkeypatch.delenv("SBX_REAL")
    with sbx.Plan(window_n=n, pad_factor=pad, hop=hop + 1, rows_h=rows, channels=1, lut=magma) as plan:
        assert not plan.kernel_path.startswith("fused-real")         # odd hop: sample pairs would be misaligned


def test_real_sequence_geometry_persistent(sbx, viridis):
    """M = 32768 mono: one 1024-thread CTA per SM, persistent grid with cp.async.bulk staging; enough frames to take it."""
    import torch
    n, pad, hop, rows, frames = 8192, 4, 1024, 256, 2 * 148 + 37
    ns = n + hop * (frames - 1)
    pcm = signals.sweeps_plus_noise(ns, 48000.0)
    p = oracle.params(window_n=n, pad_factor=pad, hop=hop, rows_h=rows)
    o_rgba, o_mags = oracle.batch(p, pcm, 1, viridis, want_mags=True, threads=8)
    with sbx.Plan(window_n=n, pad_factor=pad, hop=hop, rows_h=rows, channels=1, lut=viridis) as plan:
        assert plan.kernel_path == "fused-real"
        d_mags = torch.zeros((frames, plan.k, 2), dtype=torch.float32, device="cuda")
        assert plan.magnitudes_batch_device(torch.from_numpy(pcm).cuda(), ns, d_mags, frames) == frames
        g_rgba = plan.process_batch(pcm)
        plan.synchronize()
        g_mags = d_mags.cpu().numpy()
        peak = np.abs(o_mags).max(axis=(1, 2), keepdims=True)
        assert (np.abs(g_mags - o_mags) / peak).max() <= MAG_TOL
        assert (g_rgba != o_rgba).any(axis=2).mean() <= 1e-3
        short = plan.process_batch(pcm[: n + hop * 9])               # one CTA per frame, global loads
        assert np.array_equal(short, g_rgba[:10])


@pytest.mark.parametrize("n,pad,ch,hop,rows", [
    (4096, 2, 1, 512, 1024), (2048, 2, 2, 2048, 1024), (8192, 2, 1, 256, 1024), (8192, 2, 1, 255, 333), (8192, 4, 1, 1024, 2048),
    (16384, 1, 2, 2048, 100), (16384, 4, 1, 2048, 512), (2400, 2, 2, 58, 1024), (512, 4, 1, 64, 2048)])
def test_outputs_stay_inside_their_buffers(sbx, magma, n, pad, ch, hop, rows):
    """Every kernel family writes exactly frames x rows x 4 bytes (and frames x K x 2 floats): canary frames before
    and after the output region of an odd-length batch must survive, and no output byte may keep its canary."""
    import torch
    frames, guard = 5, 3
    ns = n + hop * (frames - 1)
    pcm = _pcm(ch, ns, seed=3)
    with sbx.Plan(window_n=n, pad_factor=pad, hop=hop, rows_h=rows, channels=ch, lut=magma) as plan:
        d_pcm = torch.from_numpy(pcm).cuda()
        rgba = torch.full((frames + 2 * guard, rows, 4), 0x5A, dtype=torch.uint8, device="cuda")
        mags = torch.full((frames + 2 * guard, plan.k, 2), -123.0, dtype=torch.float32, device="cuda")
        f16 = torch.full((frames + 2 * guard, plan.k, 2), 0x7E7E, dtype=torch.int16, device="cuda")
        assert plan.process_batch_device(d_pcm, ns, rgba[guard:], frames) == frames
        assert plan.magnitudes_batch_device(d_pcm, ns, mags[guard:], frames) == frames
        assert plan.magnitudes_batch_device_f16(d_pcm, ns, f16[guard:], frames) == frames
        plan.synchronize()
        r, m, h = rgba.cpu().numpy(), mags.cpu().numpy(), f16.cpu().numpy()
        for arr, canary in ((r, 0x5A), (m, -123.0), (h, 0x7E7E)):
            assert (arr[:guard] == canary).all() and (arr[guard + frames:] == canary).all()
        assert (r[guard:guard + frames, :, 3] == 255).all()            # every pixel written (alpha 255 != 0x5A)
        assert (m[guard:guard + frames] >= 0).all()                    # every magnitude written
        assert (h[guard:guard + frames] != 0x7E7E).any(axis=(1, 2)).all()


@pytest.mark.parametrize("n,pad,ch,hop,rows", [
    (4096, 2, 1, 512, 1024),          # cfg4: mono frame pairs; odd frame count -> a half-filled pair ends every channel
    (2048, 2, 2, 2048, 256), (8192, 2, 1, 256, 512), (8192, 2, 1, 255, 128), (8192, 4, 1, 1024, 64),
    (2400, 2, 2, 58, 128), (16384, 4, 2, 2048, 64)])
def test_planar_sequences_equal_per_sequence_batches(sbx, magma, n, pad, ch, hop, rows):
    """sbx_process_batch_device_planar: several equally long recordings in one call == one call per recording
    (every kernel family; a padded stride between the sequences)."""
    import torch
    n_seq, frames, gap = 5, 7, 6
    ns = n + hop * (frames - 1) + 1
    stride = (ns + gap + 3) // 4 * 4                   # the per-sequence reference calls need 16-byte aligned starts
    rng = np.random.Generator(np.random.PCG64(11))
    shape = (n_seq, stride) if ch == 1 else (n_seq, stride, 2)
    pcm = (0.2 * rng.standard_normal(shape)).astype(np.float32)
    with sbx.Plan(window_n=n, pad_factor=pad, hop=hop, rows_h=rows, channels=ch, lut=magma) as plan:
        d = torch.from_numpy(pcm).cuda()
        out = torch.full((n_seq + 1, frames, rows, 4), 0x5A, dtype=torch.uint8, device="cuda")
        assert plan.process_batch_device_planar(d, n_seq, stride, ns, out, n_seq * frames) == frames
        ref = torch.zeros((n_seq, frames, rows, 4), dtype=torch.uint8, device="cuda")
        for s in range(n_seq):
            assert plan.process_batch_device(d[s], ns, ref[s], frames) == frames
        plan.synchronize()
        o = out.cpu().numpy()
        assert np.array_equal(o[:n_seq], ref.cpu().numpy())
        assert (o[n_seq] == 0x5A).all()
        with pytest.raises(sbx.SbxError):
            plan.process_batch_device_planar(d, n_seq, ns - 1, ns, out, n_seq * frames)       # overlapping sequences
        with pytest.raises(sbx.SbxError):
            plan.process_batch_device_planar(d, n_seq, stride, ns, out, n_seq * frames - 1)   # capacity


def test_column_store_blit(sbx, magma):
    """a-9: the scrolling column store of SimpleSpectrogram (simple_spectrogram.rs:141-164) on the device: columns
    land in a row-major image at (cursor + f) % width; wrap-around, padded row stride, more frames than columns."""
    import torch
    n, hop, rows, width, pad_px = 1024, 256, 77, 50, 3
    with sbx.Plan(window_n=n, pad_factor=2, hop=hop, rows_h=rows, channels=1, lut=magma) as plan:
        stride = (width + pad_px) * 4
        img = torch.full((rows, width + pad_px, 4), 7, dtype=torch.uint8, device="cuda")
        ref = np.full((rows, width + pad_px, 4), 7, np.uint8)
        cursor, ref_cursor = 0, 0
        rng = np.random.Generator(np.random.PCG64(9))
        for frames in (1, 33, 50, 0, 64, 131, 2):
            ns = n + hop * (frames - 1) if frames else n - 1
            pcm = (0.3 * rng.standard_normal(ns)).astype(np.float32)
            cols = torch.zeros((max(frames, 1), rows, 4), dtype=torch.uint8, device="cuda")
            assert plan.process_batch_device(torch.from_numpy(pcm).cuda(), ns, cols, max(frames, 1)) == frames
            cursor = plan.blit_columns_device(cols, frames, img, width, stride, cursor)
            plan.synchronize()
            h = cols.cpu().numpy()
            for f in range(frames):                                    # the reference's loop, one column at a time
                ref[:, ref_cursor, :] = h[f]
                ref_cursor = (ref_cursor + 1) % width
            assert cursor == ref_cursor
            assert np.array_equal(img.cpu().numpy(), ref)
        assert (ref[:, width:, :] == 7).all()                          # the row padding is never written
        with pytest.raises(sbx.SbxError):
            plan.blit_columns_device(cols, 1, img, width, width * 4 - 4, 0)
        with pytest.raises(sbx.SbxError):
            plan.blit_columns_device(cols, 1, img, width, stride, width)


def test_gl_widget_feed_formats(sbx, magma):
    """next-4: F16F16 rows (bit-exact round-to-nearest-even of the f32 rows) and the 32x32 palette table of
    ColorScheme::lookup_table (colorscheme.rs:73-92)."""
    import torch
    n, hop, frames = 2048, 512, 9
    ns = n + hop * (frames - 1)
    pcm = signals.stereo_loopback(ns)
    with sbx.Plan(window_n=n, hop=hop, rows_h=256, channels=2, lut=magma) as plan:
        d_pcm = torch.from_numpy(pcm).cuda()
        d32 = torch.zeros((frames, plan.k, 2), dtype=torch.float32, device="cuda")
        d16 = torch.zeros((frames, plan.k, 2), dtype=torch.float16, device="cuda")
        assert plan.magnitudes_batch_device(d_pcm, ns, d32, frames) == frames
        assert plan.magnitudes_batch_device_f16(d_pcm, ns, d16, frames) == frames
        plan.synchronize()
        assert np.array_equal(d16.cpu().numpy().view(np.uint16), d32.cpu().numpy().astype(np.float16).view(np.uint16))
        tab = plan.palette_table(32)
        assert tab.shape == (32, 32, 4) and np.all(tab[..., 3] == 1.0)
        for i in (0, 7, 31):
            t = np.float32(i) / np.float32(31)
            idx = min(255, int(np.floor(float(t) * 256.0)))
            assert np.allclose(tab[i, :, :3], magma[idx] / 256.0)
    with sbx.Plan(window_n=n, hop=hop, rows_h=256, channels=2, lut=magma, scheme=1) as plan:
        tab = plan.palette_table(32)
        assert np.allclose(tab[:, 5, 3], np.arange(32, dtype=np.float32) / 31)            # alpha = magnitude
        pan = np.float32(1.0) - np.float32(5) / np.float32(31)
        assert np.allclose(tab[3, 5, :3], magma[min(255, int(np.floor(float(pan) * 256.0)))] / 256.0)


def test_reference_interface_mirror(sbx, magma):
    """The reference's own call pattern through the Python mirror: FastFourierTransform::new(sr, 0.05) (N = 2400,
    the mixed-radix path), AudioStreamTransform over a HeapRb ring with stride 2.5 s / 2048, and the column
    loop of SimpleSpectrogram::snapshot with its scrolling offset."""
    sr = 48000.0
    fft = sbx.FastFourierTransform(sr, 0.05)                        # gpu_spectrogram.rs:321-324
    assert fft.num_input_samples() == 2400 and fft.num_output_frequencies() == 2399 and fft.sample_rate() == sr
    lr = signals.stereo_loopback(4000, sr)
    assert fft.process(lr[:2399]) is None                           # fft.rs:72
    m = fft.process(lr)
    o = oracle.frame_magnitudes(lr, 2400, 2)
    assert m.shape == (2399, 2) and np.abs(m - o).max() <= MAG_TOL * np.abs(o).max()
    m_iter = fft.process((float(a), float(b)) for a, b in lr[:2400])          # any iterable of (l, r)
    assert np.array_equal(m_iter, m)

    ring = sbx.HeapRb(4096)                                         # audio_input_list_model.rs:30
    assert ring.push_iter(map(tuple, lr)) == 4000
    stream = sbx.AudioStreamTransform(ring, fft, 2.5 / 2048.0)      # gpu_spectrogram.rs:21-23 -> hop 58
    outs = list(stream.process())
    assert len(outs) == oracle.num_frames(4000, 2400, 58) == 28
    assert np.array_equal(outs[3], fft.process(lr[3 * 58:3 * 58 + 2400]))
    assert len(ring) == 4000 - 29 * 58                              # 28 frames + the failing attempt

    ring2 = sbx.HeapRb(4096)
    ring2.push_mono(signals.white_noise(3000))                      # mono device: (s, s) duplication
    cols = sbx.SpectrogramColumns(ring2, sample_rate=sr, width=16, height=128)   # period 0.05 s, stride 2/1024 s -> hop 93
    assert cols.plan.n == 2400 and cols.plan.hop == 93
    drawn = cols.snapshot()
    assert drawn == oracle.num_frames(3000, 2400, 93) == 7 and cols.offset == 7
    assert cols.buffer[:, :7, 3].min() == 255 and cols.buffer[:, 7:, 3].max() == 0
    first = cols.plan.process_column(np.stack([signals.white_noise(3000)[:2400]] * 2, 1))
    assert np.array_equal(cols.buffer[:, 0, :], first)
    cols.set_sample_rate(44100)                                     # re-plan: N = (0.05 * 44100) as usize = 2205
    assert cols.plan.n == 2205


def test_live_mono_push_and_wraparound(sbx, magma):
    """Mono pushes are duplicated to (s, s); the ring wraps; windows straddling the wrap are linearised."""
    n, hop = 1024, 256
    x = signals.white_noise(n + hop * 40)
    with sbx.Plan(window_n=n, hop=hop, rows_h=256, channels=2, lut=magma) as plan:
        live = plan.live(1500)                                      # not a multiple of anything
        want = plan.process_batch(np.stack([x, x], 1))
        # model of the reference cursor (audio_transform.rs:37-41): transform the first N buffered samples without
        # consuming them, then skip(hop) -- also after the failing attempt, which discards up to hop samples
        model, n_cols = [], 0
        for pos in range(0, len(x), 700):                           # a callback delivering 700 samples at a time
            chunk = x[pos:pos + 700]
            room = 1500 - len(model)
            assert live.push(chunk, channels=1) == min(len(chunk), room)
            model.extend(chunk[:room].tolist())
            for _ in range(64):                                     # bounded drain; a failing pop ends the tick
                col = live.pop_column()
                window = np.asarray(model[:n], np.float32)
                del model[:min(hop, len(model))]
                if col is None:
                    assert len(window) < n
                    break
                assert len(window) == n
                assert np.array_equal(col, plan.process_column(np.stack([window, window], 1)))
                n_cols += 1
            assert live.buffered() == len(model)
        assert n_cols >= 10, n_cols
        live.close()
    # a stream that always has a full window never skips and reproduces the batch exactly
    with sbx.Plan(window_n=n, hop=hop, rows_h=256, channels=2, lut=magma) as plan:
        live = plan.live(1500)
        want = plan.process_batch(np.stack([x, x], 1))
        live.push(x[:n], channels=1)
        pos = n
        for f in range(want.shape[0]):
            col = live.pop_column()
            assert col is not None and np.array_equal(col, want[f]), f
            pos += live.push(x[pos:pos + hop], channels=1)
        live.close()


def test_seeded_random_configurations(sbx, magma):
    """24 seeded random geometries (odd hops, hop > N, tiny / huge frequency ranges, one-row images, both
    interpolators and schemes, all three kernel families) through the C ABI against the oracle."""
    import torch
    rng = np.random.Generator(np.random.PCG64(2026))
    sizes = [256, 512, 1024, 2048, 4096, 100, 300, 441, 1470, 2205, 2400, 8192]
    for case in range(24):
        n = int(rng.choice(sizes))
        if n == 8192:
            pad = 4
        elif n & (n - 1):
            pad = int(rng.choice([2, 4]))
        else:
            pad = int(rng.choice([1, 2, 4]))
            if n * pad > 16384:
                pad = 2
        sr = float(rng.choice([8000.0, 22050.0, 44100.0, 48000.0, 96000.0]))
        hop = int(rng.choice([1, 3, n // 8, n // 2, n, n + 17, int(rng.integers(1, 2 * n))]))
        rows = int(rng.choice([1, 2, 7, 33, 100, 257])) if n != 8192 else int(rng.choice([16, 64]))
        f_min = float(rng.choice([1.0, 20.0, 32.0, 500.0]))
        f_max = float(rng.choice([f_min * 1.5, 4000.0, sr / 2, 22030.0, sr]))
        f_max = f_max if f_max > f_min else f_min * 2
        interp, scheme, ch, frames = int(rng.integers(0, 2)), int(rng.integers(0, 2)), int(rng.integers(1, 3)), int(rng.integers(1, 5))
        ns = n + max(1, hop) * (frames - 1)
        amp = 10.0 ** rng.uniform(-4, 0.3)
        x = (amp * rng.standard_normal(ns)).astype(np.float32)
        pcm = x if ch == 1 else np.stack([x, (0.3 * amp * rng.standard_normal(ns)).astype(np.float32)], 1)
        kw = dict(sample_rate=sr, window_n=n, pad_factor=pad, hop=max(1, hop), rows_h=rows, f_min=f_min, f_max=f_max, interp=interp)
        p = oracle.params(stereo_scheme=scheme, **kw)
        o_rgba, o_mags = oracle.batch(p, pcm, ch, magma, want_mags=True)
        with sbx.Plan(channels=ch, scheme=scheme, lut=magma, **kw) as plan:
            g_rgba = plan.process_batch(pcm)
            d = torch.zeros((frames, plan.k, 2), dtype=torch.float32, device="cuda")
            plan.magnitudes_batch_device(torch.from_numpy(pcm).cuda(), ns, d, frames)
            plan.synchronize()
            g_mags = d.cpu().numpy()
            peak = np.maximum(np.abs(o_mags).max(axis=(1, 2), keepdims=True), 1e-30)
            assert (np.abs(g_mags - o_mags) / peak).max() <= MAG_TOL, (case, kw)
            rec = np.stack([oracle.column_from_magnitudes(p, g_mags[f], magma) for f in range(frames)])
            assert np.array_equal(plan.colorize_magnitudes(g_mags), rec), (case, kw)
            assert np.array_equal(g_rgba, rec), (case, kw)


@pytest.mark.parametrize("n,pad,ch,hop", [(16384, 4, 2, 1021), (16384, 4, 1, 1024), (32768, 2, 2, 4096)])
def test_large_path_host_batch_over_several_chunks(sbx, magma, monkeypatch, n, pad, ch, hop):
    """sbx_process_batch pipelines chunks over three slot streams, and the multi-kernel path (M > 16384) works in ONE
    plan-wide scratch: the chunks' kernels must be serialised (round-1 ADVICE, high).  1 MiB chunks -> many chunks in
    flight; the bytes must equal the single-stream device batch, and repeat identically."""
    import torch
    monkeypatch.setenv("SBX_BATCH_CHUNK_MB", "1")
    rows = 1024
    frames = 3 * (1 << 20) // (rows * 4) + 37                       # > 3 chunks of 256 frames
    n_samples = n + hop * (frames - 1)
    rng = np.random.Generator(np.random.PCG64(5))
    shape = (n_samples,) if ch == 1 else (n_samples, 2)
    # a different spectrum in every stretch, so that a column computed from another chunk's scratch cannot match
    t = np.arange(n_samples, dtype=np.float64)
    tone = np.sin(2 * np.pi * (200.0 + 6000.0 * t / n_samples) * t / 48000.0)
    pcm = (0.3 * tone.reshape((-1,) + (1,) * (len(shape) - 1)) + 0.05 * rng.standard_normal(shape)).astype(np.float32)
    with sbx.Plan(window_n=n, pad_factor=pad, hop=hop, rows_h=rows, channels=ch, lut=magma) as plan:
        assert plan.kernel_path == "large"
        d_pcm = torch.from_numpy(pcm).cuda()
        d_rgba = torch.zeros((frames, rows, 4), dtype=torch.uint8, device="cuda")
        assert plan.process_batch_device(d_pcm, n_samples, d_rgba, frames) == frames
        plan.synchronize()
        want = d_rgba.cpu().numpy()
        for _ in range(3):
            got = plan.process_batch(pcm)
            assert got.shape == want.shape
            bad = (got != want).any(axis=(1, 2))
            assert not bad.any(), f"{int(bad.sum())} of {frames} columns differ, first at frame {int(np.argmax(bad))}"


def test_stereo_scheme_pan_of_huge_equal_magnitudes(sbx, viridis):
    """pan = l / (|l| + |r|) in f32: for l == r the sum overflows to inf above FLT_MAX / 2 and the reference gets
    pan 0 (index 0), not 0.5 (round-1 ADVICE: the shortcut compared against 3e38)."""
    n, rows = 1024, 64
    p = oracle.params(window_n=n, pad_factor=2, rows_h=rows, stereo_scheme=1)
    k = n - 1
    vals = np.array([1.0, 1.5e38, 1.7014117e38, 1.7014118e38, np.float32(1.7014118e38) * np.float32(1.0000002),
                     1.8e38, 2.9e38, 3.2e38, 3.4e38, 0.0], np.float32)
    mags = np.repeat(vals[:, None, None], k, axis=1).repeat(2, axis=2).astype(np.float32)       # l == r in every bin
    want = np.stack([oracle.column_from_magnitudes(p, mags[f], viridis) for f in range(len(vals))])
    with sbx.Plan(window_n=n, pad_factor=2, rows_h=rows, channels=2, lut=viridis, scheme=1) as plan:
        assert np.array_equal(plan.colorize_magnitudes(mags), want)
    # the mono-pair shortcut (pan_index_equal) sees the same values through a mono batch: colours 128 or 0
    assert (want[0, :, :3] == viridis[128]).all() and (want[5, :, :3] == viridis[0]).all()


def test_live_stream_survives_plan_close(sbx, magma):
    """Plan.close() first closes its live streams (sbx_live_destroy never dereferences the plan): no use-after-free
    when a caller closes the plan before the stream object is collected."""
    plan = sbx.Plan(window_n=1024, hop=1024, rows_h=128, lut=magma)
    live = plan.live()
    live.push(signals.stereo_loopback(1024))
    assert live.pop_column() is not None
    plan.close()
    assert live._h is None
    live.close()                                                    # idempotent
    del live


def test_tiny_power_of_two_window_message(sbx, magma):
    with pytest.raises(sbx.SbxError) as e:
        sbx.Plan(window_n=4, pad_factor=1, lut=magma)
    assert e.value.status == -2 and ">= 8" in str(e.value)


Q2_CASES = [
    # (N, channels, hop, H, interp, scheme): pad factor 2, N = 2048 .. 8192 -- the packed-lane kernel of stft_q2.cuh
    (4096, 1, 512, 1024, 0, 0),        # cfg1 / cfg4 geometry (hop = 2 S: shared PCM loads)
    (4096, 1, 256, 512, 1, 0),         # hop = S
    (4096, 1, 333, 300, 0, 1),         # ragged hop, stereo scheme on mono pairs
    (2048, 2, 2048, 1024, 0, 0),       # cfg2 geometry
    (2048, 1, 100, 128, 0, 0),
    (4096, 2, 512, 1024, 1, 1),
    (8192, 2, 1024, 1024, 0, 0),       # M = 16384: one 512-thread CTA per SM
    (8192, 1, 255, 512, 0, 0),         # odd hop: the mono pair kernel, not the real-sequence geometry
]


@pytest.mark.parametrize("n,ch,hop,rows,interp,scheme", Q2_CASES)
def test_q2_packed_lane_kernel(sbx, magma, monkeypatch, n, ch, hop, rows, interp, scheme):
    """fused_q2_kernel against the oracle, and against fused_frames_kernel on the same input (same tolerance: the two
    differ only in rounding); planar multi-sequence launches and the magnitude output included."""
    import torch
    frames = 9
    n_samples = n + hop * (frames - 1) + 3
    pcm = _pcm(ch, n_samples, seed=11)
    p = oracle.params(window_n=n, pad_factor=2, hop=hop, rows_h=rows, interp=interp, stereo_scheme=scheme)
    o_rgba, o_mags = oracle.batch(p, pcm, ch, magma, want_mags=True)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("SBX_Q2", mode)
        monkeypatch.setenv("SBX_REAL", "0")
        with sbx.Plan(window_n=n, pad_factor=2, hop=hop, rows_h=rows, channels=ch, interp=interp, scheme=scheme, lut=magma) as plan:
            assert plan.kernel_path.endswith("-q2") == (mode == "1")
            d_pcm = torch.from_numpy(pcm).cuda()
            d_mags = torch.zeros((frames, plan.k, 2), dtype=torch.float32, device="cuda")
            d_rgba = torch.zeros((frames, rows, 4), dtype=torch.uint8, device="cuda")
            assert plan.magnitudes_batch_device(d_pcm, n_samples, d_mags, frames) == frames
            assert plan.process_batch_device(d_pcm, n_samples, d_rgba, frames) == frames
            plan.synchronize()
            g_mags, g_rgba = d_mags.cpu().numpy(), d_rgba.cpu().numpy()
            peak = np.abs(o_mags).max(axis=(1, 2), keepdims=True)
            assert (np.abs(g_mags - o_mags) / peak).max() <= MAG_TOL
            assert np.array_equal(plan.colorize_magnitudes(g_mags), g_rgba)           # fused == its own magnitudes coloured
            assert np.array_equal(plan.process_batch(pcm), g_rgba)
            assert (g_rgba != o_rgba).any(axis=2).mean() <= 2e-3
            if mode == "1" and hop % 4 == 0:
                # three planar sequences in one grid == three separate batches
                seq = n_samples + (-n_samples) % 4
                planar = np.zeros((3, seq) + pcm.shape[1:], np.float32)
                for s in range(3):
                    planar[s, :n_samples] = np.roll(pcm, 17 * s, axis=0)
                d_pl = torch.from_numpy(planar).cuda()
                d_out = torch.zeros((3, frames, rows, 4), dtype=torch.uint8, device="cuda")
                assert plan.process_batch_device_planar(d_pl, 3, seq, n_samples, d_out, 3 * frames) == frames
                plan.synchronize()
                for s in range(3):
                    assert np.array_equal(d_out[s].cpu().numpy(), plan.process_batch(planar[s, :n_samples]))
            out[mode] = g_mags
    peak = np.abs(out["0"]).max(axis=(1, 2), keepdims=True)
    assert (np.abs(out["1"] - out["0"]) / peak).max() <= MAG_TOL


@pytest.mark.parametrize("n,hop,rows,interp,scheme", [(8192, 256, 1024, 0, 0), (4096, 512, 512, 1, 0), (16384, 1024, 256, 0, 1),
                                                       (8192, 1000, 300, 0, 0)])
def test_q2_real_sequence_geometry(sbx, viridis, monkeypatch, n, hop, rows, interp, scheme):
    """fused_q2_kernel<Q2Geom<.., REAL>> (one mono frame per CTA, both sub-transforms in packed lanes) against the oracle
    and against the classic real-sequence kernel."""
    import torch
    frames = 9
    n_samples = n + hop * (frames - 1) + 3
    pcm = _pcm(1, n_samples, seed=13)
    p = oracle.params(window_n=n, pad_factor=2, hop=hop, rows_h=rows, interp=interp, stereo_scheme=scheme)
    o_rgba, o_mags = oracle.batch(p, pcm, 1, viridis, want_mags=True)
    out = {}
    monkeypatch.setenv("SBX_REAL", "1")
    for mode in ("0", "1"):
        monkeypatch.setenv("SBX_Q2", mode)
        with sbx.Plan(window_n=n, pad_factor=2, hop=hop, rows_h=rows, channels=1, interp=interp, scheme=scheme, lut=viridis) as plan:
            assert plan.kernel_path == ("fused-real-q2" if mode == "1" else "fused-real")
            d_pcm = torch.from_numpy(pcm).cuda()
            d_mags = torch.zeros((frames, plan.k, 2), dtype=torch.float32, device="cuda")
            d_rgba = torch.zeros((frames, rows, 4), dtype=torch.uint8, device="cuda")
            assert plan.magnitudes_batch_device(d_pcm, n_samples, d_mags, frames) == frames
            assert plan.process_batch_device(d_pcm, n_samples, d_rgba, frames) == frames
            plan.synchronize()
            g_mags, g_rgba = d_mags.cpu().numpy(), d_rgba.cpu().numpy()
            peak = np.abs(o_mags).max(axis=(1, 2), keepdims=True)
            assert (np.abs(g_mags - o_mags) / peak).max() <= MAG_TOL
            assert np.array_equal(g_mags[..., 0], g_mags[..., 1])                      # L == R for (s, s) input
            assert np.array_equal(plan.colorize_magnitudes(g_mags), g_rgba)
            assert np.array_equal(plan.process_batch(pcm), g_rgba)
            assert (g_rgba != o_rgba).any(axis=2).mean() <= 2e-3
            out[mode] = g_mags
    peak = np.abs(out["0"]).max(axis=(1, 2), keepdims=True)
    assert (np.abs(out["1"] - out["0"]) / peak).max() <= MAG_TOL


@pytest.mark.parametrize("n,pad,ch,hop,rows,entries", [(1024, 2, 2, 256, 256, 4096), (4096, 2, 2, 512, 1024, 1024),
                                                       (4096, 2, 1, 512, 512, 4096), (2400, 2, 2, 58, 128, 4096),
                                                       (16384, 4, 2, 2048, 256, 4096)])
def test_stereo_scheme_with_a_finer_pan_gradient(sbx, n, pad, ch, hop, rows, entries):
    """sbx_plan_set_pan_gradient (all kernel families): the stereo schemes' continuous diverging gradient
    (colorscheme.rs:63-66, RED_YELLOW_BLUE is the application's default) read from a table of `entries` steps.
    Bit-exact with the oracle given identical magnitudes; the palette table follows the installed gradient."""
    lut = sbx.builtin_lut("red_yellow_blue")
    pan = sbx.builtin_gradient("red_yellow_blue", entries)
    ns = n + hop * 4
    pcm = signals.stereo_loopback(ns) if ch == 2 else signals.white_noise(ns)
    if ch == 2:
        pcm[:, 1] *= np.linspace(0, 2, ns).astype(np.float32)
    p = oracle.params(window_n=n, pad_factor=pad, hop=hop, rows_h=rows, stereo_scheme=1, pan_levels=entries)
    with sbx.Plan(window_n=n, pad_factor=pad, hop=hop, rows_h=rows, channels=ch, lut=lut, scheme=1) as plan:
        coarse = plan.process_batch(pcm)
        plan.set_pan_gradient(pan)
        g = plan.process_batch(pcm)
        o_rgba, o_mags = oracle.batch(p, pcm, ch, pan, want_mags=True)
        assert np.array_equal(plan.colorize_magnitudes(o_mags), o_rgba)
        assert (g != o_rgba).any(axis=2).mean() <= 2e-3
        assert np.array_equal(g[..., 3], coarse[..., 3])                   # alpha does not depend on the gradient table
        if ch == 2:
            assert (g[..., :3] != coarse[..., :3]).any()
            # a single frame through the single-frame and the live entry points reads the same table
            lr = pcm[:n]
            col = plan.process_column(lr)
            assert np.array_equal(col, oracle.column_from_magnitudes(p, plan.process_magnitudes(lr), pan))
            live = plan.live(max(4096, n))
            live.push(lr)
            assert np.array_equal(live.pop_column(), col)
            live.close()
        tab = plan.palette_table(8)
        assert np.allclose(tab[0, 0, :3], pan[-1] / 256.0) and np.allclose(tab[0, -1, :3], pan[0] / 256.0)
    with sbx.Plan(window_n=1024, lut=lut, scheme=1) as plan:
        with pytest.raises(sbx.SbxError):
            plan.set_pan_gradient(np.zeros((1, 3), np.uint8))


@pytest.mark.parametrize("n,pad,ch,hop,rows", [(16384, 2, 2, 2048, 512), (16384, 4, 1, 2048, 2048), (16384, 4, 1, 2047, 256),
                                               (32768, 1, 2, 4096, 256), (32768, 2, 1, 4096, 1024), (32768, 4, 2, 4096, 1024),
                                               (32768, 4, 1, 4096, 2048), (65536, 1, 2, 8192, 128), (65536, 2, 1, 8192, 2048)])
def test_cluster_path_equals_four_kernel_path(sbx, magma, monkeypatch, n, pad, ch, hop, rows):
    """M = 32768 .. 131072: the transform held in the distributed shared memory of a 2 / 4 / 8-CTA cluster
    (large_cluster_kernel) against the four-kernel path through global scratch (SBX_LARGE_CLUSTER=0) -- the same
    butterflies in the same order -- and against the oracle."""
    import torch
    frames = 9
    ns = n + hop * (frames - 1) + 3
    pcm = _pcm(ch, ns, seed=11)
    p = oracle.params(window_n=n, pad_factor=pad, hop=hop, rows_h=rows)
    o_rgba, o_mags = oracle.batch(p, pcm, ch, magma, want_mags=True)
    got = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("SBX_LARGE_CLUSTER", mode)
        with sbx.Plan(window_n=n, pad_factor=pad, hop=hop, rows_h=rows, channels=ch, lut=magma) as plan:
            assert plan.kernel_path == "large"
            d_x = torch.from_numpy(pcm).cuda()
            d_rgba = torch.zeros((frames, rows, 4), dtype=torch.uint8, device="cuda")
            d_mags = torch.zeros((frames, plan.k, 2), dtype=torch.float32, device="cuda")
            assert plan.process_batch_device(d_x, ns, d_rgba, frames) == frames
            assert plan.magnitudes_batch_device(d_x, ns, d_mags, frames) == frames
            plan.synchronize()
            got[mode] = (d_rgba.cpu().numpy(), d_mags.cpu().numpy())
    # same algorithm, separately compiled (FMA contraction differs): equal to a few ulp of the frame peak, not bit for bit
    peak = np.abs(got["0"][1]).max(axis=(1, 2), keepdims=True)
    assert float((np.abs(got["0"][1] - got["1"][1]) / peak).max()) <= 2e-6
    assert (got["0"][0] != got["1"][0]).any(axis=2).mean() <= 2e-3
    g_rgba, g_mags = got["1"]
    err = float((np.abs(g_mags - o_mags) / np.abs(o_mags).max(axis=(1, 2), keepdims=True)).max())
    assert err <= MAG_TOL, err
    assert (g_rgba != o_rgba).any(axis=2).mean() <= 2e-3
