"""CPU: the kernel's own phase functions (stft_kernels.cuh compiled for the host by tests/emu) and the
product's host tables (sbx_tables.h) against the oracle.  This checks the index math, butterflies,
swizzles, L/R split, resampling and the dB step function without a GPU; the GPU suite
(test_gpu_parity.py) checks the same things through the C ABI on the device."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle
from tests import signals
from tests.emu import emu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MAG_TOL = 1e-5

CASES = [
    # (N, P, channels, hop, H, interp)
    (256, 2, 2, 64, 64, 0),
    (512, 2, 1, 128, 128, 0),
    (1024, 4, 2, 256, 256, 1),
    (2048, 2, 2, 2048, 1024, 0),
    (4096, 2, 1, 512, 1024, 0),
    (4096, 2, 2, 512, 1024, 1),
    (8192, 2, 1, 256, 1024, 0),
    (4096, 4, 1, 512, 2048, 0),
    (2048, 1, 2, 512, 512, 0),
    (16384, 1, 2, 2048, 128, 0),
    (256, 1, 1, 100, 77, 0),
    (1024, 2, 1, 1, 33, 0),
    # not a power of two: stft_generic.cuh (the live app's 2400 @ 48 kHz / 2205 @ 44.1 kHz windows)
    (2400, 2, 2, 58, 1024, 0),
    (2205, 2, 2, 86, 1024, 0),
    (2400, 2, 1, 93, 512, 1),
    (1600, 4, 1, 200, 256, 0),
    (100, 2, 2, 10, 64, 0),
    # M > 16384: the multi-kernel path of stft_large.cuh (cfg5's N >= 8192 at P = 4)
    (8192, 4, 1, 1024, 2048, 0),
    (16384, 2, 2, 2048, 512, 1),
    (32768, 1, 1, 4096, 256, 0),
    (65536, 4, 1, 8192, 2048, 0),
]


def _pcm(ch, ns, seed=0):
    rng = np.random.Generator(np.random.PCG64(seed))
    if ch == 1:
        return (signals.log_sweep(ns, duration=ns / 48000.0) + 0.02 * rng.standard_normal(ns)).astype(np.float32)
    return signals.stereo_loopback(ns, seed=seed + 1)


@pytest.mark.parametrize("n,pad,ch,hop,rows,interp", CASES)
def test_emulated_kernel_matches_oracle(magma, n, pad, ch, hop, rows, interp):
    ns = n + hop * 2 + 2
    pcm = _pcm(ch, ns)
    frames = oracle.num_frames(ns, n, hop)
    cfg = emu.make_config(magma, window_n=n, pad_factor=pad, hop=hop, rows_h=rows, channels=ch, interp=interp)
    mags, rgba = emu.batch(cfg, pcm)
    p = oracle.params(window_n=n, pad_factor=pad, hop=hop, rows_h=rows, interp=interp)
    o_rgba, o_mags = oracle.batch(p, pcm, ch, magma, want_mags=True)
    assert mags.shape == o_mags.shape and rgba.shape == o_rgba.shape
    peak = np.abs(o_mags).max(axis=(1, 2), keepdims=True)
    assert (np.abs(mags - o_mags) / peak).max() <= MAG_TOL
    # identical magnitudes -> bit-identical columns
    rec = np.stack([oracle.column_from_magnitudes(p, mags[f], magma) for f in range(frames)])
    assert np.array_equal(rec, rgba)
    assert ((rgba != o_rgba).any(axis=2)).mean() <= 2e-3


REAL_CASES = [
    # (N, P, hop, H, interp, scheme): mono; real transform lengths 8192, 16384 and 32768, every pad factor
    (4096, 2, 512, 1024, 0, 0),
    (2048, 4, 256, 512, 0, 0),
    (8192, 1, 1024, 256, 1, 0),
    (8192, 2, 256, 1024, 0, 0),          # cfg3
    (4096, 4, 512, 2048, 0, 1),
    (16384, 1, 2048, 512, 0, 0),
    (8192, 4, 1024, 2048, 0, 0),         # cfg5, N = 8192
    (16384, 2, 2048, 1024, 1, 0),
    (32768, 1, 4096, 512, 0, 1),
]


@pytest.mark.parametrize("n,pad,hop,rows,interp,scheme", REAL_CASES)
@pytest.mark.parametrize("real", [True, False])
def test_real_sequence_geometry_matches_oracle(magma, n, pad, hop, rows, interp, scheme, real):
    """Mono batches have two kernels: a frame PAIR packed into one complex transform, and ONE frame as the even/odd
    packed sequence of half the length (Geom<.., REAL>).  Both must reproduce the oracle, whichever the plan picks."""
    ns = n + hop * 2 + 2
    pcm = _pcm(1, ns, seed=n + pad)
    frames = oracle.num_frames(ns, n, hop)
    cfg = emu.make_config(magma, window_n=n, pad_factor=pad, hop=hop, rows_h=rows, channels=1, interp=interp, scheme=scheme)
    mags, rgba = emu.batch(cfg, pcm, real=real, staged=(n == 8192))
    p = oracle.params(window_n=n, pad_factor=pad, hop=hop, rows_h=rows, interp=interp, stereo_scheme=scheme)
    o_rgba, o_mags = oracle.batch(p, pcm, 1, magma, want_mags=True)
    peak = np.abs(o_mags).max(axis=(1, 2), keepdims=True)
    assert (np.abs(mags - o_mags) / peak).max() <= MAG_TOL
    assert np.array_equal(mags[..., 0], mags[..., 1])                   # (s, s) input: L == R
    rec = np.stack([oracle.column_from_magnitudes(p, mags[f], magma) for f in range(frames)])
    assert np.array_equal(rec, rgba)
    assert ((rgba != o_rgba).any(axis=2)).mean() <= 2e-3


def test_real_sequence_geometry_needs_an_even_hop(magma):
    """hop odd: frames start on odd samples, the pair loads would be misaligned -> the frame-pair kernel runs."""
    n, pad, hop, rows = 8192, 2, 255, 256
    pcm = _pcm(1, n + 3 * hop)
    cfg = emu.make_config(magma, window_n=n, pad_factor=pad, hop=hop, rows_h=rows, channels=1)
    m_auto, c_auto = emu.batch(cfg, pcm, real=True)
    m_pair, c_pair = emu.batch(cfg, pcm, real=False)
    assert np.array_equal(m_auto, m_pair) and np.array_equal(c_auto, c_pair)


@pytest.mark.parametrize("n,pad,hop,rows,interp,scheme", [
    (256, 2, 64, 64, 0, 0), (512, 1, 128, 33, 0, 0), (1024, 4, 256, 256, 1, 0), (2048, 2, 2048, 1024, 0, 0),
    (2048, 1, 512, 512, 0, 1), (4096, 1, 1024, 256, 0, 0), (256, 4, 64, 100, 1, 1), (1024, 2, 300, 2048, 0, 0)])
def test_wide_single_frame_geometry(magma, n, pad, hop, rows, interp, scheme):
    """Single stereo frames (sbx_process_column, the live graphs) run the same phases with up to four times the
    threads per CTA (Geom<.., WIDE>): another thread -> bin / row / sample mapping, the same arithmetic, so the
    result must equal the batch geometry bit for bit (and the oracle within tolerance)."""
    pcm = signals.stereo_loopback(n + 2 * hop)
    cfg = emu.make_config(magma, window_n=n, pad_factor=pad, hop=hop, rows_h=rows, channels=2, interp=interp, scheme=scheme)
    m_wide, c_wide = emu.batch(cfg, pcm, wide=True)
    m_base, c_base = emu.batch(cfg, pcm, wide=False)
    assert np.array_equal(m_wide, m_base) and np.array_equal(c_wide, c_base)
    p = oracle.params(window_n=n, pad_factor=pad, hop=hop, rows_h=rows, interp=interp, stereo_scheme=scheme)
    _, o_mags = oracle.batch(p, pcm, 2, magma, want_mags=True)
    assert (np.abs(m_wide - o_mags) / np.abs(o_mags).max(axis=(1, 2), keepdims=True)).max() <= MAG_TOL


@pytest.mark.parametrize("name,kw,ch,lut", [
    ("cfg1_sweep_mono", dict(sample_rate=48000.0, window_n=4096, pad_factor=2, hop=512, rows_h=1024), 1, "magma"),
    ("cfg2_stereo_live", dict(sample_rate=48000.0, window_n=2048, pad_factor=2, hop=2048, rows_h=1024), 2, "magma"),
    ("cfg3_96k_viridis", dict(sample_rate=96000.0, window_n=8192, pad_factor=2, hop=256, rows_h=1024), 1, "viridis"),
    ("cfg5_pad4_n1024", dict(sample_rate=48000.0, window_n=1024, pad_factor=4, hop=128, rows_h=2048), 1, "magma"),
    ("small_cosine", dict(sample_rate=44100.0, window_n=256, pad_factor=2, hop=64, rows_h=96, interp=1), 2, "viridis"),
])
def test_emulated_kernel_matches_golden(name, kw, ch, lut):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = emu.make_config(signals.lut(lut), channels=ch, **kw)
    mags, rgba = emu.batch(cfg, g["pcm"])
    peak = np.abs(g["mags"]).max(axis=(1, 2), keepdims=True)
    assert (np.abs(mags - g["mags"]) / peak).max() <= MAG_TOL
    assert ((rgba != g["rgba"]).any(axis=2)).mean() <= 2e-3
    assert np.array_equal(rgba[..., 3], g["rgba"][..., 3])


@pytest.mark.parametrize("kw", [
    dict(window_n=4096, pad_factor=2, rows_h=1024, sample_rate=48000.0),
    dict(window_n=8192, pad_factor=2, rows_h=1024, sample_rate=96000.0),
    dict(window_n=1024, pad_factor=4, rows_h=2048, sample_rate=48000.0),
    dict(window_n=2048, pad_factor=2, rows_h=333, sample_rate=44100.0, f_min=20.0, f_max=20000.0),
])
def test_host_tables_bit_exact(magma, kw):
    """Window, bin map and period computed by the product's sbx_tables.h == the oracle's, bit for bit."""
    t = emu.tables(emu.make_config(magma, **kw))
    p = oracle.params(**kw)
    f0, step, n, period = oracle.bin_map(p)
    assert np.array_equal(t["f0"].view(np.uint32), f0.view(np.uint32))
    assert np.array_equal(t["step"].view(np.uint32), step.view(np.uint32))
    assert np.array_equal(t["n"], n)
    assert np.float32(t["period"]).view(np.uint32) == np.float32(period).view(np.uint32)
    assert np.array_equal(t["window"].view(np.uint32), oracle.hann(kw["window_n"]).view(np.uint32))


@pytest.mark.parametrize("db", [(-70.0, -10.0, 1e-7), (-90.0, 0.0, 1e-9), (-40.0, -20.0, 1e-7)])
def test_colour_thresholds_reproduce_the_db_chain(magma, db):
    """step(p) = #{thr <= p} must equal clamp(floor(t * 256)) of the dB chain for every p: checked on
    each threshold, its two neighbouring floats, and 200k random powers."""
    db_min, db_max, eps = db
    t = emu.tables(emu.make_config(magma, db_min=db_min, db_max=db_max, db_eps=eps))
    thr = t["thr"][:255]
    assert np.all(np.diff(thr) > 0)
    L = oracle.lib()

    rng = np.random.Generator(np.random.PCG64(5))
    probes = [thr, np.nextafter(thr, np.float32(0)), np.nextafter(thr, np.float32(np.inf)),
              (10.0 ** rng.uniform(-11, 1, 200000)).astype(np.float32), np.array([0.0, np.inf], np.float32)]
    p_all = np.concatenate(probes).astype(np.float32)
    got = np.searchsorted(thr, p_all, side="right")
    # the oracle's dB chain evaluated on the power value itself (colorscheme.rs:60-61,69)
    want = np.empty(p_all.size, np.int64)
    for i, pv in enumerate(p_all):
        tt = L.orc_db_chain(float(pv), db_min, db_max, eps)
        want[i] = L.orc_gradient_index(tt)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n,pad,ch,hop,rows,interp", [(1024, 2, 2, 256, 256, 0), (1024, 2, 1, 256, 256, 0),
                                                      (4096, 2, 2, 512, 1024, 1), (16384, 2, 2, 2048, 512, 0)])
def test_stereo_pan_scheme(viridis, n, pad, ch, hop, rows, interp):
    """ColorScheme::new_stereo (colorscheme.rs:63-66): colour = gradient(l / (|l| + |r|)), alpha = normalised dB."""
    ns = n + hop * 3
    if ch == 2:
        pcm = signals.stereo_loopback(ns)
        pcm[:, 1] *= np.linspace(0, 2, ns).astype(np.float32)         # sweep the pan
    else:
        pcm = signals.white_noise(ns)
    cfg = emu.make_config(viridis, window_n=n, pad_factor=pad, hop=hop, rows_h=rows, channels=ch, interp=interp, scheme=1)
    mags, rgba = emu.batch(cfg, pcm)
    p = oracle.params(window_n=n, pad_factor=pad, hop=hop, rows_h=rows, interp=interp, stereo_scheme=1)
    rec = np.stack([oracle.column_from_magnitudes(p, mags[f], viridis) for f in range(mags.shape[0])])
    assert np.array_equal(rec, rgba)
    assert rgba[..., 3].min() < 64 and rgba[..., 3].max() > 128        # alpha really varies
    if ch == 1:
        assert np.all((rgba[..., :3] == viridis[128]).all(axis=2) | (rgba[..., :3] == viridis[0]).all(axis=2))


@pytest.mark.parametrize("n,ch,hop,rows,interp", [
    (4096, 1, 512, 1024, 0), (4096, 1, 256, 256, 1), (2048, 2, 2048, 1024, 0), (8192, 2, 1000, 512, 0),
    (8192, 1, 255, 256, 0), (1024, 1, 256, 128, 0), (512, 2, 77, 64, 1), (2048, 1, 100, 128, 0)])
def test_q2_packed_lane_phases_match_oracle(n, ch, hop, rows, interp):
    """The phase functions of stft_q2.cuh (pad factor 2, both sub-transforms per thread), host-compiled: element
    layout, swizzle, fused last pass, planar partner exchange and the packed split against the oracle."""
    lut = signals.lut("magma")
    frames = 5                                                       # odd: the half-filled mono pair
    ns = n + hop * (frames - 1) + 3
    rng = np.random.Generator(np.random.PCG64(3))
    pcm = ((signals.log_sweep(ns) + 0.02 * rng.standard_normal(ns)).astype(np.float32) if ch == 1
           else signals.stereo_loopback(ns, seed=2))
    cfg = emu.make_config(lut, window_n=n, pad_factor=2, hop=hop, rows_h=rows, channels=ch, interp=interp)
    mags, rgba = emu.batch(cfg, pcm, q2=True, real=False)
    p = oracle.params(window_n=n, pad_factor=2, hop=hop, rows_h=rows, interp=interp)
    o_rgba, o_mags = oracle.batch(p, pcm, ch, lut, want_mags=True)
    peak = np.abs(o_mags).max(axis=(1, 2), keepdims=True)
    assert (np.abs(mags - o_mags) / peak).max() <= 1e-5
    assert (rgba != o_rgba).any(axis=2).mean() <= 2e-3


@pytest.mark.parametrize("n,hop,rows,interp", [(8192, 256, 1024, 0), (4096, 512, 512, 1), (16384, 1024, 256, 0), (8192, 1000, 300, 0)])
def test_q2_real_sequence_phases_match_oracle(n, hop, rows, interp):
    """stft_q2.cuh on the real-sequence geometry (one mono frame per CTA as N/2 sample pairs): the even/odd
    recombination on packed lanes, the self-paired bin M/2 and the single-channel resample."""
    lut = signals.lut("viridis")
    frames = 4
    ns = n + hop * (frames - 1) + 3
    rng = np.random.Generator(np.random.PCG64(3))
    pcm = (signals.log_sweep(ns) + 0.02 * rng.standard_normal(ns)).astype(np.float32)
    cfg = emu.make_config(lut, window_n=n, pad_factor=2, hop=hop, rows_h=rows, channels=1, interp=interp)
    mags, rgba = emu.batch(cfg, pcm, q2=True, real=True)
    p = oracle.params(window_n=n, pad_factor=2, hop=hop, rows_h=rows, interp=interp)
    o_rgba, o_mags = oracle.batch(p, pcm, 1, lut, want_mags=True)
    peak = np.abs(o_mags).max(axis=(1, 2), keepdims=True)
    assert (np.abs(mags - o_mags) / peak).max() <= 1e-5
    assert (rgba != o_rgba).any(axis=2).mean() <= 2e-3


# ---- shared-memory layout rules of the resample tables (sbx_tables.h) ---------------------------------------------
LAYOUT_CASES = [
    # N, P, H, sample rate: cfg1/cfg4, cfg3 (real-sequence geometry: 4-byte elements, groups of 32), cfg2, cfg5, the app
    (4096, 2, 1024, 48000.0), (8192, 2, 1024, 96000.0), (2048, 2, 1024, 48000.0), (4096, 4, 2048, 48000.0),
    (512, 4, 2048, 48000.0), (2400, 2, 1024, 48000.0), (1024, 1, 333, 44100.0),
]


@pytest.mark.parametrize("fill", [0, 1])
@pytest.mark.parametrize("n,pad,rows,sr", LAYOUT_CASES)
def test_sample_table_lane_groups_are_conflict_free(magma, monkeypatch, n, pad, rows, sr, fill):
    """One lane per table entry; an aligned group of `bank_group` entries is served by one shared-memory request.
    Checked here: (1) the real entries, in order, are exactly the (row, sample) list of the bin map with the right
    bin and slot; (2) with filler entries (SBX_TABLE_FILL=1; measured slower on a B200 and therefore off by default,
    profiles/README.md) within every group the bins span < bank_group (the four tap loads are conflict-free) and so
    do the slots of the entries that store; (3) filler entries repeat a bin of their group and store nothing;
    (4) the rows one request serves start at different residues, so the sequential row sums are conflict-free;
    (5) no two samples of a chunk share a slot and every slot is inside the scratch area."""
    monkeypatch.setenv("SBX_TABLE_FILL", str(fill))
    cfg = emu.make_config(magma, sample_rate=sr, window_n=n, pad_factor=pad, hop=n // 8, rows_h=rows, channels=1)
    tab = emu.sample_table(cfg)
    assert tab["ok"]
    g = tab["bank_group"]
    assert g in (16, 32)
    samp, chunks, rowinfo = tab["samp"], tab["chunks"], tab["rowinfo"]
    x1, slot = samp[:, 0] & 0x1ffff, samp[:, 0] >> 17
    filler = slot == 0x7fff
    tb = emu.tables(cfg)
    k_out = n * pad // 2 - 1
    cap = n * pad // 2
    n_rows = rowinfo & 0xffff
    aux = rowinfo >> 16
    assert np.array_equal(n_rows, tb["n"])
    # (1) real entries == the bin map's samples, row by row
    want_x, want_slot = [], []
    for py in range(rows):
        s = np.arange(n_rows[py], dtype=np.float32)
        fq = tb["f0"][py] + s * tb["step"][py]
        x = np.clip(fq * tb["period"], np.float32(0), np.float32(k_out - 1)).astype(np.float32)
        want_x.append(np.floor(x).astype(np.uint32))
        want_slot.append(aux[py] + np.arange(n_rows[py], dtype=np.uint32))
    assert np.array_equal(x1[~filler], np.concatenate(want_x))
    assert np.array_equal(slot[~filler], np.concatenate(want_slot))
    fillers = 0
    for row_begin, row_end, s_begin, s_end in chunks:
        cx, cs, cf = x1[s_begin:s_end].astype(np.int64), slot[s_begin:s_end].astype(np.int64), filler[s_begin:s_end]
        fillers += int(cf.sum())
        for lo in range(0, len(cx), g):                                      # (2), (3)
            gx, gs, gf = cx[lo:lo + g], cs[lo:lo + g], cf[lo:lo + g]
            real = gs[~gf]
            assert len(real) >= 1 and len(set(real)) == len(real)
            if fill:
                assert gx.max() - gx.min() < g, (lo, gx)
                assert real.max() - real.min() < g
            assert set(gx[gf]) <= set(gx[~gf])
        real_slots = cs[~cf]
        assert len(set(real_slots)) == len(real_slots) and real_slots.max() < cap          # (5)
        for lo in range(row_begin, row_end, g):                              # (4)
            res = (aux[lo:min(lo + g, row_end)] % g).tolist()
            assert len(set(res)) == len(res), (lo, res)
    # the fillers cost at most a tenth more lanes, and the BASELINE geometries still need a single chunk
    assert fillers <= 0.1 * (~filler).sum() + g if fill else fillers == 0
    if (n, pad) in ((4096, 2), (8192, 2), (4096, 4)):
        assert len(chunks) == 1


@pytest.mark.parametrize("n,pad,ch,hop,rows,entries", [(1024, 2, 2, 256, 256, 4096), (512, 2, 1, 128, 100, 1024), (2400, 2, 2, 58, 128, 4096)])
def test_stereo_scheme_with_a_finer_pan_gradient(n, pad, ch, hop, rows, entries):
    """sbx_plan_set_pan_gradient: the pan lookup reads a table of `entries` steps of the continuous diverging
    gradient instead of the 256 of lut_rgb; alpha is unchanged.  Bit-exact against the oracle given the GPU-side
    (here: emulated) magnitudes."""
    import spectrogram_rs_b200 as sbx
    lut = sbx.builtin_lut("red_yellow_blue")
    pan = sbx.builtin_gradient("red_yellow_blue", entries)
    cfg = emu.make_config(lut, window_n=n, pad_factor=pad, hop=hop, rows_h=rows, channels=ch, scheme=1)
    x = signals.stereo_loopback(n + hop * 4) if ch == 2 else signals.log_sweep(n + hop * 4)
    emu.set_pan_gradient(pan)
    try:
        mags, rgba = emu.batch(cfg, x)
    finally:
        emu.set_pan_gradient(None)
    p = oracle.params(window_n=n, pad_factor=pad, hop=hop, rows_h=rows, stereo_scheme=1, pan_levels=entries)
    for f in range(mags.shape[0]):
        assert np.array_equal(rgba[f], oracle.column_from_magnitudes(p, mags[f], pan)), f
    # the finer table does change pixels relative to the 256-step one (it is not a no-op) for genuinely stereo input
    if ch == 2:
        _, coarse = emu.batch(cfg, x)
        assert (coarse[..., :3] != rgba[..., :3]).any() and np.array_equal(coarse[..., 3], rgba[..., 3])


@pytest.mark.parametrize("n,pad,ch,hop,rows,interp", [(1102, 2, 2, 43, 256, 0), (551, 2, 1, 21, 128, 1), (1103, 2, 1, 100, 64, 0),
                                                      (347, 4, 2, 50, 64, 0), (5461, 2, 2, 700, 64, 0), (94, 2, 2, 9, 32, 0)])
def test_bluestein_windows_match_oracle(magma, n, pad, ch, hop, rows, interp):
    """Windows with a prime factor > 7 (cpal devices at 22.05 / 11.025 kHz: N = (0.05 s * sr) as usize = 1102, 551;
    fft.rs:19-21 takes any N): chirp-z over a power-of-two circular convolution, forward DIF -> product with the
    transformed chirp in digit-reversed order -> inverse as the reversed passes.  Against the oracle's f64 DFT."""
    cfg = emu.make_config(magma, sample_rate=22050.0, window_n=n, pad_factor=pad, hop=hop, rows_h=rows, channels=ch, interp=interp)
    ns = n + hop * 3
    lr = signals.stereo_loopback(ns, 22050.0)
    x = lr if ch == 2 else lr[:, 0].copy()
    mags, rgba = emu.batch(cfg, x)
    p = oracle.params(sample_rate=22050.0, window_n=n, pad_factor=pad, hop=hop, rows_h=rows, interp=interp)
    o_rgba, o_mags = oracle.batch(p, x, ch, magma, want_mags=True)
    err = (np.abs(mags - o_mags) / np.abs(o_mags).max(axis=(1, 2), keepdims=True)).max()
    assert err <= 1e-5, err
    assert (rgba != o_rgba).any(axis=2).mean() <= 2e-3
    for f in range(mags.shape[0]):
        assert np.array_equal(rgba[f], oracle.column_from_magnitudes(p, mags[f], magma))


@pytest.mark.parametrize("n,pad,rows,sr", LAYOUT_CASES)
def test_budgeted_fillers_never_add_a_table_round(magma, monkeypatch, n, pad, rows, sr):
    """SBX_TABLE_FILL=2: fillers are spent only while the chunk's table stays within the rounds of 2048 entries
    (threads x loads in flight) its real entries need anyway; the real entries are untouched."""
    cfg = emu.make_config(magma, sample_rate=sr, window_n=n, pad_factor=pad, hop=n // 8, rows_h=rows, channels=1)
    monkeypatch.setenv("SBX_TABLE_FILL", "0")
    dense = emu.sample_table(cfg)
    monkeypatch.setenv("SBX_TABLE_FILL", "2")
    tab = emu.sample_table(cfg)
    assert tab["ok"] and len(tab["chunks"]) == len(dense["chunks"])
    real = (tab["samp"][:, 0] >> 17) != 0x7fff
    assert np.array_equal(tab["samp"][real], dense["samp"])
    for (_, _, b, e), (_, _, b0, e0) in zip(tab["chunks"], dense["chunks"]):
        assert (int(e) - int(b) + 2047) // 2048 == (int(e0) - int(b0) + 2047) // 2048
        assert int(e) - int(b) >= int(e0) - int(b0)


@pytest.mark.parametrize("log2m", [15, 16, 17])
def test_cluster_path_column_assignment(log2m):
    """large_cluster_kernel (stft_large.cuh): the Q = M / 4096 blocks ("columns") of a transform are spread over
    Q / 4 CTAs, four slots each; the L/R split of bin kk = k Q + c needs column Q - c (0 and Q / 2 pair with themselves),
    so partners must share a CTA, in adjacent slots (the kernel reads the partner at slot ^ 1)."""
    import ctypes as C
    fn = emu.lib().emu_cluster_layout
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_void_p]
    out = np.zeros((64, 3), np.int32)
    q = fn(log2m, out.ctypes.data)
    assert q == (1 << log2m) // 4096
    owner, slot, back = out[:q, 0], out[:q, 1], out[:q, 2]
    assert np.array_equal(back, np.arange(q))                               # column_of_slot inverts the assignment
    assert sorted(zip(owner.tolist(), slot.tolist())) == [(j, s) for j in range(q // 4) for s in range(4)]
    for c in range(q):
        partner = c if c in (0, q // 2) else q - c
        assert owner[partner] == owner[c]
        assert slot[partner] == (slot[c] if partner == c else slot[c] ^ 1)


@pytest.mark.parametrize("n,pad,ch,hop,rows,interp", [(16384, 2, 2, 2048, 128, 0), (8192, 4, 2, 1000, 64, 1), (32768, 1, 1, 4095, 64, 0),
                                                      (16384, 4, 1, 2048, 256, 0), (32768, 4, 2, 4096, 64, 0)])
def test_cluster_kernel_phases_match_oracle(magma, n, pad, ch, hop, rows, interp):
    """large_cluster_kernel's own phase functions, every CTA rank and thread id in turn (the st.shared::cluster scatter
    becomes a store into the owning rank's buffer): against the oracle, and bit-identical to the four-kernel path's
    phases (same butterflies in the same order; the host build contracts nothing)."""
    frames = 3
    ns = n + hop * (frames - 1)
    lr = signals.stereo_loopback(ns)
    x = lr if ch == 2 else lr[:, 0].copy()
    cfg = emu.make_config(magma, window_n=n, pad_factor=pad, hop=hop, rows_h=rows, channels=ch, interp=interp)
    mags_c, rgba_c = emu.batch(cfg, x, cluster=True, real=False)
    mags_4, rgba_4 = emu.batch(cfg, x, cluster=False, real=False)
    assert np.array_equal(mags_c, mags_4) and np.array_equal(rgba_c, rgba_4)
    p = oracle.params(window_n=n, pad_factor=pad, hop=hop, rows_h=rows, interp=interp)
    o_rgba, o_mags = oracle.batch(p, x, ch, magma, want_mags=True)
    err = (np.abs(mags_c - o_mags) / np.abs(o_mags).max(axis=(1, 2), keepdims=True)).max()
    assert err <= 1e-5, err
    assert (rgba_c != o_rgba).any(axis=2).mean() <= 2e-3
