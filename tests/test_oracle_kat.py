"""CPU: known-answer tests that pin the oracle analytically, independent of
any FFT library and of the oracle's own golden vectors."""
import os

import numpy as np
import pytest

from oracle import oracle
from tests import signals

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sine_frame(n, pad, k0, amp=1.0):
    t = np.arange(n, dtype=np.float64)
    return (amp * np.sin(2 * np.pi * k0 * t / (n * pad))).astype(np.float32)


@pytest.mark.parametrize("n,pad,k0", [(4096, 2, 400), (1024, 2, 100), (2048, 4, 300), (256, 2, 40)])
def test_k1_unit_sine_at_exact_padded_bin(n, pad, k0):
    s = _sine_frame(n, pad, k0)
    m = oracle.frame_magnitudes(np.stack([s, s], 1), n, pad)
    assert m.shape == (n * pad // 2 - 1, 2)
    j = k0 - 1                                            # output index j <-> FFT bin j+1 (DC dropped)
    tol = 2e-6 if n >= 1024 else 1e-5                     # short windows: the negative-frequency image leaks in
    assert abs(m[j, 0] - 0.5) < tol and abs(m[j, 1] - 0.5) < tol
    if pad == 2:
        # Hann main lobe sampled at half-bin spacing: 0.5, 0.424413, 0.25, 0.084883, then ~0
        for d, want in ((1, 0.42441318), (2, 0.25), (3, 0.08488264)):
            assert abs(m[j + d, 0] - want) < tol and abs(m[j - d, 0] - want) < tol
        # beyond the main lobe the integer (unpadded) bins vanish; the half bins carry the side lobes
        assert m[j + 4:j + 30:2, 0].max() < tol and m[j - 30:j - 3:2, 0].max() < tol
        assert abs(m[j + 5, 0] - 0.012126088) < 2 * tol


def test_k2_stereo_separation():
    n, pad, k0 = 4096, 2, 400
    s = _sine_frame(n, pad, k0)
    m = oracle.frame_magnitudes(np.stack([s, np.zeros_like(s)], 1), n, pad)
    assert abs(m[k0 - 1, 0] - 0.5) < 2e-6
    assert m[:, 1].max() < 1e-6
    m = oracle.frame_magnitudes(np.stack([np.zeros_like(s), 0.5 * s], 1), n, pad)
    assert abs(m[k0 - 1, 1] - 0.25) < 2e-6 and m[:, 0].max() < 1e-6


def test_k3_dc_input_leaks_only_the_hann_main_lobe():
    n, pad = 1024, 2
    ones = np.ones((n, 2), np.float32)
    m = oracle.frame_magnitudes(ones, n, pad)
    # DC itself (k = 0) is dropped; k = 1, 2, 3 carry the main lobe of amplitude 1 (no +- split at DC)
    for j, want in ((0, 0.84882636), (1, 0.5), (2, 0.16976527)):
        assert abs(m[j, 0] - want) < 3e-6
    assert m[3::2, 0].max() < 2e-6                          # k = 4, 6, ...: integer bins of the unpadded window
    assert abs(m[4, 0] - 2 * 0.012126088) < 4e-6            # k = 5: first side lobe


def test_k4_impulse_gives_flat_spectrum():
    n, pad, i = 1024, 2, 300
    x = np.zeros((n, 2), np.float32)
    x[i] = (1.0, 1.0)
    m = oracle.frame_magnitudes(x, n, pad)
    want = oracle.hann(n)[i] * 2.0 / n
    assert np.abs(m - want).max() < 1e-6 * want + 1e-9


@pytest.mark.parametrize("n,pad", [(256, 2), (512, 4), (1024, 2), (300, 2)])
def test_k5_matches_f64_dft(n, pad):
    """Independent referee: numpy's f64 FFT of the f64-windowed signal (300: the non-power-of-two path)."""
    rng = np.random.Generator(np.random.PCG64(7))
    lr = rng.standard_normal((n, 2)).astype(np.float32)
    m = oracle.frame_magnitudes(lr, n, pad)
    w = oracle.hann(n).astype(np.float64)
    z = (lr[:, 0].astype(np.float64) + 1j * lr[:, 1].astype(np.float64)) * w
    zf = np.fft.fft(np.concatenate([z, np.zeros(n * (pad - 1))]))
    mm = n * pad
    k = np.arange(1, mm // 2)
    a, b = zf[k], zf[mm - k]
    truth = np.stack([np.abs(a + np.conj(b)) / 2, np.abs(a - np.conj(b)) / 2], 1) * (2.0 / n)
    assert np.abs(m - truth).max() <= 5e-7 * truth.max()
    # Parseval on the windowed sequence
    zf32 = oracle.fft_c2c_f32(oracle.window_pad(lr, n, pad))
    e_t = (np.abs(z) ** 2).sum()
    e_f = (zf32.astype(np.float64) ** 2).sum() / mm
    assert abs(e_t - e_f) <= 1e-5 * e_t


def test_hann_window_is_periodic_f32():
    for n in (256, 2400, 4096):
        w = oracle.hann(n)
        assert w.dtype == np.float32 and w[0] == 0.0
        assert abs(w[n // 2] - 1.0) < 1e-6
        assert np.abs(w[1:] - w[1:][::-1]).max() < 2e-6          # w[i] == w[N-i]: periodic, not symmetric
        ref = 0.5 * (1 - np.cos(2 * np.pi * np.arange(n) / n))
        assert np.abs(w - ref).max() < 2e-6


def test_k6_bin_map_invariants_and_survey_probes():
    # cfg1: SURVEY.md 8a-5 probe values
    p = oracle.params(sample_rate=48000.0, window_n=4096, pad_factor=2, rows_h=1024)
    f0, step, n, period = oracle.bin_map(p)
    assert f0[0] == np.float32(32.0)
    assert abs(float(f0[0]) * float(period) - 5.46) < 0.01
    top = oracle.lib().orc_row_edge(32.0, 22030.0, 1024, 1024)
    assert abs(top - 22030.0) <= 22030.0 * 2 ** -23
    assert abs(top * float(period) - 3758.9) < 0.1
    assert n.min() >= 1 and n.max() == 23 and int(n.sum()) == 3893
    assert np.all(np.diff(f0) > 0)
    edges = np.array([oracle.lib().orc_row_edge(32.0, 22030.0, py, 1024) for py in range(1025)])
    assert np.all(np.diff(edges) > 0)
    ratio = edges[1:] / edges[:-1]
    assert np.abs(ratio - (22030.0 / 32.0) ** (1 / 1024)).max() < 1e-6
    # cfg3 has the same indices (sr and N both doubled)
    p3 = oracle.params(sample_rate=96000.0, window_n=8192, pad_factor=2, rows_h=1024)
    assert int(oracle.bin_map(p3)[2].sum()) in (3893, 3894, 3892)
    # the live app: N = 2400 @ 48 kHz
    pl = oracle.params(sample_rate=48000.0, window_n=2400, pad_factor=2, rows_h=1024)
    assert int(oracle.bin_map(pl)[2].sum()) == 2460
    # cfg5 largest: N = 65536, H = 2048 (P = 2 in the probe)
    p5 = oracle.params(sample_rate=48000.0, window_n=65536, pad_factor=2, rows_h=2048)
    n5 = oracle.bin_map(p5)[2]
    assert n5.max() == 191 and int(n5.sum()) == 59441


def test_k7_colour_index_steps():
    """p at the centre of LUT step i maps to index i; the ends clamp."""
    for i in range(256):
        p = 10.0 ** ((-70.0 + 60.0 * (i + 0.5) / 256.0) / 10.0) - 1e-7
        assert p > 0
        assert oracle.color_index(float(np.float32(np.sqrt(p))), 0.0) == i
    assert oracle.color_index(0.0, 0.0) == 0                       # dB = 10 log10f(1e-7f) = -70 -> t <= 0 -> index 0
    assert oracle.color_index(float(np.float32(0.4)), 0.0) == 255  # p = 0.16 > 0.1: t > 1, clamped
    assert oracle.color_index(float("inf"), 0.0) == 255
    assert oracle.color_index(float("nan"), 0.0) == 0
    t = oracle.lib().orc_normalised_db(0.0, 0.0, -70.0, -10.0, 1e-7)
    assert -2e-7 < t <= 0.0
    assert oracle.lib().orc_gradient_index(float("nan")) == 0
    assert oracle.lib().orc_gradient_index(-1e-7) == 0
    assert oracle.lib().orc_gradient_index(1.0000001) == 255
    assert oracle.lib().orc_gradient_index(0.5) == 128
    assert oracle.lib().orc_gradient_index(0.49999997) == 127


def test_colour_index_is_monotone_in_power():
    """The product replaces log10 by 255 thresholds; that needs the dB chain to be monotone in p."""
    rng = np.random.Generator(np.random.PCG64(11))
    p = np.sort(np.concatenate([10.0 ** rng.uniform(-9, 1, 200000), [0.0, 1e-7, 0.1, 1.0]])).astype(np.float32)
    L = oracle.lib()
    idx = np.array([L.orc_gradient_index(L.orc_normalised_db(float(np.sqrt(v)), 0.0, -70.0, -10.0, 1e-7)) for v in p[::40]])
    assert np.all(np.diff(idx) >= 0) and idx[0] == 0 and idx[-1] == 255


def test_lut_spot_values_match_d3_ramps(magma, viridis):
    assert tuple(magma[0]) == (0, 0, 4) and tuple(magma[255]) == (252, 253, 191)          # "#000004" .. "#fcfdbf"
    assert tuple(magma[127]) == (181, 54, 122) and tuple(magma[128]) == (183, 55, 121)    # "#b5367a", "#b73779"
    assert tuple(viridis[0]) == (68, 1, 84) and tuple(viridis[255]) == (253, 231, 37)     # "#440154" .. "#fde725"
    assert tuple(signals.lut("inferno")[0]) == (0, 0, 4) and tuple(signals.lut("plasma")[0]) == (13, 8, 135)


def test_framing_and_literal_conversions():
    L = oracle.lib()
    assert L.orc_window_from_period(0.05, 48000.0) == 2400 and L.orc_window_from_period(0.05, 44100.0) == 2205
    assert L.orc_hop_from_stride(2.0 / 1024.0, 48000.0) == 93           # simple_spectrogram.rs:102
    assert L.orc_hop_from_stride(2.5 / 2048.0, 48000.0) == 58           # gpu_spectrogram.rs:21-23
    assert oracle.num_frames(2880000, 4096, 512) == 5618                # cfg1: 60 s @ 48 kHz
    assert oracle.num_frames(345600000, 8192, 256) == 1349969           # cfg3: 1 h @ 96 kHz
    assert oracle.num_frames(4095, 4096, 512) == 0 and oracle.num_frames(4096, 4096, 512) == 1
    assert oracle.frame_magnitudes(np.zeros((100, 2), np.float32), 256, 2) is None      # the reference's None


def test_column_is_flipped_and_opaque(magma):
    """Row 0 of a column is the highest frequency; alpha = 255 for mono schemes."""
    n, pad = 1024, 2
    p = oracle.params(window_n=n, pad_factor=pad, rows_h=256)
    s = _sine_frame(n, pad, 8, 0.5)                               # 187 Hz: a low tone -> bright near the bottom
    m = oracle.frame_magnitudes(np.stack([s, s], 1), n, pad)
    col = oracle.column_from_magnitudes(p, m, magma)
    assert col.shape == (256, 4) and np.all(col[:, 3] == 255)
    rows = oracle.rows_from_magnitudes(p, m)
    assert 40 < int(rows[:, 0].argmax()) < 90                      # py counts up from the lowest frequency
    for py in range(256):                                          # pixel row = H - 1 - py (simple_spectrogram.rs:150)
        idx = oracle.color_index(float(rows[py, 0]), float(rows[py, 1]))
        assert tuple(col[255 - py, :3]) == tuple(magma[idx])


@pytest.mark.parametrize("name", ["cfg1_sweep_mono", "cfg2_stereo_live", "cfg3_96k_viridis", "cfg5_pad4_n1024", "small_cosine"])
def test_k8_oracle_reproduces_golden_vectors(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    want = np.load(os.path.join(GOLDEN, name + ".npz"))
    got = mg.generate(name)
    assert np.array_equal(got["pcm"], want["pcm"])
    assert np.array_equal(got["rgba"], want["rgba"])
    assert np.array_equal(got["mags"].view(np.uint32), want["mags"].view(np.uint32))


def test_threaded_batch_equals_serial(magma):
    p = oracle.params(window_n=512, pad_factor=2, hop=128, rows_h=128)
    x = signals.white_noise(512 + 128 * 50)
    a, am = oracle.batch(p, x, 1, magma, want_mags=True, threads=1)
    b, bm = oracle.batch(p, x, 1, magma, want_mags=True, threads=4)
    assert np.array_equal(a, b) and np.array_equal(am, bm)
    c, _ = oracle.batch(p, x, 1, magma, frame_lo=10, frame_hi=20)
    assert np.array_equal(c, a[10:20])


def test_fast_build_equals_the_checker():
    """bench.py times liboracle_fast.so (-O3 -march=native, no fast-math, no contraction): same source, and the same
    results as the -O2 checker build -- columns and magnitudes bit for bit; with pocketfft as the FFT stage the
    columns agree to within a colour step on at most 0.1 % of the pixels (a different, equally valid f32 FFT)."""
    from tests import signals
    lut = signals.lut("viridis")
    for ch, kw in ((1, dict(window_n=1024, pad_factor=2, hop=256, rows_h=256)),
                   (2, dict(window_n=512, pad_factor=4, hop=100, rows_h=77, interp=1))):
        p = oracle.params(**kw)
        ns = kw["window_n"] + kw["hop"] * 40
        x = signals.log_sweep(ns) if ch == 1 else signals.stereo_loopback(ns)
        rgba, mags = oracle.batch(p, x, ch, lut, want_mags=True)
        rgba_f, mags_f = oracle.batch(p, x, ch, lut, want_mags=True, fast=True, threads=3)
        assert np.array_equal(rgba, rgba_f) and np.array_equal(mags, mags_f)
        rgba_p = oracle.batch_pocketfft(p, x, ch, lut, threads=2, chunk_frames=16)
        assert rgba_p.shape == rgba.shape
        assert (rgba_p != rgba).any(axis=2).mean() <= 1e-3
        assert np.abs(rgba_p.astype(int) - rgba.astype(int)).max() <= 40


# ---- the 19 gradients of default_color_schemes() (src/colorscheme.rs:129-149) ------------------------------------
REFERENCE_SCHEMES = [  # (colorous constant lower-cased, stereo?) in the reference's list order
    ("red_yellow_blue", True), ("magma", False), ("viridis", False), ("red_blue", True), ("spectral", True),
    ("red_yellow_green", True), ("pink_green", True), ("purple_orange", True), ("inferno", False), ("plasma", False),
    ("cividis", False), ("cubehelix", False), ("turbo", False), ("cool", False), ("reds", False), ("blues", False),
    ("greens", False), ("greys", False), ("oranges", False)]


def test_builtin_gradients_cover_the_reference_list():
    """sbx_builtin_lut / sbx_builtin_gradient know every gradient the reference lists, in its order (host-side
    functions of the product library: no GPU involved)."""
    import re
    import spectrogram_rs_b200 as sbx
    assert sbx.builtin_gradient_names() == REFERENCE_SCHEMES
    ref = "/root/reference/src/colorscheme.rs"
    if os.path.exists(ref):                                   # in the build container only: the constants as written there
        names = re.findall(r"ColorScheme::new_(mono|stereo)\((\w+),", open(ref).read())
        assert [(n.lower(), kind == "stereo") for kind, n in names] == REFERENCE_SCHEMES
    for name, _ in REFERENCE_SCHEMES:
        t = sbx.builtin_lut(name)
        assert t.shape == (256, 3) and len(np.unique(t, axis=0)) > 100
    with pytest.raises(sbx.SbxError):
        sbx.builtin_lut("jet")
    py = sbx.default_color_schemes()
    assert len(py) == 19 and [s.is_stereo for s in py] == [st for _, st in REFERENCE_SCHEMES]
    assert py[0].name == "Blue-Yellow-Red (Stereo)" and py[0].pan_gradient.shape == (4096, 3)


def test_builtin_gradients_against_second_sources():
    """[unvendored] colorous is absent; what can be cross-checked here: the four ramps equal OpenCV's matplotlib
    tables exactly; d3's polynomial Turbo / Cividis approximate the original tables OpenCV ships (their published
    accuracy: a few levels on average); the ColorBrewer splines start and end on the scheme's first / last colour;
    and a table of E steps deviates from the continuous gradient by <= 2 levels at E = 256 (6 for Turbo, whose
    blue channel is steep) and <= 1 at E = 4096 -- the numbers include/sbx.h quotes for sbx_plan_set_pan_gradient."""
    cv2 = pytest.importorskip("cv2")
    import spectrogram_rs_b200 as sbx
    ramp = np.arange(256, dtype=np.uint8).reshape(256, 1)
    for name, code, max_tol, mean_tol in (("magma", cv2.COLORMAP_MAGMA, 0, 0), ("viridis", cv2.COLORMAP_VIRIDIS, 0, 0),
                                          ("inferno", cv2.COLORMAP_INFERNO, 0, 0), ("plasma", cv2.COLORMAP_PLASMA, 0, 0),
                                          ("turbo", cv2.COLORMAP_TURBO, 30, 7), ("cividis", cv2.COLORMAP_CIVIDIS, 20, 3)):
        ref = cv2.applyColorMap(ramp, code).reshape(256, 3)[:, ::-1].astype(int)
        d = np.abs(sbx.builtin_lut(name).astype(int) - ref)
        assert d.max() <= max_tol and d.mean() <= mean_tol, (name, d.max(), d.mean())
    ends = {"red_yellow_blue": (0xa50026, 0x313695), "red_blue": (0x67001f, 0x053061), "spectral": (0x9e0142, 0x5e4fa2),
            "red_yellow_green": (0xa50026, 0x006837), "pink_green": (0x8e0152, 0x276419), "purple_orange": (0x2d004b, 0x7f3b08),
            "reds": (0xfff5f0, 0x67000d), "blues": (0xf7fbff, 0x08306b), "greens": (0xf7fcf5, 0x00441b),
            "greys": (0xffffff, 0x000000), "oranges": (0xfff5eb, 0x7f2704)}
    for name, (first, last) in ends.items():
        dense = sbx.builtin_gradient(name, 65536).astype(int)
        for got, want in ((dense[0], first), (dense[-1], last)):
            assert np.abs(got - [(want >> 16) & 255, (want >> 8) & 255, want & 255]).max() <= 1, name
    for name, _ in REFERENCE_SCHEMES:
        dense = sbx.builtin_gradient(name, 65536).astype(int)
        for entries, tol in ((256, 6 if name == "turbo" else 2), (1024, 2), (4096, 1)):
            tab = sbx.builtin_gradient(name, entries).astype(int)
            idx = np.minimum((np.arange(65536) + 0.5) / 65536 * entries, entries - 1).astype(int)
            assert np.abs(tab[idx] - dense).max() <= tol, (name, entries)
