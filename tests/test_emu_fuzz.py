"""CPU: randomised configurations (hypothesis) through the host-emulated kernel phases against the oracle:
odd hops, hops larger than the window, tiny and huge frequency ranges (indices clamped at both spectrum
edges), one-row images, both interpolators, both colour schemes, power-of-two / mixed-radix / Bluestein / large paths."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import oracle
from tests import signals
from tests.emu import emu

LUT = signals.lut("magma")

pow2_small = st.sampled_from([256, 512, 1024, 2048])
generic_n = st.sampled_from([100, 120, 300, 441, 600, 1000, 1470, 2205, 2400])
bluestein_n = st.sampled_from([22, 94, 347, 551, 1102, 1103, 2731])     # a prime factor > 7 (2731 at P = 2: L = 8192)


@st.composite
def configs(draw):
    kind = draw(st.sampled_from(["fused", "fused", "generic", "large", "real", "bluestein"]))
    if kind == "fused":
        n, pad = draw(pow2_small), draw(st.sampled_from([1, 2, 4]))
    elif kind == "real":                 # mono, real transform length 8192 .. 32768: one frame per CTA (even hops)
        n, pad = draw(st.sampled_from([(2048, 4), (4096, 2), (8192, 1), (8192, 2), (4096, 4)]))
    elif kind == "generic":
        n, pad = draw(generic_n), draw(st.sampled_from([2, 4]))
    elif kind == "bluestein":
        n = draw(bluestein_n)
        pad = draw(st.sampled_from([1, 2, 4] if n % 2 == 0 else [2, 4]))     # M = P N must be even
    else:
        n, pad = 8192, 4
    sr = draw(st.sampled_from([8000.0, 22050.0, 44100.0, 48000.0, 96000.0]))
    hop = draw(st.one_of(st.integers(1, 2 * n), st.sampled_from([n // 8, n // 2, n])))
    rows = draw(st.sampled_from([1, 2, 7, 33, 64, 100, 257] if kind not in ("large", "real") else [16, 64]))
    f_min = draw(st.sampled_from([1.0, 20.0, 32.0, 500.0]))
    f_max = draw(st.sampled_from([f_min * 1.5, 4000.0, sr / 2, 22030.0, sr]))
    if f_max <= f_min:
        f_max = f_min * 2
    return dict(kind=kind, sample_rate=sr, window_n=n, pad_factor=pad, hop=max(1, hop), rows_h=rows, f_min=f_min, f_max=f_max,
                interp=draw(st.integers(0, 1)), scheme=draw(st.integers(0, 1)),
                channels=1 if kind == "real" else draw(st.integers(1, 2)),
                db=draw(st.sampled_from([(-70.0, -10.0, 1e-7), (-100.0, 0.0, 1e-10), (-30.0, -5.0, 1e-7)])),
                seed=draw(st.integers(0, 2 ** 16)), frames=draw(st.integers(1, 3)))


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(configs())
def test_random_configurations(c):
    n, hop, ch = c["window_n"], c["hop"], c["channels"]
    ns = n + hop * (c["frames"] - 1)
    rng = np.random.Generator(np.random.PCG64(c["seed"]))
    amp = 10.0 ** rng.uniform(-4, 0.3)                                   # from near-silence to clipping-loud
    t = np.arange(ns) / c["sample_rate"]
    tone = np.sin(2 * np.pi * rng.uniform(10.0, c["sample_rate"] / 2) * t)
    x = (amp * (0.7 * tone + 0.3 * rng.standard_normal(ns))).astype(np.float32)
    pcm = x if ch == 1 else np.stack([x, (0.5 * amp * rng.standard_normal(ns)).astype(np.float32)], 1)
    db_min, db_max, eps = c["db"]
    kw = dict(sample_rate=c["sample_rate"], window_n=n, pad_factor=c["pad_factor"], hop=hop, rows_h=c["rows_h"],
              f_min=c["f_min"], f_max=c["f_max"], db_min=db_min, db_max=db_max, db_eps=eps, interp=c["interp"])
    cfg = emu.make_config(LUT, channels=ch, scheme=c["scheme"], **kw)
    mags, rgba = emu.batch(cfg, pcm, real=True if c["kind"] == "real" else None)
    p = oracle.params(stereo_scheme=c["scheme"], **kw)
    o_rgba, o_mags = oracle.batch(p, pcm, ch, LUT, want_mags=True)
    assert mags.shape == o_mags.shape and rgba.shape == o_rgba.shape
    peak = np.maximum(np.abs(o_mags).max(axis=(1, 2), keepdims=True), 1e-30)
    assert (np.abs(mags - o_mags) / peak).max() <= 1e-5
    rec = np.stack([oracle.column_from_magnitudes(p, mags[f], LUT) for f in range(mags.shape[0])])
    assert np.array_equal(rec, rgba)                                     # identical magnitudes -> identical bytes


def test_silence_and_nan_paths():
    """All-zero input: mono scheme -> LUT[0]; stereo scheme -> pan = 0/0 = NaN -> index 0, alpha from t <= 0 -> 0.
    NaN / inf samples must not crash and must agree with the oracle's saturating casts."""
    for scheme in (0, 1):
        for ch in (1, 2):
            kw = dict(window_n=256, pad_factor=2, hop=64, rows_h=32)
            pcm = np.zeros(256 + 64 * 2, np.float32) if ch == 1 else np.zeros((256 + 64 * 2, 2), np.float32)
            cfg = emu.make_config(LUT, channels=ch, scheme=scheme, **kw)
            mags, rgba = emu.batch(cfg, pcm)
            p = oracle.params(stereo_scheme=scheme, **kw)
            o_rgba, _ = oracle.batch(p, pcm, ch, LUT)
            assert np.array_equal(rgba, o_rgba) and not mags.any()
            assert np.all(rgba[..., :3] == LUT[0]) and np.all(rgba[..., 3] == (255 if scheme == 0 else 0))
    bad = np.zeros((256, 2), np.float32)
    bad[100] = (np.nan, np.inf)
    cfg = emu.make_config(LUT, channels=2, scheme=0, window_n=256, pad_factor=2, hop=256, rows_h=16)
    mags, rgba = emu.batch(cfg, bad)
    assert np.isnan(mags).all()                                         # NaN spreads over the whole spectrum
    assert np.all(rgba[..., :3] == LUT[0])                              # NaN power -> step 0 (both here and in colorous' clamp)
