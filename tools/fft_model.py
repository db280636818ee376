"""numpy model of the kernel's FFT index math (development aid; run on CPU, no GPU needed).

Models exactly what spectrogram-rs_b200/csrc/stft_kernels.cuh does:
  M = P*N point transform of a sequence whose last (P-1)*N entries are zero
  = P sub-transforms of length N over x[n]*W_M^(n*q0)            (radix-P DIF step, pruned)
  each sub-transform: in-place radix-(R1..RL) DIF, natural in, digit-reversed out
  bin kk = P*k + q0 lives at position q0*N + digitrev(k)
and checks it against numpy.fft.fft.
"""
import numpy as np


def radix_plan(n):
    lg = n.bit_length() - 1
    a, rem = divmod(lg, 4)
    return ([1 << rem] if rem else []) + [16] * a


def dif_inplace(buf, n, radices):
    """buf: P sub-blocks of n; in-place DIF on each block."""
    m = len(buf)
    stride = n
    for p, r in enumerate(radices):
        stride //= r
        last = p == len(radices) - 1
        span = stride * r
        for blk in range(0, m, n):
            for u in range(n // r):
                lo, hi = u % stride, u // stride
                idx = blk + hi * span + lo + stride * np.arange(r)
                v = buf[idx]
                y = np.array([np.sum(v * np.exp(-2j * np.pi * np.arange(r) * q / r)) for q in range(r)])
                if not last:
                    y = y * np.exp(-2j * np.pi * lo * np.arange(r) / span)
                buf[idx] = y
    return buf


def pos_of_bin(kk, n, pad, radices):
    q0, k = kk % pad, kk // pad
    pos, stride = q0 * n, n
    for r in radices:
        stride //= r
        pos += (k % r) * stride
        k //= r
    return pos


def bin_of_pos(pos, n, pad, radices):
    q0, rem = pos // n, pos % n
    k, mult, stride = 0, 1, n
    for r in radices:
        stride //= r
        d = rem // stride
        rem %= stride
        k += d * mult
        mult *= r
    return k * pad + q0


def model_fft(z, pad):
    n = len(z)
    m = n * pad
    radices = radix_plan(n)
    buf = np.empty(m, dtype=np.complex128)
    for q0 in range(pad):
        buf[q0 * n:(q0 + 1) * n] = z * np.exp(-2j * np.pi * np.arange(n) * q0 / m)
    dif_inplace(buf, n, radices)
    out = np.empty(m, dtype=np.complex128)
    for kk in range(m):
        p = pos_of_bin(kk, n, pad, radices)
        assert bin_of_pos(p, n, pad, radices) == kk
        out[kk] = buf[p]
    return out


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n in (16, 32, 64, 128, 256, 512, 1024):
        for pad in (1, 2, 4):
            z = rng.standard_normal(n) + 1j * rng.standard_normal(n)
            ref = np.fft.fft(np.concatenate([z, np.zeros(n * (pad - 1))]))
            got = model_fft(z, pad)
            err = np.abs(got - ref).max() / np.abs(ref).max()
            print(n, pad, radix_plan(n), f"{err:.2e}")
            assert err < 1e-12
