#!/usr/bin/env python
"""Numerics of the planned 4-bit operand path, checked on the CPU with numpy.

A bf16 activation block of 32 consecutive K values is written as P rows of signed base-4 digits d in {0, +-1, +-2, +-3}
(all exactly representable in e2m1) with one power-of-two scale per (row, block), the form a block-scaled
tcgen05.mma.kind::mxf4 consumes:       x[k]  ~=  sum_p  d[p, k] * 2**(E - 2*p),     E = block exponent.
The sign matrix is exactly +-1 in e2m1 with scale 1.  This script measures the error of x . S against the exact product for
P = 4..9 digit rows on activation-like data with a wide dynamic range, i.e. what the delta MMAs would accumulate.
usage: python tools/experiments/mxf4_numerics.py
"""
import numpy as np


def bf16_round(x):
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def base4_digit_rows(x, P):
    """x: [..., 32] fp32 (bf16 values).  Returns digits [P, ..., 32] in {-3..3} and exponents E [...] with
    |x| < 4 * 2**E, so that sum_p digits[p] * 2**(E - 2p) -> x from below in magnitude (truncation)."""
    ax = np.abs(x).astype(np.float64)
    mx = ax.max(axis=-1)
    E = np.where(mx > 0, np.floor(np.log2(np.maximum(mx, 1e-300))) - 1, 0.0)  # max in [2, 4) * 2**E
    sgn = np.sign(x)
    rem = ax / np.exp2(E)[..., None]  # in [0, 4)
    digits = []
    for p in range(P):
        d = np.floor(rem)            # 0..3
        digits.append(d * sgn)
        rem = (rem - d) * 4.0
    return np.stack(digits), E


def reconstruct(digits, E):
    P = digits.shape[0]
    scale = np.exp2(E[None, ..., None] - 2.0 * np.arange(P).reshape((P,) + (1,) * (digits.ndim - 1)))
    return (digits * scale).sum(axis=0)


def main():
    rng = np.random.default_rng(0)
    K, N = 4096, 256
    # activations with outliers: log-normal magnitudes spanning ~2^12 inside a 32-block now and then
    x = rng.standard_normal(K) * np.exp(rng.standard_normal(K) * 1.5)
    x[rng.integers(0, K, 16)] *= 200.0
    x = bf16_round(x).astype(np.float64)
    S = rng.integers(0, 2, (K, N)) * 2.0 - 1.0
    exact = x @ S
    xb = x.reshape(K // 32, 32)
    print(f"K={K} N={N}; |exact| mean {np.abs(exact).mean():.3f}; bf16 half-ulp relative = {2.0**-9:.2e}")
    for P in range(4, 10):
        d, E = base4_digit_rows(xb, P)
        assert np.all(np.abs(d) <= 3) and np.all(d == np.round(d))
        xr = reconstruct(d, E).reshape(K)
        got = xr @ S
        el = np.abs(xr - x).max() / np.abs(x).max()
        rel = np.abs(got - exact).mean() / np.abs(exact).mean()
        worst = np.abs(got - exact).max() / np.abs(exact).mean()
        print(f"P={P}: max element error / max|x| = {el:.2e}   product mean-rel error = {rel:.2e}   worst / mean|exact| = {worst:.2e}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
