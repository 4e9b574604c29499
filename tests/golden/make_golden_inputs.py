"""Seeded input builders shared by make_golden.py-style cases and the parity tests (no reference
import needed)."""
import numpy as np


def adversarial_stack(rng, n, h, w):
    """uint8 stack that hits every rule branch, ties at +-eps, NaN pixels and saturation."""
    assert h >= 8
    f = 4 * n + 2
    st = rng.integers(0, 256, (f, h, w)).astype(np.uint8)
    st[:, 0] = 0                                             # white+black == 0 -> NaN
    st[1, 1] = 0                                             # white == 0, black > 0
    st[:, 2] = rng.integers(0, 4, (f, w))                    # dark noise
    st[:, 3] = rng.integers(250, 256, (f, w))                # saturated
    st[:, 4] = np.where(rng.random((f, w)) < .5, 255, 0)     # extremes
    base = rng.integers(0, 256, (1, w))
    st[2:, 5] = np.clip(base + rng.integers(-2, 3, (f - 2, w)), 0, 255)   # p = q +- {0,1,2}
    rows = slice(6, h)
    nr = h - 6
    st[0, rows] = rng.integers(0, 20, (nr, w))               # well-exposed pixels
    st[1, rows] = rng.integers(180, 256, (nr, w))
    bits = rng.integers(0, 2, (2 * n, nr, w))
    hi = rng.integers(150, 256, (2 * n, nr, w))
    lo = rng.integers(0, 40, (2 * n, nr, w))
    st[2:2 + 2 * n, rows] = np.where(bits == 1, hi, lo)
    st[2 + 2 * n:, rows] = np.where(bits == 1, lo, hi)
    return st
