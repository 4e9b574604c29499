"""Property tests (hypothesis) of the CUDA decode against the oracle: random shapes, bit counts, eps values and
sample statistics, including ragged sizes.  Bit-exact."""
import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

from oracle import graycode_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu


@st.composite
def stacks(draw):
    n = draw(st.integers(3, 12))
    h = draw(st.integers(1, 40))
    w = draw(st.integers(1, 70))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    kind = draw(st.sampled_from(["uniform", "bimodal", "dark", "ties"]))
    rng = np.random.default_rng(seed)
    f = 4 * n + 2
    if kind == "uniform":
        s = rng.integers(0, 256, (f, h, w))
    elif kind == "dark":
        s = rng.integers(0, 14, (f, h, w))
    elif kind == "ties":
        base = rng.integers(0, 256, (1, h, w))
        s = np.clip(base + rng.integers(-2, 3, (f, h, w)), 0, 255)
    else:
        bits = rng.integers(0, 2, (2 * n, h, w))
        hi = rng.integers(120, 256, (2 * n, h, w)); lo = rng.integers(0, 60, (2 * n, h, w))
        s = np.empty((f, h, w), dtype=np.int64)
        s[0] = rng.integers(0, 30, (h, w)); s[1] = rng.integers(150, 256, (h, w))
        s[2:2 + 2 * n] = np.where(bits == 1, hi, lo); s[2 + 2 * n:] = np.where(bits == 1, lo, hi)
    eps = draw(st.sampled_from([0, 1, 1, 1, 2, 9, 0.5, 130]))
    return s.astype(np.uint8), eps


@settings(max_examples=40, deadline=None)
@given(stacks())
def test_decode_matches_oracle(case):
    import torch
    import slb200
    st_, eps = case
    f64 = st_.astype(np.float64)
    l_d, l_g = O.get_direct_indirect(f64)
    h, v = O.get_is_lit(f64, l_d, l_g, eps=eps)
    hp, vp = O.codes_to_pixels(h, v)
    dev = torch.from_numpy(st_).cuda()
    hc, vc, hi, vi = slb200.decode_codes(dev, eps=eps)
    np.testing.assert_array_equal(hc.cpu().numpy(), h)
    np.testing.assert_array_equal(vc.cpu().numpy(), v)
    np.testing.assert_array_equal(hi.cpu().numpy(), hp)
    hi2, vi2 = slb200.decode_pixels(dev, eps=eps)          # index-only kernel (fast pass + fallback)
    np.testing.assert_array_equal(hi2.cpu().numpy(), hp)
    np.testing.assert_array_equal(vi2.cpu().numpy(), vp)
