"""The kernels' shared per-quad byte-SWAR decode (csrc/sl_decode_quad.cuh), compiled for the host
by tests/hostcheck, against the oracle.  This checks the arithmetic the GPU executes on a machine
without a GPU; it is test infrastructure, not a product path."""
import ctypes
import os
import sys

import numpy as np
import pytest

from oracle import graycode_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hostlib():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    return ctypes.CDLL(ge.build_hostcheck())


def run(lib, st, e, codes=True):
    f, h, w = st.shape
    npix = h * w
    st = np.ascontiguousarray(st, dtype=np.uint8)
    n = (f - 2) // 4
    hi = np.empty(npix, np.int16); vi = np.empty(npix, np.int16)
    hc = np.empty((n, npix), np.int8); vc = np.empty((n, npix), np.int8)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.hostcheck_decode(p(st), f, ctypes.c_long(npix), e, p(hi), p(vi), p(hc) if codes else None, p(vc) if codes else None)
    assert rc == 0
    return hi.reshape(h, w), vi.reshape(h, w), hc.reshape(n, h, w), vc.reshape(n, h, w)


@pytest.mark.parametrize("n", [3, 7, 8, 9, 10, 11, 12, 15])
@pytest.mark.parametrize("e", [0, 1, 3, 127])
def test_swar_decode_bit_exact(hostlib, n, e):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden_inputs import adversarial_stack
    rng = np.random.default_rng(100 * n + e)
    st = adversarial_stack(rng, n, 24, 64)
    f64 = st.astype(np.float64)
    l_d, l_g = O.get_direct_indirect(f64)
    ho, vo = O.get_is_lit(f64, l_d, l_g, eps=e)
    hpo, vpo = O.codes_to_pixels(ho, vo)
    hi, vi, hc, vc = run(hostlib, st, e)
    np.testing.assert_array_equal(hc, ho)
    np.testing.assert_array_equal(vc, vo)
    np.testing.assert_array_equal(hi, hpo)
    np.testing.assert_array_equal(vi, vpo)
    hi2, vi2, _, _ = run(hostlib, st, e, codes=False)
    np.testing.assert_array_equal(hi2, hpo)
    np.testing.assert_array_equal(vi2, vpo)


def test_swar_exhaustive_thresholds(hostlib):
    """Every (white, black) pair with a sweep of L_max/L_min: the integer thresholds derived in
    fp64 reproduce the reference's fp64 comparisons for all 256 values of the compared byte."""
    n = 3
    f = 4 * n + 2
    rng = np.random.default_rng(7)
    w = 256 * 256
    st = np.zeros((f, 4, w), dtype=np.uint8)
    st[1] = np.repeat(np.arange(256), 256)[None, :]
    st[0] = np.tile(np.arange(256), 256)[None, :]
    for row in range(4):
        st[2:, row] = rng.integers(0, 256, (f - 2, w))
    f64 = st.astype(np.float64)
    ho, vo = O.get_codes(f64)
    hpo, vpo = O.codes_to_pixels(ho, vo)
    hi, vi, hc, vc = run(hostlib, st, 1)
    np.testing.assert_array_equal(hc, ho)
    np.testing.assert_array_equal(vc, vo)
    np.testing.assert_array_equal(hi, hpo)
    np.testing.assert_array_equal(vi, vpo)


def test_bgr_to_gray_quad_all_colours(hostlib):
    """gray4_from_bgr12 (byte dot products with split coefficients) on all 2^24 colours, in every one of the four
    pixel positions of a quad, against cv2's fixed-point formula (3735 B + 19235 G + 9798 R + 16384) >> 15."""
    c = np.arange(1 << 24, dtype=np.uint32)
    bgr = np.stack([c & 255, (c >> 8) & 255, c >> 16], axis=1).astype(np.uint8)
    b32 = bgr.astype(np.uint32)
    want = ((3735 * b32[:, 0] + 19235 * b32[:, 1] + 9798 * b32[:, 2] + 16384) >> 15).astype(np.uint8)
    for shift in range(4):                       # rotate so every colour visits every lane of the quad
        b = np.roll(bgr, shift, axis=0)
        got = np.empty(len(b), np.uint8)
        assert hostlib.hostcheck_bgr_to_gray(b.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(len(b)),
                                             got.ctypes.data_as(ctypes.c_void_p)) == 0
        np.testing.assert_array_equal(got, np.roll(want, shift))
