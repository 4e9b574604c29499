"""The oracle (oracle/graycode_oracle.py) against golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import graycode_oracle as O


@pytest.fixture(scope="module")
def dec(golden_dir):
    return np.load(os.path.join(golden_dir, "decode_golden.npz"))


@pytest.fixture(scope="module")
def tri(golden_dir):
    return np.load(os.path.join(golden_dir, "triangulate_golden.npz"))


@pytest.mark.parametrize("name", ["n3", "n5", "n10", "n11", "n12"])
def test_decode_matches_reference(dec, name):
    stacks = dec[f"decode_{name}_stacks"]
    hs, vs = [], []
    for r, st in enumerate(stacks):
        imgs = st.astype(np.float64)
        l_d, l_g = O.get_direct_indirect(imgs)
        np.testing.assert_array_equal(l_d, dec[f"decode_{name}_L_d"][r])     # NaN-aware, bit exact
        np.testing.assert_array_equal(l_g, dec[f"decode_{name}_L_g"][r])
        h, v = O.get_codes(imgs)
        assert h.dtype == np.int8 and v.dtype == np.int8
        np.testing.assert_array_equal(h, dec[f"decode_{name}_h_codes"][r])
        np.testing.assert_array_equal(v, dec[f"decode_{name}_v_codes"][r])
        hs.append(h); vs.append(v)
    hp, vp = O.codes_to_pixels(O.merge_runs(hs), O.merge_runs(vs))
    np.testing.assert_array_equal(hp, dec[f"decode_{name}_h_pixels"])
    np.testing.assert_array_equal(vp, dec[f"decode_{name}_v_pixels"])
    hp2, vp2 = O.decode_pixels(None, runs=[s.astype(np.float64) for s in stacks])
    np.testing.assert_array_equal(hp2, hp)
    np.testing.assert_array_equal(vp2, vp)
    # the scalar helper agrees with the vectorised form
    n = hs[0].shape[0]
    best_h = O.merge_runs(hs)
    for (y, x) in [(0, 0), (3, 5), (7, 9), (hp.shape[0] - 1, hp.shape[1] - 1)]:
        assert O.gray_to_decimal(best_h[:, y, x]) == hp[y, x]
    assert n == (stacks.shape[1] - 2) // 4


@pytest.mark.parametrize("eps", [0, 2, 0.5, 3.25])
def test_is_lit_eps(dec, eps):
    st = dec["eps_stack"].astype(np.float64)
    l_d, l_g = O.get_direct_indirect(st)
    h, v = O.get_is_lit(st, l_d, l_g, eps=eps)
    np.testing.assert_array_equal(h, dec[f"eps_{eps}_h"])
    np.testing.assert_array_equal(v, dec[f"eps_{eps}_v"])


def test_bad_frame_counts():
    for f in (1, 6, 13, 15, 16, 43):
        with pytest.raises(ValueError):
            O.get_direct_indirect(np.zeros((f, 2, 2)))


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_triangulation_matches_reference(tri, name):
    g = lambda k: tri[f"{name}_{k}"]
    proj_size = tuple(int(x) for x in g("proj_size"))
    calib_size = tuple(int(x) for x in g("calib_size"))
    cam, proj, colors = O.cam_proj_pairs(g("h_pixels"), g("v_pixels"), proj_size, g("white"))
    assert cam.dtype == np.float32 and proj.dtype == np.float32 and colors.dtype == np.float64
    np.testing.assert_array_equal(cam, g("cam_pts"))
    np.testing.assert_array_equal(proj, g("proj_pts"))
    np.testing.assert_array_equal(colors, g("colors"))
    pm = O.scale_proj_mtx(g("proj_mtx_in"), proj_size, calib_size)
    np.testing.assert_array_equal(pm, g("proj_mtx_scaled"))
    np.testing.assert_array_equal(O.undistort_points(g("cam_full"), g("cam_mtx"), g("cam_dist"), R=tri["R"]), g("und_cam"))
    np.testing.assert_array_equal(O.undistort_points(proj, pm, g("proj_dist")), g("und_proj"))
    pts = O.triangulate(g("cam_full"), proj, g("cam_mtx"), g("cam_dist"), pm, g("proj_dist"), tri["R"], tri["T"])
    assert pts.dtype == np.float64 and pts.shape == g("pts").shape
    np.testing.assert_allclose(pts, g("pts"), rtol=1e-12, atol=0, equal_nan=True)   # BLAS-order ulps only
    pts_small = O.triangulate(g("cam_pts"), proj, g("cam_mtx"), g("cam_dist"), pm, g("proj_dist"), tri["R"], tri["T"])
    np.testing.assert_allclose(pts_small, g("pts_small"), rtol=1e-12, atol=0, equal_nan=True)
    for thr, key in ((.5, "filt"), (1.25, "filt125")):
        fp, fc = O.filter_3d_pts(g("pts"), g("colors"), thr)
        np.testing.assert_array_equal(fp, g(f"{key}_pts"))
        np.testing.assert_array_equal(fc, g(f"{key}_colors"))


def test_ideal_sequence_decodes_to_identity(dec):
    """Known-answer property: the rendered pattern sequence decodes to the pixel coordinates."""
    for key in ("patterns_16x8", "patterns_32x32", "patterns_24x10"):
        seq = dec[key].astype(np.float64)
        hp, vp = O.decode_pixels(seq)
        h, w = hp.shape
        np.testing.assert_array_equal(hp, np.tile(np.arange(w), (h, 1)))
        np.testing.assert_array_equal(vp, np.tile(np.arange(h)[:, None], (1, w)))
