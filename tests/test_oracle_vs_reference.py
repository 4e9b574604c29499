"""Oracle and host-side mirrors against the reference imported live from /root/reference.
Skipped where the reference tree is not mounted (the GPU box); the golden vectors cover that case."""
import numpy as np
import pytest

from oracle import graycode_oracle as O


def test_get_codes_random(reference_modules):
    ref = reference_modules["decode"]
    rng = np.random.default_rng(5)
    for n in (3, 6, 11):
        st = rng.integers(0, 256, (4 * n + 2, 20, 28)).astype(np.float64)
        st[:, :2] = 0
        with np.errstate(all="ignore"):
            h, v = ref.get_codes(st)
            l_d, l_g = ref.get_direct_indirect(st)
        ho, vo = O.get_codes(st)
        np.testing.assert_array_equal(h, ho)
        np.testing.assert_array_equal(v, vo)
        ldo, lgo = O.get_direct_indirect(st)
        np.testing.assert_array_equal(l_d, ldo)
        np.testing.assert_array_equal(l_g, lgo)
        hp = np.array([ref.gray_to_decimal(h[:, y, x]) for y in range(20) for x in range(28)]).reshape(20, 28)
        vp = np.array([ref.gray_to_decimal(np.flip(v[:, y, x])) for y in range(20) for x in range(28)]).reshape(20, 28)
        hpo, vpo = O.codes_to_pixels(ho, vo)
        np.testing.assert_array_equal(hp, hpo)
        np.testing.assert_array_equal(vp, vpo)


def test_gray_helpers(reference_modules):
    ref = reference_modules["decode"]
    from scanner.grayCode import decode_codes as mine
    for n in list(range(0, 70)) + [1023, 2047, 4095, 65535]:
        assert ref.gray_decode(n) == O.gray_decode(n) == mine.gray_decode(n)
    rng = np.random.default_rng(1)
    for _ in range(200):
        bits = rng.integers(-1, 2, rng.integers(1, 14)).tolist()
        assert ref.gray_to_decimal(bits) == O.gray_to_decimal(bits) == mine.gray_to_decimal(bits)


@pytest.mark.parametrize("size", [(16, 8), (32, 32), (40, 24), (8, 16), (64, 48), (1024, 768)])
def test_pattern_generator(reference_modules, size):
    ref = reference_modules["generate"]
    from scanner.grayCode import generate_codes as mine
    pw, ph = size
    np.testing.assert_array_equal(ref.get_gray_codes(pw, ph), mine.get_gray_codes(pw, ph))
    a = ref.get_image_sequence(ref.get_gray_codes(pw, ph), pw, ph)
    b = mine.get_image_sequence(mine.get_gray_codes(pw, ph), pw, ph)
    assert a.dtype == b.dtype and a.shape == b.shape
    np.testing.assert_array_equal(a, b)


def test_triangulate_class(reference_modules):
    import cv2
    ref = reference_modules["triangulate"]
    from slb200 import calib_data as cd
    rng = np.random.default_rng(2)
    R, _ = cv2.Rodrigues(np.array([0.02, -0.15, 0.01]))
    T = np.array([[0.2], [0.02], [0.03]])
    h, w = 30, 44
    hp = rng.integers(-1, 1100, (h, w)); vp = rng.integers(-1, 900, (h, w))
    white = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
    pm = cd.PROJ_MTX.copy()
    t = ref.Triangulate(hp, vp, (w, h), cd.CAM_1080_MTX.copy(), cd.CAM_1080_DIST, (1024, 768), (1024, 768), pm,
                        cd.PROJ_DIST, R, T)
    cam, proj, col = t.get_cam_proj_pts(white)
    co, po, colo = O.cam_proj_pairs(hp, vp, (1024, 768), white)
    np.testing.assert_array_equal(cam, co)
    np.testing.assert_array_equal(proj, po)
    np.testing.assert_array_equal(col, colo)
    cam = cam * np.array([[40, 30]], dtype=np.float32)
    pts = t.triangulate(cam, proj)
    pto = O.triangulate(cam, proj, cd.CAM_1080_MTX, cd.CAM_1080_DIST, pm, cd.PROJ_DIST, R, T)
    np.testing.assert_allclose(pts, pto, rtol=1e-12, equal_nan=True)
    # synthetic-scene Rodrigues matches cv2
    from slb200 import synth
    np.testing.assert_allclose(synth.rodrigues([0.02, -0.15, 0.01]), R, atol=1e-15)
