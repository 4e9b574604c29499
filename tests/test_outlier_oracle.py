"""CPU checks of the cloud clean-up oracle (oracle/outlier_oracle.py; parity UNPINNED: Open3D is not installable
here, see its header) and of the host-side mirrors in scanner.utils.visualize."""
import os

import numpy as np
import pytest

from oracle import outlier_oracle as OO


def scene(rng, n_surface, n_noise, dup=0):
    xy = rng.uniform(-0.15, 0.15, (n_surface, 2))
    z = 0.6 + 0.1 * xy[:, 0] + 0.02 * np.sin(20 * xy[:, 1])
    pts = np.r_[np.c_[xy, z], rng.uniform(-0.5, 0.5, (n_noise, 3))]
    if dup:
        pts = np.r_[pts, pts[:dup]]
    return pts.astype(np.float32)


def test_tree_and_brute_force_agree_bitwise():
    pts = scene(np.random.default_rng(1), 6000, 80, dup=7)
    a = OO.knn_mean_distances(pts, 20, brute_limit=1 << 30)
    b = OO.knn_mean_distances(pts, 20, brute_limit=0)
    np.testing.assert_array_equal(a, b)


def test_known_answers():
    # 3 collinear points, k = 2: neighbours are the point itself and the nearer one
    pts = np.array([[0, 0, 0], [1, 0, 0], [3, 0, 0]], np.float32)
    np.testing.assert_array_equal(OO.knn_mean_distances(pts, 2), [0.5, 0.5, 1.0])
    # fewer points than neighbours: mean over what exists
    np.testing.assert_array_equal(OO.knn_mean_distances(pts, 20), [4 / 3, 1.0, 5 / 3])
    ind, avg, thr = OO.statistical_outliers(pts, 2, 0.5)
    # mean = 2/3, std = sqrt(((1/6)^2 * 2 + (1/3)^2) / 2)
    assert thr == pytest.approx(2 / 3 + 0.5 * np.sqrt((2 / 36 + 1 / 9) / 2), rel=1e-15)
    np.testing.assert_array_equal(ind, [0, 1])
    # coincident points have a zero mean distance: Open3D keeps 0 < avg only
    ind, avg, _ = OO.statistical_outliers(np.zeros((30, 3), np.float32), 20, 0.5)
    assert len(ind) == 0 and (avg == 0).all()
    # one point: std over (n - 1) = 0 is NaN, nothing passes
    ind, _, thr = OO.statistical_outliers(np.ones((1, 3), np.float32), 20, 0.5)
    assert len(ind) == 0 and np.isnan(thr)
    with pytest.raises(ValueError):
        OO.statistical_outliers(pts, 0, 0.5)


def test_noise_is_removed_surface_is_kept():
    rng = np.random.default_rng(2)
    pts = scene(rng, 5000, 100)
    ind, avg, thr = OO.statistical_outliers(pts, 20, 0.5)
    keep = np.zeros(len(pts), bool)
    keep[ind] = True
    assert keep[:5000].mean() > 0.9 and keep[5000:].mean() < 0.05


def test_ply_layout_roundtrip(tmp_path):
    import sys
    from scanner.utils import visualize
    rng = np.random.default_rng(3)
    xyz = rng.normal(size=(17, 3))
    col = rng.uniform(-0.1, 1.1, (17, 3))
    path = os.path.join(tmp_path, "cloud.ply")
    visualize.write_ply(path, xyz, col)
    raw = open(path, "rb").read()
    assert raw == OO.ply_bytes(xyz, col)
    head, body = raw.split(b"end_header\n")
    assert b"element vertex 17" in head and len(body) == 17 * 27
    rec = np.frombuffer(body, dtype=[("p", "<f8", 3), ("c", "u1", 3)])
    np.testing.assert_array_equal(rec["p"], xyz)
    np.testing.assert_array_equal(rec["c"], np.rint(np.clip(col, 0, 1) * 255).astype(np.uint8))


def test_color_map_and_preview_match_reference():
    """scanner.utils.visualize.create_color_map / the plot_decoded_graycodes image against the reference's code
    (its module cannot be imported: matplotlib / open3d are absent, so the two functions are exec'd from source)."""
    ref_path = "/root/reference/scanner/utils/visualize.py"
    if not os.path.exists(ref_path):
        pytest.skip("reference tree not mounted")
    import ast
    from scanner.utils import visualize
    tree = ast.parse(open(ref_path).read())
    ns = {"np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("arr_creat", "create_color_map"):
            exec(compile(ast.Module([node], []), ref_path, "exec"), ns)
    for w, h in [(16, 8), (64, 48), (7, 5)]:
        ref = ns["create_color_map"](w, h)
        got = visualize.create_color_map(w, h)
        assert got.dtype == ref.dtype and got.shape == ref.shape
        np.testing.assert_array_equal(got, ref)
    # the preview image, restated from visualize.py:75-85 on a small case
    rng = np.random.default_rng(4)
    cmap = visualize.create_color_map(16, 8)
    hp = rng.integers(-1, 20, (6, 9))
    vp = rng.integers(-1, 12, (6, 9))
    want = np.zeros((6, 9, 3), dtype=int)
    for i in range(9):
        for j in range(6):
            if hp[j, i] != -1 and vp[j, i] != -1:
                want[j, i] = cmap[min(7, vp[j, i]), min(15, hp[j, i])]
    np.testing.assert_array_equal(visualize.decoded_graycodes_image(hp, vp, cmap), want)
