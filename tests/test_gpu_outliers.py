"""GPU parity of the statistical outlier removal (sl_remove_statistical_outliers, through the C ABI) against
oracle/outlier_oracle.py.  Per-point mean neighbour distances are compared BIT-EXACTLY; the cloud mean /
standard deviation / threshold come from tree reductions on the GPU and a sequential sum in the oracle, so they
are compared to 1e-12 relative and the inlier sets may differ only at points whose mean distance lies within
1e-12 relative of the threshold."""
import os

import numpy as np
import pytest

from oracle import outlier_oracle as OO

pytestmark = pytest.mark.gpu


def scene(rng, n_surface, n_noise, dup=0):
    xy = rng.uniform(-0.15, 0.15, (n_surface, 2))
    z = 0.6 + 0.1 * xy[:, 0] + 0.02 * np.sin(20 * xy[:, 1])
    pts = np.r_[np.c_[xy, z], rng.uniform(-0.5, 0.5, (n_noise, 3))]
    if dup:
        pts = np.r_[pts, pts[:dup]]
    return pts.astype(np.float32)


def run_gpu(xyz, k=20, ratio=0.5):
    import torch
    import slb200
    cloud = torch.zeros((len(xyz), 4), dtype=torch.float32, device="cuda")
    if len(xyz):
        cloud[:, :3] = torch.from_numpy(np.ascontiguousarray(xyz))
        cloud[:, 3] = torch.arange(len(xyz), dtype=torch.float32, device="cuda")     # payload travels with the record
    out, ind, mean_dist, stats = slb200.remove_statistical_outliers(cloud, k, ratio, details=True)
    torch.cuda.synchronize()
    return out.cpu().numpy(), ind.cpu().numpy(), mean_dist.cpu().numpy(), stats.cpu().numpy()


def check(xyz, k=20, ratio=0.5):
    out, ind, mean_dist, stats = run_gpu(xyz, k, ratio)
    want_ind, want_avg, want_thr = OO.statistical_outliers(xyz, k, ratio)
    np.testing.assert_array_equal(mean_dist, want_avg)                     # bit-exact
    if np.isnan(want_thr):
        assert np.isnan(stats[2]) and len(ind) == 0
        return
    assert stats[2] == pytest.approx(want_thr, rel=1e-12)
    diff = np.setxor1d(ind, want_ind)
    assert np.all(np.abs(want_avg[diff] - want_thr) <= 1e-12 * abs(want_thr)), "inlier sets differ away from the threshold"
    assert np.all(np.diff(ind) > 0)                                        # ascending, like Open3D's `ind`
    np.testing.assert_array_equal(out[:, :3], xyz[ind])
    np.testing.assert_array_equal(out[:, 3], ind.astype(np.float32))


@pytest.mark.parametrize("n", [1, 2, 5, 19, 20, 21, 25, 300])
def test_small_clouds(n):
    check(np.random.default_rng(n).normal(size=(n, 3)).astype(np.float32))


def test_empty_cloud():
    out, ind, mean_dist, stats = run_gpu(np.zeros((0, 3), np.float32))
    assert len(out) == 0 and len(ind) == 0


def test_surface_with_noise_and_duplicates():
    check(scene(np.random.default_rng(10), 3000, 60, dup=5))


@pytest.mark.parametrize("k,ratio", [(1, 0.5), (2, 1.0), (8, 2.0), (33, 0.5), (128, 0.25)])
def test_neighbour_counts(k, ratio):
    check(scene(np.random.default_rng(11 + k), 2000, 30), k, ratio)


def test_degenerate_geometry():
    rng = np.random.default_rng(12)
    check(np.zeros((50, 3), np.float32))                                               # one location
    check(np.c_[rng.uniform(0, 1, 400), np.zeros(400), np.zeros(400)].astype(np.float32))   # a line
    check(np.c_[rng.uniform(0, 1, (900, 2)), np.full(900, 0.25)].astype(np.float32))        # a plane
    check((rng.integers(0, 6, (2000, 3)) * 0.125).astype(np.float32))                  # lattice: many exact ties
    check(np.r_[scene(rng, 500, 0), [[1e6, -1e6, 1e6]]].astype(np.float32))            # one far point stretches the box


def test_far_queries_against_a_dense_cluster():
    """Sparse far points whose neighbours are a big dense cluster: they are handled one CTA per query, with the
    threshold / candidate-list pass; 6000 coincident points make thousands of exact ties at the threshold (list
    overflow -> the per-thread-list fallback), the jittered cluster takes the list path."""
    rng = np.random.default_rng(15)
    far = rng.uniform(-3, 3, (40, 3))
    check(np.r_[np.zeros((6000, 3)), far].astype(np.float32))
    check(np.r_[rng.normal(0, 0.01, (9000, 3)), far].astype(np.float32))
    check(np.r_[np.repeat(rng.normal(0, 0.01, (30, 3)), 300, axis=0), far].astype(np.float32))   # 30 x 300 duplicates


def test_non_finite_records_are_dropped():
    rng = np.random.default_rng(13)
    xyz = scene(rng, 1500, 20)
    bad = xyz.copy()
    bad[[3, 700, 1519]] = [[np.nan, 0, 0], [0, np.inf, 0], [0, 0, -np.inf]]
    out, ind, mean_dist, stats = run_gpu(bad)
    good = np.ones(len(bad), bool)
    good[[3, 700, 1519]] = False
    want_ind, want_avg, want_thr = OO.statistical_outliers(bad[good], 20, 0.5)
    np.testing.assert_array_equal(mean_dist[good], want_avg)
    assert (mean_dist[~good] == -1).all()
    assert stats[2] == pytest.approx(want_thr, rel=1e-12)
    np.testing.assert_array_equal(ind, np.nonzero(good)[0][want_ind])


def test_scan_cloud_large():
    """The clean-up on a real pipeline output: a 720p synthetic scan -> box filter -> outlier removal
    (src/4-triangulate.py:71-74), ~0.8 M points against the KD-tree oracle."""
    import torch
    import slb200
    from slb200 import synth
    calib = synth.synthetic_calibration((1280, 720), (1024, 768))
    frames, white = synth.render_scan((1280, 720), (1024, 768), seed=5, device="cuda", calib=calib)
    bufs = slb200.scan_to_cloud(frames, white, calib)
    n = int(bufs.counts[1].item())
    cloud = slb200.filter_cloud_box(bufs.cloud[:n], 2.0)          # the synthetic plane stands 0.9 m away
    assert len(cloud) > 200_000
    out, ind, mean_dist, stats = slb200.remove_statistical_outliers(cloud, 20, 0.5, details=True)
    xyz = cloud[:, :3].cpu().numpy()
    want_ind, want_avg, want_thr = OO.statistical_outliers(xyz, 20, 0.5)
    np.testing.assert_array_equal(mean_dist.cpu().numpy(), want_avg)
    assert float(stats[2]) == pytest.approx(want_thr, rel=1e-12)
    diff = np.setxor1d(ind.cpu().numpy(), want_ind)
    assert np.all(np.abs(want_avg[diff] - want_thr) <= 1e-12 * want_thr)
    torch.testing.assert_close(out, cloud[ind], rtol=0, atol=0)


def test_plot_point_cloud_writes_ply(tmp_path, capsys):
    """scanner.utils.visualize.plot_point_cloud (reference visualize.py:91-129) end to end, headless."""
    from scanner.utils import visualize
    rng = np.random.default_rng(14)
    xyz = scene(rng, 4000, 80).astype(np.float64) + 1e-9         # not float32-representable: the mirror narrows
    colors = rng.uniform(0, 1, (len(xyz), 3))
    visualize.plot_point_cloud(xyz.T, colors, str(tmp_path))
    want_ind, _, _ = OO.statistical_outliers(xyz.astype(np.float32), 20, 0.5)
    raw = open(os.path.join(tmp_path, "cloud.ply"), "rb").read()
    assert raw == OO.ply_bytes(xyz.astype(np.float32)[want_ind].astype(np.float64), colors[want_ind])


def test_argument_errors():
    import torch
    import slb200
    cloud = torch.zeros((10, 4), dtype=torch.float32, device="cuda")
    with pytest.raises(ValueError):
        slb200.remove_statistical_outliers(cloud, 0, 0.5)
    with pytest.raises(ValueError):
        slb200.remove_statistical_outliers(cloud, 20, 0.0)
    with pytest.raises(slb200.SlError):
        slb200.remove_statistical_outliers(cloud, 129, 0.5)
    with pytest.raises(TypeError):
        slb200.remove_statistical_outliers(cloud[:, :3], 20, 0.5)


def test_full_size_cloud_properties():
    """Size-independent properties on the cloud of a 1080p scan (1.8 M points, BASELINE config 2), where the oracle
    would take minutes: (1) the per-point mean distances do not depend on the order of the records (the kernel
    sorts points into cells with atomics, every run in a different order): a shuffled copy gives bit-identical
    values; (2) same for a repeated run; (3) the inliers are exactly {0 < mean < threshold}, in ascending order;
    (4) the threshold is mean + ratio * std of the positive means (float64 on the host, 1e-12)."""
    import torch
    import slb200
    from slb200 import synth
    calib = synth.synthetic_calibration((1920, 1080), (1920, 1080))
    frames, white = synth.render_scan((1920, 1080), (1920, 1080), seed=7, device="cuda", calib=calib)
    bufs = slb200.scan_to_cloud(frames, white, calib)
    cloud = slb200.filter_cloud_box(bufs.cloud[: int(bufs.counts[1].item())], 2.0)
    n = len(cloud)
    assert n > 1_500_000
    out, ind, md, stats = slb200.remove_statistical_outliers(cloud, 20, 0.5, details=True)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    perm = torch.randperm(n, device="cuda", generator=g)
    out2, ind2, md2, stats2 = slb200.remove_statistical_outliers(cloud[perm].contiguous(), 20, 0.5, details=True)
    assert torch.equal(md2, md[perm])                                  # (1)
    _, ind3, md3, _ = slb200.remove_statistical_outliers(cloud, 20, 0.5, details=True)
    assert torch.equal(md3, md) and torch.equal(ind3, ind)             # (2)
    keep = (md > 0) & (md < stats[2])
    assert torch.equal(ind, torch.nonzero(keep).flatten())             # (3)
    assert torch.equal(out, cloud[ind])
    m = md.cpu().numpy()
    pos = m[m > 0]
    mean = pos.sum() / n
    std = np.sqrt(((pos - mean) ** 2).sum() / (n - 1))
    assert float(stats[0]) == pytest.approx(mean, rel=1e-12) and float(stats[1]) == pytest.approx(std, rel=1e-10)
    assert float(stats[2]) == pytest.approx(mean + 0.5 * std, rel=1e-11)   # (4)
    assert torch.equal(torch.sort(perm[ind2]).values, ind)             # the same points survive in the shuffled copy
