"""CPU oracle for the point-cloud clean-up that follows the hot path.

TEST INFRASTRUCTURE ONLY (same rules as ``graycode_oracle.py``: imported by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs, never by the product).

Parity status: **UNPINNED**.  The reference calls
``pcd.remove_statistical_outlier(nb_neighbors=20, std_ratio=0.5)`` and
``o3d.io.write_point_cloud`` (scanner/utils/visualize.py:104-113).  Both live in the
third-party package ``open3d==0.17.0`` (requirements.txt, not vendored in the reference, and
not installable here: no network, no wheel).  The reference has no test, golden vector or
fixture for this step.  What follows restates Open3D's published algorithm
(``PointCloud::RemoveStatisticalOutliers``, cpp/open3d/geometry/PointCloud.cpp in 0.17):

* a KD-tree k-NN query of every point against the cloud, k = ``nb_neighbors``; the query point
  itself is one of the neighbours (distance 0); squared distances accumulate x, y, z in order in
  float64; fewer than k results when the cloud has fewer than k points;
* ``avg[i]`` = mean of the square roots of those distances, summed nearest first;
* ``cloud_mean`` = sum of the positive ``avg`` / number of points; ``std_dev`` =
  sqrt(sum over positive ``avg`` of (avg - cloud_mean)^2 / (points - 1)) (Bessel);
* inliers: ``0 < avg[i] < cloud_mean + std_ratio * std_dev``, ascending index order.

The reference first narrows the points to float32 (``o3c.Tensor(Pts.T, o3c.float32)``,
visualize.py:100) -- ``statistical_outliers`` therefore takes float32 coordinates.
"""
from __future__ import annotations

import numpy as np

__all__ = ["knn_mean_distances", "statistical_outliers", "ply_bytes"]


def _sq_dists(q, pts):
    """((dx*dx + dy*dy) + dz*dz) in float64, one rounding per operation (no FMA)."""
    d = pts.astype(np.float64) - q.astype(np.float64)[None, :]
    return (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]


def knn_mean_distances(xyz32, nb_neighbors=20, brute_limit=4096):
    """avg[i] of the header; exact brute force for small clouds, scipy's cKDTree otherwise
    (candidates from the tree, distances and summation order recomputed here)."""
    pts = np.ascontiguousarray(np.asarray(xyz32, dtype=np.float32))
    n = len(pts)
    k = min(int(nb_neighbors), n)
    avg = np.full(n, -1.0)
    if n == 0 or k < 1:
        return avg
    if n <= brute_limit:
        for i in range(n):
            d2 = np.sort(_sq_dists(pts[i], pts))[:k]
            s = 0.0
            for v in np.sqrt(d2):
                s += float(v)
            avg[i] = s / k
        return avg
    from scipy.spatial import cKDTree
    p64 = pts.astype(np.float64)
    tree = cKDTree(p64)
    kk = min(n, k + 4)                     # a few spares: ties at the k-th distance
    _, idx = tree.query(p64, k=kk)
    idx = idx.reshape(n, kk)
    nb = p64[idx] - p64[:, None, :]
    d2 = (nb[..., 0] * nb[..., 0] + nb[..., 1] * nb[..., 1]) + nb[..., 2] * nb[..., 2]
    d = np.sqrt(np.sort(d2, axis=1)[:, :k])
    s = np.zeros(n)
    for j in range(k):                     # nearest first, one rounding per add
        s = s + d[:, j]
    return s / k


def statistical_outliers(xyz32, nb_neighbors=20, std_ratio=0.5):
    """Returns ``(inlier_index int64 ascending, avg float64 (N,), threshold float)``."""
    if nb_neighbors < 1 or std_ratio <= 0:
        raise ValueError("nb_neighbors and std_ratio must be positive")
    avg = knn_mean_distances(xyz32, nb_neighbors)
    n = len(avg)
    if n == 0:
        return np.zeros(0, np.int64), avg, float("nan")
    cloud_mean = 0.0
    for v in avg[avg > 0]:
        cloud_mean += float(v)
    cloud_mean /= n
    sq = 0.0
    for v in avg[avg > 0]:
        sq += (float(v) - cloud_mean) * (float(v) - cloud_mean)
    with np.errstate(divide="ignore", invalid="ignore"):
        std_dev = float(np.sqrt(np.float64(sq) / np.float64(n - 1)))
    thr = cloud_mean + std_ratio * std_dev
    keep = (avg > 0) & (avg < thr)
    return np.nonzero(keep)[0].astype(np.int64), avg, thr


def ply_bytes(xyz, colors=None):
    """Binary little-endian PLY as Open3D's ``write_point_cloud`` lays out a legacy point cloud:
    float64 x/y/z, then uint8 red/green/blue = round(clamp(c, 0, 1) * 255) when colours are given."""
    xyz = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
    n = len(xyz)
    head = ["ply", "format binary_little_endian 1.0", "comment Created by Open3D", f"element vertex {n}",
            "property double x", "property double y", "property double z"]
    fields = [("x", "<f8"), ("y", "<f8"), ("z", "<f8")]
    if colors is not None:
        head += ["property uchar red", "property uchar green", "property uchar blue"]
        fields += [("r", "u1"), ("g", "u1"), ("b", "u1")]
    head.append("end_header")
    rec = np.zeros(n, dtype=np.dtype(fields))
    rec["x"], rec["y"], rec["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    if colors is not None:
        c = np.rint(np.clip(np.asarray(colors, dtype=np.float64).reshape(-1, 3), 0.0, 1.0) * 255.0).astype(np.uint8)
        rec["r"], rec["g"], rec["b"] = c[:, 0], c[:, 1], c[:, 2]
    return ("\n".join(head) + "\n").encode("ascii") + rec.tobytes()
