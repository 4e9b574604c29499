"""Drop-in for the reference's scanner/utils/visualize.py.

The viewers themselves (matplotlib / Open3D windows) are outside the hot path and are imported lazily so
``from scanner.utils import visualize`` (src/3-capture_decode.py:8, src/4-triangulate.py:6) works on a headless
box.  The two computational steps of ``plot_point_cloud`` are not left to Open3D: the statistical outlier
removal (visualize.py:104-105) runs in the sm_100a kernels behind ``slb200.remove_statistical_outliers`` and
``cloud.ply`` (visualize.py:108) is written here in the layout Open3D gives a legacy point cloud.
"""
import os

import numpy as np


def _corner_ramp(top_left, top_right, bottom_left, bottom_right, width, height):
    # integer bilinear ramp between four corner values; row 0 holds the "bottom" pair (visualize.py:7-31)
    first_row = np.linspace(bottom_left, bottom_right, width)
    last_row = np.linspace(top_left, top_right, width)
    return np.linspace(first_row, last_row, height, dtype=int)[:, :, None]


def create_color_map(width, height):
    """``(height, width, 3)`` int colour ramp used to preview decoded indices (visualize.py:34-56)."""
    return np.concatenate([_corner_ramp(0, 255, 0, 255, width, height),
                           _corner_ramp(0, 0, 255, 0, width, height),
                           _corner_ramp(255, 255, 0, 0, width, height)], axis=2)


def decoded_graycodes_image(h_pixels, v_pixels, color_map_img):
    """The image ``plot_decoded_graycodes`` shows (visualize.py:58-89): colour-map look-up of every decoded pixel,
    indices clamped to the map, black where either index is -1."""
    h = np.asarray(h_pixels)
    v = np.asarray(v_pixels)
    bad = (h == -1) | (v == -1)
    out = color_map_img[np.minimum(v, color_map_img.shape[0] - 1), np.minimum(h, color_map_img.shape[1] - 1)].astype(int)
    out[bad] = 0
    return out


def plot_decoded_graycodes(h_pixels, v_pixels, color_map_img):
    import matplotlib.pyplot as plt
    plt.imshow(decoded_graycodes_image(h_pixels, v_pixels, color_map_img))
    plt.axis('off')
    plt.show()


def clean_point_cloud(Pts, colors, nb_neighbors=20, std_ratio=0.5):
    """Statistical outlier removal as ``plot_point_cloud`` applies it (visualize.py:100-106): the points are
    narrowed to float32 first, like ``o3c.Tensor(Pts.T, o3c.float32)``.  ``Pts`` (3,N), ``colors`` (N,3) ->
    ``points (M,3) f64`` (float32-valued), ``colors (M,3)``, ``ind (M,) int64``.  Runs on the GPU."""
    import torch
    import slb200
    xyz = np.ascontiguousarray(np.asarray(Pts).T, dtype=np.float32)
    cloud = torch.zeros((len(xyz), 4), dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
    cloud[:, :3] = torch.from_numpy(xyz)
    _, ind = slb200.remove_statistical_outliers(cloud, nb_neighbors=nb_neighbors, std_ratio=std_ratio)
    ind = ind.cpu().numpy()
    return xyz[ind].astype(np.float64), np.asarray(colors)[ind], ind


def write_ply(path, points, colors=None):
    """Binary little-endian PLY, float64 x/y/z + uint8 red/green/blue (Open3D's layout for
    ``o3d.io.write_point_cloud`` of a legacy cloud, visualize.py:108)."""
    points = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    layout = [("x", "<f8"), ("y", "<f8"), ("z", "<f8")]
    header = "ply\nformat binary_little_endian 1.0\ncomment Created by Open3D\n"
    header += "element vertex %d\nproperty double x\nproperty double y\nproperty double z\n" % len(points)
    if colors is not None:
        layout += [("red", "u1"), ("green", "u1"), ("blue", "u1")]
        header += "property uchar red\nproperty uchar green\nproperty uchar blue\n"
    header += "end_header\n"
    body = np.empty(len(points), dtype=layout)
    for i, name in enumerate("xyz"):
        body[name] = points[:, i]
    if colors is not None:
        c8 = np.rint(np.clip(np.asarray(colors, dtype=np.float64).reshape(-1, 3), 0, 1) * 255).astype(np.uint8)
        for i, name in enumerate(("red", "green", "blue")):
            body[name] = c8[:, i]
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(body.tobytes())


def plot_point_cloud(Pts, colors, save_to):
    """visualize.py:91-129: outlier removal, ``cloud.ply``, then the interactive viewer when Open3D is there."""
    print('Remove outlier in point cloud')
    points, inlier_colors, _ = clean_point_cloud(Pts, colors, nb_neighbors=20, std_ratio=0.5)
    print('Save ply file')
    write_ply(os.path.join(save_to, 'cloud.ply'), points, inlier_colors)
    try:
        import open3d as o3d
    except ImportError:
        print('open3d is not installed: viewer skipped')
        return
    pcd = o3d.geometry.PointCloud()
    pcd.points = o3d.utility.Vector3dVector(points)
    pcd.colors = o3d.utility.Vector3dVector(np.asarray(inlier_colors, dtype=np.float64))
    o3d.visualization.draw_geometries([pcd])
