"""Import-safe shim for the reference's scanner/utils/visualize.py (viewers are outside the
hot path).  matplotlib / open3d are imported lazily so `from scanner.utils import visualize`
(src/3-capture_decode.py:8, src/4-triangulate.py:6) works on a headless box."""
import numpy as np


def create_color_map(width, height):
    """RGB ramp image used to preview decoded indices (reference visualize.py:34-56)."""
    xs = np.linspace(0.0, 1.0, int(width), dtype=np.float64)[None, :]
    ys = np.linspace(0.0, 1.0, int(height), dtype=np.float64)[:, None]
    img = np.zeros((int(height), int(width), 3), dtype=np.float64)
    img[..., 0] = xs
    img[..., 1] = ys
    img[..., 2] = 1.0 - 0.5 * (xs + ys)
    return img


def plot_decoded_graycodes(h_pixels, v_pixels, color_map_img):
    import matplotlib.pyplot as plt
    h = np.clip(np.asarray(h_pixels), 0, color_map_img.shape[1] - 1)
    v = np.clip(np.asarray(v_pixels), 0, color_map_img.shape[0] - 1)
    out = color_map_img[v, h]
    out[(np.asarray(h_pixels) < 0) | (np.asarray(v_pixels) < 0)] = 0
    plt.imshow(out)
    plt.show()


def plot_point_cloud(pts_3d, colors, output_folder=None):
    import open3d as o3d
    pcd = o3d.geometry.PointCloud()
    pcd.points = o3d.utility.Vector3dVector(np.asarray(pts_3d).T)
    pcd.colors = o3d.utility.Vector3dVector(np.asarray(colors))
    o3d.visualization.draw_geometries([pcd])
