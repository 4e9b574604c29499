"""Drop-in for the reference's scanner/triangulation/triangulate.py, computed on a B200.

``Triangulate`` keeps the reference's constructor arguments, defaults and method signatures
(triangulate.py:4-122) so src/4-triangulate.py runs unchanged; ``Triangulation`` is an alias.
NumPy arrays in and out; the arithmetic runs in the sm_100a kernels behind ``slb200``.
"""
import numpy as np
import torch

import slb200


class Triangulate(object):
    def __init__(self,
                 h_pixels=None,
                 v_pixels=None,
                 cam_size=(1920, 1080),
                 cam_mtx=None,
                 cam_dist=None,
                 proj_size=(1920, 1080),
                 proj_calib_size=(1920, 1080),
                 proj_mtx=None,
                 proj_dist=None,
                 proj_R=None,
                 proj_T=None,
                 image_folder=None):
        self.h_pixels = h_pixels
        self.v_pixels = v_pixels
        self.cam_w, self.cam_h = cam_size
        self.cam_mtx = cam_mtx
        self.cam_dist = cam_dist
        self.proj_w, self.proj_h = proj_size
        self.proj_mtx = proj_mtx
        self.proj_dist = proj_dist
        # The reference rescales the caller's matrix in place (triangulate.py:28-33); callers
        # that reuse the array rely on that, so it is kept.
        if proj_mtx is not None:
            calib_w, calib_h = proj_calib_size
            self.proj_mtx[0, :] = self.proj_mtx[0, :] * (self.proj_w / calib_w)
            self.proj_mtx[1, :] = self.proj_mtx[1, :] * (self.proj_h / calib_h)
        self.proj_T = proj_T
        self.proj_R = proj_R
        self.image_folder = image_folder
        # device-side copies of what the previous call returned, keyed by the identity of the NumPy arrays handed
        # out: src/4-triangulate.py:63-71 chains get_cam_proj_pts -> triangulate -> filter_3d_pts on those very
        # arrays, so the chain does not re-upload what the GPU just produced
        self._dev_cache = {}

    # -- helpers ---------------------------------------------------------------------------
    def _calibration(self):
        return slb200.Calibration(np.asarray(self.cam_mtx, np.float64), np.asarray(self.cam_dist, np.float64),
                                  np.asarray(self.proj_mtx, np.float64), np.asarray(self.proj_dist, np.float64),
                                  np.asarray(self.proj_R, np.float64), np.asarray(self.proj_T, np.float64),
                                  (int(self.proj_w), int(self.proj_h)))

    def _indices(self):
        h = np.asarray(self.h_pixels)
        v = np.asarray(self.v_pixels)
        if h.shape != v.shape or h.ndim != 2:
            raise ValueError("h_pixels / v_pixels must be (H,W) arrays of equal shape")
        if h.shape[0] < self.cam_h or h.shape[1] < self.cam_w:
            raise IndexError("h_pixels / v_pixels smaller than cam_size")     # reference: IndexError at :54
        return h[:self.cam_h, :self.cam_w], v[:self.cam_h, :self.cam_w]

    # -- reference API ---------------------------------------------------------------------
    def get_cam_proj_pts(self, img_white=None):
        """``cam_pts (N,2) f32, proj_pts (N,2) f32, colors (N,3) f64`` in x-major order
        (reference triangulate.py:39-71)."""
        h, v = self._indices()
        white = None
        if img_white is not None:
            white = np.ascontiguousarray(np.asarray(img_white)[:self.cam_h, :self.cam_w, :3])
            if white.dtype != np.uint8:
                raise TypeError("img_white must be a uint8 image")
        cam, proj, rgb = slb200.cam_proj_pairs(np.ascontiguousarray(h), np.ascontiguousarray(v),
                                               (self.proj_w, self.proj_h), white)
        cam_pts = cam.cpu().numpy()
        proj_pts = proj.cpu().numpy()
        if img_white is None:
            # the reference falls off the end with `colors_array` unbound (triangulate.py:66-71)
            raise UnboundLocalError("cannot access local variable 'colors_array': img_white is None")
        colors_array = rgb.cpu().numpy().astype(np.float64) / 255.0
        self._dev_cache = {"pairs": (cam_pts, proj_pts, cam, proj)}
        return cam_pts, proj_pts, colors_array

    def triangulate(self, cam_pts, proj_pts):
        """``(3,N)`` float64 points (reference triangulate.py:73-97)."""
        hit = self._dev_cache.get("pairs")
        if hit is not None and hit[0] is cam_pts and hit[1] is proj_pts:
            cam_dev, proj_dev = hit[2], hit[3]                   # still on the device from get_cam_proj_pts
        else:
            cam_dev, proj_dev = np.asarray(cam_pts, dtype=np.float32), np.asarray(proj_pts, dtype=np.float32)
        pts = slb200.triangulate_points(cam_dev, proj_dev, self._calibration(), precision="fp64")
        out = pts.cpu().numpy()
        self._dev_cache["pts"] = (out, pts)
        return out

    def filter_3d_pts(self, Pts, colors, threshold=.5):
        """Strict +-threshold box filter (reference triangulate.py:99-122)."""
        dev = torch.device("cuda", torch.cuda.current_device())
        hit = self._dev_cache.get("pts")
        if hit is not None and hit[0] is Pts:
            p = hit[1]                                           # still on the device from triangulate
        else:
            p = torch.as_tensor(np.ascontiguousarray(np.asarray(Pts, dtype=np.float64))).to(dev)
        c = torch.as_tensor(np.ascontiguousarray(np.asarray(colors, dtype=np.float64))).to(dev)
        out_p, out_c = slb200.filter_box(p, c, float(threshold))
        return out_p.cpu().numpy(), out_c.cpu().numpy()

    # -- additive fast path ----------------------------------------------------------------
    def point_cloud(self, img_white, precision="fp64"):
        """get_cam_proj_pts + triangulate in one fused pass (no (N,2) round trip through the host).
        Returns ``pts (3,N) f64, colors (N,3) f64`` like the two reference calls chained."""
        h, v = self._indices()
        dev = torch.device("cuda", torch.cuda.current_device())
        h_idx = slb200.api._idx16(np.ascontiguousarray(h), dev)
        v_idx = slb200.api._idx16(np.ascontiguousarray(v), dev)
        white = torch.as_tensor(np.ascontiguousarray(np.asarray(img_white)[:self.cam_h, :self.cam_w, :3])).to(dev)
        bufs = slb200.api.CloudBuffers.allocate(1, self.cam_h, self.cam_w, dev)
        bufs.h_idx.copy_(h_idx.view(1, self.cam_h, self.cam_w))
        bufs.v_idx.copy_(v_idx.view(1, self.cam_h, self.cam_w))
        from slb200.api import _ptr, _stream, _PREC, check, lib
        import ctypes as C
        s = self._calibration().struct()
        with torch.cuda.device(dev):
            check(lib().sl_compact_offsets(_ptr(bufs.h_idx), _ptr(bufs.v_idx), 1, self.cam_h, self.cam_w,
                                           _ptr(bufs.ws), _ptr(bufs.counts), _stream()))
            check(lib().sl_emit_cloud(_ptr(bufs.h_idx), _ptr(bufs.v_idx), _ptr(white), C.byref(s), _PREC[precision],
                                      None, None, 1, self.cam_h, self.cam_w, _ptr(bufs.ws), _ptr(bufs.cloud), _stream()))
        n = int(bufs.counts[1].item())
        # 16 bytes per point cross PCIe; the float64 (3,N) / (N,3) arrays of the reference are produced on the host
        # (sl_host_unpack_cloud: exact widening, colours = uint8 / 255 as an IEEE division, triangulate.py:69)
        pts, colors = slb200.api.cloud_to_host_f64(bufs.cloud[:n])
        return pts, colors


Triangulation = Triangulate
