"""The bindings shown in INTEGRATION.md section 3, executed as written: raw ctypes against libslb200.so, NumPy
host buffers only for the pipeline call (no slb200 Python package on that path)."""
import ctypes
import os

import numpy as np
import pytest

from oracle import graycode_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "3dscanner-graycode_b200", "lib", "libslb200.so")


class SlCalib(ctypes.Structure):           # include/slb200.h
    _fields_ = [("cam_K", ctypes.c_double * 9), ("cam_dist", ctypes.c_double * 8), ("proj_K", ctypes.c_double * 9),
                ("proj_dist", ctypes.c_double * 8), ("R", ctypes.c_double * 9), ("T", ctypes.c_double * 3),
                ("proj_w", ctypes.c_int32), ("proj_h", ctypes.c_int32), ("cam_row0", ctypes.c_int32),
                ("reserved", ctypes.c_int32)]


def test_ctypes_only_host_pipeline():
    from slb200 import synth            # input generation + calibration values only
    sl = ctypes.CDLL(LIB)
    sl.sl_last_error.restype = ctypes.c_char_p
    cam_size, proj_size = (160, 96), (1024, 768)
    cal = synth.synthetic_calibration(cam_size, proj_size)
    B = 2
    scans = [synth.render_scan(cam_size, proj_size, seed=s, device="cpu") for s in range(B)]
    frames = np.ascontiguousarray(np.stack([s[0].numpy() for s in scans]))
    white = np.ascontiguousarray(np.stack([s[1].numpy() for s in scans]))
    F, H, W = frames.shape[1:]
    c = SlCalib()
    c.cam_K[:] = cal.cam_mtx.ravel(); c.proj_K[:] = cal.proj_mtx.ravel(); c.R[:] = cal.proj_R.ravel()
    c.cam_dist[:] = list(cal.cam_dist.ravel()) + [0.0] * 3
    c.proj_dist[:] = list(cal.proj_dist.ravel()) + [0.0] * 3
    c.T[:] = cal.proj_T.ravel()
    c.proj_w, c.proj_h = proj_size
    pipe = ctypes.c_void_p()
    assert sl.sl_pipeline_create(ctypes.byref(pipe), 0, 4, F, H, W) == 0, sl.sl_last_error()
    cloud = np.empty((B * H * W, 4), np.float32)
    counts = np.empty(B + 1, np.int64)
    h_idx = np.empty((B, H, W), np.int16); v_idx = np.empty((B, H, W), np.int16)
    rc = sl.sl_scan_to_cloud_host(pipe, ctypes.c_void_p(frames.ctypes.data), ctypes.c_void_p(white.ctypes.data), B,
                                  ctypes.byref(c), 1, ctypes.c_double(1.0), ctypes.c_void_p(h_idx.ctypes.data),
                                  ctypes.c_void_p(v_idx.ctypes.data), ctypes.c_void_p(cloud.ctypes.data),
                                  ctypes.c_int64(cloud.shape[0]), ctypes.c_void_p(counts.ctypes.data))
    assert rc == 0, sl.sl_last_error()
    # the streaming form shown in INTEGRATION.md: one scan per call, results valid after the flush
    clouds = [np.zeros((H * W, 4), np.float32) for _ in range(B)]
    cnts = [np.zeros(1, np.int64) for _ in range(B)]
    for i in range(B):
        rc = sl.sl_pipeline_submit(pipe, ctypes.c_void_p(frames[i].ctypes.data), ctypes.c_void_p(white[i].ctypes.data),
                                   ctypes.byref(c), 1, ctypes.c_double(1.0), ctypes.c_void_p(clouds[i].ctypes.data),
                                   ctypes.c_int64(clouds[i].shape[0]), ctypes.c_void_p(cnts[i].ctypes.data))
        assert rc == 0, sl.sl_last_error()
    assert sl.sl_pipeline_flush(pipe) == 0, sl.sl_last_error()
    off = 0
    for i in range(B):
        assert int(cnts[i][0]) == int(counts[i])
        np.testing.assert_array_equal(clouds[i][: int(counts[i])].view(np.int32),
                                      cloud[off: off + int(counts[i])].view(np.int32))
        off += int(counts[i])
    sl.sl_pipeline_destroy(pipe)
    start = 0
    for b in range(B):
        hp, vp = O.decode_pixels(frames[b].astype(np.float64))
        np.testing.assert_array_equal(h_idx[b], hp); np.testing.assert_array_equal(v_idx[b], vp)
        cam, proj, col = O.cam_proj_pairs(hp, vp, proj_size, white[b])
        ref = O.triangulate(cam, proj, cal.cam_mtx, cal.cam_dist, cal.proj_mtx, cal.proj_dist, cal.proj_R, cal.proj_T)
        n = int(counts[b])
        assert n == ref.shape[1]
        pts_3d = cloud[start:start + n, :3].T.astype(np.float64)                 # what src/4-triangulate.py:67 saves
        colors = cloud[start:start + n, 3].copy().view(np.uint8).reshape(-1, 4)[:, :3] / 255.0
        np.testing.assert_allclose(pts_3d, ref, rtol=2e-7)                        # fp64 mode, f32 records
        np.testing.assert_array_equal(colors, col)
        start += n
    assert counts[B] == start
