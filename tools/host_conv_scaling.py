"""Thread scaling of the host-side conversions of the reference-API path (sl_host_*), on whatever box runs it.
One 1080p scan: 46 float64 frames -> uint8 (763 MB read), 1.8 M records -> (3,N) + (N,3) float64."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "3dscanner-graycode_b200"))
from slb200._lib import lib  # noqa: E402

L = lib()
F, H, W = 46, 1080, 1920
rng = np.random.default_rng(0)
u = rng.integers(0, 256, (F, H, W), dtype=np.uint8)
stack = u.astype(np.float64)
dst = np.zeros_like(u)
n = 1_816_581
cloud = rng.standard_normal((n, 4)).astype(np.float32)
out = {"cores": os.cpu_count(), "narrow_ms": {}, "unpack_ms": {}, "numpy_astype_ms": None}
t = time.perf_counter(); stack.astype(np.uint8); out["numpy_astype_ms"] = (time.perf_counter() - t) * 1e3
for th in (1, 2, 4, 8, 16, 32, 0):
    best = 1e9
    for _ in range(3):
        t = time.perf_counter()
        assert L.sl_host_narrow_u8(stack.ctypes.data, 0, stack.size, dst.ctypes.data, th) == 0
        best = min(best, time.perf_counter() - t)
    out["narrow_ms"][str(th)] = round(best * 1e3, 2)
    pts = np.empty((3, n)); col = np.empty((n, 3))
    best = 1e9
    for _ in range(3):
        t = time.perf_counter()
        L.sl_host_unpack_cloud(cloud.ctypes.data, n, pts.ctypes.data, col.ctypes.data, th)
        best = min(best, time.perf_counter() - t)
    out["unpack_ms"][str(th)] = round(best * 1e3, 2)
out["narrow_read_gbs_auto"] = round(stack.nbytes / (out["narrow_ms"]["0"] * 1e-3) / 1e9, 1)
print(json.dumps(out))
