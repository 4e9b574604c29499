"""Times the UNMODIFIED reference (imported from /root/reference; build container only) on BASELINE config[0]:
get_codes + the two per-pixel gray_to_decimal loops exactly as src/3-capture_decode.py:75,99-100 writes them, then
Triangulate.get_cam_proj_pts + triangulate as src/4-triangulate.py:63-64.  Single thread by construction.
    python tools/time_reference_faithful.py > profiles/r02_reference_faithful.json
The reference cannot travel to the GPU box, so this number is recorded with its provenance instead of being part of
bench.py (whose CPU arms time the oracle port, ~25x faster than these loops)."""
import importlib.util, json, os, platform, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3dscanner-graycode_b200")); sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
REF = "/root/reference"

def load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m); return m

dc = load("_ref_dc", "scanner/grayCode/decode_codes.py")
tr = load("_ref_tr", "scanner/triangulation/triangulate.py")
from slb200 import synth
cam_size, proj_size = (1280, 720), (1024, 768)
calib = synth.synthetic_calibration(cam_size, proj_size)
frames, white = synth.render_scan(cam_size, proj_size, seed=0, device="cpu")
images = frames.numpy().astype(np.float64)          # src/3-capture_decode.py:69
img_white = white.numpy()
t = {}
t0 = time.perf_counter(); h_codes, v_codes = dc.get_codes(images); t["get_codes_s"] = time.perf_counter() - t0
H, W = images.shape[1:]
t0 = time.perf_counter()
h_pixels = np.array([[dc.gray_to_decimal(h_codes[:, y, x]) for x in range(W)] for y in range(H)])              # :99
v_pixels = np.array([[dc.gray_to_decimal(np.flip(v_codes[:, y, x])) for x in range(W)] for y in range(H)])     # :100
t["gray_to_decimal_loops_s"] = time.perf_counter() - t0
tri = tr.Triangulate(h_pixels, v_pixels, cam_size, calib.cam_mtx, calib.cam_dist, proj_size, proj_size,
                     calib.proj_mtx.copy(), calib.proj_dist, calib.proj_R, calib.proj_T)
t0 = time.perf_counter(); cam_pts, proj_pts, colors = tri.get_cam_proj_pts(img_white); t["get_cam_proj_pts_s"] = time.perf_counter() - t0
t0 = time.perf_counter(); pts = tri.triangulate(cam_pts, proj_pts); t["triangulate_s"] = time.perf_counter() - t0
total = sum(t.values())
print(json.dumps({"what": "unmodified reference, BASELINE config[0] (1280x720 camera, 1024x768 projector, F = 42), one synthetic scan",
                  "Mpx_per_s": W * H / 1e6 / total, "seconds_per_scan": total, "stages": t, "points": int(pts.shape[1]),
                  "cores_used": 1, "host": platform.processor() or platform.machine(), "cpu_count": os.cpu_count(),
                  "where": "build container (no GPU box access to /root/reference)", "python": platform.python_version(),
                  "numpy": np.__version__}))
