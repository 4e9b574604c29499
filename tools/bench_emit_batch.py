"""Emit of B decoded scans: ONE batched launch against B single-scan launches spread over S streams (us per scan)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3dscanner-graycode_b200")); sys.path.insert(0, ROOT)
import torch, slb200
from slb200 import synth
from slb200.api import _ptr, _stream, check, lib, CloudBuffers
cam, proj = (1920, 1080), (1920, 1080); dev = torch.device("cuda", 0)
calib = synth.synthetic_calibration(cam, proj); pats = synth.pattern_stack(proj, dev)
tables = slb200.LensTables(calib, cam, dev, precision="fp32")
B = 16
scans = [synth.render_scan(cam, proj, seed=i, device=dev, patterns=pats, calib=calib) for i in range(8)]
frames = torch.stack([scans[i % 8][0] for i in range(B)]).contiguous(); white = torch.stack([scans[i % 8][1] for i in range(B)]).contiguous()
big = CloudBuffers.allocate(B, cam[1], cam[0], dev)
slb200.scan_to_cloud(frames, white, calib, buffers=big, tables=tables)
singles = []
for i in range(B):
    b = CloudBuffers.allocate(1, cam[1], cam[0], dev)
    slb200.scan_to_cloud(frames[i], white[i], calib, buffers=b, tables=tables)
    singles.append(b)
torch.cuda.synchronize()
s = calib.struct(); L = lib(); H, W = cam[1], cam[0]
def emit(b, wh, nb):
    check(L.sl_emit_cloud(_ptr(b.h_idx), _ptr(b.v_idx), _ptr(wh), C.byref(s), 0, _ptr(tables.cam), _ptr(tables.proj_bt), nb, H, W, _ptr(b.ws), _ptr(b.cloud), _stream()))
def offsets(b, nb):
    check(L.sl_compact_offsets_decoded(nb, H, W, _ptr(b.ws), _ptr(b.counts), _stream()))
def timeit(fn, iters=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3 / B
print(f"batched emit, one launch of {B} scans: {timeit(lambda: emit(big, white, B)):.2f} us per scan")
print(f"batched offsets + emit: {timeit(lambda: (offsets(big, B), emit(big, white, B))):.2f} us per scan")
for S in (1, 2, 4, 8):
    sts = [torch.cuda.Stream(dev) for _ in range(S)]
    def fn():
        cur = torch.cuda.current_stream()
        for st in sts: st.wait_stream(cur)
        for i in range(B):
            with torch.cuda.stream(sts[i % S]): emit(singles[i], white[i], 1)
        for st in sts: cur.wait_stream(st)
    print(f"{B} single-scan emits over {S} stream(s): {timeit(fn):.2f} us per scan")
