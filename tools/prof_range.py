"""DRAM traffic of the PIPELINED bench step (independent scans over several streams), which per-kernel ncu cannot see
because it serialises kernels:  ncu --replay-mode range --profile-from-start off --metrics ... python tools/prof_range.py
One profiled range = `--scans` scans launched round-robin over `--streams` streams (plain launches, no graphs)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3dscanner-graycode_b200")); sys.path.insert(0, ROOT)
import torch, slb200
from slb200 import synth
from slb200.api import CloudBuffers
ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=4); ap.add_argument("--scans", type=int, default=48)
ap.add_argument("--ring", type=int, default=12); ap.add_argument("--sm-reserve", type=int, default=8)
ap.add_argument("--what", default="step", choices=["step", "decode", "stage2"])
ap.add_argument("--graphs", action="store_true", help="replay pre-captured CUDA graphs inside the range (the decode launch encodes a tensor map, which range replay cannot capture)")
a = ap.parse_args()
cam, proj = (1920, 1080), (1920, 1080); dev = torch.device("cuda", 0)
calib = synth.synthetic_calibration(cam, proj); pats = synth.pattern_stack(proj, dev)
tables = slb200.LensTables(calib, cam, dev, precision="fp32"); cs = calib.struct()
scans = [synth.render_scan(cam, proj, seed=i, device=dev, patterns=pats, calib=calib) for i in range(a.ring)]
bufs = [CloudBuffers.allocate(1, cam[1], cam[0], dev) for _ in range(a.streams)]
sts = [torch.cuda.Stream(dev) for _ in range(a.streams)]
from slb200.api import _ptr, _stream, check, lib
import ctypes as C
L = lib(); H, W = cam[1], cam[0]
def one(i):
    k = i % a.streams
    fr, wh = scans[i % a.ring]
    with torch.cuda.stream(sts[k]):
        if a.what == "step":
            slb200.scan_to_cloud(fr, wh, calib, buffers=bufs[k], calib_struct=cs, tables=tables, sm_reserve=a.sm_reserve)
        elif a.what == "decode":
            slb200.decode_pixels(fr, out=(bufs[k].h_idx[0], bufs[k].v_idx[0]), sm_reserve=a.sm_reserve)
        else:
            b = bufs[k]
            check(L.sl_compact_offsets(_ptr(b.h_idx), _ptr(b.v_idx), 1, H, W, _ptr(b.ws), _ptr(b.counts), _stream()))
            check(L.sl_emit_cloud(_ptr(b.h_idx), _ptr(b.v_idx), _ptr(wh), C.byref(cs), 0, _ptr(tables.cam), _ptr(tables.proj_bt), 1, H, W, _ptr(b.ws), _ptr(b.cloud), _stream()))
for i in range(2 * a.ring): one(i)
torch.cuda.synchronize()
if a.graphs:
    graphs = []
    for i in range(a.ring):
        k = i % a.streams
        with torch.cuda.stream(sts[k]):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=sts[k]):
                one(i)
            graphs.append(g)
    torch.cuda.synchronize()
    plain = one
    def one(i):
        with torch.cuda.stream(sts[i % a.ring % a.streams]):
            graphs[i % a.ring].replay()
    for i in range(a.ring): one(i)
    torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
for i in range(a.scans): one(i)
for s in sts: torch.cuda.current_stream().wait_stream(s)
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"{a.what}: {a.scans} scans over {a.streams} streams, {e0.elapsed_time(e1) / a.scans * 1e3:.2f} us per scan (events, plain launches)")
