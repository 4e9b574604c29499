"""Stage-2-only timing (offsets + emit) on decoded 1080p scans; honours SLB200_LIB."""
import argparse, ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3dscanner-graycode_b200")); sys.path.insert(0, ROOT)
import torch, slb200
from slb200 import synth
from slb200.api import _ptr, _stream, check, lib, CloudBuffers
ap = argparse.ArgumentParser(); ap.add_argument("--iters", type=int, default=300); ap.add_argument("--ring", type=int, default=6); ap.add_argument("--streams", type=int, default=1)
a = ap.parse_args()
cam, proj = (1920, 1080), (1920, 1080); dev = torch.device("cuda", 0)
calib = synth.synthetic_calibration(cam, proj); pats = synth.pattern_stack(proj, dev)
tables = slb200.LensTables(calib, cam, dev)
sets = []
for i in range(a.ring):
    fr, wh = synth.render_scan(cam, proj, seed=i, device=dev, patterns=pats, calib=calib)
    b = CloudBuffers.allocate(1, cam[1], cam[0], dev)
    slb200.scan_to_cloud(fr, wh, calib, buffers=b, tables=tables)       # leaves h/v + bitmap in b
    sets.append((b, wh)); del fr
torch.cuda.synchronize()
s = calib.struct(); L = lib(); H, W = cam[1], cam[0]
def offsets(b): check(L.sl_compact_offsets_decoded(1, H, W, _ptr(b.ws), _ptr(b.counts), _stream()))
def emit(b, wh): check(L.sl_emit_cloud(_ptr(b.h_idx), _ptr(b.v_idx), _ptr(wh), C.byref(s), 0, _ptr(tables.cam), _ptr(tables.proj_bt), 1, H, W, _ptr(b.ws), _ptr(b.cloud), _stream()))
def timeit(fn, name):
    sts = [torch.cuda.Stream(dev) for _ in range(a.streams)]; graphs = []
    for i, (b, wh) in enumerate(sets):
        with torch.cuda.stream(sts[i % a.streams]):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=sts[i % a.streams]): fn(b, wh)
            graphs.append(g)
    torch.cuda.synchronize()
    main = torch.cuda.current_stream()
    def run(n):
        for i in range(n):
            k = i % len(graphs)
            with torch.cuda.stream(sts[k % a.streams]): graphs[k].replay()
    run(2 * len(graphs)); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for st in sts: st.wait_event(e0)
    run(a.iters)
    for st in sts:
        ev = torch.cuda.Event(); ev.record(st); main.wait_event(ev)
    e1.record(main); torch.cuda.synchronize()
    print(f"{os.environ.get('SLB200_LIB','default').split('/')[-1]:45s} streams={a.streams} {name:14s} {e0.elapsed_time(e1) / a.iters * 1e3:8.2f} us")
timeit(lambda b, wh: offsets(b), "offsets")
timeit(lambda b, wh: emit(b, wh), "emit")
timeit(lambda b, wh: (offsets(b), emit(b, wh)), "offsets+emit")
