"""Does stage 2 make progress beside a running decode?  Decode of scan i on stream A and offsets + emit of an already
decoded scan j on stream B, launched together, against each of them alone (CUDA graphs, events over both streams)."""
import argparse, ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3dscanner-graycode_b200")); sys.path.insert(0, ROOT)
import torch, slb200
from slb200 import synth
from slb200.api import _ptr, _stream, check, lib, CloudBuffers
ap = argparse.ArgumentParser(); ap.add_argument("--iters", type=int, default=200); ap.add_argument("--ring", type=int, default=6)
ap.add_argument("--sm-reserve", type=int, default=0)
a = ap.parse_args()
cam, proj = (1920, 1080), (1920, 1080); dev = torch.device("cuda", 0)
calib = synth.synthetic_calibration(cam, proj); pats = synth.pattern_stack(proj, dev)
tables = slb200.LensTables(calib, cam, dev, precision="fp32")
frames, sets = [], []
for i in range(a.ring):
    fr, wh = synth.render_scan(cam, proj, seed=i, device=dev, patterns=pats, calib=calib)
    b = CloudBuffers.allocate(1, cam[1], cam[0], dev)
    slb200.scan_to_cloud(fr, wh, calib, buffers=b, tables=tables)
    frames.append(fr); sets.append((b, wh))
dec_out = [CloudBuffers.allocate(1, cam[1], cam[0], dev) for _ in range(2)]
torch.cuda.synchronize()
s = calib.struct(); L = lib(); H, W = cam[1], cam[0]
sa, sb = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def stage2(b, wh):
    check(L.sl_compact_offsets_decoded(1, H, W, _ptr(b.ws), _ptr(b.counts), _stream()))
    check(L.sl_emit_cloud(_ptr(b.h_idx), _ptr(b.v_idx), _ptr(wh), C.byref(s), 0, _ptr(tables.cam), _ptr(tables.proj_bt), 1, H, W, _ptr(b.ws), _ptr(b.cloud), _stream()))
def emit_only(b, wh):
    check(L.sl_emit_cloud(_ptr(b.h_idx), _ptr(b.v_idx), _ptr(wh), C.byref(s), 0, _ptr(tables.cam), _ptr(tables.proj_bt), 1, H, W, _ptr(b.ws), _ptr(b.cloud), _stream()))
def cap(stream, fn):
    with torch.cuda.stream(stream):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream): fn()
    return g
gd = [cap(sa, lambda i=i: slb200.decode_pixels(frames[i], out=(dec_out[i % 2].h_idx[0], dec_out[i % 2].v_idx[0]), sm_reserve=a.sm_reserve)) for i in range(a.ring)]
g2 = [cap(sb, lambda i=i: stage2(*sets[i])) for i in range(a.ring)]
ge = [cap(sb, lambda i=i: emit_only(*sets[i])) for i in range(a.ring)]
torch.cuda.synchronize()
main = torch.cuda.current_stream()
def timed(ga, gb, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main); sa.wait_event(e0); sb.wait_event(e0)
    for i in range(n):
        if ga is not None:
            with torch.cuda.stream(sa): ga[i % a.ring].replay()
        if gb is not None:
            with torch.cuda.stream(sb): gb[i % a.ring].replay()
    for st in (sa, sb):
        ev = torch.cuda.Event(); ev.record(st); main.wait_event(ev)
    e1.record(main); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for name, ga, gb in (("decode alone", gd, None), ("offsets+emit alone", None, g2), ("emit alone", None, ge),
                     ("decode || offsets+emit", gd, g2), ("decode || emit", gd, ge)):
    timed(ga, gb, 20)
    print(f"reserve={a.sm_reserve} {name:24s} {timed(ga, gb, a.iters):8.2f} us per pair")
