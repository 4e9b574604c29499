"""Decode-only timing (kernel tuning): python tools/bench_decode.py [--size 1080p] [--iters 300]
Honours SLB200_LIB to time another build of the library."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3dscanner-graycode_b200")); sys.path.insert(0, ROOT)
import torch, slb200
from slb200 import synth
SIZES = {"720p": ((1280, 720), (1024, 768)), "1080p": ((1920, 1080), (1920, 1080)), "4k": ((3840, 2160), (3840, 2160))}
ap = argparse.ArgumentParser(); ap.add_argument("--size", default="1080p"); ap.add_argument("--iters", type=int, default=300)
ap.add_argument("--ring", type=int, default=8); ap.add_argument("--noise", action="store_true"); ap.add_argument("--batch", type=int, default=1); ap.add_argument("--streams", type=int, default=1)
a = ap.parse_args()
cam, proj = SIZES[a.size]; dev = torch.device("cuda", 0)
calib = synth.synthetic_calibration(cam, proj); pats = synth.pattern_stack(proj, dev)
ring = [synth.render_scan(cam, proj, seed=i, device=dev, patterns=pats, calib=calib)[0] for i in range(a.ring)]
if a.noise:
    ring = [torch.randint(0, 256, r.shape, dtype=torch.uint8, device=dev) for r in ring]
F = ring[0].shape[0]
if a.batch > 1:
    ring = [torch.stack([ring[(i + k) % len(ring)] for k in range(a.batch)]).contiguous() for i in range(max(2, a.ring // a.batch))]
    a.ring = len(ring)
shape = (a.batch, cam[1], cam[0]) if a.batch > 1 else (cam[1], cam[0])
outs = [(torch.empty(shape, dtype=torch.int16, device=dev), torch.empty(shape, dtype=torch.int16, device=dev)) for _ in range(a.streams)]
for r in ring: slb200.decode_pixels(r, out=outs[0])
torch.cuda.synchronize()
streams = [torch.cuda.Stream(dev) for _ in range(a.streams)]
graphs = []
for i, r in enumerate(ring):
    st = streams[i % a.streams]
    with torch.cuda.stream(st):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st): slb200.decode_pixels(r, out=outs[i % a.streams])
        graphs.append(g)
torch.cuda.synchronize()
def run(n):
    for i in range(n):
        k = i % a.ring
        with torch.cuda.stream(streams[k % a.streams]): graphs[k].replay()
run(20); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
main = torch.cuda.current_stream()
e0.record(main)
for st in streams: st.wait_event(e0)
run(a.iters)
for st in streams:
    ev = torch.cuda.Event(); ev.record(st); main.wait_event(ev)
e1.record(main); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / a.iters * 1e3
gb = a.batch * (F + 4) * cam[0] * cam[1] / 1e9
print(f"{os.environ.get('SLB200_LIB', 'default').split('/')[-1]:60s} {a.size} B={a.batch} streams={a.streams} noise={a.noise} {us:8.2f} us  {gb / (us * 1e-6):8.1f} GB/s  {gb / (us * 1e-6) / 6535.4:6.3f} of peak")
