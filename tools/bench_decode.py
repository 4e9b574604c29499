"""Decode-only timing (kernel tuning): python tools/bench_decode.py [--size 1080p] [--iters 300]
Honours SLB200_LIB to time another build of the library."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3dscanner-graycode_b200")); sys.path.insert(0, ROOT)
import torch, slb200
from slb200 import synth
SIZES = {"720p": ((1280, 720), (1024, 768)), "1080p": ((1920, 1080), (1920, 1080)), "4k": ((3840, 2160), (3840, 2160))}
ap = argparse.ArgumentParser(); ap.add_argument("--size", default="1080p"); ap.add_argument("--iters", type=int, default=300)
ap.add_argument("--ring", type=int, default=8); ap.add_argument("--noise", action="store_true")
a = ap.parse_args()
cam, proj = SIZES[a.size]; dev = torch.device("cuda", 0)
calib = synth.synthetic_calibration(cam, proj); pats = synth.pattern_stack(proj, dev)
ring = [synth.render_scan(cam, proj, seed=i, device=dev, patterns=pats, calib=calib)[0] for i in range(a.ring)]
if a.noise:
    ring = [torch.randint(0, 256, r.shape, dtype=torch.uint8, device=dev) for r in ring]
F = ring[0].shape[0]
h = torch.empty((cam[1], cam[0]), dtype=torch.int16, device=dev); v = torch.empty_like(h)
for r in ring: slb200.decode_pixels(r, out=(h, v))
torch.cuda.synchronize()
graphs = []; side = torch.cuda.Stream(dev)
with torch.cuda.stream(side):
    for r in ring:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side): slb200.decode_pixels(r, out=(h, v))
        graphs.append(g)
torch.cuda.synchronize()
for i in range(20): graphs[i % a.ring].replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.iters): graphs[i % a.ring].replay()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / a.iters * 1e3
gb = (F + 4) * cam[0] * cam[1] / 1e9
print(f"{os.environ.get('SLB200_LIB', 'default').split('/')[-1]:60s} {a.size} noise={a.noise} {us:8.2f} us  {gb / (us * 1e-6):8.1f} GB/s  {gb / (us * 1e-6) / 6535.4:6.3f} of peak")
