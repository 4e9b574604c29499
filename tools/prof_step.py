"""Small driver for ncu: a few plain (non-graph) steps of the bench workload on one GPU.
    ncu ... python tools/prof_step.py [--steps 6] [--size 1080p|720p|4k] [--batch B]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3dscanner-graycode_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import slb200  # noqa: E402
from slb200 import synth  # noqa: E402

SIZES = {"720p": ((1280, 720), (1024, 768)), "1080p": ((1920, 1080), (1920, 1080)), "4k": ((3840, 2160), (3840, 2160))}

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--size", default="1080p")
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--precision", default="fp32")
ap.add_argument("--lens", default="tables")
a = ap.parse_args()
cam, proj = SIZES[a.size]
dev = torch.device("cuda", 0)
calib = synth.synthetic_calibration(cam, proj)
pats = synth.pattern_stack(proj, dev)
scans = [synth.render_scan(cam, proj, seed=i, device=dev, patterns=pats, calib=calib) for i in range(max(2, a.batch))]
if a.batch > 1:
    frames = [torch.stack([s[0] for s in scans[:a.batch]]).contiguous()]
    white = [torch.stack([s[1] for s in scans[:a.batch]]).contiguous()]
else:
    frames = [s[0] for s in scans]
    white = [s[1] for s in scans]
bufs = slb200.api.CloudBuffers.allocate(a.batch, cam[1], cam[0], dev)
tables = slb200.LensTables(calib, cam, dev) if a.lens == "tables" else None
torch.cuda.synchronize()
for i in range(a.steps):
    slb200.scan_to_cloud(frames[i % len(frames)], white[i % len(white)], calib, precision=a.precision, buffers=bufs, tables=tables)
torch.cuda.synchronize()
print("valid fraction", float(bufs.counts[-1]) / (a.batch * cam[0] * cam[1]))
