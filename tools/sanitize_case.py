"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3dscanner-graycode_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch, slb200
from slb200 import synth
for cam in ((160, 96), (131, 67)):
    proj = (1024, 768)
    calib = synth.synthetic_calibration(cam, proj)
    fr, wh = synth.render_scan(cam, proj, seed=2, device="cpu")
    fr, wh = fr.cuda(), wh.cuda()
    tab = slb200.LensTables(calib, cam, "cuda")
    for t in (tab, None):
        b = slb200.scan_to_cloud(fr, wh, calib, tables=t)
    hc, vc, hi, vi = slb200.decode_codes(fr)
    slb200.codes_to_pixels(torch.stack([hc, hc]).contiguous(), torch.stack([vc, vc]).contiguous())
    cam_p, proj_p, rgb = slb200.cam_proj_pairs(hi, vi, proj, wh)
    pts = slb200.triangulate_points(cam_p, proj_p, calib)
    slb200.filter_box(pts, rgb.to(torch.float64), 5.0)
    slb200.decode_pixels(fr, eps=0.5)
    slb200.bgr_to_gray(wh); slb200.pair_diff_counts(fr)
    torch.cuda.synchronize()
    print(cam, "points", int(b.counts[-1]))
print("done")
