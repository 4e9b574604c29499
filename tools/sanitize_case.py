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
        for prec in ("fp32", "fp64"):
            b = slb200.scan_to_cloud(fr, wh, calib, tables=t, precision=prec)
    # a batch: the offsets kernel's cross-strip / cross-scan published totals
    slb200.scan_to_cloud(torch.stack([fr, fr, fr]).contiguous(), torch.stack([wh, wh, wh]).contiguous(), calib, tables=tab)
    # both emissions explicitly (the row-mapped one needs W % 32 == 0), the boxed compaction, a row band
    if cam[0] % 32 == 0:
        for em in ("tiles", "rows"):
            slb200.scan_to_cloud(torch.stack([fr, fr]).contiguous(), torch.stack([wh, wh]).contiguous(), calib, tables=tab, emit=em)
            slb200.scan_to_cloud(fr, None, calib, tables=tab, emit=em, box=0.4)
            slb200.scan_to_cloud(fr[:, 32:], wh[32:].contiguous(), calib, tables=tab, emit=em, row0=32)
    slb200.scan_to_cloud(fr, wh, calib, box=0.4)
    slb200.decode_stats(fr)
    b = slb200.scan_to_cloud(fr, wh, calib, tables=tab)
    hc, vc, hi, vi = slb200.decode_codes(fr)
    slb200.codes_to_pixels(torch.stack([hc, hc]).contiguous(), torch.stack([vc, vc]).contiguous())
    cam_p, proj_p, rgb = slb200.cam_proj_pairs(hi, vi, proj, wh)
    pts = slb200.triangulate_points(cam_p, proj_p, calib)
    slb200.filter_box(pts, rgb.to(torch.float64), 5.0)
    slb200.decode_pixels(fr, eps=0.5)
    slb200.bgr_to_gray(wh); slb200.pair_diff_counts(fr)
    # decode straight from BGR frames (TMA path for the aligned size, fp64-compare kernel for the ragged one)
    bgr = fr[..., None].repeat(1, 1, 1, 3).contiguous()
    bgr[..., 0] += 3
    hb, vb = slb200.decode_pixels_bgr(bgr)
    slb200.scan_to_cloud(torch.stack([bgr, bgr]).contiguous(), torch.stack([wh, wh]).contiguous(), calib, tables=tab, bgr=True)
    # cloud clean-up: every level of the grid (surface + sparse far points + coincident points + a NaN record)
    n = int(b.counts[-1])
    cloud = torch.cat([b.cloud[:n], torch.rand((40, 4), device="cuda") * 8 - 4, torch.zeros((50, 4), device="cuda"),
                       torch.full((1, 4), float("nan"), device="cuda")])
    inl, ind = slb200.remove_statistical_outliers(cloud, 20, 0.5)
    slb200.remove_statistical_outliers(cloud, 40, 1.0)
    slb200.filter_cloud_box(cloud, 0.5)
    torch.cuda.synchronize()
    print(cam, "points", n, "inliers", len(ind))
print("done")
