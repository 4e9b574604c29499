"""One statistical outlier removal on the cloud of a 1080p synthetic scan (for ncu launch lists)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3dscanner-graycode_b200")); sys.path.insert(0, ROOT)
import torch, slb200
from slb200 import synth
dev = torch.device("cuda", 0)
cam = proj = (1920, 1080)
calib = synth.synthetic_calibration(cam, proj)
frames, white = synth.render_scan(cam, proj, seed=3, device=dev, calib=calib)
bufs = slb200.scan_to_cloud(frames, white, calib, tables=slb200.LensTables(calib, cam, dev))
cloud = slb200.filter_cloud_box(bufs.cloud[: int(bufs.counts[1].item())], 2.0)
torch.cuda.synchronize()
inl, ind = slb200.remove_statistical_outliers(cloud, 20, 0.5)
torch.cuda.synchronize()
print(len(cloud), len(inl))
