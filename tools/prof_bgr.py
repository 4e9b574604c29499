"""A few decodes of a 1080p BGR stack (for ncu captures of decode_bgr_tma_kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3dscanner-graycode_b200")); sys.path.insert(0, ROOT)
import torch, slb200
from slb200 import synth
dev = torch.device("cuda", 0)
cam = proj = (1920, 1080)
calib = synth.synthetic_calibration(cam, proj)
frames, white = synth.render_scan(cam, proj, seed=3, device=dev, calib=calib)
g = torch.Generator(device=dev); g.manual_seed(1)
bgr = (frames[..., None].to(torch.int16) + torch.randint(-4, 5, (*frames.shape, 3), generator=g, device=dev, dtype=torch.int16))
bgr = bgr.clamp(0, 255).to(torch.uint8).contiguous()
for _ in range(3):
    h, v = slb200.decode_pixels_bgr(bgr)
torch.cuda.synchronize()
print(float((h >= 0).float().mean()))
