"""Seeded synthetic scans: a textured, tilted plane about 0.9 m from a
camera/projector pair, captured while the projector shows the Gray-code sequence.

Input generation only -- not part of the decode/triangulate path.  Written with torch ops so
large batches can be rendered directly in HBM; runs on CPU tensors as well (tests).
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np
import torch

from . import calib_data
from .api import Calibration


CAM_ZOOM = 4.0


def rodrigues(rvec) -> np.ndarray:
    r = np.asarray(rvec, dtype=np.float64).reshape(3)
    theta = float(np.linalg.norm(r))
    if theta < 1e-15:
        return np.eye(3)
    k = r / theta
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + math.sin(theta) * K + (1 - math.cos(theta)) * (K @ K)


def synthetic_calibration(cam_size: Tuple[int, int], proj_size: Tuple[int, int]) -> Calibration:
    """cam_1080 intrinsics scaled to ``cam_size`` with a CAM_ZOOM x longer focal length (the
    shipped projector lens throws a +-11 degree beam; the zoom puts the camera's field of view
    inside the projected patch, as in a real scan, so ~7/8 of the pixels decode); shipped projector
    intrinsics (valid for a 1024x768 panel) scaled to ``proj_size`` as Triangulate.__init__ does;
    synthetic stereo pose with the camera 0.2 m beside the projector, toed in on the patch."""
    cam = calib_data.CAM_1080_MTX.copy()
    cam[0, :] *= cam_size[0] / 1920.0
    cam[1, :] *= cam_size[1] / 1080.0
    cam[0, 0] *= CAM_ZOOM
    cam[1, 1] *= CAM_ZOOM
    R = rodrigues([-0.025, -0.237, 0.01])
    T = np.array([[0.20], [0.02], [0.03]])
    return Calibration.from_reference(cam, calib_data.CAM_1080_DIST, calib_data.PROJ_MTX, calib_data.PROJ_DIST,
                                      R, T, proj_size=proj_size, proj_calib_size=calib_data.PROJ_CALIB_SIZE)


def pattern_stack(proj_size: Tuple[int, int], device) -> torch.Tensor:
    import os, sys
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if here not in sys.path:
        sys.path.insert(0, here)
    from scanner.grayCode.generate_codes import get_gray_codes, get_image_sequence
    pw, ph = proj_size
    seq = get_image_sequence(get_gray_codes(pw, ph), pw, ph)
    return torch.from_numpy(seq).to(device)


def render_scan(cam_size: Tuple[int, int], proj_size: Tuple[int, int], seed: int = 0, device="cpu",
                patterns: torch.Tensor = None, calib: Calibration = None, scene: str = "plane"):
    """Returns ``frames (F,H,W) uint8``, ``white_rgb (H,W,3) uint8`` on ``device``.

    frame = clip(ambient + albedo * 0.85 * pattern[v,u] + N(0,1.5)) where (u,v) is the projector
    pixel the camera ray hits on the plane; pixels that miss the panel see ambient light only; the
    left eighth of the image is a dark noise band that exercises the invalid path.

    ``scene="hard"``: the same geometry under conditions that defeat the decoder's single-threshold fast pass --
    dark surface (albedo U[0.08, 0.45]), pattern contrast 0.5, an inter-reflection term (0.3 x the 31-pixel
    box-blurred pattern: light from neighbouring stripes, strongest for the fine patterns) and N(0, 4) noise.
    """
    W, H = cam_size
    pw, ph = proj_size
    dev = torch.device(device)
    if patterns is None:
        patterns = pattern_stack(proj_size, dev)
    if calib is None:
        calib = synthetic_calibration(cam_size, proj_size)
    F = patterns.shape[0]
    g = torch.Generator(device=dev)
    g.manual_seed(1000003 * int(seed) + 17)
    f64 = torch.float64
    Kc = torch.tensor(calib.cam_mtx, dtype=f64, device=dev)
    Kp = torch.tensor(calib.proj_mtx, dtype=f64, device=dev)
    R = torch.tensor(calib.proj_R, dtype=f64, device=dev)
    T = torch.tensor(np.asarray(calib.proj_T).reshape(3), dtype=f64, device=dev)
    # plane n.X = d in the camera frame; the pose varies a little with the seed (turntable views)
    rs = np.random.default_rng(seed)
    nrm = np.array([0.1, -0.05, 1.0]) + (rs.uniform(-0.05, 0.05, 3) * [1, 1, 0] if seed else 0)
    dist = 0.9 + (rs.uniform(-0.03, 0.03) if seed else 0.0)
    n_t = torch.tensor(nrm, dtype=f64, device=dev)

    ys, xs = torch.meshgrid(torch.arange(H, dtype=f64, device=dev), torch.arange(W, dtype=f64, device=dev), indexing="ij")
    dx = (xs - Kc[0, 2]) / Kc[0, 0]
    dy = (ys - Kc[1, 2]) / Kc[1, 1]
    t = dist / (n_t[0] * dx + n_t[1] * dy + n_t[2])
    X = torch.stack([dx * t, dy * t, t], dim=0)                    # camera frame
    Xp = torch.einsum("ij,jhw->ihw", R, X) + T[:, None, None]       # projector frame
    u = torch.floor(Kp[0, 0] * Xp[0] / Xp[2] + Kp[0, 2]).long()
    v = torch.floor(Kp[1, 1] * Xp[1] / Xp[2] + Kp[1, 2]).long()
    inside = (u >= 0) & (u < pw) & (v >= 0) & (v < ph) & (Xp[2] > 0)
    uc, vc = u.clamp(0, pw - 1), v.clamp(0, ph - 1)
    lit = patterns[:, vc, uc].to(torch.float32)                     # (F,H,W)
    lit = lit * inside.to(torch.float32)[None]
    if scene == "hard":
        albedo = 0.08 + 0.37 * torch.rand((H, W), generator=g, device=dev, dtype=torch.float32)
        noise = 4.0 * torch.randn((F, H, W), generator=g, device=dev, dtype=torch.float32)
        glob = torch.nn.functional.avg_pool2d(lit[None], 31, stride=1, padding=15)[0]
        frames = (10.0 + albedo[None] * (0.5 * lit + 0.3 * glob) + noise).round().clamp(0, 255).to(torch.uint8)
        del glob
    elif scene == "plane":
        albedo = 0.3 + 0.7 * torch.rand((H, W), generator=g, device=dev, dtype=torch.float32)
        noise = 1.5 * torch.randn((F, H, W), generator=g, device=dev, dtype=torch.float32)
        frames = (10.0 + albedo[None] * 0.85 * lit + noise).round().clamp(0, 255).to(torch.uint8)
    else:
        raise ValueError("scene must be 'plane' or 'hard'")
    band = W // 8
    frames[:, :, :band] = torch.randint(0, 12, (F, H, band), generator=g, device=dev, dtype=torch.uint8)
    white = frames[1]
    tint = torch.rand((H, W, 3), generator=g, device=dev, dtype=torch.float32) * 0.3 + 0.7
    white_rgb = (white[..., None].to(torch.float32) * tint).round().clamp(0, 255).to(torch.uint8)
    return frames.contiguous(), white_rgb.contiguous()
