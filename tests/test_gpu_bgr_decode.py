"""GPU parity of the decode that reads the camera's BGR frames directly (sl_decode_bgr_u8): bit-exact against
`cv2-formula gray -> oracle decode` and against the two-kernel path (sl_bgr_to_gray_u8 + sl_decode_u8)."""
import os
import sys

import numpy as np
import pytest

from oracle import graycode_oracle as O

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_inputs import adversarial_stack  # noqa: E402


def gray_of(bgr):
    """cv2.cvtColor(BGR2GRAY) on uint8 (pinned to cv2 by tests/golden, test_gray_formula_matches_cv2_golden)."""
    b = bgr.astype(np.uint32)
    return ((3735 * b[..., 0] + 19235 * b[..., 1] + 9798 * b[..., 2] + 16384) >> 15).astype(np.uint8)


def colourise(rng, stack):
    """A BGR stack whose gray image is close to `stack` (so the decode sees realistic patterns) but whose three
    channels differ, plus pure random colours in the upper rows."""
    f, h, w = stack.shape
    bgr = np.repeat(stack[..., None], 3, axis=-1).astype(np.int32)
    bgr += rng.integers(-25, 26, bgr.shape)
    bgr = np.clip(bgr, 0, 255).astype(np.uint8)
    bgr[:, : max(1, h // 4)] = rng.integers(0, 256, (f, max(1, h // 4), w, 3), dtype=np.uint8)
    return bgr


def oracle_pixels(gray, eps=1):
    f64 = gray.astype(np.float64)
    l_d, l_g = O.get_direct_indirect(f64)
    hc, vc = O.get_is_lit(f64, l_d, l_g, eps=eps)
    return O.codes_to_pixels(hc, vc)


@pytest.mark.parametrize("n", [3, 8, 10, 11, 12, 15])
def test_bgr_decode_bit_exact(n):
    import torch
    import slb200
    rng = np.random.default_rng(300 + n)
    bgr = colourise(rng, adversarial_stack(rng, n, 40, 96))
    hp, vp = oracle_pixels(gray_of(bgr))
    t = torch.from_numpy(bgr).cuda()
    h, v = slb200.decode_pixels_bgr(t)
    np.testing.assert_array_equal(h.cpu().numpy(), hp)
    np.testing.assert_array_equal(v.cpu().numpy(), vp)
    h2, v2 = slb200.decode_pixels(slb200.bgr_to_gray(t))
    assert torch.equal(h, h2) and torch.equal(v, v2)


@pytest.mark.parametrize("shape", [(8, 16), (9, 17), (13, 31), (1, 1), (3, 5), (64, 1030), (129, 257)])
def test_bgr_decode_ragged_shapes(shape):
    """Pixel counts that are not multiples of 4 / 128 (tail pixels and partial tiles), unaligned frame pitches
    (the fp64-compare kernel converts on load)."""
    import torch
    import slb200
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    bgr = rng.integers(0, 256, (22, *shape, 3), dtype=np.uint8)
    hp, vp = oracle_pixels(gray_of(bgr))
    h, v = slb200.decode_pixels_bgr(torch.from_numpy(bgr).cuda())
    np.testing.assert_array_equal(h.cpu().numpy(), hp)
    np.testing.assert_array_equal(v.cpu().numpy(), vp)


@pytest.mark.parametrize("eps", [0, 3, 0.5])
def test_bgr_decode_eps(eps):
    import torch
    import slb200
    rng = np.random.default_rng(17)
    bgr = colourise(rng, adversarial_stack(rng, 9, 32, 64))
    hp, vp = oracle_pixels(gray_of(bgr), eps=eps)
    h, v = slb200.decode_pixels_bgr(torch.from_numpy(bgr).cuda(), eps=eps)
    np.testing.assert_array_equal(h.cpu().numpy(), hp)
    np.testing.assert_array_equal(v.cpu().numpy(), vp)


def test_bgr_batch_and_cloud_match_gray_path():
    """A batch of 1080p scans: the BGR path and the gray path give identical indices and identical clouds."""
    import torch
    import slb200
    from slb200 import synth
    calib = synth.synthetic_calibration((1920, 1080), (1920, 1080))
    pats = synth.pattern_stack((1920, 1080), "cuda")
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    grays, bgrs, whites = [], [], []
    for seed in (1, 2, 3):
        frames, white = synth.render_scan((1920, 1080), (1920, 1080), seed=seed, device="cuda", patterns=pats, calib=calib)
        bgr = frames[..., None].to(torch.int16) + torch.randint(-20, 21, (*frames.shape, 3), generator=g, device="cuda",
                                                                dtype=torch.int16)
        bgr = bgr.clamp(0, 255).to(torch.uint8).contiguous()
        bgrs.append(bgr)
        grays.append(slb200.bgr_to_gray(bgr))
        whites.append(white)
    bgr_b, gray_b, white_b = torch.stack(bgrs), torch.stack(grays), torch.stack(whites)
    h, v = slb200.decode_pixels_bgr(bgr_b)
    h2, v2 = slb200.decode_pixels(gray_b)
    assert torch.equal(h, h2) and torch.equal(v, v2)
    assert float((h >= 0).float().mean()) > 0.5
    tables = slb200.LensTables(calib, (1920, 1080), "cuda")
    a = slb200.scan_to_cloud(bgr_b, white_b, calib, tables=tables, bgr=True)
    b = slb200.scan_to_cloud(gray_b, white_b, calib, tables=tables)
    assert torch.equal(a.counts, b.counts)
    n = int(a.counts[3].item())
    assert torch.equal(a.cloud[:n].view(torch.int32), b.cloud[:n].view(torch.int32))


def test_bgr_argument_errors():
    import torch
    import slb200
    with pytest.raises(TypeError):
        slb200.decode_pixels_bgr(torch.zeros((14, 8, 8), dtype=torch.uint8, device="cuda"))
    with pytest.raises(slb200.SlError):
        slb200.decode_pixels_bgr(torch.zeros((13, 8, 8, 3), dtype=torch.uint8, device="cuda"))


def test_bgr_identity_full_size_4k():
    """BASELINE config 3 size (3840x2160, n = 12, F = 50; 1.24 GB of BGR frames): the projector sequence seen through
    an identity camera, painted in colour (every pixel keeps its gray level under cv2's formula only for grey
    colours, so the lit level is a grey 255 and the dark level a saturated blue whose gray value is 29), decodes
    to the identity map -- the reference's own known-answer property, at full size, through the BGR kernel."""
    import torch
    import slb200
    pats = slb200.pattern_sequence((3840, 2160))                       # (50, 2160, 3840) u8, 0 / 255
    lit = pats > 0
    bgr = torch.empty((*pats.shape, 3), dtype=torch.uint8, device="cuda")
    bgr[..., 0] = 255                                                  # B: 255 either way
    bgr[..., 1] = torch.where(lit, 255, 0)
    bgr[..., 2] = torch.where(lit, 255, 0)
    del lit
    h, v = slb200.decode_pixels_bgr(bgr)
    xs = torch.arange(3840, device="cuda", dtype=torch.int16)[None, :].expand(2160, 3840)
    ys = torch.arange(2160, device="cuda", dtype=torch.int16)[:, None].expand(2160, 3840)
    assert torch.equal(h, xs) and torch.equal(v, ys)


def test_pipeline_submit_bgr_matches_gray_submit():
    """sl_pipeline_submit_bgr (the camera's BGR frames in host memory, more scans than the pipeline has slots, mixed with
    gray submits on the same pipeline) == cv2-formula gray on the host + sl_pipeline_submit, bit for bit."""
    import torch
    import slb200
    from slb200 import synth
    cam_size, proj_size = (320, 200), (1024, 768)
    calib = synth.synthetic_calibration(cam_size, proj_size)
    rng = np.random.default_rng(11)
    scans = [synth.render_scan(cam_size, proj_size, seed=s, device="cpu") for s in range(5)]
    bgrs = [torch.from_numpy(colourise(rng, s[0].numpy())).pin_memory() for s in scans]
    grays = [torch.from_numpy(gray_of(b.numpy())).pin_memory() for b in bgrs]
    whites = [s[1].contiguous().pin_memory() for s in scans]
    f = bgrs[0].shape[0]
    pipe = slb200.Pipeline(f, 200, 320, max_batch=1)
    cl_b = [torch.zeros((200 * 320, 4), dtype=torch.float32).pin_memory() for _ in scans]
    cl_g = [torch.zeros((200 * 320, 4), dtype=torch.float32).pin_memory() for _ in scans]
    n_b = [torch.zeros(1, dtype=torch.int64) for _ in scans]
    n_g = [torch.zeros(1, dtype=torch.int64) for _ in scans]
    for i in range(5):
        pipe.submit(bgrs[i], whites[i], calib, cl_b[i], n_b[i], bgr=True)
        pipe.submit(grays[i], whites[i], calib, cl_g[i], n_g[i])
    pipe.flush()
    for i in range(5):
        n = int(n_g[i])
        assert int(n_b[i]) == n and n > 0
        assert torch.equal(cl_b[i][:n].view(torch.int32), cl_g[i][:n].view(torch.int32))
    hp, vp = oracle_pixels(grays[0].numpy())
    assert int(n_g[0]) == int(((hp != -1) & (vp != -1)).sum())
    with pytest.raises(ValueError):
        pipe.submit(grays[0], whites[0], calib, cl_b[0], n_b[0], bgr=True)        # a gray stack handed in as BGR
    with pytest.raises(ValueError):
        pipe.submit(bgrs[0], whites[0], calib, cl_b[0], n_b[0])
    pipe.close()
