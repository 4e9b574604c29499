"""CPU oracle for the Gray-code decode + triangulation hot path.

TEST INFRASTRUCTURE ONLY.  This module is a NumPy restatement of the reference
algorithm (guillaume-charron/3DScanner-GrayCode).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker or the timed CPU
baseline -- never as a product code path.  The product (``slb200`` /
``scanner``) calls the CUDA extension and fails loudly when it is missing.

Parity status: PINNED.  Every function below is checked against the unmodified
reference imported from /root/reference (``tests/test_oracle_vs_reference.py``,
runs where the reference is mounted) and against golden vectors generated from
that reference and committed under ``tests/golden/`` (``make_golden.py``).

All ``file:line`` citations are relative to the reference repository root.
Third-party arithmetic on the path: ``cv2.undistortPoints`` (pinned
opencv-contrib-python==4.8.0.76 in requirements.txt:2; not vendored in the
reference).  ``undistort_points`` restates OpenCV's published iterative
algorithm; it is pinned against the installed cv2 in the golden vectors.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "get_direct_indirect", "get_is_lit", "get_codes", "merge_runs",
    "gray_decode", "gray_to_decimal", "codes_to_pixels", "decode_pixels",
    "cam_proj_pairs", "undistort_points", "scale_proj_mtx", "triangulate",
    "filter_3d_pts", "scan_to_cloud",
]


# --------------------------------------------------------------------------
# stage 1: decode
# --------------------------------------------------------------------------
def _n_bits(num_frames: int) -> int:
    if num_frames < 14 or (num_frames - 2) % 4:
        raise ValueError("frame stack must hold 4n+2 frames with n >= 3")
    return (num_frames - 2) // 4


def get_direct_indirect(images):
    """Direct / global illumination estimate per pixel.

    scanner/grayCode/decode_codes.py:90-122.  ``images`` is the float64
    ``(4n+2, H, W)`` stack; frame 0 is black, frame 1 white (:106-107).  L_max is
    taken over the three finest column patterns and their inverses (:110,116),
    L_min over the three finest row patterns and their inverses (:111,117).
    ``white + black == 0`` yields NaN exactly as in the reference (:113).
    """
    images = np.asarray(images)
    n = _n_bits(len(images))
    black, white = images[0], images[1]
    pats = images[2:]
    hi_ids = [2 * n - 2, 2 * n - 4, 2 * n - 6, 4 * n - 2, 4 * n - 4, 4 * n - 6]
    lo_ids = [1, 3, 5, 2 * n + 1, 2 * n + 3, 2 * n + 5]
    with np.errstate(divide="ignore", invalid="ignore"):
        b_inv = white / (white + black)
        L_max = pats[hi_ids].max(axis=0)
        L_min = pats[lo_ids].min(axis=0)
        L_d = (L_max - L_min) * b_inv
        L_g = 2.0 * (L_max - L_d) * b_inv
    return L_d, L_g


def _classify(norm, inv, L_d, L_g, eps):
    """Tri-state bit per (bit, pixel); closed form of decode_codes.py:162-182.

    The reference assigns in sequence, so the LAST matching rule wins:
    r4 -> 1, else r3 -> 0, else r2 -> 0, else r1 -> 1, else -1.  The
    ``L_d < m`` assignment (:169-170) writes -1 over -1 and is dead.
    """
    with np.errstate(invalid="ignore"):
        a = L_d > (L_g + eps)
        r1 = a & (norm > (inv + eps))                      # :172-173
        r2 = a & ((norm + eps) < inv)                      # :175-176
        r3 = ((norm + eps) < L_d) & (inv > (L_g + eps))    # :178-179
        r4 = (norm > (L_g + eps)) & ((inv + eps) < L_d)    # :181-182
    out = np.where(r4, 1, np.where(r3 | r2, 0, np.where(r1, 1, -1)))
    return out.astype(np.int8)


def get_is_lit(images, L_d, L_g, eps=1, m=10):
    """decode_codes.py:125-186.  ``m`` is accepted and has no effect (see above).

    Even pattern indices are the column ("h") bits MSB-first, odd indices the
    row ("v") bits LSB-first (:154-160); inverses follow after 2n frames (:151-152).
    """
    images = np.asarray(images)
    n = _n_bits(len(images))
    pats = images[2:]
    h_norm, v_norm = pats[0:2 * n:2], pats[1:2 * n:2]
    h_inv, v_inv = pats[2 * n::2], pats[2 * n + 1::2]
    return (_classify(h_norm, h_inv, L_d[None], L_g[None], eps),
            _classify(v_norm, v_inv, L_d[None], L_g[None], eps))


def get_codes(images):
    """decode_codes.py:231-248."""
    L_d, L_g = get_direct_indirect(images)
    return get_is_lit(images, L_d, L_g)


def merge_runs(code_list):
    """Per-bit max over capture runs (1 > 0 > -1); src/3-capture_decode.py:95-96."""
    return np.max(np.asarray(code_list), axis=0)


def gray_decode(n: int) -> int:
    """Gray code -> binary by XOR prefix; decode_codes.py:189-207."""
    n = int(n)
    shift = 1
    while (n >> shift) and shift < 64:
        n ^= n >> shift
        shift <<= 1
    return n


def gray_to_decimal(bits) -> int:
    """decode_codes.py:209-229: -1 if any bit is -1, else MSB-first Gray -> int."""
    value = 0
    for b in bits:
        b = int(b)
        if b < 0:
            return -1
        value = (value << 1) | b
    return gray_decode(value)


def codes_to_pixels(h_codes, v_codes):
    """Vectorised form of the two per-pixel loops at src/3-capture_decode.py:99-100.

    ``h`` bits are MSB-first; ``v`` bits are stored LSB-first and flipped by the
    caller (:100).  Returns int64 ``(H, W)`` arrays with -1 where any bit is -1.
    """
    def one(codes, msb_first):
        codes = np.asarray(codes)
        n = codes.shape[0]
        bad = (codes < 0).any(axis=0)
        order = range(n) if msb_first else range(n - 1, -1, -1)
        g = np.zeros(codes.shape[1:], dtype=np.int64)
        for k in order:
            g = (g << 1) | (codes[k] > 0)
        s = 1
        while s < 64:
            g ^= g >> s
            s <<= 1
        g[bad] = -1
        return g
    return one(h_codes, True), one(v_codes, False)


def decode_pixels(frames, runs=None):
    """get_codes [+ run merge] + codes_to_pixels for a ``(F,H,W)`` stack, or a
    list of stacks (``runs``) merged as src/3-capture_decode.py:75-100 does."""
    stacks = [frames] if runs is None else list(runs)
    hs, vs = zip(*(get_codes(np.asarray(s, dtype=np.float64)) for s in stacks))
    return codes_to_pixels(merge_runs(hs), merge_runs(vs))


# --------------------------------------------------------------------------
# stage 2: triangulation
# --------------------------------------------------------------------------
def cam_proj_pairs(h_pixels, v_pixels, proj_size, img_white=None):
    """Vectorised triangulate.py:39-71.  x-major order (outer loop over camera
    x, inner over y, :52-53); projector coordinates clamped to the panel (:60-61);
    colour = img_white[y, x] / 255 as float64 (:64,69)."""
    h = np.asarray(h_pixels)
    v = np.asarray(v_pixels)
    proj_w, proj_h = proj_size
    valid_t = ((h != -1) & (v != -1)).T          # (W, H): x-major
    xs, ys = np.nonzero(valid_t)
    cam = np.stack([xs, ys], axis=1).astype(np.float32)
    proj = np.stack([np.minimum(proj_w - 1, h[ys, xs]),
                     np.minimum(proj_h - 1, v[ys, xs])], axis=1).astype(np.float32)
    colors = None
    if img_white is not None:
        colors = np.asarray(img_white)[ys, xs, :].astype(np.float64) / 255.0
    return cam.reshape(-1, 2), proj.reshape(-1, 2), colors


def undistort_points(pts, K, dist, R=None, iters=5):
    """OpenCV ``undistortPoints`` (default criteria: 5 fixed-point iterations).

    Called at triangulate.py:84-85.  float32 points in, float64 arithmetic,
    float32 normalised points out.  ``dist`` = (k1,k2,p1,p2[,k3[,k4,k5,k6]]).
    """
    pts = np.asarray(pts, dtype=np.float32).reshape(-1, 2).astype(np.float64)
    K = np.asarray(K, dtype=np.float64)
    k = np.zeros(8)
    d = np.asarray(dist, dtype=np.float64).ravel()
    k[:min(8, d.size)] = d[:8]
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    ifx, ify = 1.0 / fx, 1.0 / fy
    x0 = (pts[:, 0] - cx) * ifx
    y0 = (pts[:, 1] - cy) * ify
    x, y = x0.copy(), y0.copy()
    live = np.ones(x.shape, dtype=bool)
    with np.errstate(all="ignore"):
        for _ in range(iters):
            r2 = x * x + y * y
            icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / \
                     (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2)
            bail = live & (icdist < 0)
            x = np.where(bail, x0, x)
            y = np.where(bail, y0, y)
            live &= ~bail
            dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x)
            dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y
            xn = (x0 - dx) * icdist
            yn = (y0 - dy) * icdist
            x = np.where(live, xn, x)
            y = np.where(live, yn, y)
    if R is not None:
        R = np.asarray(R, dtype=np.float64)
        xx = R[0, 0] * x + R[0, 1] * y + R[0, 2]
        yy = R[1, 0] * x + R[1, 1] * y + R[1, 2]
        ww = 1.0 / (R[2, 0] * x + R[2, 1] * y + R[2, 2])
        x, y = xx * ww, yy * ww
    return np.stack([x, y], axis=1).astype(np.float32)


def scale_proj_mtx(proj_mtx, proj_size, proj_calib_size):
    """Rows 0/1 of the projector matrix scaled to the panel size; triangulate.py:28-33
    (the reference does this in place on the caller's array; this returns a copy)."""
    out = np.array(proj_mtx, dtype=np.float64, copy=True)
    out[0, :] *= proj_size[0] / proj_calib_size[0]
    out[1, :] *= proj_size[1] / proj_calib_size[1]
    return out


def triangulate(cam_pts, proj_pts, cam_mtx, cam_dist, proj_mtx_scaled, proj_dist, proj_R, proj_T):
    """Law-of-sines triangulation; triangulate.py:73-97.

    dtype flow follows the reference: undistorted homogeneous rays are float32
    (:84-85), the normalised camera ray stays float32 (:90), the angles and the
    result are float64 (:91-95).  Output ``(3, N)`` float64.
    """
    c2 = undistort_points(cam_pts, cam_mtx, cam_dist, R=proj_R)
    p2 = undistort_points(proj_pts, proj_mtx_scaled, proj_dist)
    one = np.ones((1, c2.shape[0]), dtype=np.float32)
    c = np.concatenate([c2.T, one], axis=0)               # (3,N) f32
    p = np.concatenate([p2.T, one], axis=0)
    T = np.asarray(proj_T, dtype=np.float64)[:, 0]
    t_len = np.linalg.norm(T)
    with np.errstate(all="ignore"):
        normed = c / np.linalg.norm(c, axis=0)            # f32
        alpha = np.arccos(np.dot(-T, normed) / t_len)
        beta = np.arccos(np.dot(T, p) / (t_len * np.linalg.norm(p, axis=0)))
        gamma = np.pi - alpha - beta
        p_len = t_len * np.sin(beta) / np.sin(gamma)
        return normed * p_len


def filter_3d_pts(pts, colors, threshold=.5):
    """Strict box filter |x|,|y|,|z| < threshold; triangulate.py:99-122."""
    pts = np.asarray(pts)
    keep = np.all((pts < threshold) & (pts > -threshold), axis=0)
    return pts[:, keep], np.asarray(colors)[keep]


def scan_to_cloud(frames, img_white, calib, runs=None):
    """Whole path for one scan: decode -> pairs -> triangulate.

    ``calib`` is a dict with cam_mtx, cam_dist, proj_mtx (unscaled), proj_dist,
    proj_R, proj_T, proj_size, proj_calib_size.  Mirrors src/3-capture_decode.py:75-100
    followed by src/4-triangulate.py:50-64.
    """
    h_px, v_px = decode_pixels(frames, runs)
    cam, proj, colors = cam_proj_pairs(h_px, v_px, calib["proj_size"], img_white)
    pm = scale_proj_mtx(calib["proj_mtx"], calib["proj_size"], calib["proj_calib_size"])
    pts = triangulate(cam, proj, calib["cam_mtx"], calib["cam_dist"], pm,
                      calib["proj_dist"], calib["proj_R"], calib["proj_T"])
    return h_px, v_px, pts, colors
