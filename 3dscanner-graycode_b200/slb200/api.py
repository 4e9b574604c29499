"""Torch-tensor front end of the C ABI.

PyTorch is used for device memory, streams and dtype plumbing only; every function
here ends in a call into ``libslb200.so``.  Reference citations are relative to the
reference repository (guillaume-charron/3DScanner-GrayCode).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from ._lib import SL_PREC_FP32, SL_PREC_FP64, SlCalib, check, lib

_PREC = {"fp32": SL_PREC_FP32, "fp64": SL_PREC_FP64, SL_PREC_FP32: SL_PREC_FP32, SL_PREC_FP64: SL_PREC_FP64}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("slb200 needs a CUDA device (sm_100a); there is no CPU fallback")


def launch_count() -> int:
    """Kernels launched by the extension in this process."""
    return int(lib().sl_launch_count())


# --------------------------------------------------------------------------- calibration
@dataclass
class Calibration:
    """Calibration in the form the kernels take.  ``proj_mtx`` is ALREADY scaled to the
    panel size (scanner/triangulation/triangulate.py:28-33)."""
    cam_mtx: np.ndarray
    cam_dist: np.ndarray
    proj_mtx: np.ndarray
    proj_dist: np.ndarray
    proj_R: np.ndarray
    proj_T: np.ndarray
    proj_size: Tuple[int, int]

    @classmethod
    def from_reference(cls, cam_mtx, cam_dist, proj_mtx, proj_dist, proj_R, proj_T,
                       proj_size=(1920, 1080), proj_calib_size=(1920, 1080)) -> "Calibration":
        pm = np.array(proj_mtx, dtype=np.float64, copy=True)
        pm[0, :] = pm[0, :] * (proj_size[0] / proj_calib_size[0])     # triangulate.py:30-33
        pm[1, :] = pm[1, :] * (proj_size[1] / proj_calib_size[1])
        return cls(np.asarray(cam_mtx, np.float64), np.asarray(cam_dist, np.float64), pm,
                   np.asarray(proj_dist, np.float64), np.asarray(proj_R, np.float64),
                   np.asarray(proj_T, np.float64), (int(proj_size[0]), int(proj_size[1])))

    def struct(self, row0: int = 0) -> SlCalib:
        s = SlCalib()
        s.cam_row0 = int(row0)

        def fill(dst, src, n, what):
            a = np.asarray(src, dtype=np.float64).ravel()
            if a.size > n:
                raise ValueError(f"{what}: at most {n} values supported, got {a.size}")
            for i in range(n):
                dst[i] = float(a[i]) if i < a.size else 0.0

        def mat(dst, src, what):
            a = np.asarray(src, dtype=np.float64)
            if a.shape != (3, 3):
                raise ValueError(f"{what} must be 3x3")
            for i in range(9):
                dst[i] = float(a.flat[i])

        mat(s.cam_K, self.cam_mtx, "cam_mtx")
        mat(s.proj_K, self.proj_mtx, "proj_mtx")
        mat(s.R, self.proj_R, "proj_R")
        fill(s.cam_dist, self.cam_dist, 8, "cam_dist")
        fill(s.proj_dist, self.proj_dist, 8, "proj_dist")
        t = np.asarray(self.proj_T, dtype=np.float64).reshape(-1)
        if t.size != 3:
            raise ValueError("proj_T must hold 3 values")
        for i in range(3):
            s.T[i] = float(t[i])
        s.proj_w, s.proj_h = int(self.proj_size[0]), int(self.proj_size[1])
        return s


# --------------------------------------------------------------------------- inputs
def as_u8_stack(images, device=None) -> torch.Tensor:
    """Frame stack -> contiguous uint8 CUDA tensor.

    The reference feeds float64 frames holding integer grey levels
    (src/3-capture_decode.py:66-70).  Float input is accepted only if every value is an
    integer in 0..255 -- anything else would silently diverge from the reference's fp64
    arithmetic, so it raises instead.
    """
    _require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if isinstance(images, torch.Tensor):
        t = images
    else:
        a = np.asarray(images)
        if a.dtype == object:
            raise TypeError("frame stack must be numeric")
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.dtype == torch.uint8:
        return t.to(dev).contiguous()
    if t.dtype == torch.bool:
        raise TypeError("boolean frame stack")
    # Narrow where the data is: a float64 stack is 8x the bytes of the uint8 samples it holds (763 MB for one 1080p
    # scan), so a host stack is validated and narrowed frame by frame on the host (torch CPU ops, multi-threaded)
    # into one pinned uint8 buffer and only that crosses PCIe.
    if not t.is_cuda and t.dtype in (torch.float64, torch.float32) and t.is_contiguous():
        # the reference's own stack dtype: one multi-threaded host pass (convert + validate) into pinned staging
        u = _pinned_u8(t.shape)
        rc = lib().sl_host_narrow_u8(t.data_ptr(), 0 if t.dtype == torch.float64 else 1, t.numel(), u.data_ptr(), 0)
        if rc:
            raise ValueError("frame stack must hold integer grey levels in 0..255 "
                             "(the B200 path computes on uint8 samples)")
        return u.to(dev).contiguous()   # blocking: the pinned staging buffer is reused by the next call
    u = torch.empty(t.shape, dtype=torch.uint8, device=t.device) if t.is_cuda else _pinned_u8(t.shape)
    planes = t.reshape(-1, *t.shape[-2:]) if t.dim() >= 3 else t.reshape(1, -1)
    u_planes = u.reshape(planes.shape)
    for k in range(planes.shape[0]):
        src = planes[k]
        dst = u_planes[k]
        dst.copy_(src)          # plain cast: a value that is not an integer in 0..255 (NaN, 300, -1, 0.5) fails the
        if not torch.equal(dst.to(src.dtype), src):        # round trip below whatever the cast made of it
            raise ValueError("frame stack must hold integer grey levels in 0..255 "
                             "(the B200 path computes on uint8 samples)")
    return u.to(dev).contiguous()       # blocking: the pinned staging buffer is reused by the next call


_PINNED_STAGE = __import__("threading").local()


def _pinned_u8(shape) -> torch.Tensor:
    """Pinned uint8 staging buffer of the drop-in wrappers (one per host thread, grown on demand: pinning 95 MB costs
    tens of milliseconds, a capture loop narrows a stack of the same size every scan)."""
    n = 1
    for d in shape:
        n *= int(d)
    buf = getattr(_PINNED_STAGE, "buf", None)
    if buf is None or buf.numel() < n:
        buf = torch.empty(max(n, 1), dtype=torch.uint8, pin_memory=True)
        _PINNED_STAGE.buf = buf
    return buf[:n].view(*shape)


def _pinned_stage(name: str, nbytes: int) -> torch.Tensor:
    """Named pinned byte buffer of the calling host thread (grown on demand) for device -> host results."""
    buf = getattr(_PINNED_STAGE, name, None)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=True)
        setattr(_PINNED_STAGE, name, buf)
    return buf[:nbytes]


def idx_to_host_int64(idx: torch.Tensor) -> np.ndarray:
    """Device int16 index image(s) -> the reference's int64 NumPy array (decode_codes.py:189-229 returns Python ints
    that the script stores in int64 arrays): int16 over PCIe into pinned staging, widened by sl_host_widen_idx."""
    idx = idx.contiguous()
    stage = _pinned_stage("idx", idx.numel() * 2).view(torch.int16)
    stage.copy_(idx.reshape(-1))                    # blocking (pinned destination, synchronises the stream)
    out = np.empty(tuple(idx.shape), dtype=np.int64)
    check(lib().sl_host_widen_idx(stage.data_ptr(), idx.numel(), out.ctypes.data, 0))
    return out


def cloud_to_host_f64(cloud: torch.Tensor, colors: bool = True):
    """Device ``(n,4)`` packed records -> ``pts (3,n) float64[, colors (n,3) float64]`` NumPy arrays, the return types
    of Triangulate.triangulate / get_cam_proj_pts (triangulate.py:64-69,95): 16 bytes per point over PCIe into pinned
    staging, unpacked by sl_host_unpack_cloud (exact widening; colours = channel / 255 as an IEEE division)."""
    n = int(cloud.shape[0])
    stage = _pinned_stage("cloud", n * 16).view(torch.float32).view(n, 4)
    stage.copy_(cloud)
    pts = np.empty((3, n), dtype=np.float64)
    col = np.empty((n, 3), dtype=np.float64) if colors else None
    check(lib().sl_host_unpack_cloud(stage.data_ptr(), n, pts.ctypes.data, col.ctypes.data if colors else None, 0))
    return (pts, col) if colors else pts


def _stack_dims(frames: torch.Tensor):
    if frames.dim() == 3:
        f, h, w = frames.shape
        return 1, f, h, w, False
    if frames.dim() == 4:
        b, f, h, w = frames.shape
        return b, f, h, w, True
    raise ValueError("frames must be (F,H,W) or (B,F,H,W)")


def _pitches(frames: torch.Tensor):
    """(frame_pitch, scan_pitch) in bytes.  Row bands of a resident stack (``stack[:, y0:y1]``) are
    accepted as they are: rows must be dense, frames / scans may be strided."""
    b, f, h, w, batched = _stack_dims(frames)
    st = frames.stride()
    frame_pitch = st[-3] if f > 1 else h * w
    scan_pitch = (st[0] if b > 1 else f * frame_pitch) if batched else f * frame_pitch
    return int(frame_pitch), int(scan_pitch)


def _check_stack(frames: torch.Tensor):
    if not (isinstance(frames, torch.Tensor) and frames.is_cuda and frames.dtype == torch.uint8):
        raise TypeError("frames must be a CUDA uint8 tensor (use as_u8_stack)")
    b, f, h, w, batched = _stack_dims(frames)
    st = frames.stride()
    rows_dense = (st[-1] == 1 or w == 1) and (st[-2] == w or h == 1)
    fp, sp = _pitches(frames)
    if not rows_dense or fp < h * w or (batched and b > 1 and sp < f * fp):
        raise ValueError("frames must be contiguous, or a row band of a contiguous stack")


# --------------------------------------------------------------------------- capture front end
def pattern_sequence(proj_size: Tuple[int, int], device=None) -> torch.Tensor:
    """The projector's Gray-code frame sequence ``(4n+2, proj_h, proj_w)`` uint8, generated on the device
    (same frames as generate_codes.get_image_sequence, reference generate_codes.py:34-81)."""
    _require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    pw, ph = int(proj_size[0]), int(proj_size[1])
    n = max(1, int(np.ceil(np.log2(max(pw, ph)))))
    out = torch.empty((4 * n + 2, ph, pw), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib().sl_pattern_sequence_u8(pw, ph, _ptr(out), _stream()))
    return out


def bgr_to_gray(bgr: torch.Tensor) -> torch.Tensor:
    """cv2.cvtColor(BGR2GRAY) of uint8 CUDA frames ``(..., 3)`` -> ``(...)`` uint8
    (src/3-capture_decode.py:66, decode_codes.py:86)."""
    if not (bgr.is_cuda and bgr.dtype == torch.uint8 and bgr.shape[-1] == 3):
        raise TypeError("bgr must be a CUDA uint8 tensor with a trailing dimension of 3")
    bgr = bgr.contiguous()
    gray = torch.empty(bgr.shape[:-1], dtype=torch.uint8, device=bgr.device)
    with torch.cuda.device(bgr.device):
        check(lib().sl_bgr_to_gray_u8(_ptr(bgr), gray.numel(), _ptr(gray), _stream()))
    return gray


def pair_diff_counts(frames: torch.Tensor, thresh: int = 50) -> torch.Tensor:
    """Number of elements with ``|frames[k+1] - frames[k]| > thresh`` for every consecutive pair
    (decode_codes.py:49-54).  frames: CUDA uint8 ``(N, ...)``; returns int64 ``(N-1,)`` (CUDA)."""
    if not (frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() >= 2):
        raise TypeError("frames must be a CUDA uint8 tensor (N, ...)")
    frames = frames.contiguous()
    n = frames.shape[0]
    elems = frames[0].numel()
    counts = torch.zeros(max(n - 1, 0), dtype=torch.int64, device=frames.device)
    if n >= 2:
        with torch.cuda.device(frames.device):
            check(lib().sl_pair_diff_counts(_ptr(frames), n, elems, elems, int(thresh), _ptr(counts), _stream()))
    return counts


# --------------------------------------------------------------------------- stage 1
def decode_pixels(frames: torch.Tensor, eps: float = 1.0, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                  sm_reserve: int = 0):
    """Fused decode: ``(F,H,W)`` / ``(B,F,H,W)`` uint8 -> ``h_pixels, v_pixels`` int16, -1 invalid.

    Replaces get_codes + the per-pixel gray_to_decimal loops
    (scanner/grayCode/decode_codes.py:231-248, src/3-capture_decode.py:99-100).
    """
    _check_stack(frames)
    b, f, h, w, batched = _stack_dims(frames)
    shape = (b, h, w) if batched else (h, w)
    if out is None:
        h_idx = torch.empty(shape, dtype=torch.int16, device=frames.device)
        v_idx = torch.empty(shape, dtype=torch.int16, device=frames.device)
    else:
        h_idx, v_idx = out
    with torch.cuda.device(frames.device):
        fp, sp = _pitches(frames)
        check(lib().sl_decode_u8(_ptr(frames), b, f, h, w, fp, sp, float(eps),
                                 _ptr(h_idx), _ptr(v_idx), None, None, None, int(sm_reserve), _stream()))
    return h_idx, v_idx


def decode_stats(frames: torch.Tensor, eps: float = 1.0) -> dict:
    """Diagnostics of the fused decode on this input (``sl_decode_u8_stats``): how many 4-pixel groups needed the
    full evaluation of the four rules instead of the single-threshold fast pass.  Synchronises."""
    _check_stack(frames)
    b, f, h, w, batched = _stack_dims(frames)
    shape = (b, h, w) if batched else (h, w)
    h_idx = torch.empty(shape, dtype=torch.int16, device=frames.device)
    v_idx = torch.empty(shape, dtype=torch.int16, device=frames.device)
    stats = torch.zeros(2, dtype=torch.int64, device=frames.device)
    with torch.cuda.device(frames.device):
        fp, sp = _pitches(frames)
        check(lib().sl_decode_u8_stats(_ptr(frames), b, f, h, w, fp, sp, float(eps), _ptr(h_idx), _ptr(v_idx),
                                       _ptr(stats), _stream()))
    quads, full = (int(x) for x in stats.tolist())
    return {"quads": quads, "full_rule_quads": full, "full_rule_fraction": (full / quads) if quads else None,
            "fast_pass_fraction": (1.0 - full / quads) if quads else None}


def _bgr_view(frames_bgr: torch.Tensor):
    """``([B,]F,H,W,3)`` uint8 CUDA -> (b, f, h, w, batched, frame_pitch, scan_pitch); dense frames only."""
    if not (isinstance(frames_bgr, torch.Tensor) and frames_bgr.is_cuda and frames_bgr.dtype == torch.uint8
            and frames_bgr.dim() in (4, 5) and frames_bgr.shape[-1] == 3):
        raise TypeError("frames_bgr must be a CUDA uint8 tensor ([B,]F,H,W,3)")
    if not frames_bgr.is_contiguous():
        raise ValueError("frames_bgr must be contiguous")
    batched = frames_bgr.dim() == 5
    b = frames_bgr.shape[0] if batched else 1
    f, h, w = frames_bgr.shape[-4:-1]
    return b, f, h, w, batched, 3 * h * w, 3 * h * w * f


def decode_pixels_bgr(frames_bgr: torch.Tensor, eps: float = 1.0):
    """decode_pixels on the camera's BGR frames ``([B,]F,H,W,3)``: the per-frame
    ``cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)`` of the capture loop (src/3-capture_decode.py:66) is fused into the
    decode's load, no gray stack is materialised."""
    b, f, h, w, batched, fp, sp = _bgr_view(frames_bgr)
    shape = (b, h, w) if batched else (h, w)
    h_idx = torch.empty(shape, dtype=torch.int16, device=frames_bgr.device)
    v_idx = torch.empty(shape, dtype=torch.int16, device=frames_bgr.device)
    with torch.cuda.device(frames_bgr.device):
        check(lib().sl_decode_bgr_u8(_ptr(frames_bgr), b, f, h, w, fp, sp, float(eps), _ptr(h_idx), _ptr(v_idx),
                                     None, 0, _stream()))
    return h_idx, v_idx


def decode_codes(frames: torch.Tensor, eps: float = 1.0):
    """get_codes (decode_codes.py:231-248): tri-state ``h_codes, v_codes`` int8 ``([B,]n,H,W)``
    plus the decoded indices (same launch)."""
    _check_stack(frames)
    b, f, h, w, batched = _stack_dims(frames)
    n = (f - 2) // 4
    cshape = (b, n, h, w) if batched else (n, h, w)
    ishape = (b, h, w) if batched else (h, w)
    dev = frames.device
    h_codes = torch.empty(cshape, dtype=torch.int8, device=dev)
    v_codes = torch.empty(cshape, dtype=torch.int8, device=dev)
    h_idx = torch.empty(ishape, dtype=torch.int16, device=dev)
    v_idx = torch.empty(ishape, dtype=torch.int16, device=dev)
    with torch.cuda.device(dev):
        fp, sp = _pitches(frames)
        check(lib().sl_decode_u8(_ptr(frames), b, f, h, w, fp, sp, float(eps),
                                 _ptr(h_idx), _ptr(v_idx), _ptr(h_codes), _ptr(v_codes), None, 0, _stream()))
    return h_codes, v_codes, h_idx, v_idx


def direct_indirect(frames: torch.Tensor):
    """get_direct_indirect (decode_codes.py:90-122): ``L_d, L_g`` float64 ``(H,W)``."""
    _check_stack(frames)
    b, f, h, w, batched = _stack_dims(frames)
    if batched:
        raise ValueError("direct_indirect takes one (F,H,W) stack")
    l_d = torch.empty((h, w), dtype=torch.float64, device=frames.device)
    l_g = torch.empty((h, w), dtype=torch.float64, device=frames.device)
    with torch.cuda.device(frames.device):
        check(lib().sl_direct_indirect(_ptr(frames), f, h, w, h * w, _ptr(l_d), _ptr(l_g), _stream()))
    return l_d, l_g


def is_lit(frames: torch.Tensor, l_d: torch.Tensor, l_g: torch.Tensor, eps: float = 1.0):
    """get_is_lit (decode_codes.py:125-186) with caller-supplied L_d / L_g."""
    _check_stack(frames)
    b, f, h, w, batched = _stack_dims(frames)
    if batched:
        raise ValueError("is_lit takes one (F,H,W) stack")
    n = (f - 2) // 4
    l_d = l_d.to(device=frames.device, dtype=torch.float64).contiguous()
    l_g = l_g.to(device=frames.device, dtype=torch.float64).contiguous()
    if tuple(l_d.shape) != (h, w) or tuple(l_g.shape) != (h, w):
        raise ValueError("L_d / L_g must be (H, W)")
    h_codes = torch.empty((n, h, w), dtype=torch.int8, device=frames.device)
    v_codes = torch.empty((n, h, w), dtype=torch.int8, device=frames.device)
    with torch.cuda.device(frames.device):
        check(lib().sl_is_lit(_ptr(frames), f, h, w, h * w, _ptr(l_d), _ptr(l_g), float(eps),
                              _ptr(h_codes), _ptr(v_codes), _stream()))
    return h_codes, v_codes


def codes_to_pixels(h_codes: torch.Tensor, v_codes: torch.Tensor):
    """Run merge + Gray->decimal (src/3-capture_decode.py:95-100).

    ``h_codes, v_codes``: int8 CUDA ``(n,H,W)`` or ``(R,n,H,W)`` (R capture runs).
    """
    for t in (h_codes, v_codes):
        if not (t.is_cuda and t.dtype == torch.int8 and t.is_contiguous()):
            raise TypeError("codes must be contiguous CUDA int8 tensors")
    if h_codes.shape != v_codes.shape:
        raise ValueError("h_codes and v_codes differ in shape")
    if h_codes.dim() == 3:
        r, (n, h, w) = 1, h_codes.shape
    elif h_codes.dim() == 4:
        r, n, h, w = h_codes.shape
    else:
        raise ValueError("codes must be (n,H,W) or (R,n,H,W)")
    h_idx = torch.empty((h, w), dtype=torch.int16, device=h_codes.device)
    v_idx = torch.empty((h, w), dtype=torch.int16, device=h_codes.device)
    with torch.cuda.device(h_codes.device):
        check(lib().sl_codes_to_pixels(_ptr(h_codes), _ptr(v_codes), r, n, h * w, n * h * w,
                                       _ptr(h_idx), _ptr(v_idx), _stream()))
    return h_idx, v_idx


# --------------------------------------------------------------------------- stage 2
def _idx16(t, device) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        a = np.asarray(t)
        if a.dtype == np.int64 and a.flags.c_contiguous and a.size:
            # the reference's h_pixels / v_pixels dtype: range check + narrowing in one multi-threaded host pass
            out = np.empty(a.shape, dtype=np.int16)
            if lib().sl_host_narrow_idx(a.ctypes.data, a.size, out.ctypes.data, 0):
                raise ValueError("decoded indices must lie in -1..32767")
            return torch.from_numpy(out).to(device).contiguous()
        if a.size and (a.min() < -1 or a.max() > 32767):
            raise ValueError("decoded indices must lie in -1..32767")
        t = torch.from_numpy(np.ascontiguousarray(a.astype(np.int16)))
    elif t.dtype != torch.int16:
        if t.numel() and (int(t.min()) < -1 or int(t.max()) > 32767):
            raise ValueError("decoded indices must lie in -1..32767")
        t = t.to(torch.int16)
    return t.to(device).contiguous()


class _Offsets:
    """Workspace + counts of one ordered compaction."""

    def __init__(self, h_idx: torch.Tensor, v_idx: torch.Tensor):
        if h_idx.shape != v_idx.shape:
            raise ValueError("h_pixels and v_pixels differ in shape")
        if h_idx.dim() == 2:
            self.b, (self.h, self.w) = 1, h_idx.shape
        elif h_idx.dim() == 3:
            self.b, self.h, self.w = h_idx.shape
        else:
            raise ValueError("indices must be (H,W) or (B,H,W)")
        dev = h_idx.device
        nbytes = lib().sl_cloud_workspace_bytes(self.b, self.h, self.w)
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.counts = torch.empty(self.b + 1, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            check(lib().sl_compact_offsets(_ptr(h_idx), _ptr(v_idx), self.b, self.h, self.w,
                                           _ptr(self.ws), _ptr(self.counts), _stream()))


def cam_proj_pairs(h_pixels, v_pixels, proj_size, img_white=None, device=None):
    """Triangulate.get_cam_proj_pts (triangulate.py:39-71) for one image.

    Returns CUDA tensors ``cam_pts (N,2) f32, proj_pts (N,2) f32, rgb (N,3) u8 | None``
    in the reference's x-major order.
    """
    _require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    h_idx, v_idx = _idx16(h_pixels, dev), _idx16(v_pixels, dev)
    if h_idx.dim() != 2:
        raise ValueError("cam_proj_pairs takes one (H,W) image")
    white = None
    if img_white is not None:
        white = img_white if isinstance(img_white, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(img_white))
        if white.dtype != torch.uint8 or tuple(white.shape) != (*h_idx.shape, 3):
            raise ValueError("img_white must be uint8 (H,W,3)")
        white = white.to(dev).contiguous()
    off = _Offsets(h_idx, v_idx)
    n = int(off.counts[1].item())
    pairs = torch.empty((n, 4), dtype=torch.float32, device=dev)
    rgb32 = torch.empty((n,), dtype=torch.int32, device=dev) if white is not None else None
    if n:
        with torch.cuda.device(dev):
            check(lib().sl_emit_pairs(_ptr(h_idx), _ptr(v_idx), _ptr(white), 1, off.h, off.w,
                                      int(proj_size[0]), int(proj_size[1]), _ptr(off.ws), _ptr(pairs),
                                      _ptr(rgb32), _stream()))
    rgb = unpack_rgb(rgb32) if rgb32 is not None else None
    return pairs[:, :2].contiguous(), pairs[:, 2:].contiguous(), rgb


def triangulate_points(cam_pts, proj_pts, calib: Calibration, precision="fp64", device=None) -> torch.Tensor:
    """Triangulate.triangulate (triangulate.py:73-97): ``(N,2)`` f32 pairs -> ``(3,N)`` f64 (CUDA)."""
    _require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    def pts(a):
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32)))
        t = t.to(device=dev, dtype=torch.float32).reshape(-1, 2).contiguous()
        return t

    cam, proj = pts(cam_pts), pts(proj_pts)
    if cam.shape != proj.shape:
        raise ValueError("cam_pts and proj_pts differ in length")
    n = cam.shape[0]
    out = torch.empty((3, n), dtype=torch.float64, device=dev)
    s = calib.struct()
    with torch.cuda.device(dev):
        check(lib().sl_triangulate_points(_ptr(cam), _ptr(proj), n, C.byref(s), _PREC[precision], _ptr(out), _stream()))
    return out


def filter_box(pts: torch.Tensor, colors: Optional[torch.Tensor], threshold: float = .5):
    """Triangulate.filter_3d_pts (triangulate.py:99-122) on CUDA tensors ``(3,N)`` f64 / ``(N,3)`` f64."""
    if not (pts.is_cuda and pts.dtype == torch.float64 and pts.dim() == 2 and pts.shape[0] == 3):
        raise TypeError("pts must be a CUDA float64 (3,N) tensor")
    pts = pts.contiguous()
    n = pts.shape[1]
    dev = pts.device
    if colors is not None:
        colors = colors.to(device=dev, dtype=torch.float64).contiguous()
        if tuple(colors.shape) != (n, 3):
            raise ValueError("colors must be (N,3)")
    out_pts = torch.empty((3, n), dtype=torch.float64, device=dev)
    out_col = torch.empty((n, 3), dtype=torch.float64, device=dev) if colors is not None else None
    cnt = torch.empty(1, dtype=torch.int64, device=dev)
    nbytes = lib().sl_filter_workspace_bytes(n)
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib().sl_filter_box(_ptr(pts), _ptr(colors), n, float(threshold), _ptr(out_pts), _ptr(out_col),
                                  _ptr(cnt), _ptr(ws), nbytes, _stream()))
    m = int(cnt.item())
    return out_pts[:, :m], (out_col[:m] if out_col is not None else None)


def filter_cloud_box(cloud: torch.Tensor, threshold: float = .5) -> torch.Tensor:
    """Triangulate.filter_3d_pts (triangulate.py:99-122) on a packed XYZRGB cloud ``(N,4)`` f32 (CUDA):
    keeps records with |x|,|y|,|z| < threshold, order preserved."""
    if not (cloud.is_cuda and cloud.dtype == torch.float32 and cloud.dim() == 2 and cloud.shape[1] == 4):
        raise TypeError("cloud must be a CUDA float32 (N,4) tensor")
    cloud = cloud.contiguous()
    n = cloud.shape[0]
    out = torch.empty_like(cloud)
    cnt = torch.empty(1, dtype=torch.int64, device=cloud.device)
    nbytes = lib().sl_filter_workspace_bytes(n)
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=cloud.device)
    with torch.cuda.device(cloud.device):
        check(lib().sl_filter_cloud_box(_ptr(cloud), n, float(threshold), _ptr(out), _ptr(cnt), _ptr(ws), nbytes, _stream()))
    return out[: int(cnt.item())]


def remove_statistical_outliers(cloud: torch.Tensor, nb_neighbors: int = 20, std_ratio: float = 0.5,
                                details: bool = False):
    """``pcd.remove_statistical_outlier(nb_neighbors, std_ratio)`` (scanner/utils/visualize.py:104-105, Open3D 0.17)
    on a packed XYZRGB cloud ``(N,4)`` f32 (CUDA).  Returns ``(inlier_cloud (M,4), ind (M,) int64)``; with
    ``details`` also ``mean_dist (N,) f64`` and ``stats (16,) f64`` = cloud mean, standard deviation, threshold, finite records,
    queries per grid level."""
    if not (cloud.is_cuda and cloud.dtype == torch.float32 and cloud.dim() == 2 and cloud.shape[1] == 4):
        raise TypeError("cloud must be a CUDA float32 (N,4) tensor")
    if nb_neighbors < 1 or not std_ratio > 0:
        raise ValueError("nb_neighbors and std_ratio must be positive")        # Open3D's own check
    cloud = cloud.contiguous()
    n, dev = cloud.shape[0], cloud.device
    out = torch.empty_like(cloud)
    ind = torch.empty(n, dtype=torch.int64, device=dev)
    cnt = torch.empty(1, dtype=torch.int64, device=dev)
    mean_dist = torch.empty(n, dtype=torch.float64, device=dev) if details else None
    stats = torch.empty(16, dtype=torch.float64, device=dev) if details else None
    nbytes = lib().sl_outlier_workspace_bytes(n)
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib().sl_remove_statistical_outliers(_ptr(cloud), n, int(nb_neighbors), float(std_ratio), _ptr(out),
                                                   _ptr(ind), _ptr(mean_dist), _ptr(stats), _ptr(cnt), _ptr(ws),
                                                   nbytes, _stream()))
    m = int(cnt.item())
    if details:
        return out[:m], ind[:m], mean_dist, stats
    return out[:m], ind[:m]


def unpack_rgb(word: torch.Tensor) -> torch.Tensor:
    """Colour words (R | G<<8 | B<<16, int32 / float32 bit patterns) -> (N,3) uint8."""
    flat = torch.empty(word.shape, dtype=word.dtype, device=word.device)      # fresh, unit-stride storage
    flat.copy_(word)
    return flat.view(torch.uint8).reshape(-1, 4)[:, :3]


class _RgbView:
    """Lazy ``rgb[:n]`` accessor over the colour word of the records (avoids unpacking the whole buffer)."""

    def __init__(self, cloud):
        self._cloud = cloud

    def __getitem__(self, idx) -> torch.Tensor:
        return unpack_rgb(self._cloud[idx][..., 3])


@dataclass
class CloudBuffers:
    """Pre-allocated device buffers so a scan_to_cloud call allocates nothing.

    ``cloud`` holds one 16-byte XYZRGB record per point: ``[x, y, z, bits(R | G<<8 | B<<16)]``;
    ``xyz`` / ``rgb`` give the two halves as (strided) views / unpacked copies."""
    h_idx: torch.Tensor
    v_idx: torch.Tensor
    ws: torch.Tensor
    counts: torch.Tensor
    cloud: torch.Tensor

    @property
    def xyz(self) -> torch.Tensor:
        return self.cloud[:, :3]

    @property
    def rgb(self) -> torch.Tensor:
        return _RgbView(self.cloud)


    def column_counts(self) -> torch.Tensor:
        """Per-column point counts ``(B, W)`` int32 of the last compaction (view into the workspace)."""
        b, h, w = self.h_idx.shape
        ptr = lib().sl_workspace_col_totals(_ptr(self.ws), b, h, w)
        off = int(ptr) - self.ws.data_ptr()
        return self.ws[off: off + 4 * b * w].view(torch.int32).view(b, w)

    @classmethod
    def allocate(cls, b: int, h: int, w: int, device, cap_points: Optional[int] = None) -> "CloudBuffers":
        cap = b * h * w if cap_points is None else int(cap_points)
        shape = (b, h, w)
        return cls(torch.empty(shape, dtype=torch.int16, device=device),
                   torch.empty(shape, dtype=torch.int16, device=device),
                   torch.empty(lib().sl_cloud_workspace_bytes(b, h, w), dtype=torch.uint8, device=device),
                   torch.empty(b + 1, dtype=torch.int64, device=device),
                   torch.empty((cap, 4), dtype=torch.float32, device=device))


# scan_to_cloud(emit="auto") takes the row-mapped emission where it applies (fp32, lens tables, W % 32 == 0: 33.4 -> 29.6 us
# per pipelined 1080p scan); SLB200_EMIT=tiles in the environment keeps sl_emit_cloud everywhere (A/B runs)
_AUTO_EMIT_ROWS = __import__("os").environ.get("SLB200_EMIT", "rows") != "tiles"


class LensTables:
    """Undistorted rays of the camera and projector pixel lattices for one calibration
    (``sl_lens_tables``); build once, reuse for every scan.  ``cam``: the camera rays (rotated by R);
    ``proj``: the exact projector rays (what the fp64 mode looks up); ``proj_bt``: ``(|T x p|, T.p)`` of those rays
    (what the fp32 mode looks up).  ``precision`` restricts the build to one mode's projector table."""

    def __init__(self, calib: Calibration, cam_size: Tuple[int, int], device, precision=None):
        w, h = int(cam_size[0]), int(cam_size[1])
        pw, ph = calib.proj_size
        self.cam_size, self.proj_size = (w, h), (pw, ph)
        self.cam = torch.empty((h, w, 2), dtype=torch.float32, device=device)
        self.proj = self.proj_bt = None
        s = calib.struct()
        want = (SL_PREC_FP64, SL_PREC_FP32) if precision is None else (_PREC[precision],)
        with torch.cuda.device(device):
            for prec in want:
                t = torch.empty((ph, pw, 2), dtype=torch.float32, device=device)
                check(lib().sl_lens_tables(C.byref(s), w, h, prec, _ptr(self.cam), _ptr(t), _stream()))
                if prec == SL_PREC_FP64:
                    self.proj = t
                else:
                    self.proj_bt = t
            self._transposed = None
            if self.proj_bt is not None and _AUTO_EMIT_ROWS:
                self.transposed()          # now, not inside somebody's CUDA graph capture

    def transposed(self):
        """``(cam_t (W,H,2), proj_bt_t (pw,ph,2))``: the fp32 mode's tables stored column-major, what the row-mapped
        emission (``sl_emit_cloud_rows``) reads; built on first use, once per calibration."""
        if getattr(self, "_transposed", None) is None:
            if self.proj_bt is None:
                raise ValueError("lens tables were built for the other triangulation precision")
            h, w = self.cam.shape[:2]
            ph, pw = self.proj_bt.shape[:2]
            cam_t = torch.empty((w, h, 2), dtype=torch.float32, device=self.cam.device)
            proj_t = torch.empty((pw, ph, 2), dtype=torch.float32, device=self.cam.device)
            with torch.cuda.device(self.cam.device):
                check(lib().sl_transpose_table(_ptr(self.cam), h, w, _ptr(cam_t), _stream()))
                check(lib().sl_transpose_table(_ptr(self.proj_bt), ph, pw, _ptr(proj_t), _stream()))
            self._transposed = (cam_t, proj_t)
        return self._transposed

    def projector_table(self, precision) -> torch.Tensor:
        t = self.proj if _PREC[precision] == SL_PREC_FP64 else self.proj_bt
        if t is None:
            raise ValueError("lens tables were built for the other triangulation precision")
        return t


def scan_to_cloud(frames: torch.Tensor, white_rgb: Optional[torch.Tensor], calib: Calibration,
                  precision="fp32", eps: float = 1.0, buffers: Optional[CloudBuffers] = None,
                  calib_struct: Optional[SlCalib] = None, tables: Optional[LensTables] = None,
                  row0: int = 0, bgr: bool = False, sm_reserve: int = 0, box: Optional[float] = None,
                  emit: str = "auto") -> CloudBuffers:
    """Whole device-resident path for ``(F,H,W)`` / ``(B,F,H,W)`` uint8 stacks:
    decode -> ordered compaction -> undistort + triangulate.  No host synchronisation.

    ``tables`` (LensTables of the same calibration / sizes) replaces the per-pixel undistortion
    iterations by look-ups of exact precomputed rays; without it the iterations run in the kernel.
    ``row0``: the stack is the row band ``[row0, row0 + H)`` of a larger capture (row-band sharding);
    camera coordinates and ray look-ups use absolute rows, ``tables`` must cover the whole image.

    ``bgr``: ``frames`` holds the camera's BGR frames ``([B,]F,H,W,3)``; the gray conversion of
    src/3-capture_decode.py:66 happens inside the decode kernel.

    ``sm_reserve``: SMs the persistent decode grid leaves to the kernels of other streams (see ``sl_decode_u8``);
    8 on a B200 when independent scans are pipelined over several CUDA streams, else 0.

    ``box``: Triangulate.filter_3d_pts (triangulate.py:99-122; src/4-triangulate.py:71 uses 0.5) folded into the
    compaction (``sl_compact_offsets_boxed``): only points with |x|, |y|, |z| < box are counted and emitted, identical
    to ``filter_cloud_box`` of the unfiltered cloud, which is never materialised.

    ``emit``: ``"tiles"`` = ``sl_emit_cloud``; ``"rows"`` = ``sl_emit_cloud_rows`` (fp32, tables, W % 32 == 0; the same
    records bit for bit from transposed tables, fewer L1 wavefronts per pixel); ``"auto"`` picks.

    Results are in the returned buffers: ``cloud[:counts[B]]`` (XYZRGB records; ``.xyz`` / ``.rgb``
    views) packed scan after scan, each scan in the reference's x-major order; ``counts[b]`` points per scan.
    Mirrors src/3-capture_decode.py:75-100 + src/4-triangulate.py:50-64.
    """
    if bgr:
        b, f, h, w, _, fp, sp = _bgr_view(frames)
    else:
        _check_stack(frames)
        b, f, h, w, _ = _stack_dims(frames)
        fp, sp = _pitches(frames)
    dev = frames.device
    if buffers is None:
        buffers = CloudBuffers.allocate(b, h, w, dev)
    if white_rgb is not None:
        if not (white_rgb.is_cuda and white_rgb.dtype == torch.uint8 and white_rgb.is_contiguous()):
            raise TypeError("white_rgb must be a contiguous CUDA uint8 tensor")
        if white_rgb.numel() != b * h * w * 3:
            raise ValueError("white_rgb must be ([B,]H,W,3)")
    if calib_struct is None:
        s = calib.struct(row0)
    elif int(calib_struct.cam_row0) != int(row0):
        s = SlCalib.from_buffer_copy(calib_struct)       # a cached struct is reused for another row band
        s.cam_row0 = int(row0)
    else:
        s = calib_struct
    if tables is not None and (tables.cam_size[0] != w or tables.cam_size[1] < row0 + h
                               or tables.proj_size != tuple(calib.proj_size)):
        raise ValueError("lens tables were built for another image / panel size")
    L = lib()
    st = _stream()
    cam_t = _ptr(tables.cam) if tables is not None else None
    proj_t = _ptr(tables.projector_table(precision)) if tables is not None else None
    if emit not in ("auto", "tiles", "rows"):
        raise ValueError("emit must be 'auto', 'tiles' or 'rows'")
    rows_ok = (tables is not None and _PREC[precision] == SL_PREC_FP32 and w % 32 == 0 and tables.proj_bt is not None
               and buffers.h_idx.data_ptr() % 16 == 0 and buffers.v_idx.data_ptr() % 16 == 0
               and (white_rgb is None or white_rgb.data_ptr() % 16 == 0))
    if emit == "rows" and not rows_ok:
        raise ValueError("the row-mapped emission needs fp32 lens tables, W % 32 == 0 and 16-byte aligned buffers")
    use_rows = rows_ok and (emit == "rows" or (emit == "auto" and _AUTO_EMIT_ROWS))

    def emit_cloud():
        if use_rows:
            cam_tt, proj_tt = tables.transposed()
            check(L.sl_emit_cloud_rows(_ptr(buffers.h_idx), _ptr(buffers.v_idx), _ptr(white_rgb), C.byref(s), _ptr(cam_tt),
                                       int(cam_tt.shape[1]), _ptr(proj_tt), b, h, w, _ptr(buffers.ws), _ptr(buffers.cloud), st))
        else:
            check(L.sl_emit_cloud(_ptr(buffers.h_idx), _ptr(buffers.v_idx), _ptr(white_rgb), C.byref(s),
                                  _PREC[precision], cam_t, proj_t, b, h, w, _ptr(buffers.ws), _ptr(buffers.cloud), st))

    with torch.cuda.device(dev):
        if bgr and box is not None:
            raise ValueError("box is not offered together with bgr")
        if bgr:
            check(L.sl_decode_bgr_u8(_ptr(frames), b, f, h, w, fp, sp, float(eps),
                                     _ptr(buffers.h_idx), _ptr(buffers.v_idx), _ptr(buffers.ws), int(sm_reserve), st))
            check(L.sl_compact_offsets_decoded(b, h, w, _ptr(buffers.ws), _ptr(buffers.counts), st))
            emit_cloud()
        elif box is not None:
            check(L.sl_decode_u8(_ptr(frames), b, f, h, w, fp, sp, float(eps), _ptr(buffers.h_idx), _ptr(buffers.v_idx),
                                 None, None, None, int(sm_reserve), st))
            check(L.sl_compact_offsets_boxed(_ptr(buffers.h_idx), _ptr(buffers.v_idx), C.byref(s), _PREC[precision], cam_t,
                                             proj_t, float(box), b, h, w, _ptr(buffers.ws), _ptr(buffers.counts), st))
            emit_cloud()
        elif use_rows:
            check(L.sl_decode_u8(_ptr(frames), b, f, h, w, fp, sp, float(eps), _ptr(buffers.h_idx), _ptr(buffers.v_idx),
                                 None, None, _ptr(buffers.ws), int(sm_reserve), st))
            check(L.sl_compact_offsets_decoded(b, h, w, _ptr(buffers.ws), _ptr(buffers.counts), st))
            emit_cloud()
        else:
            check(L.sl_decode_triangulate_fused(_ptr(frames), b, f, h, w, fp, sp, float(eps), _ptr(white_rgb), C.byref(s),
                                                _PREC[precision], cam_t, proj_t, _ptr(buffers.h_idx), _ptr(buffers.v_idx),
                                                _ptr(buffers.ws), _ptr(buffers.cloud), _ptr(buffers.counts),
                                                int(sm_reserve), st))
    return buffers


# --------------------------------------------------------------------------- host-buffer pipeline
class Pipeline:
    """Host-buffer path (``sl_scan_to_cloud_host``): frames in host memory (pinned memory makes
    the copies asynchronous), H2D + kernels + D2H overlapped on a ring of CUDA streams."""

    def __init__(self, f: int, h: int, w: int, max_batch: int = 1, device: Optional[int] = None):
        _require_cuda()
        self.f, self.h, self.w, self.max_batch = int(f), int(h), int(w), int(max_batch)
        self.device = torch.cuda.current_device() if device is None else int(device)
        self._h = C.c_void_p()
        check(lib().sl_pipeline_create(C.byref(self._h), self.device, self.max_batch, self.f, self.h, self.w))

    def set_sm_reserve(self, sms: int):
        """SMs this pipeline's decode launches leave to the kernels of its other streams (``sl_pipeline_set_sm_reserve``)."""
        check(lib().sl_pipeline_set_sm_reserve(self._h, int(sms)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().sl_pipeline_destroy(self._h)
            self._h = None

    __del__ = close

    @staticmethod
    def _host_ptr(a):
        if a is None:
            return None
        if isinstance(a, torch.Tensor):
            if a.is_cuda:
                raise TypeError("host buffer expected")
            return C.c_void_p(a.data_ptr())
        return C.c_void_p(a.ctypes.data)

    @staticmethod
    def _host_check(a, name, dtype, shape=None, min_elems=None, nullable=False):
        """The C ABI takes raw host pointers: dtype, shape and contiguity are checked here, because a wrong one is a
        silent misread or an out-of-bounds write by an asynchronous copy.  ``shape`` entries of None are free."""
        if a is None:
            if nullable:
                return None
            raise TypeError(f"{name} is required")
        if isinstance(a, torch.Tensor):
            if a.is_cuda:
                raise TypeError(f"{name}: host buffer expected, got a CUDA tensor")
            ok_dtype = a.dtype == getattr(torch, dtype)
            contiguous = a.is_contiguous()
        elif isinstance(a, np.ndarray):
            ok_dtype = a.dtype == np.dtype(dtype)
            contiguous = bool(a.flags["C_CONTIGUOUS"])
        else:
            raise TypeError(f"{name}: numpy array or CPU torch tensor expected, got {type(a).__name__}")
        if not ok_dtype:
            raise TypeError(f"{name} must be {dtype}, got {a.dtype}")
        if not contiguous:
            raise ValueError(f"{name} must be C-contiguous")
        if shape is not None:
            got = tuple(a.shape)
            if len(got) != len(shape) or any(s is not None and int(s) != int(g) for s, g in zip(shape, got)):
                raise ValueError(f"{name} must have shape {tuple('*' if s is None else s for s in shape)}, got {got}")
        if min_elems is not None:
            n = 1
            for d in a.shape:
                n *= int(d)
            if n < min_elems:
                raise ValueError(f"{name} needs at least {min_elems} elements, got {n}")
        return a

    def scan_to_cloud(self, frames, white_rgb, calib: Calibration, precision="fp32", eps: float = 1.0,
                      cloud=None, counts=None, h_idx=None, v_idx=None):
        """``frames`` host uint8 ``(B,F,H,W)`` (numpy or CPU torch, ideally pinned); the XYZRGB records
        are written into the host tensor ``cloud`` ``(cap,4)`` f32 (allocated pinned when not
        supplied).  Returns ``(cloud[:n], counts)``; ``cloud[:, :3]`` is xyz, ``unpack_rgb(cloud[:, 3])``
        the colours."""
        shape = tuple(getattr(frames, "shape", ()))
        if len(shape) == 3:
            shape = (1, *shape)
        if len(shape) != 4 or shape[1:] != (self.f, self.h, self.w):
            raise ValueError(f"frames must be (B,{self.f},{self.h},{self.w})")
        b = shape[0]
        self._host_check(frames, "frames", "uint8")
        self._host_check(white_rgb, "white_rgb", "uint8", min_elems=b * self.h * self.w * 3, nullable=True)
        if white_rgb is not None and tuple(white_rgb.shape)[-3:] != (self.h, self.w, 3):
            raise ValueError(f"white_rgb must be ([B,]{self.h},{self.w},3)")
        cap = b * self.h * self.w
        if cloud is None:
            cloud = torch.empty((cap, 4), dtype=torch.float32).pin_memory()
        if counts is None:
            counts = torch.empty(b + 1, dtype=torch.int64)
        self._host_check(cloud, "cloud", "float32", shape=(None, 4))
        self._host_check(counts, "counts", "int64", min_elems=b + 1)
        self._host_check(h_idx, "h_idx", "int16", shape=(b, self.h, self.w), nullable=True)
        self._host_check(v_idx, "v_idx", "int16", shape=(b, self.h, self.w), nullable=True)
        s = calib.struct()
        check(lib().sl_scan_to_cloud_host(self._h, self._host_ptr(frames), self._host_ptr(white_rgb), b,
                                          C.byref(s), _PREC[precision], float(eps), self._host_ptr(h_idx),
                                          self._host_ptr(v_idx), self._host_ptr(cloud),
                                          int(cloud.shape[0]), self._host_ptr(counts)))
        n = int(counts[b])
        return cloud[:n], counts

    def submit(self, frames, white_rgb, calib: Calibration, cloud, count, precision="fp32", eps: float = 1.0,
               bgr: bool = False):
        """Streaming form (``sl_pipeline_submit``): queue ONE scan ``(F,H,W)`` uint8 (host, ideally pinned) and
        return; its records go to the host tensor ``cloud`` ``(cap,4)`` f32 and its point count to ``count``
        (int64 host tensor / array of one element).  Results are valid after :meth:`flush`; the buffers must
        stay alive until then.  ``bgr=True`` (``sl_pipeline_submit_bgr``): ``frames`` is ``(F,H,W,3)`` 8-bit BGR as
        the camera delivers it; cv2's BGR2GRAY happens inside the decode load, bit-identical."""
        fshape = (self.f, self.h, self.w, 3) if bgr else (self.f, self.h, self.w)
        self._host_check(frames, "frames", "uint8", shape=fshape)
        self._host_check(white_rgb, "white_rgb", "uint8", shape=(self.h, self.w, 3), nullable=True)
        self._host_check(cloud, "cloud", "float32", shape=(None, 4))
        self._host_check(count, "count", "int64", min_elems=1)
        s = calib.struct()
        fn = lib().sl_pipeline_submit_bgr if bgr else lib().sl_pipeline_submit
        check(fn(self._h, self._host_ptr(frames), self._host_ptr(white_rgb), C.byref(s), _PREC[precision], float(eps),
                 self._host_ptr(cloud), int(cloud.shape[0]), self._host_ptr(count)))

    def flush(self):
        """Wait until every submitted scan's cloud and count are in host memory (``sl_pipeline_flush``)."""
        check(lib().sl_pipeline_flush(self._h))

    def decode(self, frames, eps: float = 1.0, codes: bool = False):
        """Host uint8 ``(B,F,H,W)`` -> host int16 ``h_pixels, v_pixels`` [+ int8 tri-state codes]."""
        shape = tuple(getattr(frames, "shape", ()))
        squeeze = len(shape) == 3
        if squeeze:
            shape = (1, *shape)
        if len(shape) != 4 or shape[1:] != (self.f, self.h, self.w):
            raise ValueError(f"frames must be (B,{self.f},{self.h},{self.w})")
        b = shape[0]
        self._host_check(frames, "frames", "uint8")
        n = (self.f - 2) // 4
        h_idx = np.empty((b, self.h, self.w), dtype=np.int16)
        v_idx = np.empty((b, self.h, self.w), dtype=np.int16)
        h_codes = np.empty((b, n, self.h, self.w), dtype=np.int8) if codes else None
        v_codes = np.empty((b, n, self.h, self.w), dtype=np.int8) if codes else None
        check(lib().sl_decode_host(self._h, self._host_ptr(frames), b, float(eps), self._host_ptr(h_idx),
                                   self._host_ptr(v_idx), self._host_ptr(h_codes), self._host_ptr(v_codes)))
        if squeeze:
            h_idx, v_idx = h_idx[0], v_idx[0]
            if codes:
                h_codes, v_codes = h_codes[0], v_codes[0]
        return (h_idx, v_idx, h_codes, v_codes) if codes else (h_idx, v_idx)
