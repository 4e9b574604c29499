"""Drop-in for the reference's scanner/grayCode/decode_codes.py, computed on a B200.

Same function names, argument order, defaults and return types as the reference module that
src/3-capture_decode.py:7 imports; NumPy arrays in and out.  The arithmetic runs in the
sm_100a kernels behind ``slb200`` -- there is no NumPy fallback: without the extension or a
GPU these functions raise.

Additive fast entry points (``decode_pixels``, ``codes_to_pixels``) replace the script's
2*H*W per-pixel ``gray_to_decimal`` calls (src/3-capture_decode.py:99-100); see INTEGRATION.md.
"""
import numpy as np
import torch

import slb200


def _stack(images):
    return slb200.as_u8_stack(images)


# ---- capture front end (reference decode_codes.py:6-87) ------------------------------------
def read_images(folder):
    """All images of a folder, sorted by name length like the reference (decode_codes.py:6-32).
    Host file I/O (cv2.imread); returns ``(images float64 (N,H,W,3), file_names)``."""
    import os
    import cv2
    files = sorted(os.listdir(folder), key=len)
    images = None
    for i, name in enumerate(files):
        im = cv2.imread(os.path.join(folder, name))
        if images is None:
            images = np.empty((len(files),) + im.shape)
        images[i] = im
    return images, np.array(files)


def _selection_from_counts(d):
    """The reference's transition-frame state machine (decode_codes.py:49-68) on the consecutive-pair
    difference counts d[k] = count(|images[k+1] - images[k]| > 50)."""
    n_images = len(d) + 1
    if n_images < 3:
        raise IndexError("remove_bad_images needs at least 3 images")   # the reference indexes images[2]
    diff1, diff2 = int(d[0]), int(d[1])
    keep = []
    for i in range(max(n_images - 3, 0)):            # `for i, image in enumerate(images[2:-1])`
        diff3 = int(d[i + 2])                        # absdiff(images[i+3], images[i+2])
        if diff1 < diff2 and diff1 < diff3 and diff2 < diff3 and i + 1 not in keep:
            if len(keep) == 0 or (keep[-1] != i and keep[-1] != i + 2):
                keep.append(i + 1)
        elif diff2 <= diff1 and diff2 <= diff3 and i + 2 not in keep:
            if len(keep) == 0 or (keep[-1] != i + 1 and keep[-1] != i + 3):
                keep.append(i + 2)
                diff3 = -1
                diff2 = -1
        diff1 = diff2
        diff2 = diff3
    return keep


def remove_bad_images(images):
    """Indexes of the frames to keep from a captured video sequence (transition frames dropped);
    reference decode_codes.py:34-68.  The per-pair difference counts run on the GPU."""
    d = slb200.pair_diff_counts(slb200.as_u8_stack(images)).cpu().numpy()
    return _selection_from_counts(d)


def to_gray(images):
    """``(N,H,W,3)`` BGR frames -> ``(N,H,W)`` float64 grey levels (reference decode_codes.py:70-87:
    ``astype(uint8)`` then cv2 BGR2GRAY).  Float input is truncated to uint8 like the reference does."""
    a = np.asarray(images)
    u8 = a if a.dtype == np.uint8 else a.astype(np.uint8)
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
    if dev is None:
        raise RuntimeError("slb200 needs a CUDA device (sm_100a); there is no CPU fallback")
    gray = slb200.bgr_to_gray(torch.from_numpy(np.ascontiguousarray(u8)).to(dev))
    return gray.cpu().numpy().astype(np.float64)


def get_direct_indirect(images):
    """L_d, L_g float64 ``(H,W)``; reference decode_codes.py:90-122."""
    l_d, l_g = slb200.direct_indirect(_stack(images))
    return l_d.cpu().numpy(), l_g.cpu().numpy()


def get_is_lit(images, L_d, L_g, eps=1, m=10):
    """Tri-state codes ``(n,H,W)`` int8 x2; reference decode_codes.py:125-186.

    ``m`` is accepted for signature compatibility; in the reference the ``L_d < m`` assignment
    is overwritten by the later rules and has no effect (decode_codes.py:169-182).
    """
    u8 = _stack(images)
    h_codes, v_codes = slb200.is_lit(u8, torch.as_tensor(np.asarray(L_d, dtype=np.float64)),
                                     torch.as_tensor(np.asarray(L_g, dtype=np.float64)), eps=float(eps))
    return h_codes.cpu().numpy(), v_codes.cpu().numpy()


def get_codes(images):
    """``h_codes, v_codes`` int8 ``(n,H,W)`` in {-1,0,1}; reference decode_codes.py:231-248."""
    h_codes, v_codes, _, _ = slb200.decode_codes(_stack(images))
    return h_codes.cpu().numpy(), v_codes.cpu().numpy()


def gray_decode(n):
    """Gray code -> integer (reference decode_codes.py:189-207)."""
    n = int(n)
    shift = 1
    while n >> shift:
        n ^= n >> shift
        shift <<= 1
    return n


def gray_to_decimal(gray_code_list):
    """One pixel's MSB-first tri-state bits -> index, -1 if any bit is -1
    (reference decode_codes.py:209-229).  Scalar helper kept for API compatibility; whole
    images go through ``codes_to_pixels`` / ``decode_pixels``."""
    value = 0
    for bit in gray_code_list:
        bit = int(bit)
        if bit < 0:
            return -1
        value = (value << 1) | (bit & 1)
    return gray_decode(value)


# ---- additive fast entry points ---------------------------------------------------------
def codes_to_pixels(h_codes, v_codes):
    """Whole-image replacement of src/3-capture_decode.py:95-100.

    ``h_codes, v_codes``: int8 ``(n,H,W)`` (already merged) or ``(R,n,H,W)`` / list of R runs,
    merged by per-bit max exactly like ``np.max(cache, axis=0)``.  Returns int64 ``(H,W)``
    ``h_pixels, v_pixels`` with -1 where undecodable -- the arrays the script saves as .npy.
    """
    dev = torch.device("cuda", torch.cuda.current_device())
    h = torch.as_tensor(np.ascontiguousarray(np.asarray(h_codes, dtype=np.int8))).to(dev)
    v = torch.as_tensor(np.ascontiguousarray(np.asarray(v_codes, dtype=np.int8))).to(dev)
    h_idx, v_idx = slb200.codes_to_pixels(h, v)
    return slb200.api.idx_to_host_int64(h_idx), slb200.api.idx_to_host_int64(v_idx)


def decode_pixels(images, runs=None, eps=1):
    """Frame stack(s) -> ``h_pixels, v_pixels`` int64 ``(H,W)`` in one fused pass.

    ``images``: one ``(F,H,W)`` stack; or pass ``runs=[stack0, stack1, ...]`` to merge capture
    runs as the script does (src/3-capture_decode.py:75-100).
    """
    if runs is None:
        h_idx, v_idx = slb200.decode_pixels(_stack(images), eps=float(eps))
    else:
        codes = [slb200.decode_codes(_stack(r), eps=float(eps))[:2] for r in runs]
        h_idx, v_idx = slb200.codes_to_pixels(torch.stack([c[0] for c in codes]).contiguous(),
                                              torch.stack([c[1] for c in codes]).contiguous())
    return slb200.api.idx_to_host_int64(h_idx), slb200.api.idx_to_host_int64(v_idx)
