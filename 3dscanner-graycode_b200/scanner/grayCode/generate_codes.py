"""Projector-side Gray-code pattern synthesis (host, NumPy).

Same module path and signatures as the reference's scanner/grayCode/generate_codes.py
(imported by src/3-capture_decode.py:6).  Pattern synthesis runs once per session on the
projector side and is outside the per-pixel hot path, but it fixes the frame layout the
decode kernels assume (see csrc/sl_decode_quad.cuh) and feeds the synthetic renderer.
"""
import numpy as np


def get_gray_codes(width, height):
    """Gray-code table ``(max(width,height), n)`` uint8, bits MSB-first, n = ceil(log2(max)).

    generate_codes.py:5-32.
    """
    size = max(int(width), int(height))
    n_bits = int(np.ceil(np.log2(size)))
    idx = np.arange(size, dtype=np.uint32)
    gray = idx ^ (idx >> 1)
    shifts = np.arange(n_bits - 1, -1, -1, dtype=np.uint32)
    return ((gray[:, None] >> shifts[None, :]) & 1).astype(np.uint8)


def get_image_sequence(gray_codes, width, height):
    """``(4n+2, height, width)`` uint8 projector frames; generate_codes.py:34-81.

    frame 0 black, 1 white; 2+2j column bit j (MSB first); 3+2k row bit k (LSB first);
    +2n the inverse of each.  As in the reference the stripe width is
    ``width // len(gray_codes)``: 1 for landscape panels, 0 (all-black patterns) for
    portrait ones.
    """
    gray_codes = np.asarray(gray_codes)
    size, n = gray_codes.shape
    images = np.zeros((4 * n + 2, height, width), dtype=np.uint8)
    images[1] = 255
    if width // size == 0:
        return images
    col_bits = gray_codes[:width]                 # stripe i covers column i
    row_bits = gray_codes[:height]
    for j in range(n):
        id_v = 2 * j + 2
        id_h = 2 * (n - (j + 1)) + 3
        col = (col_bits[:, j] * 255).astype(np.uint8)
        images[id_v, :, :col.size] = col[None, :]
        images[id_v + 2 * n, :, :col.size] = 255 - col[None, :]
        row = (row_bits[:, j] * 255).astype(np.uint8)
        images[id_h, :row.size, :] = row[:, None]
        images[id_h + 2 * n, :row.size, :] = 255 - row[:, None]
    return images


def display_gray_code(gray_codes, width, height, repeat_n, fps, write_video_seq=False):
    """Projector-side playback of the pattern sequence (reference generate_codes.py:83-120).  Out of scope of the
    B200 hot path (it needs a window system and a projector): this shim only keeps the name the capture script
    imports (src/3-capture_decode.py:6).  It hands the frames, generated here, to a minimal viewer loop; the
    reference's optional mp4 dump is not reproduced."""
    if write_video_seq:
        raise NotImplementedError("video export of the pattern sequence is outside the B200 hot path; "
                                  "use get_image_sequence() and an encoder of your choice")
    import cv2
    title = "patterns"
    cv2.namedWindow(title, cv2.WINDOW_NORMAL)
    cv2.setWindowProperty(title, cv2.WND_PROP_FULLSCREEN, cv2.WINDOW_FULLSCREEN)
    delay_ms = max(1, int(round(1000.0 / fps)))
    frames = get_image_sequence(gray_codes, width, height)
    try:
        for _ in range(int(repeat_n)):
            for frame in frames:
                cv2.imshow(title, frame)
                cv2.waitKey(delay_ms)
    finally:
        cv2.destroyWindow(title)
