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
    """Fullscreen display / mp4 dump of the sequence (generate_codes.py:83-120).  Needs a GUI."""
    import cv2
    images = get_image_sequence(gray_codes, width, height)
    writer = None
    if write_video_seq:
        writer = cv2.VideoWriter('./data/gray_sequence.mp4', cv2.VideoWriter_fourcc(*'mp4v'), fps,
                                 (width, height), isColor=False)
    for rep in range(repeat_n):
        for frame in images:
            cv2.namedWindow('Gray Code Pattern', cv2.WINDOW_NORMAL)
            cv2.setWindowProperty('Gray Code Pattern', cv2.WND_PROP_FULLSCREEN, cv2.WINDOW_FULLSCREEN)
            cv2.imshow("Gray Code Pattern", frame)
            cv2.waitKey(int(1000 / fps))
            if writer is not None and rep == 0:
                writer.write(frame)
    cv2.destroyAllWindows()
    if writer is not None:
        writer.release()
