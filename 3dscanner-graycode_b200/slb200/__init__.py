"""slb200 -- B200 (sm_100a) Gray-code decode + triangulation hot path.

Host side is Python/PyTorch (device memory, streams, torch.distributed); the
arithmetic runs in hand-written CUDA kernels behind a C ABI (include/slb200.h).
"""
from ._lib import SlCalib, SlError, SL_PREC_FP32, SL_PREC_FP64, lib, LIB_PATH  # noqa: F401
from .api import (  # noqa: F401
    Calibration, decode_pixels, decode_codes, direct_indirect, is_lit, codes_to_pixels,
    cam_proj_pairs, triangulate_points, filter_box, scan_to_cloud, Pipeline, launch_count,
    as_u8_stack, LensTables, CloudBuffers, unpack_rgb, bgr_to_gray, pair_diff_counts, filter_cloud_box, pattern_sequence,
    remove_statistical_outliers, decode_pixels_bgr, decode_stats,
)
