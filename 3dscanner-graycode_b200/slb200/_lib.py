"""ctypes binding of ``lib/libslb200.so`` (C ABI: ``include/slb200.h``).

There is no fallback: if the shared library is missing or a call fails, a Python
exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SLB200_LIB selects another build of the same library (kernel tuning experiments only)
LIB_PATH = os.environ.get("SLB200_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libslb200.so")

SL_PREC_FP32 = 0
SL_PREC_FP64 = 1


class SlCalib(C.Structure):
    _fields_ = [
        ("cam_K", C.c_double * 9), ("cam_dist", C.c_double * 8),
        ("proj_K", C.c_double * 9), ("proj_dist", C.c_double * 8),
        ("R", C.c_double * 9), ("T", C.c_double * 3),
        ("proj_w", C.c_int32), ("proj_h", C.c_int32), ("cam_row0", C.c_int32), ("reserved", C.c_int32),
    ]


class SlError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"slb200 error {code}: {msg}")
        self.code = code


_vp, _i, _sz, _d, _i64 = C.c_void_p, C.c_int, C.c_size_t, C.c_double, C.c_int64

# name -> (restype, argtypes); every symbol include/slb200.h declares
SIGNATURES = {
    "sl_last_error": (C.c_char_p, []),
    "sl_version": (_i, []),
    "sl_device_check": (_i, [_i]),
    "sl_decode_u8": (_i, [_vp, _i, _i, _i, _i, _sz, _sz, _d, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "sl_decode_u8_stats": (_i, [_vp, _i, _i, _i, _i, _sz, _sz, _d, _vp, _vp, _vp, _vp]),
    "sl_decode_bgr_u8": (_i, [_vp, _i, _i, _i, _i, _sz, _sz, _d, _vp, _vp, _vp, _i, _vp]),
    "sl_direct_indirect": (_i, [_vp, _i, _i, _i, _sz, _vp, _vp, _vp]),
    "sl_is_lit": (_i, [_vp, _i, _i, _i, _sz, _vp, _vp, _d, _vp, _vp, _vp]),
    "sl_codes_to_pixels": (_i, [_vp, _vp, _i, _i, _i64, _sz, _vp, _vp, _vp]),
    "sl_pattern_sequence_u8": (_i, [_i, _i, _vp, _vp]),
    "sl_bgr_to_gray_u8": (_i, [_vp, _i64, _vp, _vp]),
    "sl_pair_diff_counts": (_i, [_vp, _i, _i64, _sz, _i, _vp, _vp]),
    "sl_cloud_workspace_bytes": (_sz, [_i, _i, _i]),
    "sl_compact_offsets": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "sl_compact_offsets_decoded": (_i, [_i, _i, _i, _vp, _vp, _vp]),
    "sl_compact_offsets_boxed": (_i, [_vp, _vp, C.POINTER(SlCalib), _i, _vp, _vp, _d, _i, _i, _i, _vp, _vp, _vp]),
    "sl_workspace_col_totals": (_vp, [_vp, _i, _i, _i]),
    "sl_emit_pairs": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "sl_lens_tables": (_i, [C.POINTER(SlCalib), _i, _i, _i, _vp, _vp, _vp]),
    "sl_emit_cloud": (_i, [_vp, _vp, _vp, C.POINTER(SlCalib), _i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "sl_decode_triangulate_fused": (_i, [_vp, _i, _i, _i, _i, _sz, _sz, _d, _vp, C.POINTER(SlCalib), _i, _vp, _vp, _vp, _vp,
                                         _vp, _vp, _vp, _i, _vp]),
    "sl_triangulate_points": (_i, [_vp, _vp, _i64, C.POINTER(SlCalib), _i, _vp, _vp]),
    "sl_filter_box": (_i, [_vp, _vp, _i64, _d, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sl_filter_workspace_bytes": (_sz, [_i64]),
    "sl_filter_cloud_box": (_i, [_vp, _i64, _d, _vp, _vp, _vp, _sz, _vp]),
    "sl_remove_statistical_outliers": (_i, [_vp, _i64, _i, _d, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sl_outlier_workspace_bytes": (_sz, [_i64]),
    "sl_pipeline_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i]),
    "sl_pipeline_destroy": (None, [_vp]),
    "sl_pipeline_set_sm_reserve": (_i, [_vp, _i]),
    "sl_scan_to_cloud_host": (_i, [_vp, _vp, _vp, _i, C.POINTER(SlCalib), _i, _d, _vp, _vp, _vp, _i64, _vp]),
    "sl_pipeline_submit": (_i, [_vp, _vp, _vp, C.POINTER(SlCalib), _i, _d, _vp, _i64, _vp]),
    "sl_pipeline_submit_bgr": (_i, [_vp, _vp, _vp, C.POINTER(SlCalib), _i, _d, _vp, _i64, _vp]),
    "sl_pipeline_flush": (_i, [_vp]),
    "sl_decode_host": (_i, [_vp, _vp, _i, _d, _vp, _vp, _vp, _vp]),
    "sl_peer_alloc": (_i, [_sz, C.POINTER(_vp)]),
    "sl_peer_free": (_i, [_vp]),
    "sl_peer_export": (_i, [_vp, _vp]),
    "sl_peer_open": (_i, [_vp, C.POINTER(_vp)]),
    "sl_peer_close": (_i, [_vp]),
    "sl_exchange_ctrl_bytes": (_sz, [_i, _i, _i]),
    "sl_exchange_create": (_i, [C.POINTER(_vp), _i, _i, C.POINTER(_vp), _i, _i, _d]),
    "sl_exchange_destroy": (None, [_vp]),
    "sl_exchange_offsets_by_scan": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "sl_exchange_offsets_row_band": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "sl_emit_cloud_multi": (_i, [_vp, _vp, _vp, C.POINTER(SlCalib), _i, _vp, _vp, _i, _i, _i, _vp, C.POINTER(_vp), _i, _i64,
                                _vp]),
    "sl_transpose_table": (_i, [_vp, _i, _i, _vp, _vp]),
    "sl_emit_cloud_rows": (_i, [_vp, _vp, _vp, C.POINTER(SlCalib), _vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "sl_emit_cloud_rows_multi": (_i, [_vp, _vp, _vp, C.POINTER(SlCalib), _vp, _i, _vp, _i, _i, _i, _vp, _vp, _i, _vp]),
    "sl_emit_cloud_at": (_i, [_vp, _vp, _vp, C.POINTER(SlCalib), _i, _vp, _vp, _i, _i, _i, _vp, _vp, _i64, _vp]),
    "sl_exchange_push": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i64, _i, _vp]),
    "sl_exchange_arrive": (_i, [_vp, C.c_uint32, _i, _vp]),
    "sl_exchange_status": (_i, [_vp, C.POINTER(_i), C.POINTER(_i64), C.POINTER(_i64)]),
    "sl_host_narrow_u8": (_i, [_vp, _i, _sz, _vp, _i]),
    "sl_host_widen_idx": (_i, [_vp, _sz, _vp, _i]),
    "sl_host_narrow_idx": (_i, [_vp, _sz, _vp, _i]),
    "sl_host_unpack_cloud": (_i, [_vp, _i64, _vp, _vp, _i]),
    "sl_launch_count": (_i64, []),
}

_lib = None


def lib() -> C.CDLL:
    """Load the extension once.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"slb200 CUDA extension missing at {LIB_PATH}; build it with "
                "`python __graft_entry__.py` (there is no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(code: int) -> None:
    if code != 0:
        raise SlError(code, lib().sl_last_error().decode("utf-8", "replace"))
