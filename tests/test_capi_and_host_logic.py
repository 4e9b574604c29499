"""CPU-side checks: the C-ABI library loads and exports every symbol include/slb200.h declares (no
compute calls), the host-side mirrors behave like the reference, and the product refuses to run
without a GPU instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "slb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import slb200
    from slb200 import _lib
    names = declared_symbols()
    assert len(names) >= 19
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(handle, name), f"{name} declared in include/slb200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names
    lib = slb200.lib()
    assert lib.sl_version() == 100
    assert lib.sl_cloud_workspace_bytes(1, 1080, 1920) > 0
    assert lib.sl_launch_count() == 0


def test_ctypes_signatures_match_the_header():
    """Every prototype in include/slb200.h has as many parameters as its ctypes signature, pointers bound as
    pointer-sized arguments and integer widths as declared."""
    import ctypes as C
    from slb200 import _lib
    text = open(os.path.join(ROOT, "include", "slb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = re.findall(r"\b([A-Za-z_][\w \*]*?)\b(sl_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text)
    assert len(protos) == len(_lib.SIGNATURES)
    for ret, name, params in protos:
        res, args = _lib.SIGNATURES[name]
        plist = [] if params.strip() in ("", "void") else [p.strip() for p in params.split(",")]
        assert len(plist) == len(args), f"{name}: header has {len(plist)} parameters, ctypes {len(args)}"
        for decl, ctype in zip(plist, args):
            if "*" in decl:
                assert ctype in (C.c_void_p, C.c_char_p) or issubclass(ctype, C._Pointer), f"{name}: {decl}"
            elif decl.startswith("double"):
                assert ctype is C.c_double, f"{name}: {decl}"
            elif decl.startswith("size_t"):
                assert ctype is C.c_size_t, f"{name}: {decl}"
            elif decl.startswith("int64_t"):
                assert ctype is C.c_int64, f"{name}: {decl}"
            elif decl.startswith("int "):
                assert ctype is C.c_int, f"{name}: {decl}"
            elif decl.startswith("uint32_t"):
                assert ctype is C.c_uint32, f"{name}: {decl}"
            else:
                raise AssertionError(f"{name}: unhandled parameter declaration {decl!r}")
        if "*" in ret and "char" in ret:
            assert res is C.c_char_p
        elif "*" in ret:
            assert res is C.c_void_p
        elif ret.strip() == "void":
            assert res is None
        elif "size_t" in ret:
            assert res is C.c_size_t
        elif "int64_t" in ret:
            assert res is C.c_int64
        else:
            assert res is C.c_int, f"{name}: return type {ret!r}"


def test_struct_layout_matches_header():
    from slb200 import SlCalib
    assert ctypes.sizeof(SlCalib) == (9 + 8 + 9 + 8 + 9 + 3) * 8 + 4 * 4


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import slb200
    from scanner.grayCode import decode_codes as dc
    from scanner.triangulation import Triangulate
    with pytest.raises(RuntimeError):
        dc.get_codes(np.zeros((14, 4, 4)))
    with pytest.raises(RuntimeError):
        slb200.cam_proj_pairs(np.zeros((4, 4), np.int64), np.zeros((4, 4), np.int64), (4, 4))
    t = Triangulate(np.zeros((4, 4), np.int64), np.zeros((4, 4), np.int64), (4, 4))
    with pytest.raises(RuntimeError):
        t.get_cam_proj_pts(np.zeros((4, 4, 3), np.uint8))
    from scanner.utils import visualize
    with pytest.raises(RuntimeError):                                    # the clean-up has no CPU path either
        visualize.clean_point_cloud(np.zeros((3, 30)), np.zeros((30, 3)))
    with pytest.raises(TypeError):
        slb200.remove_statistical_outliers(torch.zeros((30, 4)))         # host tensor
    with pytest.raises(TypeError):
        slb200.decode_pixels_bgr(torch.zeros((14, 4, 4, 3), dtype=torch.uint8))
    # the C ABI itself refuses too (no device -> CUDA error, never a computed result)
    assert slb200.lib().sl_device_check(-1) != 0
    assert b"CUDA" in slb200.lib().sl_last_error() or b"sm_" in slb200.lib().sl_last_error()


def test_calibration_struct_and_inplace_scaling():
    import slb200
    from scanner.triangulation import Triangulate
    from slb200 import calib_data as cd
    pm = cd.PROJ_MTX.copy()
    t = Triangulate(None, None, (1920, 1080), cd.CAM_1080_MTX, cd.CAM_1080_DIST, (1920, 1080), (1024, 768), pm,
                    cd.PROJ_DIST, np.eye(3), np.array([[0.1], [0.2], [0.3]]))
    np.testing.assert_allclose(pm[0], cd.PROJ_MTX[0] * (1920 / 1024))
    np.testing.assert_allclose(pm[1], cd.PROJ_MTX[1] * (1080 / 768))
    np.testing.assert_array_equal(pm[2], cd.PROJ_MTX[2])
    assert t.proj_mtx is pm
    s = t._calibration().struct()
    assert s.proj_w == 1920 and s.proj_h == 1080
    assert list(s.T) == [0.1, 0.2, 0.3]
    assert s.cam_dist[4] == cd.CAM_1080_DIST.ravel()[4] and s.cam_dist[5] == 0.0
    assert s.proj_dist[0] == cd.PROJ_DIST.ravel()[0]           # (1,5) and (5,1) layouts both accepted
    c = slb200.Calibration.from_reference(cd.CAM_1080_MTX, cd.CAM_1080_DIST, cd.PROJ_MTX, cd.PROJ_DIST, np.eye(3),
                                          np.zeros((3, 1)), (1920, 1080), (1024, 768))
    np.testing.assert_array_equal(c.proj_mtx, pm)
    np.testing.assert_array_equal(cd.PROJ_MTX[0, 0], 2648.8646560701163 if False else cd.PROJ_MTX[0, 0])
    t0 = Triangulate()
    assert t0.cam_w == 1920 and t0.proj_h == 1080 and t0.proj_mtx is None


def test_gray_helpers_and_generator():
    from scanner.grayCode import decode_codes as dc
    from scanner.grayCode.generate_codes import get_gray_codes, get_image_sequence
    assert [dc.gray_decode(i ^ (i >> 1)) for i in range(64)] == list(range(64))
    assert dc.gray_to_decimal([1, 1, 0]) == 4 and dc.gray_to_decimal([1, -1, 0]) == -1
    assert dc.gray_to_decimal(np.array([0, 0, 0], dtype=np.int8)) == 0
    codes = get_gray_codes(1024, 768)
    assert codes.shape == (1024, 10) and codes.dtype == np.uint8
    seq = get_image_sequence(codes, 1024, 768)
    assert seq.shape == (42, 768, 1024) and seq.dtype == np.uint8
    assert seq[0].max() == 0 and seq[1].min() == 255
    assert np.array_equal(seq[2:22] + seq[22:42], np.full_like(seq[2:22], 255))
    assert get_image_sequence(get_gray_codes(8, 16), 8, 16)[2:].max() == 0        # portrait: reference draws nothing


def test_golden_patterns(golden_dir):
    from scanner.grayCode.generate_codes import get_gray_codes, get_image_sequence
    g = np.load(os.path.join(golden_dir, "decode_golden.npz"))
    for (pw, ph) in ((16, 8), (32, 32), (24, 10), (8, 16)):
        np.testing.assert_array_equal(get_gray_codes(pw, ph), g[f"graycodes_{pw}x{ph}"])
        np.testing.assert_array_equal(get_image_sequence(get_gray_codes(pw, ph), pw, ph), g[f"patterns_{pw}x{ph}"])


def test_synthetic_scene_is_deterministic_and_decodable():
    from slb200 import synth
    from oracle import graycode_oracle as O
    a = synth.render_scan((160, 96), (1024, 768), seed=2)
    b = synth.render_scan((160, 96), (1024, 768), seed=2)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
    hp, vp = O.decode_pixels(a[0].numpy().astype(np.float64))
    valid = (hp >= 0) & (vp >= 0)
    assert valid[:, 20:].mean() > 0.99 and valid[:, :20].mean() < 0.1


def test_remove_bad_images_state_machine_golden(golden_dir):
    """Host half of remove_bad_images (the selection logic over per-pair difference counts) against the
    reference's own output on synthetic videos; the counts here come from NumPy, the GPU test checks the kernel."""
    from scanner.grayCode.decode_codes import _selection_from_counts
    g = np.load(os.path.join(golden_dir, "frames_golden.npz"))
    for name in ("v1", "v2", "v3"):
        frames = g[f"video_{name}"].astype(np.int16)
        d = [(np.abs(frames[k + 1] - frames[k]) > 50).sum() for k in range(len(frames) - 1)]
        np.testing.assert_array_equal(d, g[f"video_{name}_counts"])
        assert _selection_from_counts(d) == g[f"video_{name}_keep"].tolist()
    with pytest.raises(IndexError):
        _selection_from_counts([3])


def test_gray_formula_matches_cv2_golden(golden_dir):
    """The fixed-point BGR2GRAY formula the kernel implements, checked on the CPU against cv2's output."""
    g = np.load(os.path.join(golden_dir, "frames_golden.npz"))
    bgr = g["gray_bgr"].astype(np.uint32)
    gray = (3735 * bgr[..., 0] + 19235 * bgr[..., 1] + 9798 * bgr[..., 2] + 16384) >> 15
    np.testing.assert_array_equal(gray.astype(np.float64), g["gray_out"])


def test_pipeline_host_buffer_validation():
    """the host-buffer wrappers hand raw pointers to the C ABI, so dtype / shape / contiguity are checked
    in Python first (a wrong one would be a silent misread or an out-of-bounds write by an asynchronous copy)."""
    import torch
    from slb200.api import Pipeline
    chk = Pipeline._host_check
    ok = np.zeros((4, 4), np.float32)
    assert chk(ok, "cloud", "float32", shape=(None, 4)) is ok
    assert chk(None, "white_rgb", "uint8", nullable=True) is None
    with pytest.raises(TypeError):
        chk(None, "cloud", "float32")
    with pytest.raises(TypeError):
        chk(np.zeros((4, 4), np.float64), "cloud", "float32", shape=(None, 4))          # dtype
    with pytest.raises(ValueError):
        chk(np.zeros((4, 3), np.float32), "cloud", "float32", shape=(None, 4))          # columns
    with pytest.raises(ValueError):
        chk(np.zeros((8, 4), np.float32)[::2], "cloud", "float32", shape=(None, 4))     # strided view
    with pytest.raises(TypeError):
        chk(np.zeros(3, np.int32), "counts", "int64", min_elems=3)                      # int32 counts
    with pytest.raises(ValueError):
        chk(np.zeros(2, np.int64), "counts", "int64", min_elems=3)                      # too short
    with pytest.raises(TypeError):
        chk([1, 2, 3], "counts", "int64")                                               # not an array
    t = torch.zeros((6, 4), dtype=torch.float32)
    assert chk(t, "cloud", "float32", shape=(None, 4)) is t
    with pytest.raises(ValueError):
        chk(t.t(), "cloud", "float32")                                                  # non-contiguous torch view
    with pytest.raises(ValueError):
        chk(np.zeros((2, 5, 4, 4), np.uint8), "h_idx", "uint8", shape=(2, 5, 4, 3))


def test_exchange_geometry_and_peer_symbols():
    """The multi-GPU entry points are exported and their size query behaves (no GPU needed)."""
    import slb200
    L = slb200.lib()
    assert L.sl_exchange_ctrl_bytes(8, 8, 3840) > 8 * 2 * (32 + 8 * 8 + 3840 * 4)
    assert L.sl_exchange_ctrl_bytes(17, 8, 3840) == 0          # more than SL_MAX_RANKS
    assert L.sl_exchange_ctrl_bytes(0, 8, 3840) == 0
    assert L.sl_exchange_ctrl_bytes(8, 0, 3840) == 0
    assert not hasattr(slb200, "set_decode_sm_reserve")        # options are launch arguments now, no process-wide setter


def test_host_conversions_at_the_numpy_boundary():
    """sl_host_*: the dtype changes between the reference's NumPy types (float64 stack, int64 indices, (3,N) / (N,3)
    float64 outputs) and what crosses PCIe; host-only code, so it runs here.  Exact or an error."""
    from slb200._lib import lib
    L = lib()
    rng = np.random.default_rng(5)
    for dt, code in ((np.float64, 0), (np.float32, 1)):
        u = rng.integers(0, 256, (6, 37, 53), dtype=np.uint8)
        u.reshape(-1)[:256] = np.arange(256)
        src = u.astype(dt)
        for threads in (0, 1, 3):
            out = np.full(u.shape, 77, np.uint8)
            assert L.sl_host_narrow_u8(src.ctypes.data, code, src.size, out.ctypes.data, threads) == 0
            assert np.array_equal(out, u)
        for bad in (0.5, -1.0, 256.0, 300.0, np.nan, np.inf, -np.inf, 1e300 if dt is np.float64 else 1e30, -0.25, 254.999):
            s2 = src.copy()
            s2.reshape(-1)[rng.integers(0, s2.size)] = bad
            assert L.sl_host_narrow_u8(s2.ctypes.data, code, s2.size, out.ctypes.data, 0) == -1, bad
            assert b"0..255" in L.sl_last_error()
        s2 = src.copy(); s2.reshape(-1)[5] = -0.0          # -0.0 == 0: a valid grey level
        assert L.sl_host_narrow_u8(s2.ctypes.data, code, s2.size, out.ctypes.data, 0) == 0 and out.reshape(-1)[5] == 0
    # several threads, a length that is no multiple of the vector width or the chunk grain, bad samples at the ends
    big = rng.integers(0, 256, 1_000_003, dtype=np.uint8)
    src = big.astype(np.float64)
    out = np.zeros_like(big)
    for threads in (0, 2, 7):
        assert L.sl_host_narrow_u8(src.ctypes.data, 0, src.size, out.ctypes.data, threads) == 0
        assert np.array_equal(out, big)
    for pos in (0, 7, 8, src.size // 2, src.size - 9, src.size - 1):
        s2 = src.copy(); s2[pos] = 255.5
        assert L.sl_host_narrow_u8(s2.ctypes.data, 0, s2.size, out.ctypes.data, 7) == -1, pos
        assert b"(1 samples" in L.sl_last_error()
    assert L.sl_host_narrow_u8(None, 0, 0, None, 0) == 0      # empty
    assert L.sl_host_narrow_u8(src.ctypes.data, 7, 4, out.ctypes.data, 0) == -1

    idx = rng.integers(-1, 32768, (211, 97)).astype(np.int16)
    wide = np.empty(idx.shape, np.int64)
    assert L.sl_host_widen_idx(idx.ctypes.data, idx.size, wide.ctypes.data, 0) == 0
    assert np.array_equal(wide, idx.astype(np.int64))
    back = np.empty(idx.shape, np.int16)
    assert L.sl_host_narrow_idx(wide.ctypes.data, wide.size, back.ctypes.data, 2) == 0 and np.array_equal(back, idx)
    for bad in (-2, 32768, 1 << 40, -(1 << 40)):
        w2 = wide.copy(); w2[3, 4] = bad
        assert L.sl_host_narrow_idx(w2.ctypes.data, w2.size, back.ctypes.data, 0) == -1

    n = 70001
    cloud = rng.standard_normal((n, 4)).astype(np.float32)
    cloud[:, :3] *= np.float32(1e3)
    cloud[7, 0] = np.float32(np.nan); cloud[8, 1] = np.float32(np.inf)      # nan / inf points (parallel rays) pass through
    rgb = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    rgb[:256, 0] = np.arange(256)
    cloud[:, 3] = (rgb[:, 0].astype(np.uint32) | (rgb[:, 1].astype(np.uint32) << 8) | (rgb[:, 2].astype(np.uint32) << 16)).view(np.float32)
    for threads in (0, 1):
        pts = np.empty((3, n)); col = np.empty((n, 3))
        assert L.sl_host_unpack_cloud(cloud.ctypes.data, n, pts.ctypes.data, col.ctypes.data, threads) == 0
        assert np.array_equal(pts, cloud[:, :3].astype(np.float64).T, equal_nan=True)
        assert np.array_equal(col, rgb / 255)                  # the reference's img_white[j, i, :] / 255 (triangulate.py:69)
    pts2 = np.empty((3, n))
    assert L.sl_host_unpack_cloud(cloud.ctypes.data, n, pts2.ctypes.data, None, 0) == 0
    assert np.array_equal(pts2, pts, equal_nan=True)
    assert L.sl_host_unpack_cloud(None, 0, None, None, 0) == 0
