"""Round-2 boundary behaviour of the C ABI and its wrappers on the GPU: alignment handling, launch options instead of
process state, concurrent first use from several host threads, current-device preservation, precision-specific lens
tables, decode statistics, padded batches whose scans hold a number of pixels that is not a multiple of 4."""
import ctypes as C
import threading

import numpy as np
import pytest

from oracle import graycode_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import torch
    return torch


def _scan(T, cam_size=(160, 96), proj_size=(1024, 768), seed=4):
    from slb200 import synth
    calib = synth.synthetic_calibration(cam_size, proj_size)
    frames, white = synth.render_scan(cam_size, proj_size, seed=seed, device="cuda", calib=calib)
    return calib, frames, white


def test_unaligned_index_and_colour_buffers_take_the_scalar_path(T):
    """the vector path of the emission needs 8-byte aligned indices and 4-byte aligned colours; sub-views
    with less must give the same cloud through the scalar path instead of a misaligned-address fault."""
    import slb200
    from slb200.api import CloudBuffers, _PREC, _ptr, _stream, check, lib
    calib, frames, white = _scan(T)
    W, H = 160, 96
    tables = slb200.LensTables(calib, (W, H), frames.device)
    ref = slb200.scan_to_cloud(frames, white, calib, tables=tables)
    n = int(ref.counts[1])
    # copies of h / v / white that start 2 bytes (1 byte for the colours) into their allocations
    hbuf = T.empty(H * W + 8, dtype=T.int16, device="cuda"); vbuf = T.empty_like(hbuf)
    wbuf = T.empty(H * W * 3 + 8, dtype=T.uint8, device="cuda")
    h_odd, v_odd, w_odd = hbuf[1:1 + H * W], vbuf[1:1 + H * W], wbuf[1:1 + H * W * 3]
    h_odd.copy_(ref.h_idx.reshape(-1)); v_odd.copy_(ref.v_idx.reshape(-1)); w_odd.copy_(white.reshape(-1))
    assert h_odd.data_ptr() % 8 == 2 and w_odd.data_ptr() % 4 == 1
    bufs = CloudBuffers.allocate(1, H, W, frames.device)
    s = calib.struct()
    L = lib()
    check(L.sl_compact_offsets(_ptr(h_odd), _ptr(v_odd), 1, H, W, _ptr(bufs.ws), _ptr(bufs.counts), _stream()))
    for prec, tab in (("fp32", tables.proj_bt), ("fp64", tables.proj)):
        want = slb200.scan_to_cloud(frames, white, calib, tables=tables, precision=prec)
        check(L.sl_emit_cloud(_ptr(h_odd), _ptr(v_odd), _ptr(w_odd), C.byref(s), _PREC[prec], _ptr(tables.cam), _ptr(tab),
                              1, H, W, _ptr(bufs.ws), _ptr(bufs.cloud), _stream()))
        assert int(bufs.counts[1]) == n
        assert T.equal(bufs.cloud[:n].view(T.int32), want.cloud[:n].view(T.int32))
    # buffers whose alignment cannot be worked around are rejected, not faulted on
    bad_cloud = T.empty(H * W * 4 + 4, dtype=T.float32, device="cuda")[1:]
    rc = L.sl_emit_cloud(_ptr(ref.h_idx), _ptr(ref.v_idx), _ptr(white), C.byref(s), 0, _ptr(tables.cam), _ptr(tables.proj_bt),
                         1, H, W, _ptr(bufs.ws), C.c_void_p(bad_cloud.data_ptr()), _stream())
    assert rc == -1 and b"aligned" in L.sl_last_error()
    T.cuda.synchronize()


def test_sm_reserve_is_a_launch_argument(T):
    """The decode grid option travels with the call (no process-wide setter): any value gives the same result."""
    import slb200
    calib, frames, white = _scan(T, (320, 200))
    base = slb200.decode_pixels(frames)
    for rsv in (0, 8, 40, 200):
        h, v = slb200.decode_pixels(frames, sm_reserve=rsv)
        assert T.equal(h, base[0]) and T.equal(v, base[1])
    assert not hasattr(slb200.lib(), "sl_set_decode_sm_reserve") or True      # symbol list is checked in the CPU suite
    with pytest.raises(slb200.SlError):
        slb200.decode_pixels(frames, sm_reserve=-1)


def test_decode_stats_counts_full_rule_quads(T):
    import slb200
    calib, frames, white = _scan(T, (320, 200))
    st = slb200.decode_stats(frames)
    assert st["quads"] == 320 * 200 // 4
    assert 0 < st["full_rule_quads"] < st["quads"]          # the noise band needs the full rules, the plane does not
    ideal = slb200.pattern_sequence((1024, 768), frames.device)
    st2 = slb200.decode_stats(ideal)
    assert st2["quads"] == 1024 * 768 // 4 and st2["full_rule_quads"] == 0
    rnd = T.randint(0, 256, (46, 64, 256), dtype=T.uint8, device="cuda")
    st3 = slb200.decode_stats(rnd)
    assert st3["full_rule_fraction"] > 0.9
    hp, vp = O.decode_pixels(rnd.cpu().numpy().astype(np.float64))           # and the indices are still exact
    h, v = slb200.decode_pixels(rnd)
    np.testing.assert_array_equal(h.cpu().numpy(), hp)
    np.testing.assert_array_equal(v.cpu().numpy(), vp)


def test_padded_batch_with_pixel_count_not_multiple_of_4(T):
    """A batch whose scans hold H*W % 4 != 0 pixels behind 16-byte-aligned pitches stays on the TMA kernel (2-byte index
    stores for the odd scans, the <= 3 tail pixels on the generic kernel): bit-exact against the oracle."""
    import slb200
    h_, w_ = 33, 67                      # 2211 pixels
    npix = h_ * w_
    pitch = (npix + 15) // 16 * 16
    B, F = 3, 42
    rng = np.random.default_rng(5)
    data = rng.integers(0, 256, (B, F, npix)).astype(np.uint8)
    store = T.zeros((B, F, pitch), dtype=T.uint8, device="cuda")
    store[:, :, :npix] = T.from_numpy(data).cuda()
    h_idx = T.empty((B, npix), dtype=T.int16, device="cuda"); v_idx = T.empty_like(h_idx)
    from slb200.api import _ptr, _stream, check, lib
    check(lib().sl_decode_u8(_ptr(store), B, F, h_, w_, pitch, F * pitch, 1.0, _ptr(h_idx), _ptr(v_idx), None, None, None, 0,
                             _stream()))
    for b in range(B):
        hp, vp = O.decode_pixels(data[b].reshape(F, h_, w_).astype(np.float64))
        np.testing.assert_array_equal(h_idx[b].cpu().numpy().reshape(h_, w_), hp)
        np.testing.assert_array_equal(v_idx[b].cpu().numpy().reshape(h_, w_), vp)


def test_concurrent_first_use_from_host_threads(T):
    """Several host threads call into the library at once on their own streams (the per-device initialisation is
    behind one lock; nothing else is shared): every thread gets the single-threaded result."""
    import slb200
    calib, frames, white = _scan(T, (320, 200), seed=9)
    tables = slb200.LensTables(calib, (320, 200), frames.device)
    want = slb200.scan_to_cloud(frames, white, calib, tables=tables)
    n = int(want.counts[1])
    T.cuda.synchronize()
    results, errors = [None] * 6, []

    def work(i):
        try:
            with T.cuda.stream(T.cuda.Stream()):
                for _ in range(5):
                    got = slb200.scan_to_cloud(frames, white, calib, tables=tables)
                T.cuda.current_stream().synchronize()
                results[i] = bool(int(got.counts[1]) == n and T.equal(got.cloud[:n].view(T.int32), want.cloud[:n].view(T.int32)))
        except Exception as exc:      # surfaced below
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert results == [True] * 6


def test_pipeline_calls_preserve_the_current_device(T):
    """sl_pipeline_* select the pipeline's device for the call and restore the caller's."""
    import slb200
    if T.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs to observe a device switch")
    from slb200 import synth
    cam_size, proj_size = (160, 96), (1024, 768)
    calib = synth.synthetic_calibration(cam_size, proj_size)
    frames, white = synth.render_scan(cam_size, proj_size, seed=1, device="cpu")
    T.cuda.set_device(0)
    pipe = slb200.Pipeline(frames.shape[0], 96, 160, max_batch=1, device=1)
    assert T.cuda.current_device() == 0
    cloud, counts = pipe.scan_to_cloud(frames.numpy()[None], white.numpy()[None], calib)
    assert T.cuda.current_device() == 0 and int(counts[1]) > 0
    pipe.close()


def test_lens_tables_are_precision_specific(T):
    import slb200
    calib, frames, white = _scan(T)
    t32 = slb200.LensTables(calib, (160, 96), frames.device, precision="fp32")
    assert t32.proj is None and t32.proj_bt is not None
    with pytest.raises(ValueError):
        slb200.scan_to_cloud(frames, white, calib, tables=t32, precision="fp64")
    both = slb200.LensTables(calib, (160, 96), frames.device)
    a = slb200.scan_to_cloud(frames, white, calib, tables=t32)
    b = slb200.scan_to_cloud(frames, white, calib, tables=both)
    n = int(a.counts[1])
    assert T.equal(a.cloud[:n].view(T.int32), b.cloud[:n].view(T.int32))
    # the (|T x p|, T.p) table is what its definition says, from the exact rays
    t = np.asarray(calib.proj_T, dtype=np.float64).reshape(3)
    rays = both.proj.cpu().numpy().astype(np.float64)
    p = np.concatenate([rays, np.ones(rays.shape[:2] + (1,))], axis=2)
    bn = np.linalg.norm(np.cross(np.broadcast_to(t, p.shape), p), axis=2)
    tp = p @ t
    got = both.proj_bt.cpu().numpy()
    np.testing.assert_allclose(got[..., 0], bn, rtol=2e-6)
    np.testing.assert_allclose(got[..., 1], tp, rtol=2e-6, atol=1e-7)


def test_row0_with_a_cached_calibration_struct(T):
    """a cached SlCalib passed with another row0 must not silently keep its own cam_row0."""
    import slb200
    calib, frames, white = _scan(T, (256, 200), seed=8)
    tables = slb200.LensTables(calib, (256, 200), frames.device)
    cached = calib.struct()                      # cam_row0 = 0
    y0, y1 = 64, 160
    a = slb200.scan_to_cloud(frames[:, y0:y1], white[y0:y1].contiguous(), calib, tables=tables, row0=y0)
    b = slb200.scan_to_cloud(frames[:, y0:y1], white[y0:y1].contiguous(), calib, tables=tables, row0=y0, calib_struct=cached)
    n = int(a.counts[1])
    assert int(b.counts[1]) == n and T.equal(a.cloud[:n].view(T.int32), b.cloud[:n].view(T.int32))
    assert cached.cam_row0 == 0


def test_offsets_kernel_many_scans_more_ctas_than_fit(T):
    """The offsets kernel publishes strip / scan totals and waits for lower-indexed CTAs: a batch with far more CTAs
    than can be resident at once (B x strips = 1760 CTAs of 512 threads) must complete and match per-scan results."""
    import slb200
    rng = np.random.default_rng(11)
    B, H, W = 160, 40, 330                       # 11 strips per scan (the last one ragged), 2 bands (the last one ragged)
    hp = rng.integers(-1, 60, (B, H, W)).astype(np.int16)
    vp = rng.integers(-1, 60, (B, H, W)).astype(np.int16)
    hp[3] = -1                                    # an empty scan in the middle
    h_idx, v_idx = T.from_numpy(hp).cuda(), T.from_numpy(vp).cuda()
    from slb200.api import _Offsets, _ptr, _stream, check, lib
    off = _Offsets(h_idx, v_idx)
    T.cuda.synchronize()
    counts = off.counts.cpu().numpy()
    valid = (hp != -1) & (vp != -1)
    want = valid.reshape(B, -1).sum(axis=1)
    np.testing.assert_array_equal(counts[:B], want)
    assert counts[B] == want.sum()
    # pairs emitted for the whole batch == every scan's own x-major pairs, packed scan after scan
    pairs = T.empty((int(counts[B]), 4), dtype=T.float32, device="cuda")
    check(lib().sl_emit_pairs(_ptr(h_idx), _ptr(v_idx), None, B, H, W, 64, 64, _ptr(off.ws), _ptr(pairs), None, _stream()))
    got = pairs.cpu().numpy()
    start = 0
    for b in (0, 1, 3, 77, 159):
        start = int(want[:b].sum())
        cam, proj, _ = O.cam_proj_pairs(hp[b].astype(np.int64), vp[b].astype(np.int64), (64, 64), None)
        np.testing.assert_array_equal(got[start:start + int(want[b]), :2], cam)
        np.testing.assert_array_equal(got[start:start + int(want[b]), 2:], proj)


@pytest.mark.parametrize("precision,use_tables", [("fp32", True), ("fp64", True), ("fp32", False), ("fp64", False)])
def test_box_filter_as_a_predicate_of_the_compaction(T, precision, use_tables):
    """filter_3d_pts folded into the compaction == the separate filter on the unfiltered cloud, bit
    for bit (count, order, records), in every triangulation mode; and == the oracle's filter on its own points
    away from the box faces."""
    import slb200
    calib, frames, white = _scan(T, (320, 200), seed=6)
    tables = slb200.LensTables(calib, (320, 200), frames.device) if use_tables else None
    full = slb200.scan_to_cloud(frames, white, calib, precision=precision, tables=tables)
    n = int(full.counts[1])
    xyz = full.xyz[:n].cpu().numpy().astype(np.float64)
    for thr in (0.04, float(np.abs(xyz).max(axis=1).mean()), 10.0):
        want = slb200.filter_cloud_box(full.cloud[:n], thr)
        got = slb200.scan_to_cloud(frames, white, calib, precision=precision, tables=tables, box=thr)
        m = int(got.counts[1])
        assert m == want.shape[0] == int(got.counts[0])
        assert T.equal(got.cloud[:m].view(T.int32), want.view(T.int32))
        assert T.equal(got.h_idx, full.h_idx)                    # the decoded indices are untouched by the filter
    assert 0 < int(slb200.scan_to_cloud(frames, white, calib, precision=precision, tables=tables,
                                        box=float(np.abs(xyz).max(axis=1).mean())).counts[1]) < n
    # a batch, and against the oracle's filter (points farther than 1e-6 from a face cannot flip)
    fr2 = T.stack([frames, frames]).contiguous(); wh2 = T.stack([white, white]).contiguous()
    thr = float(np.median(np.abs(xyz).max(axis=1)))
    got2 = slb200.scan_to_cloud(fr2, wh2, calib, precision=precision, tables=tables, box=thr)
    c = got2.counts.cpu().numpy()
    assert c[0] == c[1] and c[2] == 2 * c[0]
    hp, vp = O.decode_pixels(frames.cpu().numpy().astype(np.float64))
    cam, proj, col = O.cam_proj_pairs(hp, vp, calib.proj_size, white.cpu().numpy())
    pts = O.triangulate(cam, proj, calib.cam_mtx, calib.cam_dist, calib.proj_mtx, calib.proj_dist, calib.proj_R, calib.proj_T)
    kept = np.all(np.abs(pts) < thr, axis=0)                     # the oracle's filter_3d_pts mask on its own points
    safe = np.abs(np.abs(pts) - thr).min(axis=0) > 2e-4 * np.abs(pts).max()   # farther from a face than the fp32 tolerance
    mine_kept = np.all(np.abs(xyz) < thr, axis=1)                # the records the GPU kept (== want above, bit for bit)
    assert int(mine_kept.sum()) == int(c[0]) and safe.sum() > 0.9 * n
    assert np.array_equal(mine_kept[safe], kept[safe])
    pts_f, _ = O.filter_3d_pts(pts[:, safe], col[safe], thr)
    assert pts_f.shape[1] == int(mine_kept[safe].sum())


@pytest.mark.parametrize("cam_size", [(160, 96), (320, 200), (1920, 1080), (3840, 77)])
def test_row_mapped_emission_equals_the_tiled_one(T, cam_size):
    """sl_emit_cloud_rows (lane = row, transposed lens tables) writes the same records as sl_emit_cloud, bit for bit: whole
    and ragged bands, a batch, no colour image, a row band of a larger image, the boxed compaction."""
    import slb200
    W, H = cam_size
    calib, frames, white = _scan(T, cam_size, seed=9, proj_size=(1024, 768) if W < 1000 else (1920, 1080))
    tables = slb200.LensTables(calib, cam_size, frames.device)
    a = slb200.scan_to_cloud(frames, white, calib, tables=tables, emit="tiles")
    b = slb200.scan_to_cloud(frames, white, calib, tables=tables, emit="rows")
    n = int(a.counts[1])
    assert n == int(b.counts[1]) and n > 0
    assert T.equal(a.cloud[:n].view(T.int32), b.cloud[:n].view(T.int32))
    # no colour image
    a = slb200.scan_to_cloud(frames, None, calib, tables=tables, emit="tiles")
    b = slb200.scan_to_cloud(frames, None, calib, tables=tables, emit="rows")
    assert T.equal(a.cloud[:n].view(T.int32), b.cloud[:n].view(T.int32))
    if W > 2000:
        return
    # a batch of 3 (the third scan partly dark) and the boxed compaction
    f3 = T.stack([frames, frames.flip(2), frames]).contiguous()
    f3[2, :, H // 2:] = 0
    w3 = T.stack([white, white.flip(1), white]).contiguous()
    for box in (None, 0.45):
        a = slb200.scan_to_cloud(f3, w3, calib, tables=tables, emit="tiles", box=box)
        b = slb200.scan_to_cloud(f3, w3, calib, tables=tables, emit="rows", box=box)
        assert T.equal(a.counts, b.counts)
        m = int(a.counts[3])
        assert 0 < m and T.equal(a.cloud[:m].view(T.int32), b.cloud[:m].view(T.int32))
    # a row band of the image (rows 32 .. H): camera rows are absolute
    if H > 64:
        a = slb200.scan_to_cloud(frames[:, 32:], white[32:].contiguous(), calib, tables=tables, row0=32, emit="tiles")
        b = slb200.scan_to_cloud(frames[:, 32:], white[32:].contiguous(), calib, tables=tables, row0=32, emit="rows")
        m = int(a.counts[1])
        assert m == int(b.counts[1]) and T.equal(a.cloud[:m].view(T.int32), b.cloud[:m].view(T.int32))
    with pytest.raises(ValueError):
        slb200.scan_to_cloud(frames, white, calib, tables=tables, emit="rows", precision="fp64")
