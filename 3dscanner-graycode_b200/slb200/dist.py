"""Multi-GPU plumbing: one process per GPU (torch.distributed), work sharded across ranks.

The path has no inter-pixel dependence, so ranks never talk during decode / triangulation.  The only
exchange is the gather of the compacted clouds (BASELINE.json north_star), offered two ways:

* ``CloudExchange`` -- the product path on an NVLink box: every rank's emit kernel writes its records straight
  into the gathering rank's cloud through a CUDA-IPC peer mapping, at their final packed offsets (the counts are
  exchanged on the device first, ``sl_exchange_*`` in include/slb200.h).  No host synchronisation, no padding,
  no concatenation; works for scans sharded over ranks and for row bands of one capture.
* ``gather_clouds`` -- the same result with NCCL (gloo on CPU tensors in the tests): one all-gather of the counts,
  then grouped send/recv of every rank's records directly into a pre-sized output at their final offsets.

``merge_row_bands`` is the eager restatement of the row-band merge; the tests use it as the oracle of the
device-side placement.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition [lo, hi) of n_items over `world` ranks (sizes differ by <= 1)."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def row_bands(height: int, world: int, align: int = 32) -> List[Tuple[int, int]]:
    """Row bands [y0, y1) for splitting one capture over `world` GPUs; boundaries are multiples of
    `align` rows (the compaction band height) so each band's columns compact independently."""
    blocks = (height + align - 1) // align
    out = []
    for r in range(world):
        lo, hi = shard_range(blocks, r, world)
        out.append((min(height, lo * align), min(height, hi * align)))
    return out


# --------------------------------------------------------------------------- NCCL / gloo gather
def gather_clouds(cloud: torch.Tensor, counts: torch.Tensor, group=None, dst: Optional[int] = None,
                  scans_per_rank: Optional[Sequence[int]] = None, out: Optional[torch.Tensor] = None):
    """Gather-v of per-rank packed clouds into ONE pre-sized tensor, every rank's records at their final offset.

    cloud (n_r, C) (XYZRGB records, C = 4 f32), counts (B_r+1,) int64 with counts[-1] == n_r.
    ``dst=None``: all-gather (every rank gets everything); ``dst=r``: only rank r receives (the others return
    ``(None, counts_per_rank)``).  ``scans_per_rank`` (B_r of every rank, e.g. from ``shard_range``) saves the
    collective that would otherwise exchange the lengths of the count vectors.
    Returns (cloud_all, counts_per_rank): clouds in rank order (scan order is preserved because scans are
    block-partitioned); counts_per_rank is a list of the per-rank count tensors (host, int64, (B_r+1,)).
    Collectives: one all-gather of the (padded) counts, one grouped send/recv of the payload; one host read of the
    counts in between (the receive sizes are host arguments of NCCL).
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = cloud.device
    if scans_per_rank is None:
        n_local = torch.tensor([counts.numel() - 1], dtype=torch.int64, device=dev)
        sizes = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sizes, n_local, group=group)
        scans_per_rank = [int(v) for v in sizes.tolist()]
    if len(scans_per_rank) != world or counts.numel() != scans_per_rank[rank] + 1:
        raise ValueError("scans_per_rank does not match the world size / this rank's count vector")
    ccap = max(scans_per_rank) + 1
    mine = torch.zeros(ccap, dtype=torch.int64, device=dev)
    mine[: counts.numel() - 1] = counts[:-1]
    mine[ccap - 1] = counts[-1]                                   # the rank's total sits in the last slot
    cnt_all = torch.empty(world * ccap, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(cnt_all, mine, group=group)
    host = cnt_all.cpu().view(world, ccap)                         # the one host read
    n_pts = [int(host[r, ccap - 1]) for r in range(world)]
    counts_per_rank = [torch.cat([host[r, : scans_per_rank[r]], host[r, ccap - 1:]]) for r in range(world)]
    if cloud.shape[0] != n_pts[rank]:
        raise ValueError("cloud length does not match counts[-1]")
    offs = [0]
    for n in n_pts:
        offs.append(offs[-1] + n)
    receiver = dst is None or rank == dst
    tail = tuple(cloud.shape[1:])
    if receiver:
        if out is None:
            out = torch.empty((offs[-1],) + tail, dtype=cloud.dtype, device=dev)
        elif out.shape[0] < offs[-1] or tuple(out.shape[1:]) != tail or out.dtype != cloud.dtype:
            raise ValueError("out is too small or of another record type")
    ops = []
    for r in range(world):
        if r == rank:
            continue
        if receiver and n_pts[r] > 0:
            ops.append(dist.P2POp(dist.irecv, out[offs[r]: offs[r + 1]], r if group is None else dist.get_global_rank(group, r), group))
        if (dst is None or r == dst) and n_pts[rank] > 0:
            ops.append(dist.P2POp(dist.isend, cloud, r if group is None else dist.get_global_rank(group, r), group))
    if receiver and n_pts[rank] > 0:
        out[offs[rank]: offs[rank + 1]].copy_(cloud)
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return (out[: offs[-1]] if receiver else None), counts_per_rank


# --------------------------------------------------------------------------- row-band merge (oracle of the placement)
def merge_row_bands(band_clouds: Sequence[torch.Tensor], band_cols: Sequence[torch.Tensor]):
    """Merge the clouds of row bands (top to bottom) of ONE capture into the reference's x-major
    order (scanner/triangulation/triangulate.py:52-53): for every camera column, band 0's points,
    then band 1's, ...  band_cols[k] is the int64 (W,) per-column point count of band k (columns in
    a band's cloud are already x-major).  Returns the merged (N, C) cloud.  Eager PyTorch: the tests' oracle of
    what CloudExchange places on the device."""
    cols = torch.stack(list(band_cols))                        # (K, W)
    k, w = cols.shape
    # destination of each (band, column) run: columns first, bands inside a column
    flat = cols.t().contiguous().reshape(-1)                   # order: col0 band0..K-1, col1 ...
    start = torch.cumsum(flat, 0) - flat
    start = start.reshape(w, k).t()                            # (K, W) start offset of each run
    total = int(flat.sum())
    first = band_clouds[0]
    out = torch.empty((total,) + tuple(first.shape[1:]), dtype=first.dtype, device=first.device)
    for b in range(k):
        n = int(cols[b].sum())
        src_start = torch.cumsum(cols[b], 0) - cols[b]          # (W,) start of each column run in the band cloud
        col_of_pt = torch.repeat_interleave(torch.arange(w, device=cols.device), cols[b])
        within = torch.arange(n, device=cols.device) - src_start[col_of_pt]
        dst = start[b][col_of_pt] + within
        out[dst] = band_clouds[b][:n]
    return out


def scan_to_cloud_row_bands(frames: torch.Tensor, white_rgb: Optional[torch.Tensor], calib, tables,
                            bands: Sequence[Tuple[int, int]], precision="fp32", eps: float = 1.0):
    """One capture cut into row bands on ONE GPU, each decoded + triangulated on its own and merged with
    ``merge_row_bands`` -- the single-process restatement of the row-band flow (tests).
    Returns (merged cloud (N,4), list of per-band clouds, list of per-band column counts)."""
    import slb200
    clouds, cols = [], []
    for (y0, y1) in bands:
        if y1 <= y0:
            clouds.append(torch.zeros((0, 4), dtype=torch.float32, device=frames.device))
            cols.append(torch.zeros((frames.shape[-1],), dtype=torch.int64, device=frames.device))
            continue
        band = frames[:, y0:y1]
        white = white_rgb[y0:y1].contiguous() if white_rgb is not None else None
        bufs = slb200.scan_to_cloud(band, white, calib, precision=precision, eps=eps, tables=tables, row0=y0)
        n = int(bufs.counts[1])
        clouds.append(bufs.cloud[:n].clone())
        cols.append(bufs.column_counts()[0].to(torch.int64).clone())
    return merge_row_bands(clouds, cols), clouds, cols


# --------------------------------------------------------------------------- peer memory (CUDA IPC)
class _DeviceMemoryView:
    """``__cuda_array_interface__`` over a raw device allocation, so torch can view it without a copy."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class PeerBuffer:
    """Device memory of THIS process that the other ranks map (``sl_peer_alloc`` / ``sl_peer_export``)."""

    def __init__(self, nbytes: int):
        from ._lib import check, lib
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        check(lib().sl_peer_alloc(self.nbytes, C.byref(p)))
        self.ptr = int(p.value)

    def handle(self) -> bytes:
        from ._lib import check, lib
        buf = C.create_string_buffer(64)
        check(lib().sl_peer_export(C.c_void_p(self.ptr), buf))
        return bytes(buf.raw)

    def tensor(self, dtype=torch.uint8) -> torch.Tensor:
        t = torch.as_tensor(_DeviceMemoryView(self.ptr, self.nbytes), device=torch.device("cuda", torch.cuda.current_device()))
        return t.view(dtype)

    def free(self):
        from ._lib import lib
        if self.ptr:
            lib().sl_peer_free(C.c_void_p(self.ptr))
            self.ptr = 0


def _open_peer(handle: bytes) -> int:
    from ._lib import check, lib
    p = C.c_void_p()
    check(lib().sl_peer_open(C.create_string_buffer(handle, 64), C.byref(p)))
    return int(p.value)


class CloudExchange:
    """Gather of the compacted clouds over NVLink peer memory, fused into the emission.

    One object per process group; every rank constructs it with the same arguments (collective: the IPC handles are
    exchanged with ``all_gather_object``).  ``gather_to``: the rank that receives the cloud, or ``"all"``.
    ``max_points``: capacity (records) of the gathered cloud; ``max_scans``: most scans a rank handles per step;
    ``max_width``: widest image (row-band mode).  All ranks must issue the same sequence of steps, each on one stream.

    After ``scan_to_cloud`` / ``band_to_cloud`` (asynchronous, no host synchronisation) the receiving rank(s) hold
    ``cloud[:counts_all[-1]]`` = the packed cloud of the whole job, identical to what one GPU produces for the whole
    batch / the whole image: scan-major, every scan in the reference's x-major order (triangulate.py:52-53).
    """

    def __init__(self, max_scans: int, max_width: int, max_points: int, group=None, gather_to=0,
                 timeout_s: float = 20.0, transport: str = "emit", push_ctas: int = 0):
        from ._lib import check, lib
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world > 16:
            raise ValueError("CloudExchange supports up to 16 ranks")
        if transport not in ("emit", "push"):
            raise ValueError("transport must be 'emit' or 'push'")
        # "emit": the emission writes through the peer mapping (both sharding modes).  "push" (scan-sharded steps
        # only; row bands interleave per column and always use "emit"): emit locally, then sl_exchange_push copies
        # the rank's block with push_ctas CTAs -- the better choice when the next batch's compute can overlap it.
        self.transport, self.push_ctas = transport, int(push_ctas)
        # the emission through the peer mappings: row-mapped kernel where it applies (SLB200_PEER_EMIT=tiles: always tiled)
        self.peer_rows = os.environ.get("SLB200_PEER_EMIT", "tiles") == "rows"
        self._stage = None
        self.max_points = int(max_points)
        self.max_scans, self.max_width = int(max_scans), int(max_width)
        self.dsts = list(range(self.world)) if gather_to == "all" else [int(gather_to)]
        if any(d < 0 or d >= self.world for d in self.dsts):
            raise ValueError("gather_to outside the group")
        self.is_dst = self.rank in self.dsts
        self.device = torch.device("cuda", torch.cuda.current_device())
        L = lib()
        nbytes = L.sl_exchange_ctrl_bytes(self.world, self.max_scans, self.max_width)
        if nbytes == 0:
            raise ValueError("bad exchange geometry")
        self._ctrl = PeerBuffer(nbytes)
        self._cloud = PeerBuffer(self.max_points * 16) if self.is_dst else None
        mine = {"ctrl": self._ctrl.handle(), "cloud": self._cloud.handle() if self._cloud else None}
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        self._opened = []
        ctrl_ptrs, cloud_ptrs = [], {}
        for r, h in enumerate(everyone):
            if r == self.rank:
                ctrl_ptrs.append(self._ctrl.ptr)
                if self._cloud:
                    cloud_ptrs[r] = self._cloud.ptr
                continue
            p = _open_peer(h["ctrl"])
            self._opened.append(p)
            ctrl_ptrs.append(p)
            if r in self.dsts:
                q = _open_peer(h["cloud"])
                self._opened.append(q)
                cloud_ptrs[r] = q
        # this rank's own cloud first (local stores keep their L2 policy ordering irrelevant; order is free)
        order = ([self.rank] if self.is_dst else []) + [d for d in self.dsts if d != self.rank]
        self._cloud_ptr_list = [cloud_ptrs[d] for d in order]
        self._clouds_c = (C.c_void_p * len(order))(*self._cloud_ptr_list)
        others = [cloud_ptrs[d] for d in self.dsts if d != self.rank]
        self._other_ptr_list = others
        self._others_c = (C.c_void_p * max(1, len(others)))(*others)
        arr = (C.c_void_p * self.world)(*ctrl_ptrs)
        self._x = C.c_void_p()
        check(L.sl_exchange_create(C.byref(self._x), self.rank, self.world, arr, self.max_scans, self.max_width,
                                   float(timeout_s)))
        self.dst_mask = 0
        for d in self.dsts:
            self.dst_mask |= 1 << d
        self.counts_all = torch.zeros(max(self.world * self.max_scans, self.world) + 1, dtype=torch.int64, device=self.device)
        self.cloud = self._cloud.tensor(torch.float32).view(-1, 4) if self._cloud else None
        dist.barrier(group=group)           # every mapping exists before anyone's first step

    # -- pieces (each asynchronous on the current stream)
    def offsets_by_scan(self, buffers, b: int, h: int, w: int):
        from ._lib import check, lib
        from .api import _ptr, _stream
        check(lib().sl_exchange_offsets_by_scan(self._x, b, h, w, _ptr(buffers.ws), _ptr(buffers.counts),
                                                _ptr(self.counts_all), _stream()))

    def offsets_row_band(self, buffers, h: int, w: int):
        from ._lib import check, lib
        from .api import _ptr, _stream
        check(lib().sl_exchange_offsets_row_band(self._x, h, w, _ptr(buffers.ws), _ptr(buffers.counts),
                                                 _ptr(self.counts_all), _stream()))

    def emit(self, buffers, white_rgb, calib_struct, precision, tables, b: int, h: int, w: int):
        from ._lib import check, lib
        from .api import _PREC, _ptr, _stream
        if (self.peer_rows and tables is not None and precision == "fp32" and w % 32 == 0 and tables.proj_bt is not None
                and (white_rgb is None or white_rgb.data_ptr() % 16 == 0)):
            cam_tt, proj_tt = tables.transposed()
            check(lib().sl_emit_cloud_rows_multi(_ptr(buffers.h_idx), _ptr(buffers.v_idx), _ptr(white_rgb),
                                                 C.byref(calib_struct), _ptr(cam_tt), int(cam_tt.shape[1]), _ptr(proj_tt),
                                                 b, h, w, _ptr(buffers.ws), self._clouds_c, len(self._cloud_ptr_list), _stream()))
            return
        cam_t = _ptr(tables.cam) if tables is not None else None
        proj_t = _ptr(tables.projector_table(precision)) if tables is not None else None
        check(lib().sl_emit_cloud_multi(_ptr(buffers.h_idx), _ptr(buffers.v_idx), _ptr(white_rgb), C.byref(calib_struct),
                                        _PREC[precision], cam_t, proj_t, b, h, w, _ptr(buffers.ws), self._clouds_c,
                                        len(self._cloud_ptr_list), self.max_points, _stream()))

    def counts_by_scan(self, buffers, b: int, h: int, w: int):
        """Count exchange without rebasing the local offsets (``sl_exchange_offsets_by_scan`` with a null workspace)."""
        from ._lib import check, lib
        from .api import _ptr, _stream
        check(lib().sl_exchange_offsets_by_scan(self._x, b, h, w, None, _ptr(buffers.counts), _ptr(self.counts_all),
                                                _stream()))

    def _emit_and_push(self, buffers, white_rgb, calib_struct, precision, tables, b: int, h: int, w: int):
        """transport="push": local emission, then the rank's block goes to the gathering rank(s) as one stream."""
        from ._lib import check, lib
        from .api import _PREC, _ptr, _stream
        L, st = lib(), _stream()
        cam_t = _ptr(tables.cam) if tables is not None else None
        proj_t = _ptr(tables.projector_table(precision)) if tables is not None else None
        total = C.c_void_p(buffers.counts.data_ptr() + 8 * b)
        rows = (tables is not None and precision == "fp32" and w % 32 == 0 and tables.proj_bt is not None
                and (white_rgb is None or white_rgb.data_ptr() % 16 == 0))

        def emit_local(cloud_ptr, max_points):
            # the offsets in the workspace say where (local or rebased positions); both emissions just follow them
            if rows:
                cam_tt, proj_tt = tables.transposed()
                check(L.sl_emit_cloud_rows(_ptr(buffers.h_idx), _ptr(buffers.v_idx), _ptr(white_rgb), C.byref(calib_struct),
                                           _ptr(cam_tt), int(cam_tt.shape[1]), _ptr(proj_tt), b, h, w, _ptr(buffers.ws),
                                           cloud_ptr, st))
            elif max_points:
                check(L.sl_emit_cloud_at(_ptr(buffers.h_idx), _ptr(buffers.v_idx), _ptr(white_rgb), C.byref(calib_struct),
                                         _PREC[precision], cam_t, proj_t, b, h, w, _ptr(buffers.ws), cloud_ptr, max_points, st))
            else:
                check(L.sl_emit_cloud(_ptr(buffers.h_idx), _ptr(buffers.v_idx), _ptr(white_rgb), C.byref(calib_struct),
                                      _PREC[precision], cam_t, proj_t, b, h, w, _ptr(buffers.ws), cloud_ptr, st))
        if self.is_dst:
            # my own records go straight to their place in my gathered cloud; the other receivers get a copy of it
            self.offsets_by_scan(buffers, b, h, w)
            emit_local(C.c_void_p(self._cloud.ptr), self.max_points)
            if self._other_ptr_list:
                check(L.sl_exchange_push(self._x, C.c_void_p(self._cloud.ptr), 1, total, self._others_c,
                                         len(self._other_ptr_list), self.max_points, self.push_ctas, st))
        else:
            self.counts_by_scan(buffers, b, h, w)
            need = b * h * w
            if self._stage is None or self._stage.shape[0] < need:
                self._stage = torch.empty((need, 4), dtype=torch.float32, device=self.device)
            emit_local(_ptr(self._stage), 0)
            check(L.sl_exchange_push(self._x, _ptr(self._stage), 0, total, self._others_c, len(self._other_ptr_list),
                                     self.max_points, self.push_ctas, st))

    def arrive(self):
        from ._lib import check, lib
        from .api import _stream
        check(lib().sl_exchange_arrive(self._x, self.dst_mask, 1 if self.is_dst else 0, _stream()))

    # -- whole steps
    def scan_to_cloud(self, frames: torch.Tensor, white_rgb: Optional[torch.Tensor], calib, precision="fp32",
                      eps: float = 1.0, buffers=None, calib_struct=None, tables=None, sm_reserve: int = 0):
        """This rank's block of scans ``(B_r,F,H,W)``: decode -> offsets -> count exchange -> emission into the
        gathering rank's cloud -> arrival flags.  ``counts_all[:total_scans + 1]`` = per-scan counts + total."""
        from . import api
        from ._lib import check, lib
        api._check_stack(frames)
        b, f, h, w, _ = api._stack_dims(frames)
        if b > self.max_scans:
            raise ValueError(f"{b} scans per step, the exchange was built for {self.max_scans}")
        fp, sp = api._pitches(frames)
        if buffers is None:
            buffers = api.CloudBuffers.allocate(b, h, w, frames.device, cap_points=1)
        s = calib_struct if calib_struct is not None else calib.struct()
        L, st = lib(), api._stream()
        check(L.sl_decode_u8(api._ptr(frames), b, f, h, w, fp, sp, float(eps), api._ptr(buffers.h_idx),
                             api._ptr(buffers.v_idx), None, None, api._ptr(buffers.ws), int(sm_reserve), st))
        check(L.sl_compact_offsets_decoded(b, h, w, api._ptr(buffers.ws), api._ptr(buffers.counts), st))
        if self.transport == "push":
            self._emit_and_push(buffers, white_rgb, s, precision, tables, b, h, w)
        else:
            self.offsets_by_scan(buffers, b, h, w)
            self.emit(buffers, white_rgb, s, precision, tables, b, h, w)
        self.arrive()
        return buffers

    def band_to_cloud(self, band_frames: torch.Tensor, band_white: Optional[torch.Tensor], calib, row0: int,
                      precision="fp32", eps: float = 1.0, buffers=None, tables=None, sm_reserve: int = 0):
        """This rank's band of rows ``(F,H_r,W)`` (rows ``[row0, row0 + H_r)`` of the capture, boundaries multiples
        of 32 as ``row_bands`` gives them): the merged, x-major cloud of the whole image lands on the gathering
        rank; ``counts_all[:world + 1]`` = points per band + total.  An empty band still takes part (pass a
        ``(F,0,W)`` tensor)."""
        from . import api
        from ._lib import check, lib
        f, h, w = band_frames.shape
        if w > self.max_width:
            raise ValueError(f"width {w}, the exchange was built for {self.max_width}")
        L, st = lib(), api._stream()
        if h == 0:
            # nothing to decode: publish zero column totals so the other ranks' offsets are right
            if buffers is None:
                buffers = api.CloudBuffers.allocate(1, 32, w, band_frames.device, cap_points=1)
            buffers.ws.zero_()
            buffers.counts.zero_()
            self.offsets_row_band(buffers, 32, w)
            self.arrive()
            return buffers
        api._check_stack(band_frames)
        fp, sp = api._pitches(band_frames)
        if buffers is None:
            buffers = api.CloudBuffers.allocate(1, h, w, band_frames.device, cap_points=1)
        s = calib.struct(row0)
        check(L.sl_decode_u8(api._ptr(band_frames), 1, f, h, w, fp, sp, float(eps), api._ptr(buffers.h_idx),
                             api._ptr(buffers.v_idx), None, None, api._ptr(buffers.ws), int(sm_reserve), st))
        check(L.sl_compact_offsets_decoded(1, h, w, api._ptr(buffers.ws), api._ptr(buffers.counts), st))
        self.offsets_row_band(buffers, h, w)
        self.emit(buffers, band_white, s, precision, tables, 1, h, w)
        self.arrive()
        return buffers

    def status(self):
        """Synchronous: raises if a wait for a peer timed out; returns (this rank's first record, total records)
        of the last step."""
        from ._lib import check, lib
        err, base, total = C.c_int(), C.c_int64(), C.c_int64()
        check(lib().sl_exchange_status(self._x, C.byref(err), C.byref(base), C.byref(total)))
        if err.value:
            if err.value == 3:
                raise RuntimeError("cloud exchange: the gathered cloud (max_points) is too small for this step")
            raise RuntimeError(f"cloud exchange: a wait for a peer rank timed out (stage {err.value})")
        return int(base.value), int(total.value)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def close(self):
        """Collective: every rank calls it (nobody unmaps while a peer may still write)."""
        from ._lib import lib
        L = lib()
        if getattr(self, "_x", None):
            torch.cuda.synchronize()
            try:
                dist.barrier(group=self.group)      # nobody unmaps while a peer may still write
            except Exception:
                pass
            L.sl_exchange_destroy(self._x)
            self._x = None
            self.cloud = None
            self._stage = None
            for p in self._opened:
                L.sl_peer_close(C.c_void_p(p))
            self._opened = []
            try:
                dist.barrier(group=self.group)
            except Exception:
                pass
            self._ctrl.free()
            if self._cloud:
                self._cloud.free()
