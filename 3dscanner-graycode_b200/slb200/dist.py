"""Multi-GPU plumbing: one process per GPU (torch.distributed), scans sharded across ranks.

The path has no inter-pixel dependence, so ranks never talk during decode / triangulation.  The
only exchange is the final gather of the compacted clouds (all-gather-v over NCCL / NVLink; gloo on
CPU tensors in the tests).  Row-band sharding of one huge capture is expressed the same way: each
band is decoded as its own scan and the bands' clouds are merged by `merge_row_bands`.
"""
from __future__ import annotations

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


def gather_clouds(cloud: torch.Tensor, counts: torch.Tensor, group=None):
    """All-gather-v of per-rank packed clouds.

    cloud (n_r, C) (XYZRGB records, C = 4 f32), counts (B_r+1,) int64 with counts[-1] == n_r.
    Returns (cloud_all, counts_per_rank): clouds concatenated in rank order (scan order is preserved
    because scans are block-partitioned); counts_per_rank is a list of the per-scan count tensors.
    Three collectives: sizes, payload (padded to the largest rank), counts.
    """
    world = dist.get_world_size(group)
    dev = cloud.device
    n_local = torch.tensor([cloud.shape[0], counts.numel()], dtype=torch.int64, device=dev)
    sizes = [torch.empty_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    n_pts = [int(s[0]) for s in sizes]
    n_cnt = [int(s[1]) for s in sizes]
    cap, ccap = max(n_pts), max(n_cnt)

    def allgather(t, rows):
        shape = (rows,) + tuple(t.shape[1:])
        mine = torch.zeros(shape, dtype=t.dtype, device=dev)
        mine[: t.shape[0]] = t
        out = torch.empty((world * rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(out, mine, group=group)
        return out

    all_pts = allgather(cloud, cap)
    cloud_cat = torch.cat([all_pts[r * cap: r * cap + n_pts[r]] for r in range(world)])
    cnt_all = allgather(counts, ccap)
    counts_per_rank = [cnt_all[r * ccap: r * ccap + n_cnt[r]] for r in range(world)]
    return cloud_cat, counts_per_rank


def merge_row_bands(band_clouds: Sequence[torch.Tensor], band_cols: Sequence[torch.Tensor]):
    """Merge the clouds of row bands (top to bottom) of ONE capture into the reference's x-major
    order (scanner/triangulation/triangulate.py:52-53): for every camera column, band 0's points,
    then band 1's, ...  band_cols[k] is the int64 (W,) per-column point count of band k (columns in
    a band's cloud are already x-major).  Returns the merged (N, C) cloud."""
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
    """One capture cut into row bands, each decoded + triangulated on its own (this is what every rank
    of a row-band-sharded job does for its band), then merged back into x-major order.

    On a multi-GPU job call it with the single band of this rank and all-gather the per-band results
    (`gather_clouds` for the records, `all_gather` for the (W,) column counts) before `merge_row_bands`;
    with a list of bands on one GPU it is the reference implementation of that flow (tests).
    Returns (merged cloud (N,4), list of per-band clouds, list of per-band column counts).
    """
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
