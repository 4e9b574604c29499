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


def gather_clouds(xyz: torch.Tensor, rgb: Optional[torch.Tensor], counts: torch.Tensor, group=None):
    """All-gather-v of per-rank packed clouds.

    xyz (n_r,3) f32, rgb (n_r,3) u8 (or None), counts (B_r+1,) int64 with counts[-1] == n_r.
    Returns (xyz_all, rgb_all, counts_per_rank) where clouds are concatenated in rank order (scan
    order is preserved because scans are block-partitioned) and counts_per_rank is a list of the
    per-scan count tensors.  Two collectives: one for the counts, one per payload (padded to the
    largest rank).
    """
    world = dist.get_world_size(group)
    dev = xyz.device
    n_local = torch.tensor([xyz.shape[0], counts.numel()], dtype=torch.int64, device=dev)
    sizes = [torch.empty_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    n_pts = [int(s[0]) for s in sizes]
    n_cnt = [int(s[1]) for s in sizes]
    cap, ccap = max(n_pts), max(n_cnt)

    def padded(t, rows, width, dtype):
        buf = torch.zeros((rows, width) if width else (rows,), dtype=dtype, device=dev)
        buf[: t.shape[0]] = t
        return buf

    def allgather(t, rows, width, dtype):
        mine = padded(t, rows, width, dtype)
        out = torch.empty((world * rows, width) if width else (world * rows,), dtype=dtype, device=dev)
        dist.all_gather_into_tensor(out, mine, group=group)
        return out

    xyz_all = allgather(xyz, cap, 3, torch.float32)
    xyz_cat = torch.cat([xyz_all[r * cap: r * cap + n_pts[r]] for r in range(world)])
    rgb_cat = None
    if rgb is not None:
        rgb_all = allgather(rgb, cap, 3, torch.uint8)
        rgb_cat = torch.cat([rgb_all[r * cap: r * cap + n_pts[r]] for r in range(world)])
    cnt_all = allgather(counts, ccap, 0, torch.int64)
    counts_per_rank = [cnt_all[r * ccap: r * ccap + n_cnt[r]] for r in range(world)]
    return xyz_cat, rgb_cat, counts_per_rank


def merge_row_bands(band_xyz: Sequence[torch.Tensor], band_rgb: Sequence[Optional[torch.Tensor]],
                    band_cols: Sequence[torch.Tensor]):
    """Merge the clouds of row bands (top to bottom) of ONE capture into the reference's x-major
    order (scanner/triangulation/triangulate.py:52-53): for every camera column, band 0's points,
    then band 1's, ...  band_cols[k] is the int64 (W,) per-column point count of band k (columns in
    a band's cloud are already x-major).  Returns (xyz, rgb)."""
    cols = torch.stack(list(band_cols))                        # (K, W)
    k, w = cols.shape
    # destination of each (band, column) run: columns first, bands inside a column
    flat = cols.t().contiguous().reshape(-1)                   # order: col0 band0..K-1, col1 ...
    start = torch.cumsum(flat, 0) - flat
    start = start.reshape(w, k).t()                            # (K, W) start offset of each run
    total = int(flat.sum())
    xyz = torch.empty((total, 3), dtype=band_xyz[0].dtype, device=band_xyz[0].device)
    rgb = None
    if band_rgb[0] is not None:
        rgb = torch.empty((total, 3), dtype=band_rgb[0].dtype, device=band_rgb[0].device)
    for b in range(k):
        n = int(cols[b].sum())
        src_start = torch.cumsum(cols[b], 0) - cols[b]          # (W,) start of each column run in the band cloud
        col_of_pt = torch.repeat_interleave(torch.arange(w, device=cols.device), cols[b])
        within = torch.arange(n, device=cols.device) - src_start[col_of_pt]
        dst = start[b][col_of_pt] + within
        xyz[dst] = band_xyz[b][:n]
        if rgb is not None:
            rgb[dst] = band_rgb[b][:n]
    return xyz, rgb
