"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: scan sharding, the all-gather-v of
packed clouds, and the row-band merge back into the reference's x-major order."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "3dscanner-graycode_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from slb200.dist import gather_clouds, shard_range
    n_scans = 5
    rng = np.random.default_rng(0)
    per_scan = [int(x) for x in rng.integers(0, 50, n_scans)]
    clouds = [torch.from_numpy(rng.random((n, 4)).astype(np.float32)) for n in per_scan]
    lo, hi = shard_range(n_scans, rank, world)
    mine = torch.cat(clouds[lo:hi]) if hi > lo else torch.zeros((0, 4))
    counts = torch.tensor(per_scan[lo:hi] + [sum(per_scan[lo:hi])], dtype=torch.int64)
    cloud_all, cnts = gather_clouds(mine, counts)
    ok = torch.equal(cloud_all, torch.cat(clouds))
    got = [int(c) for r in cnts for c in r[:-1]]
    ok = ok and got == per_scan
    # gather to one rank, block sizes known up front (no length exchange), into a caller-supplied pre-sized cloud
    per_rank = [shard_range(n_scans, r, world)[1] - shard_range(n_scans, r, world)[0] for r in range(world)]
    out = torch.full((sum(per_scan) + 3, 4), -1.0) if rank == 1 else None
    got_cloud, cnts2 = gather_clouds(mine, counts, dst=1, scans_per_rank=per_rank, out=out)
    if rank == 1:
        ok = ok and torch.equal(got_cloud, torch.cat(clouds)) and got_cloud.data_ptr() == out.data_ptr()
        ok = ok and bool((out[sum(per_scan):] == -1).all())          # nothing written past the gathered records
    else:
        ok = ok and got_cloud is None
    ok = ok and [int(c) for r in cnts2 for c in r[:-1]] == per_scan and [int(r[-1]) for r in cnts2] == [sum(per_scan[a:b]) for a, b in (shard_range(n_scans, r, world) for r in range(world))]
    # a rank with nothing to contribute
    empty = torch.zeros((0, 4)) if rank == 0 else mine
    ecounts = torch.tensor([0, 0], dtype=torch.int64) if rank == 0 else counts
    g3, _ = gather_clouds(empty, ecounts)
    want3 = torch.cat(clouds[shard_range(n_scans, 1, world)[0]:])
    ok = ok and torch.equal(g3, want3)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_clouds_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_range_and_bands():
    from slb200.dist import row_bands, shard_range
    for n in (0, 1, 7, 64):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    for h in (2160, 1080, 100):
        for w in (2, 4, 8):
            bands = row_bands(h, w)
            assert bands[0][0] == 0 and bands[-1][1] == h
            assert all(bands[i][1] == bands[i + 1][0] for i in range(w - 1))
            assert all(a % 32 == 0 for a, b in bands if b > a)


def test_merge_row_bands_restores_x_major_order():
    """Row-band clouds merged == the single-image x-major cloud (oracle ordering)."""
    from oracle import graycode_oracle as O
    from slb200.dist import merge_row_bands, row_bands
    rng = np.random.default_rng(3)
    h, w = 100, 37
    hp = rng.integers(-1, 50, (h, w)); vp = rng.integers(-1, 50, (h, w))
    white = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
    cam, proj, col = O.cam_proj_pairs(hp, vp, (64, 64), white)
    full = np.concatenate([cam, proj, np.round(col * 255)], axis=1).astype(np.float32)
    bx, bcols = [], []
    for (y0, y1) in row_bands(h, 3):
        c, p, cl = O.cam_proj_pairs(hp[y0:y1], vp[y0:y1], (64, 64), white[y0:y1])
        c = c + np.array([[0, y0]], dtype=np.float32)
        bx.append(torch.from_numpy(np.concatenate([c, p, np.round(cl * 255)], axis=1).astype(np.float32)))
        valid = (hp[y0:y1] != -1) & (vp[y0:y1] != -1)
        bcols.append(torch.from_numpy(valid.sum(axis=0).astype(np.int64)))
    merged = merge_row_bands(bx, bcols)
    np.testing.assert_array_equal(merged.numpy(), full)
