"""Two-rank NCCL run of both sharding modes (needs >= 2 GPUs; skipped otherwise):
  * by scan: each rank processes its block of scans, all-gather-v of the packed clouds == single-GPU batch;
  * by row band: each rank decodes + triangulates one band of the same capture, clouds and per-column counts
    are gathered and merged back into the reference's x-major order == whole-image result."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "3dscanner-graycode_b200"))
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import slb200
    from slb200 import synth
    from slb200.dist import gather_clouds, merge_row_bands, row_bands, shard_range
    cam_size, proj_size = (256, 200), (1024, 768)
    calib = synth.synthetic_calibration(cam_size, proj_size)
    tables = slb200.LensTables(calib, cam_size, dev)
    ok = True
    # ---- by scan
    n_scans = 5
    scans = [synth.render_scan(cam_size, proj_size, seed=s, device="cpu") for s in range(n_scans)]
    lo, hi = shard_range(n_scans, rank, world)
    frames = torch.stack([scans[i][0] for i in range(lo, hi)]).to(dev)
    white = torch.stack([scans[i][1] for i in range(lo, hi)]).to(dev)
    mine = slb200.scan_to_cloud(frames, white, calib, tables=tables)
    n = int(mine.counts[-1])
    cloud_all, counts = gather_clouds(mine.cloud[:n], mine.counts)
    if rank == 0:
        ref = slb200.scan_to_cloud(torch.stack([s[0] for s in scans]).to(dev), torch.stack([s[1] for s in scans]).to(dev),
                                   calib, tables=tables)
        nr = int(ref.counts[-1])
        ok &= cloud_all.shape[0] == nr and torch.equal(cloud_all.view(torch.int32), ref.cloud[:nr].view(torch.int32))
        got = [int(c) for r in counts for c in r[:-1]]
        ok &= got == [int(c) for c in ref.counts[:-1]]
    # ---- by row band
    frames1, white1 = scans[0][0].to(dev), scans[0][1].to(dev)
    y0, y1 = row_bands(cam_size[1], world)[rank]
    band = slb200.scan_to_cloud(frames1[:, y0:y1], white1[y0:y1].contiguous(), calib, tables=tables, row0=y0)
    nb = int(band.counts[-1])
    cols = band.column_counts()[0].to(torch.int64).contiguous()
    all_cols = [torch.empty_like(cols) for _ in range(world)]
    dist.all_gather(all_cols, cols)
    cloud_cat, _ = gather_clouds(band.cloud[:nb], band.counts)
    sizes = [int(c.sum()) for c in all_cols]
    parts, start = [], 0
    for s_ in sizes:
        parts.append(cloud_cat[start:start + s_]); start += s_
    merged = merge_row_bands(parts, all_cols)
    if rank == 0:
        whole = slb200.scan_to_cloud(frames1, white1, calib, tables=tables)
        nw = int(whole.counts[-1])
        ok &= merged.shape[0] == nw and torch.equal(merged.view(torch.int32), whole.cloud[:nw].view(torch.int32))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_nccl_sharding():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]
