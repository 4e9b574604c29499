"""Multi-rank runs on real GPUs (2, 4 and 8 ranks where the box has them; skipped below 2 GPUs) of both sharding
modes, through both gathers:
  * by scan: each rank processes its block of scans; the gathered cloud == the single-GPU batch, bit for bit
      - CloudExchange, both transports ("emit": records emitted straight into the gathering rank's cloud over NVLink
        peer memory; "push": local emission, then sl_exchange_push copies the rank's block), gather to rank 0 / to the
        last rank and all-gather, several steps in a row (the control blocks are double-buffered by step parity)
      - gather_clouds over NCCL (all-gather and gather to rank 0)
  * by row band: each rank decodes + triangulates one band of the same capture
      - CloudExchange.band_to_cloud: per-column counts exchanged on the device, records land at their x-major
        position of the whole image == whole-image result (incl. ranks whose band is empty)
      - NCCL gather + merge_row_bands (the eager oracle of that placement)."""
import os
import socket
import sys

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
    from slb200.dist import CloudExchange, gather_clouds, merge_row_bands, row_bands, shard_range
    fails = []

    def expect(cond, what):
        if not bool(cond):
            fails.append(what)

    def same(a, b):
        return a.shape == b.shape and torch.equal(a.contiguous().view(torch.int32), b.contiguous().view(torch.int32))

    cam_size, proj_size = (256, 200), (1024, 768)
    W, H = cam_size
    calib = synth.synthetic_calibration(cam_size, proj_size)
    tables = slb200.LensTables(calib, cam_size, dev)
    n_scans = 2 * world + 1                      # uneven blocks
    scans = [synth.render_scan(cam_size, proj_size, seed=s, device="cpu") for s in range(n_scans + 2)]
    per_rank = [shard_range(n_scans, r, world)[1] - shard_range(n_scans, r, world)[0] for r in range(world)]
    lo, hi = shard_range(n_scans, rank, world)

    def batch(shift):                            # this rank's block / the whole batch, inputs rotated by `shift`
        ids = [(i + shift) % len(scans) for i in range(n_scans)]
        fr = torch.stack([scans[i][0] for i in ids]).to(dev)
        wh = torch.stack([scans[i][1] for i in ids]).to(dev)
        return fr, wh

    # ---- by scan, peer-memory exchange: gather to rank 0, then all-gather; 3 steps each
    last = world - 1
    for mode, transport in ((0, "emit"), ("all", "emit"), (0, "push"), ("all", "push"), (last, "push")):
        ex = CloudExchange(max_scans=max(per_rank), max_width=W, max_points=n_scans * H * W, gather_to=mode,
                           transport=transport, push_ctas=(3 if mode == last else 0))
        mode = f"{mode}/{transport}"
        for step in range(3):
            fr, wh = batch(step)
            ex.scan_to_cloud(fr[lo:hi].contiguous(), wh[lo:hi].contiguous(), calib, tables=tables)
            torch.cuda.synchronize()
            base, total = ex.status()
            ref = slb200.scan_to_cloud(fr, wh, calib, tables=tables)
            nr = int(ref.counts[-1])
            expect(total == nr, f"exchange[{mode}] step {step}: total {total} != {nr}")
            expect(base == int(ref.counts[:lo].sum()), f"exchange[{mode}] step {step}: rank base")
            expect(torch.equal(ex.counts_all[: n_scans + 1], ref.counts), f"exchange[{mode}] step {step}: counts_all")
            if ex.is_dst:
                expect(same(ex.cloud[:nr], ref.cloud[:nr]), f"exchange[{mode}] step {step}: gathered cloud != single-GPU batch")
            dist.barrier()                       # the reference run above must not overlap the next step's writes
        ex.close()

    # ---- two exchanges on two streams, batches alternating between them (the gather of one batch beside the compute of
    # the next): both gathered clouds are complete and exact
    exs = [CloudExchange(max_scans=max(per_rank), max_width=W, max_points=n_scans * H * W, gather_to=0,
                         transport=("emit", "push")[k]) for k in range(2)]
    sts = [torch.cuda.Stream(dev) for _ in range(2)]
    refs = []
    for step in range(4):
        fr, wh = batch(step)
        refs.append(slb200.scan_to_cloud(fr, wh, calib, tables=tables))
    torch.cuda.synchronize()
    for step in range(4):
        fr, wh = batch(step)
        sts[step & 1].wait_stream(torch.cuda.current_stream())       # the inputs were produced on the default stream
        with torch.cuda.stream(sts[step & 1]):
            exs[step & 1].scan_to_cloud(fr[lo:hi].contiguous(), wh[lo:hi].contiguous(), calib, tables=tables)
            fr.record_stream(sts[step & 1]); wh.record_stream(sts[step & 1])
        if step >= 1 and step % 2 == 1:          # after every pair: both exchanges hold a finished batch
            torch.cuda.synchronize()
            for k in range(2):
                ref = refs[step - 1 + k]
                nr = int(ref.counts[-1])
                base, total = exs[k].status()
                expect(total == nr, f"two-stream exchange step {step - 1 + k}: total")
                if rank == 0:
                    expect(same(exs[k].cloud[:nr], ref.cloud[:nr]), f"two-stream exchange step {step - 1 + k}: cloud")
            dist.barrier()
    for e_ in exs:
        e_.close()

    # ---- by scan, NCCL
    fr, wh = batch(0)
    ref = slb200.scan_to_cloud(fr, wh, calib, tables=tables)
    nr = int(ref.counts[-1])
    mine = slb200.scan_to_cloud(fr[lo:hi].contiguous(), wh[lo:hi].contiguous(), calib, tables=tables)
    n = int(mine.counts[-1])
    for dst in (None, 0):
        for spr in (None, per_rank):
            cloud_all, counts = gather_clouds(mine.cloud[:n], mine.counts, dst=dst, scans_per_rank=spr)
            got = [int(c) for r in counts for c in r[:-1]]
            expect(got == [int(c) for c in ref.counts[:-1]], f"nccl dst={dst}: per-scan counts")
            if dst is None or rank == dst:
                expect(same(cloud_all, ref.cloud[:nr]), f"nccl dst={dst}: gathered cloud != single-GPU batch")
            else:
                expect(cloud_all is None, "non-receiving rank got a cloud")

    # ---- by row band
    frames1, white1 = scans[0][0].to(dev), scans[0][1].to(dev)
    whole = slb200.scan_to_cloud(frames1, white1, calib, tables=tables)
    nw = int(whole.counts[-1])
    y0, y1 = row_bands(H, world)[rank]
    ex = CloudExchange(max_scans=1, max_width=W, max_points=H * W, gather_to=0)
    for step in range(2):
        ex.band_to_cloud(frames1[:, y0:y1], white1[y0:y1].contiguous(), calib, row0=y0, tables=tables)
        torch.cuda.synchronize()
        base, total = ex.status()
        expect(total == nw, f"row bands step {step}: total {total} != {nw}")
        if rank == 0:
            expect(same(ex.cloud[:nw], whole.cloud[:nw]), f"row bands step {step}: placed cloud != whole-image cloud")
            if step == 0:
                ex.cloud.zero_()
        dist.barrier()
    ex.close()
    # NCCL + eager merge (the oracle of the placement)
    if y1 > y0:
        band = slb200.scan_to_cloud(frames1[:, y0:y1], white1[y0:y1].contiguous(), calib, tables=tables, row0=y0)
        nb = int(band.counts[-1])
        cols = band.column_counts()[0].to(torch.int64).contiguous()
        bcloud, bcounts = band.cloud[:nb], band.counts
    else:
        cols = torch.zeros(W, dtype=torch.int64, device=dev)
        bcloud, bcounts = torch.zeros((0, 4), dtype=torch.float32, device=dev), torch.zeros(2, dtype=torch.int64, device=dev)
    all_cols = [torch.empty_like(cols) for _ in range(world)]
    dist.all_gather(all_cols, cols)
    cloud_cat, _ = gather_clouds(bcloud, bcounts)
    sizes = [int(c.sum()) for c in all_cols]
    parts, start = [], 0
    for s_ in sizes:
        parts.append(cloud_cat[start:start + s_]); start += s_
    merged = merge_row_bands(parts, all_cols)
    expect(same(merged, whole.cloud[:nw]), "row bands over NCCL + merge_row_bands != whole-image cloud")

    q.put((rank, fails))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_rank_sharding(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=600) for _ in procs]
    finally:
        for p in procs:
            p.join(timeout=120)
            if p.is_alive():
                p.kill()
    for p in procs:
        assert p.exitcode == 0
    assert sorted(res) == [(r, []) for r in range(world)], res
