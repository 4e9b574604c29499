#!/usr/bin/env python
"""Benchmark of the Gray-code decode + triangulation hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (N=1): BASELINE config[1] -- one 1920x1080 camera scan of a 1920x1080 projector sequence
(n = 11 bits, F = 46 frames), synthetic textured plane (slb200/synth.py).  A step = one scan through
decode -> ordered compaction -> undistort + triangulate.  Steps rotate over a ring of resident
scans larger than L2 so no step finds its input in cache.  N > 1: every rank runs the same
per-rank workload on its own scans (weak scaling by scan, no data-path collective); the
all-gather-v of the compacted clouds is timed separately and reported under "gather".

Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e` the same metric through
the host-buffer C-ABI call (sl_scan_to_cloud_host: pinned host frames in, host cloud out), `roofline`
the fused decode kernel against the measured HBM peak, `cpu_baseline` the oracle port timed on this
box's host cores.  `--impl reference` times the reference algorithm's CPU port (oracle/) on all
host cores for the driver's ratio.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "3dscanner-graycode_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

CAM_SIZE = (1920, 1080)
PROJ_SIZE = (1920, 1080)
N_BITS = 11
N_FRAMES = 4 * N_BITS + 2
MPX = CAM_SIZE[0] * CAM_SIZE[1] / 1e6
METRIC = "Mpx/s decoded+triangulated"
WORKLOAD = "config[1]: single 1920x1080 camera scan of a 1920x1080 projector Gray sequence (11+11 bits + inverses + white/black)"
RING = 12         # resident scans per rank: 12 x 95.4 MB = 1.14 GB > 126 MB L2
STREAMS = 4       # independent scans are pipelined over 4 CUDA streams (own buffers each)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """Polls SM clock / throttle reasons with NVML while the timed regions run."""

    def __init__(self, index: int, period_s: float = 0.02):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self.period = period_s
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80, "sync_boost": 0x10,
                 "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# --------------------------------------------------------------------------- CPU side (oracle port)
def _oracle_calib(calib):
    return dict(cam_mtx=calib.cam_mtx, cam_dist=calib.cam_dist, proj_mtx=calib.proj_mtx, proj_dist=calib.proj_dist,
                proj_R=calib.proj_R, proj_T=calib.proj_T, proj_size=PROJ_SIZE, proj_calib_size=PROJ_SIZE)


def _oracle_band(args):
    """One bounded sample: rows [y0, y1) of a scan through the oracle's whole path."""
    import numpy as np
    from oracle import graycode_oracle as oracle
    frames, white, cal = args
    h_px, v_px, pts, colors = oracle.scan_to_cloud(frames.astype(np.float64), white, cal)
    return int(pts.shape[1])


def ncu_traffic_per_launch():
    """dram__bytes_read + dram__bytes_write of the decode kernel per launch, from the committed ncu launch
    list of the same workload (profiles/r01_launches.csv); None if the file is missing."""
    path = os.path.join(ROOT, "profiles", "r01_launches.csv")
    try:
        rd = wr = None
        with open(path) as f:
            for line in f:
                if "decode_tma_kernel" not in line:
                    continue
                parts = line.strip().split(",")
                if "dram__bytes_read.sum" in line:
                    rd = float(parts[-2])
                elif "dram__bytes_write.sum" in line:
                    wr = float(parts[-2])
        return None if rd is None or wr is None else rd + wr
    except Exception:
        return None


def cpu_baseline_single_core(frames_np, white_np, calib, budget_s=15.0):
    """Oracle port, one thread, on a bounded sample of the 1080p scan (whole rows)."""
    import numpy as np
    cal = _oracle_calib(calib)
    rows = 135                                   # 1/8 of the scan per call
    t0 = time.perf_counter()
    done_px, calls = 0, 0
    y = 0
    while True:
        band = slice(y, y + rows)
        _oracle_band((frames_np[:, band], white_np[band], cal))
        done_px += rows * CAM_SIZE[0]
        calls += 1
        y = (y + rows) % (CAM_SIZE[1] - rows + 1)
        el = time.perf_counter() - t0
        if el > budget_s or calls >= 96:
            break
    return {"value": done_px / 1e6 / el, "unit": "Mpx/s", "cores": 1, "kind": "port",
            "sample": f"{calls} bands of {rows} rows x 1920 px of the 1080p scan ({done_px / 1e6:.2f} Mpx, {el:.1f} s), "
                      "oracle/graycode_oracle.py scan_to_cloud (NumPy restatement of the reference)"}


def run_reference(args):
    """`--impl reference`: the reference algorithm's CPU port (oracle/) on all host cores.  The
    reference itself is pure Python (no build) and cannot travel to the GPU box; the oracle is its
    NumPy restatement, pinned bit-exact against it (tests/test_oracle_*).  Each step = one band of rows
    of the 1080p scan per worker process."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    import numpy as np
    from slb200 import synth
    cores = os.cpu_count() or 1
    calib = synth.synthetic_calibration(CAM_SIZE, PROJ_SIZE)
    frames, white = synth.render_scan(CAM_SIZE, PROJ_SIZE, seed=0, device="cpu")
    frames_np, white_np = frames.numpy(), white.numpy()
    cal = _oracle_calib(calib)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        # size the per-step sample so warmup + steps fit in ~100 s
        probe_rows = 8
        t0 = time.perf_counter()
        pool.map(_oracle_band, [(frames_np[:, i * probe_rows:(i + 1) * probe_rows], white_np[i * probe_rows:(i + 1) * probe_rows], cal)
                                for i in range(cores)])
        probe = time.perf_counter() - t0
        per_row = probe / probe_rows
        budget = 100.0 / max(1, args.steps + args.warmup)
        rows = int(max(2, min(CAM_SIZE[1] // cores, budget / max(per_row, 1e-6))))

        def step(k):
            jobs = []
            for i in range(cores):
                y = ((k * cores + i) * rows) % (CAM_SIZE[1] - rows + 1)
                jobs.append((frames_np[:, y:y + rows], white_np[y:y + rows], cal))
            pool.map(_oracle_band, jobs)

        for k in range(args.warmup):
            step(k)
        t0 = time.perf_counter()
        for k in range(args.steps):
            step(k)
        el = time.perf_counter() - t0
    px = args.steps * cores * rows * CAM_SIZE[0]
    value = px / 1e6 / el
    sample = f"per step: {cores} bands of {rows} rows x 1920 px (one per worker process) of the 1080p scan"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpx/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Mpx/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    global RING
    RING = max(2, args.ring)
    import numpy as np
    import torch
    import torch.distributed as dist
    import slb200
    from slb200 import synth
    from slb200.api import CloudBuffers

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    slb200.lib()                                           # fail loudly if the extension is missing
    W, H = CAM_SIZE
    K, Wm = args.steps, args.warmup
    calib = synth.synthetic_calibration(CAM_SIZE, PROJ_SIZE)
    cstruct = calib.struct()

    # ---- synthetic inputs, resident in HBM: a ring of scans (different seeds per rank)
    pats = synth.pattern_stack(PROJ_SIZE, dev)
    ring_frames, ring_white = [], []
    for i in range(RING):
        fr, wh = synth.render_scan(CAM_SIZE, PROJ_SIZE, seed=rank * RING + i, device=dev, patterns=pats, calib=calib)
        ring_frames.append(fr); ring_white.append(wh)
    del pats
    S = max(1, min(args.streams, RING))
    # the persistent decode grid leaves a few SMs to the offsets / emit kernels of the scans in flight on the other
    # streams (sl_set_decode_sm_reserve; launch geometry is fixed when the graphs below are captured)
    sm_reserve = (8 if S > 1 else 0) if args.sm_reserve is None else max(0, args.sm_reserve)
    slb200.set_decode_sm_reserve(sm_reserve)
    bufsets = [CloudBuffers.allocate(1, H, W, dev) for _ in range(S)]
    bufs = bufsets[0]
    tables = None if args.lens == "iterate" else slb200.LensTables(calib, CAM_SIZE, dev)   # once per calibration
    torch.cuda.synchronize()

    def step(i, k):
        slb200.scan_to_cloud(ring_frames[i % RING], ring_white[i % RING], calib, precision="fp32",
                             buffers=bufsets[k], calib_struct=cstruct, tables=tables)

    # One CUDA graph per ring slot (the 3 launches of a step replay without host launch latency); slot i
    # always runs on stream i % S with that stream's buffers, so consecutive scans overlap: the small
    # compaction / emit kernels of one scan run beside the decode of the next.
    before = slb200.launch_count()
    step(0, 0)
    kernels_per_step = slb200.launch_count() - before
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(dev) for _ in range(S)]
    graphs = []
    for i in range(RING):
        with torch.cuda.stream(streams[i % S]):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=streams[i % S]):
                step(i, i % S)
            graphs.append(g)
    torch.cuda.synchronize()
    main = torch.cuda.current_stream()

    def replay(graph_list, n):
        for i in range(n):
            k = i % RING
            with torch.cuda.stream(streams[k % S]):
                graph_list[k].replay()

    def timed(graph_list, n):
        """n replays bracketed by CUDA events: the start event gates every stream, the end event waits for all."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for st in streams:
            st.wait_event(e0)
        replay(graph_list, n)
        for st in streams:
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)
        e1.record(main)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    sampler = ClockSampler(local)
    sampler.start()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: EXACTLY K steps
    replay(graphs, Wm)
    barrier()
    ms = torch.tensor([timed(graphs, K)], dtype=torch.float64, device=dev)
    barrier()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    valid_pts = int(bufs.counts[1].item())

    # ---- roofline of the dominant kernel (fused decode): decode-only launches over the same ring, timed
    # with CUDA events (a) in the configuration the step uses (pipelined over the S streams) and
    # (b) isolated: back to back on one stream, every launch waiting for the previous one.
    kr = max(60, min(K, 600))
    dgraphs = []
    for i in range(RING):
        with torch.cuda.stream(streams[i % S]):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=streams[i % S]):
                slb200.decode_pixels(ring_frames[i], out=(bufsets[i % S].h_idx[0], bufsets[i % S].v_idx[0]))
            dgraphs.append(g)
    torch.cuda.synchronize()
    replay(dgraphs, 10)
    torch.cuda.synchronize()
    decode_ms = timed(dgraphs, kr) / kr
    # (b) isolated: ONE graph holding RING decode launches in sequence on one stream (kernel after kernel, no
    # overlap between launches; the graph only removes the host launch latency).  Nothing else runs, so the
    # grid covers every SM.
    slb200.set_decode_sm_reserve(0)
    with torch.cuda.stream(streams[0]):
        chain = torch.cuda.CUDAGraph()
        with torch.cuda.graph(chain, stream=streams[0]):
            for i in range(RING):
                slb200.decode_pixels(ring_frames[i], out=(bufs.h_idx[0], bufs.v_idx[0]))
        chain.replay()
        reps = max(3, kr // RING)
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record(streams[0])
        for _ in range(reps):
            chain.replay()
        d1.record(streams[0])
    torch.cuda.synchronize()
    decode_solo_ms = d0.elapsed_time(d1) / (reps * RING)
    slb200.set_decode_sm_reserve(sm_reserve)
    peak, peak_src = load_peaks()
    algo_bytes = (N_FRAMES + 4) * W * H                      # F u8 reads + two int16 writes per pixel
    achieved = algo_bytes / (decode_ms * 1e-3) / 1e9
    achieved_solo = algo_bytes / (decode_solo_ms * 1e-3) / 1e9

    # ---- e2e through the host-buffer C-ABI call: pinned host frames in, host cloud out
    e2e = None
    if not args.no_e2e:
        eb = min(12, RING)                                   # distinct host scans (pinned), streamed round-robin
        host_frames = torch.empty((eb, N_FRAMES, H, W), dtype=torch.uint8).pin_memory()
        host_white = torch.empty((eb, H, W, 3), dtype=torch.uint8).pin_memory()
        for i in range(eb):
            host_frames[i].copy_(ring_frames[i]); host_white[i].copy_(ring_white[i])
        pipe = slb200.Pipeline(N_FRAMES, H, W, max_batch=eb, device=local)
        cloud_h = torch.empty((eb * H * W, 4), dtype=torch.float32).pin_memory()
        cnt_h = torch.empty(eb + 1, dtype=torch.int64)
        calls = max(2, min(5, (K + eb - 1) // eb))           # 24..60 scans streamed
        # streaming form of the host API: one scan per call, copies and kernels of consecutive scans overlap on the
        # pipeline's streams, flush() returns when every cloud is in host memory
        cnts = [torch.zeros(1, dtype=torch.int64) for _ in range(eb)]
        clouds = [cloud_h[i * H * W:(i + 1) * H * W] for i in range(eb)]

        def stream_scans(n):
            for j in range(n):
                i = j % eb
                pipe.submit(host_frames[i], host_white[i], calib, clouds[i], cnts[i])
            pipe.flush()

        stream_scans(eb)                                                                  # warm-up
        barrier()
        t0 = time.perf_counter()
        stream_scans(calls * eb)
        torch.cuda.synchronize()
        el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        cnt_h[eb] = sum(int(c) for c in cnts)
        n_pts = int(cnt_h[eb])
        e2e = {"value": world * calls * eb * MPX / float(el.item()), "unit": "Mpx/s",
               "h2d_bytes_per_step": N_FRAMES * H * W + H * W * 3,
               "d2h_bytes_per_step": int(n_pts / eb * 16 + 16),
               "scans_per_call": 1, "distinct_host_scans": eb, "steps": calls * eb, "ms_per_scan": float(el.item()) / (calls * eb) * 1e3,
               "api": "sl_pipeline_submit per scan + sl_pipeline_flush (slb200.Pipeline.submit / flush), pinned host buffers"}
        pipe.close()

    # ---- gather of the compacted clouds (the only collective on the path), timed separately
    gather = None
    if world > 1:
        from slb200.dist import gather_clouds
        cloud = bufs.cloud[:valid_pts]
        gather_clouds(cloud, bufs.counts)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        reps = 5
        for _ in range(reps):
            out = gather_clouds(cloud, bufs.counts)
        g1.record()
        torch.cuda.synchronize()
        gms = torch.tensor([g0.elapsed_time(g1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(gms, op=dist.ReduceOp.MAX)
        gather = {"ms": float(gms.item()), "points_per_rank": valid_pts, "bytes_per_rank": valid_pts * 16,
                  "what": "all-gather-v (NCCL) of one scan's compacted XYZRGB cloud (16 B records) per rank"}
    clocks = sampler.stop()

    # ---- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_single_core(ring_frames[0].cpu().numpy(), ring_white[0].cpu().numpy(), calib)

    if rank == 0:
        line = {
            "metric": METRIC, "value": world * K * MPX / (ms_total * 1e-3), "unit": "Mpx/s",
            "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_total / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_scan": N_FRAMES, "n_bits": N_BITS,
                       "scans_per_step_per_gpu": 1, "valid_fraction": valid_pts / (W * H),
                       "triangulation_precision": "fp32",
                       "lens": ("exact fp64 ray tables built once per calibration (sl_lens_tables)" if tables is not None
                                else "5 Brown-Conrady iterations per valid pixel in the emit kernel"),
                       "parallelism": f"scan-sharded x{world}",
                       "l2": f"inputs rotate over a ring of {RING} resident scans ({RING * N_FRAMES * W * H / 1e6:.0f} MB > 126 MB L2)",
                       "launch": f"CUDA graph per scan (decode, strip-prefix + column scan, emit: 3 kernels); independent scans pipelined over {S} CUDA streams",
                       "decode_grid": f"one persistent CTA per SM, {sm_reserve} SMs left to the other streams' kernels"},
            "scans_per_s": world * K / (ms_total * 1e-3),
            "roofline": {"bound": "hbm", "kernel": "decode_tma_kernel<11>", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic_per_launch(),
                         "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/r01_launches.csv",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": algo_bytes, "us_per_launch": decode_ms * 1e3,
                         "frac_of_nominal_8000": achieved / 8000.0,
                         "timing": f"{kr} decode-only launches over the ring, CUDA events, pipelined over {S} streams like the step",
                         "isolated": {"us_per_launch": decode_solo_ms * 1e3, "achieved": achieved_solo,
                                      "frac": achieved_solo / peak,
                                      "timing": "the same launches strictly one after another on one stream (one CUDA graph of RING launches)"}},
            "cpu_baseline": cpu, "e2e": e2e, "gather": gather, "clocks": clocks,
            "gpu_launches": int(K * kernels_per_step),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lens", default="tables", choices=["tables", "iterate"])
    ap.add_argument("--streams", type=int, default=STREAMS)
    ap.add_argument("--sm-reserve", type=int, default=None,
                    help="SMs the decode grid leaves to the other streams' kernels (default: 8 when pipelining, else 0)")
    ap.add_argument("--ring", type=int, default=RING)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
