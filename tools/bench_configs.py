"""All BASELINE.json configs on one GPU (informational; bench.py is the contract benchmark).
   C1 720p/1024x768 scan, C2 1080p scan, C3 4K scan (one GPU holds the whole capture; the row-band split is
   exercised by tests/test_multigpu_nccl.py), C4 64-view 1080p batch in one call, C5 720p streaming from pinned
   host memory (throughput and per-scan latency).  Prints one JSON object."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3dscanner-graycode_b200")); sys.path.insert(0, ROOT)
import torch, slb200
from slb200 import synth
from slb200.api import CloudBuffers

dev = torch.device("cuda", 0)
PEAK = 6535.4
out = {}


def scan_set(cam, proj, n, seed0=0):
    calib = synth.synthetic_calibration(cam, proj)
    pats = synth.pattern_stack(proj, dev)
    scans = [synth.render_scan(cam, proj, seed=seed0 + i, device=dev, patterns=pats, calib=calib) for i in range(n)]
    return calib, scans


def time_graphs(fns, iters, streams=1):
    sts = [torch.cuda.Stream(dev) for _ in range(streams)]
    graphs = []
    for i, fn in enumerate(fns):
        with torch.cuda.stream(sts[i % streams]):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=sts[i % streams]):
                fn()
            graphs.append(g)
    torch.cuda.synchronize()
    main = torch.cuda.current_stream()

    def run(n):
        for i in range(n):
            k = i % len(graphs)
            with torch.cuda.stream(sts[k % streams]):
                graphs[k].replay()
    run(2 * len(graphs)); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for st in sts: st.wait_event(e0)
    run(iters)
    for st in sts:
        ev = torch.cuda.Event(); ev.record(st); main.wait_event(ev)
    e1.record(main); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3      # us per call


def single_scan(name, cam, proj, ring, iters, streams, sm_reserve=8):
    calib, scans = scan_set(cam, proj, ring)
    tables = slb200.LensTables(calib, cam, dev)
    cs = calib.struct()
    bufs = [CloudBuffers.allocate(1, cam[1], cam[0], dev) for _ in range(streams)]
    for b in bufs: slb200.scan_to_cloud(scans[0][0], scans[0][1], calib, buffers=b, tables=tables, calib_struct=cs)
    torch.cuda.synchronize()
    res = {}
    for s_ in sorted({1, streams}):
        rsv = sm_reserve if s_ > 1 else 0     # pipelined: leave SMs to the other streams' kernels
        fns = [(lambda i=i: slb200.scan_to_cloud(scans[i][0], scans[i][1], calib, buffers=bufs[i % s_], tables=tables,
                                                 calib_struct=cs, sm_reserve=rsv)) for i in range(ring)]
        us = time_graphs(fns, iters, s_)
        dec = [(lambda i=i: slb200.decode_pixels(scans[i][0], out=(bufs[i % s_].h_idx[0], bufs[i % s_].v_idx[0]),
                                                 sm_reserve=rsv)) for i in range(ring)]
        dus = time_graphs(dec, iters, s_)
        F = scans[0][0].shape[0]
        mpx = cam[0] * cam[1] / 1e6
        gbs = (F + 4) * cam[0] * cam[1] / (dus * 1e-6) / 1e9
        res[f"streams_{s_}"] = {"us_per_scan": us, "Mpx_per_s": mpx / (us * 1e-6), "scans_per_s": 1e6 / us,
                                "decode_us": dus, "decode_GBps": gbs, "decode_frac_of_measured_peak": gbs / PEAK}
    res["valid_fraction"] = float(bufs[0].counts[1]) / (cam[0] * cam[1])
    out[name] = res
    del scans, bufs
    torch.cuda.empty_cache()


single_scan("C1_720p_1024x768", (1280, 720), (1024, 768), 24, 600, 4)
single_scan("C2_1080p", (1920, 1080), (1920, 1080), 12, 600, 4)
single_scan("C3_4k_single_gpu", (3840, 2160), (3840, 2160), 3, 90, 3, sm_reserve=0)   # measured: 134 us per scan with 0, 150 us with 8

# ---- C4: 64-view batch of 1080p scans in ONE call (6.1 GB of frames resident)
cam, proj = (1920, 1080), (1920, 1080)
calib, scans = scan_set(cam, proj, 8)
B = 64
frames = torch.empty((B, scans[0][0].shape[0], cam[1], cam[0]), dtype=torch.uint8, device=dev)
white = torch.empty((B, cam[1], cam[0], 3), dtype=torch.uint8, device=dev)
for b in range(B):
    frames[b].copy_(scans[b % 8][0]); white[b].copy_(scans[b % 8][1])
tables = slb200.LensTables(calib, cam, dev)
cs = calib.struct()
bufs = CloudBuffers.allocate(B, cam[1], cam[0], dev)
slb200.scan_to_cloud(frames, white, calib, buffers=bufs, tables=tables, calib_struct=cs)
torch.cuda.synchronize()
us = time_graphs([lambda: slb200.scan_to_cloud(frames, white, calib, buffers=bufs, tables=tables, calib_struct=cs)], 10)
dus = time_graphs([lambda: slb200.decode_pixels(frames, out=(bufs.h_idx, bufs.v_idx))], 10)
F = frames.shape[1]
gbs = B * (F + 4) * cam[0] * cam[1] / (dus * 1e-6) / 1e9
out["C4_64view_1080p_one_call"] = {"ms_per_batch": us / 1e3, "scans_per_s": B / (us * 1e-6), "Mpx_per_s": B * 2.0736 / (us * 1e-6),
                                   "decode_ms": dus / 1e3, "decode_GBps": gbs, "decode_frac_of_measured_peak": gbs / PEAK,
                                   "points": int(bufs.counts[-1])}
del frames, white, bufs
torch.cuda.empty_cache()

# ---- C5: 720p streaming from pinned host memory through sl_scan_to_cloud_host
cam, proj = (1280, 720), (1024, 768)
calib, scans = scan_set(cam, proj, 8)
F = scans[0][0].shape[0]
hf = torch.stack([s[0] for s in scans]).cpu().pin_memory()
hw = torch.stack([s[1] for s in scans]).cpu().pin_memory()
pipe = slb200.Pipeline(F, cam[1], cam[0], max_batch=8)
cloud = torch.empty((8 * cam[0] * cam[1], 4), dtype=torch.float32).pin_memory()
cnt = torch.empty(9, dtype=torch.int64)
pipe.scan_to_cloud(hf, hw, calib, cloud=cloud, counts=cnt)
t0 = time.perf_counter()
reps = 20
for _ in range(reps):
    pipe.scan_to_cloud(hf, hw, calib, cloud=cloud, counts=cnt)
el = time.perf_counter() - t0
lat = []
for _ in range(20):
    t1 = time.perf_counter(); pipe.scan_to_cloud(hf[:1], hw[:1], calib, cloud=cloud, counts=cnt[:2]); lat.append(time.perf_counter() - t1)
lat.sort()
out["C5_720p_streaming_host"] = {"scans_per_s": reps * 8 / el, "Mpx_per_s": reps * 8 * 0.9216 / el, "ms_per_scan_pipelined": el / (reps * 8) * 1e3,
                                 "single_scan_latency_ms_median": lat[len(lat) // 2] * 1e3,
                                 "h2d_bytes_per_scan": F * cam[0] * cam[1] + cam[0] * cam[1] * 3, "d2h_bytes_per_scan": int(cnt[8]) // 8 * 16}
pipe.close()
print(json.dumps(out))
