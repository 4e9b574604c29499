"""A/B of the two emissions (sl_emit_cloud tiles vs sl_emit_cloud_rows) on decoded 1080p scans: equality bit for bit and
time per scan, isolated chain (one scan per launch) and batch (16 scans per launch)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3dscanner-graycode_b200"))
import ctypes as C  # noqa: E402
import torch  # noqa: E402
import slb200  # noqa: E402
from slb200 import synth  # noqa: E402
from slb200.api import CloudBuffers, _ptr, _stream, check, lib  # noqa: E402

cam, proj = (1920, 1080), (1920, 1080)
W, H = cam
dev = torch.device("cuda", 0)
calib = synth.synthetic_calibration(cam, proj)
pats = synth.pattern_stack(proj, dev)
tables = slb200.LensTables(calib, cam, dev)
cam_t, proj_t = tables.transposed()
s = calib.struct()
L = lib()


def run(B, reps):
    scans = [synth.render_scan(cam, proj, seed=i, device=dev, patterns=pats, calib=calib) for i in range(min(B, 4))]
    frames = torch.stack([scans[i % len(scans)][0] for i in range(B)]).contiguous()
    white = torch.stack([scans[i % len(scans)][1] for i in range(B)]).contiguous()
    bufs = CloudBuffers.allocate(B, H, W, dev)
    slb200.scan_to_cloud(frames, white, calib, buffers=bufs, tables=tables, emit="tiles")
    torch.cuda.synchronize()
    n = int(bufs.counts[-1])
    ref = bufs.cloud[:n].clone()
    out2 = torch.zeros_like(bufs.cloud)

    def tiles():
        check(L.sl_emit_cloud(_ptr(bufs.h_idx), _ptr(bufs.v_idx), _ptr(white), C.byref(s), 0, _ptr(tables.cam), _ptr(tables.proj_bt),
                              B, H, W, _ptr(bufs.ws), _ptr(bufs.cloud), _stream()))

    def rows():
        check(L.sl_emit_cloud_rows(_ptr(bufs.h_idx), _ptr(bufs.v_idx), _ptr(white), C.byref(s), _ptr(cam_t), H, _ptr(proj_t),
                                   B, H, W, _ptr(bufs.ws), _ptr(out2), _stream()))
    rows(); torch.cuda.synchronize()
    same = torch.equal(out2[:n].view(torch.int32), ref.view(torch.int32))
    res = {}
    for name, fn in (("tiles", tiles), ("rows", rows)):
        g = torch.cuda.CUDAGraph()
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            fn(); fn()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=st):
                for _ in range(6):
                    fn()
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(reps):
                g.replay()
            e1.record(st)
            torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / (reps * 6) * 1e3 / B
    print(f"B={B}: equal={same} n={n}  us per scan: tiles {res['tiles']:.2f}  rows {res['rows']:.2f}")


run(1, 50)
run(16, 5)
