"""How fast does one emit launch run while the other stream keeps the decode kernel busy back to back, and how much does
that slow the decodes?  (events around the emit on stream B; a chain of 12 decodes on stream A)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3dscanner-graycode_b200")); sys.path.insert(0, ROOT)
import torch, slb200
from slb200 import synth
from slb200.api import _ptr, _stream, check, lib, CloudBuffers
cam, proj = (1920, 1080), (1920, 1080); dev = torch.device("cuda", 0)
calib = synth.synthetic_calibration(cam, proj); pats = synth.pattern_stack(proj, dev)
tables = slb200.LensTables(calib, cam, dev, precision="fp32")
frames, sets = [], []
for i in range(6):
    fr, wh = synth.render_scan(cam, proj, seed=i, device=dev, patterns=pats, calib=calib)
    b = CloudBuffers.allocate(1, cam[1], cam[0], dev)
    slb200.scan_to_cloud(fr, wh, calib, buffers=b, tables=tables)
    frames.append(fr); sets.append((b, wh))
dec_out = CloudBuffers.allocate(1, cam[1], cam[0], dev)
torch.cuda.synchronize()
s = calib.struct(); L = lib(); H, W = cam[1], cam[0]
sa, sb = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def emit_only(b, wh):
    check(L.sl_emit_cloud(_ptr(b.h_idx), _ptr(b.v_idx), _ptr(wh), C.byref(s), 0, _ptr(tables.cam), _ptr(tables.proj_bt), 1, H, W, _ptr(b.ws), _ptr(b.cloud), _stream()))
for rsv in (0, 8):
    with torch.cuda.stream(sa):
        gd = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gd, stream=sa):
            for i in range(12):
                slb200.decode_pixels(frames[i % 6], out=(dec_out.h_idx[0], dec_out.v_idx[0]), sm_reserve=rsv)
    torch.cuda.synchronize()
    # decode chain alone
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(sa):
        gd.replay(); e0.record(sa); gd.replay(); e1.record(sa)
    torch.cuda.synchronize()
    alone = e0.elapsed_time(e1) * 1e3 / 12
    # emit alone
    with torch.cuda.stream(sb):
        emit_only(*sets[0]); f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(sb); emit_only(*sets[1]); f1.record(sb)
    torch.cuda.synchronize()
    emit_alone = f0.elapsed_time(f1) * 1e3
    # emit launched in the middle of a running decode chain
    res = []
    for trial in range(5):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        mid = torch.cuda.Event()
        with torch.cuda.stream(sa):
            a0.record(sa); gd.replay(); a1.record(sa)
        with torch.cuda.stream(sb):
            sb.wait_event(a0)                      # start together with the chain
            f0.record(sb); emit_only(*sets[trial % 6]); f1.record(sb)
        torch.cuda.synchronize()
        res.append((f0.elapsed_time(f1) * 1e3, a0.elapsed_time(a1) * 1e3 / 12))
    print(f"reserve={rsv}: decode alone {alone:.2f} us each; emit alone {emit_alone:.2f} us; emit beside the decode chain "
          f"{sorted(r[0] for r in res)[2]:.2f} us (median), decodes then {sorted(r[1] for r in res)[2]:.2f} us each")
