"""Clip throughput of ofxcv_farneback_sequence_u8 (frames resident in HBM). usage: seq_rate.py W H levels pairs [lanes]"""
import importlib, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
p = importlib.import_module("openfx-opencv_b200"); s = importlib.import_module("openfx-opencv_b200.synth")
W, H, levels, npairs = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
lanes = int(sys.argv[5]) if len(sys.argv) > 5 else 2
ctx = p.Context(0)
ctx.farneback_set_lanes(lanes)
base = s.gray(s.texture(H, W, seed=2000))
frames = np.stack([s.shift_bilinear(base, 2.5 * f, -1.5 * f) for f in range(npairs + 1)])
d_frames = ctx.to_device(frames); d_flows = ctx.alloc(W * H * 8 * npairs)
par = p.FbParams(levels=levels)
for _ in range(2):
    ctx.farneback_sequence_dev(d_frames.ptr, W, H, npairs + 1, d_flows.ptr, par)
ctx.synchronize()
t = time.perf_counter(); n = 4
for _ in range(n):
    ctx.farneback_sequence_dev(d_frames.ptr, W, H, npairs + 1, d_flows.ptr, par)
ctx.synchronize(); dt = (time.perf_counter() - t) / (n * npairs)
print("%dx%d levels %d lanes %d: %.3f ms/pair  %.1f pairs/s" % (W, H, levels, lanes, dt * 1e3, 1 / dt))
