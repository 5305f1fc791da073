"""One call of every body at 4K (ncu target for the per-kernel DRAM table). usage: run_all_once.py [W H]"""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
p = importlib.import_module("openfx-opencv_b200"); s = importlib.import_module("openfx-opencv_b200.synth")
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
ctx = p.Context(0)
base = s.gray(s.texture(H, W, seed=2000)); nxt = s.shift_bilinear(base, 2.5, -1.5)
d0, d1, df = ctx.to_device(base), ctx.to_device(nxt), ctx.alloc(W * H * 8)
ctx.farneback_dev(d0.ptr, d1.ptr, W, H, df.ptr); ctx.synchronize()
ctx.tvl1_dev(d0.ptr, d1.ptr, W, H, df.ptr, p.Tvl1Params(epsilon=0.0, nscales=2, warps=1, outer_iterations=1, iterations=3)); ctx.synchronize()
ctx.farneback_dev(d0.ptr, d1.ptr, W, H, df.ptr); ctx.synchronize()
img = s.texture(H, W, seed=4); mask = s.iid_mask(H, W, 1000, 0.10)
di, dm, do = ctx.to_device(img), ctx.to_device(mask), ctx.alloc(W * H * 3)
for m in (p.INPAINT_NS, p.INPAINT_TELEA):
    ctx.inpaint_dev(di.ptr, 3, dm.ptr, do.ptr, W, H, 3.0, m); ctx.synchronize()
mk = s.seed_markers(H, W, 256, 5); dk = ctx.to_device(mk)
ctx.watershed_dev(di.ptr, dk.ptr, W, H, 1); ctx.synchronize()
rgba = np.zeros((H, W, 4), np.float32); rgba[..., :3] = (img / 255.0) ** 2.2; rgba[..., 3] = 1
L = p.lib()
dr, dg, d8, dF = ctx.to_device(rgba), ctx.alloc(W * H), ctx.alloc(W * H * 4), ctx.alloc(W * H * 16)
L.ofxcv_rgba32f_to_srgb_gray8(ctx.h, None, dr.ptr, W * 16, 4, dg.ptr, W, W, H)
L.ofxcv_rgba32f_to_srgb8_packed(ctx.h, None, dr.ptr, W * 16, 4, d8.ptr, W * 4, 4, W, H)
L.ofxcv_srgb8_packed_to_rgba32f(ctx.h, None, d8.ptr, W * 4, dF.ptr, W * 16, 4, W, H)
import ctypes as C
sel = (C.c_int * 4)(0, 1, -1, -1)
L.ofxcv_flow_to_rgba32f(ctx.h, None, df.ptr, W * 8, dF.ptr, W * 16, W, H, sel, 1.0, 1.0)
drgb, dmask = ctx.alloc(W * H * 3), ctx.alloc(W * H)
L.ofxcv_rgba8_to_rgb8_mask(ctx.h, None, d8.ptr, W * 4, drgb.ptr, W * 3, dmask.ptr, W, W, H, 1)
L.ofxcv_rgb8_to_rgba8(ctx.h, None, drgb.ptr, W * 3, d8.ptr, W * 4, W, H)
ctx.synchronize()
print("done")
