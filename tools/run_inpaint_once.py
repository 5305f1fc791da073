"""One device-resident inpaint call at a given size (profiling target). usage: run_inpaint_once.py W H method(0=NS,1=Telea) [frac]"""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
p = importlib.import_module("openfx-opencv_b200"); s = importlib.import_module("openfx-opencv_b200.synth")
W, H, method = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
frac = float(sys.argv[4]) if len(sys.argv) > 4 else 0.10
ctx = p.Context(0)
img = s.texture(H, W, seed=4); mask = s.iid_mask(H, W, 1000, frac)
d_img, d_mask, d_out = ctx.to_device(img), ctx.to_device(mask), ctx.alloc(W * H * 3)
for _ in range(2):
    ctx.inpaint_dev(d_img.ptr, 3, d_mask.ptr, d_out.ptr, W, H, 3.0, method)
ctx.synchronize()
print(ctx.inpaint_stats())
