"""One short Dual TV-L1 run that cannot stop early (ncu target for tv_iter). usage: tvl1_once.py [W H]"""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
p = importlib.import_module("openfx-opencv_b200"); s = importlib.import_module("openfx-opencv_b200.synth")
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
ctx = p.Context(0)
base = s.gray(s.texture(H, W, seed=2000)); nxt = s.shift_bilinear(base, 2.5, -1.5)
a, b, f = ctx.to_device(base), ctx.to_device(nxt), ctx.alloc(W * H * 8)
ctx.tvl1_dev(a.ptr, b.ptr, W, H, f.ptr, p.Tvl1Params(epsilon=0.0, nscales=1, warps=1, outer_iterations=1, iterations=6)); ctx.synchronize()
print("done")
