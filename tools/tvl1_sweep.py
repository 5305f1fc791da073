"""tv_iter launch time at W x H for the current OFXCV_TV_* environment (development helper)."""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
p = importlib.import_module("openfx-opencv_b200"); s = importlib.import_module("openfx-opencv_b200.synth")
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
ctx = p.Context(0)
base = s.gray(s.texture(H, W, seed=2000)); nxt = s.shift_bilinear(base, 2.5, -1.5)
a, b, f = ctx.to_device(base), ctx.to_device(nxt), ctx.alloc(W * H * 8)
par = p.Tvl1Params(epsilon=0.0, nscales=1, warps=1, outer_iterations=2, iterations=15)
ctx.tvl1_dev(a.ptr, b.ptr, W, H, f.ptr, par); ctx.synchronize()
ctx.timing(True)
for _ in range(3):
    ctx.tvl1_dev(a.ptr, b.ptr, W, H, f.ptr, par)
ctx.synchronize()
n, ms = ctx.kernel_time_ms(1)
us = ms * 1e3 / n
print("%dx%d RB=%s MINB=%s: %.1f us/launch, %.0f GB/s" % (W, H, os.environ.get("OFXCV_TV_RB", "auto"), os.environ.get("OFXCV_TV_MINB", "4"), us, p.lib().ofxcv_tvl1_iter_bytes(W, H) / us / 1e3))
