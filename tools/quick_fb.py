"""Quick device-resident timing of the Farneback path (development helper, not the bench)."""
import importlib, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
p = importlib.import_module("openfx-opencv_b200"); s = importlib.import_module("openfx-opencv_b200.synth")
W, H = int(sys.argv[1]), int(sys.argv[2]); levels = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = p.Context(0)
prev, nxt = s.flow_pair(H, W)
dp, dn, df = ctx.to_device(prev), ctx.to_device(nxt), ctx.alloc(W * H * 8)
par = p.FbParams(levels=levels)
for _ in range(3): ctx.farneback_dev(dp.ptr, dn.ptr, W, H, df.ptr, par)
ctx.synchronize()
ctx.timing(True)
t = time.perf_counter(); n = 10
for _ in range(n): ctx.farneback_dev(dp.ptr, dn.ptr, W, H, df.ptr, par)
ctx.synchronize(); dt = (time.perf_counter() - t) / n
nl, ms = ctx.kernel_time_ms(0)
ab = p.farneback_algorithmic_bytes(W, H, par); ib = p.farneback_iter_bytes(W, H, par)
print("%dx%d L%d: %.3f ms/pair  %.1f pair/s  eff %.0f GB/s (%.3f of 6530)  iter-kernels %.3f ms/pair -> %.0f GB/s" % (W, H, levels, dt * 1e3, 1 / dt, ab / dt / 1e9, ab / dt / 1e9 / 6530.6, ms / n, ib / (ms / n * 1e-3) / 1e9))
