"""Device-resident timing of Dual TV-L1 (development helper). usage: tvl1_rate.py [W H]"""
import importlib, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
p = importlib.import_module("openfx-opencv_b200"); s = importlib.import_module("openfx-opencv_b200.synth")
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
ctx = p.Context(0)
base = s.gray(s.texture(H, W, seed=3)); nxt = s.shift_bilinear(base, 2.5, -1.5)
a, b, f = ctx.to_device(base), ctx.to_device(nxt), ctx.alloc(W * H * 8)
L = p.lib()
for name, par in (("defaults", p.Tvl1Params()), ("no early exit (epsilon 0)", p.Tvl1Params(epsilon=0.0))):
    ctx.tvl1_dev(a.ptr, b.ptr, W, H, f.ptr, par); ctx.synchronize()
    t = time.perf_counter()
    ctx.tvl1_dev(a.ptr, b.ptr, W, H, f.ptr, par); ctx.synchronize()
    dt = time.perf_counter() - t
    it = L.ofxcv_tvl1_iterations_run(ctx.h)
    print("%dx%d %s: %.1f ms/pair, %d inner iterations run" % (W, H, name, dt * 1e3, it))
    ctx.timing(True)
    ctx.tvl1_dev(a.ptr, b.ptr, W, H, f.ptr, par); ctx.synchronize()
    n, ms = ctx.kernel_time_ms(1)
    ctx.timing(False)
    if n:
        us = ms * 1e3 / n
        print("   full-res inner iteration (2 launches, incl. skipped ones): %d timed, %.1f us avg -> %.0f GB/s algorithmic" % (n, us, L.ofxcv_tvl1_iter_bytes(W, H) / us / 1e3))
    ctx.prof(True)
    ctx.tvl1_dev(a.ptr, b.ptr, W, H, f.ptr, par); ctx.synchronize()
    for r in sorted(ctx.prof_report(), key=lambda r: (r[1], r[0])):
        print("    scale %d %-12s n=%4d %9.3f ms" % (r[1], r[0], r[2], r[3]))
    ctx.prof(False)
