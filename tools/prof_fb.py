"""Per-kernel / per-scale device times of the Farneback path from the library's own event profile
(development helper; the numbers bench.py reports come from bench.py).  usage: prof_fb.py W H [levels] [pairs]"""
import importlib, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
p = importlib.import_module("openfx-opencv_b200"); s = importlib.import_module("openfx-opencv_b200.synth")
W, H = int(sys.argv[1]), int(sys.argv[2]); levels = int(sys.argv[3]) if len(sys.argv) > 3 else 3
npairs = int(sys.argv[4]) if len(sys.argv) > 4 else 4
ctx = p.Context(0)
base = s.gray(s.texture(H, W, seed=2000))
frames = [ctx.to_device(s.shift_bilinear(base, 2.5 * f, -1.5 * f)) for f in range(npairs + 1)]
flows = [ctx.alloc(W * H * 8) for _ in range(npairs)]
par = p.FbParams(levels=levels)


def run():
    for i in range(npairs):
        ctx.farneback_dev(frames[i].ptr, frames[i + 1].ptr, W, H, flows[i].ptr, par)


for _ in range(2):
    run()
ctx.synchronize()
t = time.perf_counter()
for _ in range(3):
    run()
ctx.synchronize()
dt = (time.perf_counter() - t) / (3 * npairs)
ab = p.farneback_algorithmic_bytes(W, H, par)
print("%dx%d L%d: %.3f ms/pair %.1f pair/s whole-pair %.0f GB/s" % (W, H, levels, dt * 1e3, 1 / dt, ab / dt / 1e9))
ctx.prof(True)
run()
ctx.synchronize()
rows = ctx.prof_report()
ctx.prof(False)
tot = sum(r[3] for r in rows)
print("profiled (events around every launch): %.3f ms/pair" % (tot / npairs))
for name, tag, n, ms in sorted(rows, key=lambda r: (-r[1], r[0])):
    print("  scale %d %-12s launches/pair %4d  %8.1f us/pair  %7.2f us/launch  %5.1f%%" % (tag, name, n // npairs, ms * 1e3 / npairs, ms * 1e3 / n, 100 * ms / tot))
