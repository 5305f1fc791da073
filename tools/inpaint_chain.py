"""Per-step latency and dependency-free throughput of the inpaint fill kernel (development measurement)."""
import importlib, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
p = importlib.import_module("openfx-opencv_b200"); s = importlib.import_module("openfx-opencv_b200.synth")
W, H = 3840, 2160
ctx = p.Context(0)
img = s.texture(H, W, seed=4)
d_img, d_out = ctx.to_device(img), ctx.alloc(W * H * 3)
def run(mask, name):
    d_mask = ctx.to_device(mask)
    for method, mname in ((p.INPAINT_NS, "NS"), (p.INPAINT_TELEA, "Telea")):
        ctx.inpaint_dev(d_img.ptr, 3, d_mask.ptr, d_out.ptr, W, H, 3.0, method); ctx.synchronize()
        ctx.prof(True)
        ctx.inpaint_dev(d_img.ptr, 3, d_mask.ptr, d_out.ptr, W, H, 3.0, method); ctx.synchronize()
        rows = {r[0]: r[3] for r in ctx.prof_report()}
        ctx.prof(False)
        n = int((mask != 0).sum())
        print("%-28s %-5s holes %7d  fill %.3f ms  -> %.3f us per hole pixel" % (name, mname, n, rows["ip_fill"], rows["ip_fill"] * 1e3 / n))
m = np.zeros((H, W), np.uint8); m[1000, 400:3400] = 255
run(m, "one horizontal line (chain)")
m = np.zeros((H, W), np.uint8); m[10:-10:12, 10:-10:12] = 255
run(m, "isolated grid (no deps)")
m = np.zeros((H, W), np.uint8); m[200:2000:40, 400:3400] = 255
run(m, "45 lines (45 chains)")
run(s.iid_mask(H, W, 1000, 0.10), "iid 10 %")
