"""In-order (incremental) fill against the ready-queue fill on masks of different shapes (development measurement;
OFXCV_IP_READYQ=0/1 forces a scheduler, unset = the library's choice)."""
import importlib, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
p = importlib.import_module("openfx-opencv_b200"); s = importlib.import_module("openfx-opencv_b200.synth")
W, H = 3840, 2160
ctx = p.Context(0)
img = s.texture(H, W, seed=4)
d_img, d_out = ctx.to_device(img), ctx.alloc(W * H * 3)
tag = "READYQ=%s" % os.environ.get("OFXCV_IP_READYQ", "auto")
def run(mask, name):
    d_mask = ctx.to_device(mask)
    for method, mname in ((p.INPAINT_NS, "NS"), (p.INPAINT_TELEA, "Telea")):
        ctx.inpaint_dev(d_img.ptr, 3, d_mask.ptr, d_out.ptr, W, H, 3.0, method); ctx.synchronize()
        ctx.prof(True)
        ctx.inpaint_dev(d_img.ptr, 3, d_mask.ptr, d_out.ptr, W, H, 3.0, method); ctx.synchronize()
        rows = {r[0]: r[3] for r in ctx.prof_report()}
        ctx.prof(False)
        print("%-11s %-30s %-5s holes %7d  fill %8.3f ms" % (tag, name, mname, int((mask != 0).sum()), rows["ip_fill"]))
m = np.zeros((H, W), np.uint8); m[1000, 400:3400] = 255
run(m, "one 1-px line")
m = np.zeros((H, W), np.uint8); m[200:2000:40, 400:3400] = 255
run(m, "45 1-px lines")
m = np.zeros((H, W), np.uint8)
for k in range(45): m[200 + 40 * k:203 + 40 * k, 400:3400] = 255
run(m, "45 3-px lines")
run(s.blob_mask(H, W, 7, nblobs=400, rmax=20), "400 blobs r<=20")
run(s.blob_mask(H, W, 8, nblobs=40, rmax=80), "40 blobs r<=80")
rng = np.random.default_rng(5); m = np.zeros((H, W), np.uint8)
for k in range(60):   # diagonal scratches, 2 px wide
    x0, y0 = int(rng.integers(0, W - 700)), int(rng.integers(0, H - 700)); L = int(rng.integers(200, 700)); sl = rng.uniform(-1, 1)
    for t in range(L):
        yy = int(y0 + 350 + sl * t * 0.5); m[yy:yy + 2, x0 + t] = 255
run(m, "60 diagonal scratches")
run(s.iid_mask(H, W, 1000, 0.10), "iid 10 %")
