"""Fill-kernel time and whole-frame time of the inpaint body on the 4K iid 10 % mask (development measurement; the fill
variant is chosen by OFXCV_IP_FILL_INC / OFXCV_IP_FILL_WAITALL in the environment)."""
import importlib, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
p = importlib.import_module("openfx-opencv_b200"); s = importlib.import_module("openfx-opencv_b200.synth")
W, H = 3840, 2160
ctx = p.Context(0)
img = s.texture(H, W, seed=4)
mask = s.iid_mask(H, W, 1000, 0.10)
d_img, d_out, d_mask = ctx.to_device(img), ctx.alloc(W * H * 3), ctx.to_device(mask)
tag = "INC=%s WAITALL=%s" % (os.environ.get("OFXCV_IP_FILL_INC", "-"), os.environ.get("OFXCV_IP_FILL_WAITALL", "-"))
for method, mname in ((p.INPAINT_NS, "NS"), (p.INPAINT_TELEA, "Telea")):
    for _ in range(2):
        ctx.inpaint_dev(d_img.ptr, 3, d_mask.ptr, d_out.ptr, W, H, 3.0, method); ctx.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); ctx.inpaint_dev(d_img.ptr, 3, d_mask.ptr, d_out.ptr, W, H, 3.0, method); ctx.synchronize(); ts.append(time.perf_counter() - t0)
    ctx.prof(True)
    ctx.inpaint_dev(d_img.ptr, 3, d_mask.ptr, d_out.ptr, W, H, 3.0, method); ctx.synchronize()
    rows = {r[0]: r[3] for r in ctx.prof_report()}
    ctx.prof(False)
    print("%-22s %-5s fill %.3f ms, frame %.3f ms (min of 5; %.1f fps)" % (tag, mname, rows["ip_fill"], min(ts) * 1e3, 1.0 / min(ts)))
