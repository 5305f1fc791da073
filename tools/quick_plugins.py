"""Device-resident timing of the inpaint and watershed bodies at BASELINE.json's 4K configs (development helper).
usage: quick_plugins.py [W H]"""
import importlib, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
p = importlib.import_module("openfx-opencv_b200"); s = importlib.import_module("openfx-opencv_b200.synth")
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
ctx = p.Context(0)
img = s.texture(H, W, seed=4)


def timeit(fn, n=3):
    fn(); ctx.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    ctx.synchronize()
    return (time.perf_counter() - t) / n


# inpaint: 10 % iid mask, radius 3 (C4), both methods
mask = s.iid_mask(H, W, 1000, 0.10)
d_img, d_mask, d_out = ctx.to_device(img), ctx.to_device(mask), ctx.alloc(W * H * 3)
for name, method in (("NS", p.INPAINT_NS), ("Telea", p.INPAINT_TELEA)):
    dt = timeit(lambda: ctx.inpaint_dev(d_img.ptr, 3, d_mask.ptr, d_out.ptr, W, H, 3.0, method))
    print("inpaint %s %dx%d 10%%: %.2f ms/frame %.1f fps  %s" % (name, W, H, dt * 1e3, 1 / dt, ctx.inpaint_stats()))
    ctx.prof(True)
    ctx.inpaint_dev(d_img.ptr, 3, d_mask.ptr, d_out.ptr, W, H, 3.0, method); ctx.synchronize()
    for r in ctx.prof_report():
        print("    ", r)
    ctx.prof(False)

# watershed: 256 seeds (C3), 1 frame and batches
mk = s.seed_markers(H, W, 256, 5)
d_rgb = ctx.to_device(img)
for nf in (1, 8, 32, 128):
    d_rgbs = ctx.alloc(W * H * 3 * nf)
    d_mks = ctx.alloc(W * H * 4 * nf)
    L = p.lib()
    for f in range(nf):
        L.ofxcv_upload(ctx.h, None, d_rgbs.ptr + f * W * H * 3, img.ctypes.data, W * H * 3)
    ctx.synchronize()

    def run():
        for f in range(nf):
            L.ofxcv_upload(ctx.h, None, d_mks.ptr + f * W * H * 4, mk.ctypes.data, W * H * 4)
        ctx.watershed_dev(d_rgbs.ptr, d_mks.ptr, W, H, nf)
    run(); ctx.synchronize()
    t = time.perf_counter(); run(); ctx.synchronize(); dt = time.perf_counter() - t
    # subtract nothing: uploads of the markers are part of resetting the in/out map
    print("watershed %dx%d 256 seeds, %d frames in flight: %.1f ms total, %.2f fps  %s" % (W, H, nf, dt * 1e3, nf / dt, ctx.watershed_stats()))
    d_rgbs.free(); d_mks.free()
