"""One watershed frame through the device entry point. usage: ws_once.py W H seeds [image] [mode]"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("openfx-opencv_b200")
synth = importlib.import_module("openfx-opencv_b200.synth")
w, h, ns = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
kind = sys.argv[4] if len(sys.argv) > 4 else "texture"
if len(sys.argv) > 5:
    os.environ["OFXCV_WS_MODE"] = sys.argv[5]
rng = np.random.default_rng(7)
img = synth.texture(h, w, seed=4) if kind == "texture" else rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
mk = synth.seed_markers(h, w, ns, 5)
ctx = pkg.Context(0)
for rep in range(2):
    d_rgb, d_mk = ctx.to_device(img), ctx.to_device(mk)
    ctx.synchronize()
    t = time.perf_counter(); ctx.watershed_dev(d_rgb.ptr, d_mk.ptr, w, h, 1); ctx.synchronize(); dt = time.perf_counter() - t
    s = (pkg.C.c_int64 * 4)(); pkg.lib().ofxcv_watershed_last_stats(ctx.h, s)
    print("%dx%d %s: %.2f ms pops %d rounds %d passes %d" % (w, h, kind, dt * 1e3, s[0], s[2], s[3]), flush=True)
if os.environ.get("WS_CHECK"):
    import oracle
    ref, pops = oracle.watershed(img, mk)
    print("differing", int((d_mk.download((h, w), np.int32) != ref).sum()), "pops", pops)
