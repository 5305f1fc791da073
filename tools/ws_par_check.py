"""Exact parallel watershed vs the CPU oracle on several image classes, with timings. usage: ws_par_check.py [big]"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("openfx-opencv_b200")
synth = importlib.import_module("openfx-opencv_b200.synth")
import oracle

def images(h, w, seed):
    rng = np.random.default_rng(seed)
    yield "texture", synth.texture(h, w, seed=4)
    yield "noise", rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    yield "flat", np.full((h, w, 3), 77, np.uint8)
    g = np.zeros((h, w, 3), np.uint8); g[..., 0] = (np.arange(w) % 256)[None, :]; g[..., 1] = (np.arange(h) % 256)[:, None]
    yield "gradient", g
    b = rng.integers(0, 256, (h // 16 + 1, w // 16 + 1, 3), dtype=np.uint8)
    yield "blocks", np.ascontiguousarray(np.kron(b, np.ones((16, 16, 1), np.uint8))[:h, :w])

def main():
    big = len(sys.argv) > 1
    ctx = pkg.Context(0)
    sizes = [(60, 80, 5), (270, 480, 40), (1080, 1920, 100)] + ([(2160, 3840, 256)] if big else [])
    bad = 0
    for h, w, ns in sizes:
        mk = synth.seed_markers(h, w, ns, 5)
        for name, img in images(h, w, 7):
            t = time.perf_counter(); ref, pops = oracle.watershed(img, mk); t_cpu = time.perf_counter() - t
            for mode in ("par", "seq"):
                if mode == "seq" and h > 1080: continue
                os.environ["OFXCV_WS_MODE"] = mode
                d_rgb, d_mk = ctx.to_device(img), ctx.to_device(mk)
                ctx.synchronize()
                t = time.perf_counter(); ctx.watershed_dev(d_rgb.ptr, d_mk.ptr, w, h, 1); ctx.synchronize(); dt = time.perf_counter() - t
                got = d_mk.download((h, w), np.int32)
                s = (pkg.C.c_int64 * 4)(); pkg.lib().ofxcv_watershed_last_stats(ctx.h, s)
                nd = int((got != ref).sum())
                bad += nd != 0 or (s[0] != pops)
                print("%4dx%-4d %-8s %s: %8.2f ms (cpu oracle %7.1f ms)  pops %d/%d rounds %d passes %d  differing %d" % (w, h, name, mode, dt * 1e3, t_cpu * 1e3, s[0], pops, s[2], s[3], nd), flush=True)
                d_rgb.free(); d_mk.free()
    print("FAILED" if bad else "ALL EXACT")
    return 1 if bad else 0

if __name__ == "__main__":
    sys.exit(main())
