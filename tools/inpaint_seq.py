"""Sequence throughput of the inpaint body: K frames in flight (one context + host thread each), frames resident in HBM.
usage: inpaint_seq.py K blocks_per_sm [method] [frames_per_thread]"""
import importlib, sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
p = importlib.import_module("openfx-opencv_b200"); s = importlib.import_module("openfx-opencv_b200.synth")
K, bps = int(sys.argv[1]), int(sys.argv[2])
method = int(sys.argv[3]) if len(sys.argv) > 3 else p.INPAINT_NS
nper = int(sys.argv[4]) if len(sys.argv) > 4 else 4
W, H = 3840, 2160
img = s.texture(H, W, seed=4)
ctxs = [p.Context(0) for _ in range(K)]
bufs = []
for k, c in enumerate(ctxs):
    c.inpaint_set_fill_blocks(bps)
    bufs.append((c.to_device(img), c.to_device(s.iid_mask(H, W, 1000 + k, 0.10)), c.alloc(W * H * 3)))
def work(k, n):
    c = ctxs[k]; a, m, o = bufs[k]
    for _ in range(n):
        c.inpaint_dev(a.ptr, 3, m.ptr, o.ptr, W, H, 3.0, method)
    c.synchronize()
def run(n):
    th = [threading.Thread(target=work, args=(k, n)) for k in range(K)]
    t = time.perf_counter()
    for x in th: x.start()
    for x in th: x.join()
    return time.perf_counter() - t
run(1)
dt = run(nper)
print("inpaint %s 4K 10%%: %d frames in flight, %d fill CTAs/SM each: %.1f frames/s (%.1f ms per frame per context)" % (
    "NS" if method == 0 else "Telea", K, bps, K * nper / dt, dt / nper * 1e3))
