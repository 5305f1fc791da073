"""Row-transfer pipeline of the C ABI (ofxcv_upload_rows / ofxcv_download_rows) on a 4K float RGBA frame in pageable memory."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("openfx-opencv_b200")
W, H = 3840, 2160
ctx = pkg.Context(0)
a = np.random.default_rng(0).random((H, W, 4), dtype=np.float32)
b = np.zeros_like(a)
d = ctx.alloc(a.nbytes)
L = pkg.lib()
for rep in range(4):
    t0 = time.perf_counter(); L.ofxcv_upload_rows(ctx.h, None, d.ptr, a.ctypes.data, W * 16, W * 16, H); ctx.synchronize(); t1 = time.perf_counter()
    L.ofxcv_download_rows(ctx.h, None, b.ctypes.data, W * 16, d.ptr, W * 16, H); t2 = time.perf_counter()
    print("threads %s: upload %.2f ms (%.1f GB/s)  download %.2f ms (%.1f GB/s)  equal %s" % (os.environ.get("OFXCV_XFER_THREADS", "auto"), (t1 - t0) * 1e3, a.nbytes / (t1 - t0) / 1e9, (t2 - t1) * 1e3, a.nbytes / (t2 - t1) / 1e9, bool((a == b).all())))
t0 = time.perf_counter(); c = a.copy(); t1 = time.perf_counter()
print("numpy copy of the frame (1 thread): %.2f ms (%.1f GB/s), cores %d" % ((t1 - t0) * 1e3, a.nbytes / (t1 - t0) / 1e9, os.cpu_count()))
