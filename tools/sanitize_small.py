"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck).
usage: compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
p = importlib.import_module("openfx-opencv_b200"); s = importlib.import_module("openfx-opencv_b200.synth")
ctx = p.Context(0)
for (h, w, lv, it) in ((7, 29, 0, 2), (61, 62, 2, 3), (135, 241, 3, 4), (270, 480, 3, 3), (600, 1100, 2, 2)):
    a, b = s.flow_pair(h, w, seed=3)
    f = ctx.farneback(a, b, p.FbParams(levels=lv, iterations=it))
    assert np.isfinite(f).all()
frames = [s.shift_bilinear(s.gray(s.texture(90, 130, 5)), 1.0 * i, 0.5 * i) for i in range(4)]
ctx.farneback_sequence(frames, p.FbParams(levels=2, iterations=3))
img = s.texture(60, 80, 1)
for m in (p.INPAINT_NS, p.INPAINT_TELEA):
    ctx.inpaint(img, s.iid_mask(60, 80, 2, 0.1), 3, m)
    ctx.inpaint(img, s.blob_mask(60, 80, 3), 5, m)
    ctx.inpaint(img, s.iid_mask(60, 80, 4, 0.3), 4, m)                  # incremental fill, three tap rounds
    lines = np.zeros((120, 400), np.uint8); lines[10:110:9, 20:380] = 255  # chains laid out back to back: the ready-queue fill
    ctx.inpaint(s.texture(120, 400, 2), lines, 3, m)
ctx.watershed(img, s.seed_markers(60, 80, 5, 5))                       # one frame: the round-synchronous parallel flood
noise = np.random.default_rng(9).integers(0, 256, (70, 90, 3), dtype=np.uint8)
ctx.watershed(noise, s.seed_markers(70, 90, 6, 3))                    # levels above 32: the 256-level instantiation
os.environ["OFXCV_WS_MODE"] = "seq"
ctx.watershed(img, s.seed_markers(60, 80, 5, 5))                       # the one-thread flood (long clips)
del os.environ["OFXCV_WS_MODE"]
big = np.random.default_rng(3).random((300, 260, 4), dtype=np.float32)   # row-transfer pipeline (chunked) through the C ABI
d_big = ctx.alloc(big.nbytes); back = np.zeros_like(big)
ctx.upload_rows(d_big.ptr, big.ctypes.data, 260 * 16, 300, 260 * 16); ctx.download_rows(back.ctypes.data, 260 * 16, d_big.ptr, 260 * 16, 300)
assert (big == back).all()
rgba = np.random.default_rng(0).random((33, 47, 4), dtype=np.float32)
ctx.rgba32f_to_srgb_gray8(rgba); ctx.rgba32f_to_srgb8_packed(rgba, 4); ctx.srgb8_packed_to_rgba32f((rgba * 255).astype(np.uint8))
for (h, w) in ((20, 18), (61, 97), (130, 300)):
    a, b = s.flow_pair(h, w, seed=4)
    f, it = ctx.tvl1(a, b, p.Tvl1Params(nscales=3, warps=2, iterations=4, outer_iterations=2))
    assert np.isfinite(f).all() and it > 0
r8 = (np.random.default_rng(1).random((36, 48, 4)) * 255).astype(np.uint8)   # widths that take the 4-pixel kernels
ctx.rgba8_to_rgb8_mask(r8, 2); ctx.rgb8_to_rgba8(np.ascontiguousarray(r8[..., :3]))
ctx.rgb8_to_rgba8_noise(np.ascontiguousarray(r8[..., :3]), (r8[..., 0] > 200).astype(np.uint8) * 255, 2, 5)
rg = np.random.default_rng(2).random((9, 260, 4), dtype=np.float32)
ctx.rgba32f_to_srgb_gray8(rg); ctx.rgba32f_to_srgb8_packed(rg, 4); ctx.srgb8_packed_to_rgba32f((rg * 255).astype(np.uint8))
ctx.close()
print("sanitize_small: done")
