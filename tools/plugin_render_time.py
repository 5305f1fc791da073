"""Render times through the drop-in .ofx bundles themselves (mini-host, 4K): the path an OFX host takes.
VectorGenerator: default parameters = forward AND backward flow per render (two Farneback solves), float RGBA clips in
host memory and as device pointers (OfxImageEffectPropCudaEnabled); inpaint / segment: RGBA8 clips in host memory."""
import importlib, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("openfx-opencv_b200"); synth = importlib.import_module("openfx-opencv_b200.synth")
mh = importlib.import_module("openfx-opencv_b200.minihost")
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
ctx = pkg.Context(0)
base = synth.gray(synth.texture(H, W, seed=2000))
lin = (np.arange(256, dtype=np.float32) / 255.0) ** 2.2      # any monotone byte -> linear float ramp will do here
nfr = 6
frames = {}
for t in range(nfr):
    g = synth.shift_bilinear(base, 2.5 * t, -1.5 * t)
    f = np.empty((H, W, 4), np.float32); f[..., 0] = f[..., 1] = f[..., 2] = lin[g]; f[..., 3] = 1.0
    frames[t] = f
p = mh.Plugin("VectorGenerator"); assert p.create_instance() == 0
for t, f in frames.items():
    p.set_image("Source", t, f)
dst = np.zeros((H, W, 4), np.float32)
times = []
for t in range(1, nfr - 1):
    p.clear_images("Output"); p.set_image("Output", t, dst)
    t0 = time.perf_counter(); assert p.render(t, (0, 0, W, H)) == 0; times.append(time.perf_counter() - t0)
print("VectorGenerator %dx%d host-memory clips, default params (fwd+bwd): renders %s ms" % (W, H, ["%.1f" % (x * 1e3) for x in times]))
d = {t: ctx.to_device(f) for t, f in frames.items()}
dout = ctx.alloc(W * H * 16)
p.clear_images("Source"); p.clear_images("Output")
for t in d:
    p.set_device_image("Source", t, d[t].ptr, W, H, 4, np.float32)
times = []
for t in range(1, nfr - 1):
    p.clear_images("Output"); p.set_device_image("Output", t, dout.ptr, W, H, 4, np.float32)
    t0 = time.perf_counter(); assert p.render(t, (0, 0, W, H), cuda_enabled=1) == 0; times.append(time.perf_counter() - t0)
print("VectorGenerator %dx%d CUDA-enabled clips (device pointers), fwd+bwd: renders %s ms" % (W, H, ["%.1f" % (x * 1e3) for x in times]))
p.close()
rgb = synth.texture(H, W, seed=4)
mask = synth.iid_mask(H, W, 1000, 0.10)
rgba = np.dstack([np.maximum(rgb, 1), np.full((H, W), 255, np.uint8)]).astype(np.uint8); rgba[mask != 0, :3] = 0
for name in ("inpaint", "segment"):
    q = mh.Plugin(name); assert q.create_instance() == 0
    out = np.zeros((H, W, 4), np.uint8)
    q.set_image("Source", 0, rgba); q.set_image("Output", 0, out)
    ts = []
    for _ in range(2):
        t0 = time.perf_counter(); assert q.render(0, (0, 0, W, H)) == 0; ts.append(time.perf_counter() - t0)
    print("%s %dx%d RGBA8 host-memory clips: renders %s ms" % (name, W, H, ["%.1f" % (x * 1e3) for x in ts]))
    q.close()
