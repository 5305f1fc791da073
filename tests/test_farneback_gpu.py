"""Parity of the CUDA Farneback path (through the C ABI, host-buffer entry point) against the CPU oracle.

Tolerance: mean |d| <= 1e-3 px, frac(|d|inf > 1e-2) <= 1e-3,
frac(|d|inf > 1) <= 2e-4.  The kernels reproduce the oracle's arithmetic types, so in practice almost every
pixel is bit-identical; STRICT_* below pins that much tighter so that regressions show.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

STRICT_MEAN = 1e-5
STRICT_FRAC_1E2 = 1e-4


def _stats(a, b):
    d = np.abs(a - b).max(axis=2)
    return float(d.mean()), float((d > 1e-2).mean()), float((d > 1).mean()), float((d == 0).mean())


@pytest.mark.parametrize("h,w,levels,iters", [(96, 128, 1, 1), (96, 128, 1, 3), (120, 160, 3, 15), (270, 480, 3, 15), (135, 241, 2, 4)])
def test_farneback_vs_oracle(ctx, pkg, oracle, synth, h, w, levels, iters):
    prev, nxt = synth.flow_pair(h, w, seed=3)
    p = pkg.FbParams(levels=levels, iterations=iters)
    got = ctx.farneback(prev, nxt, p)
    ref = oracle.farneback(prev, nxt, levels=levels, iters=iters)
    mean, f2, f0, same = _stats(got, ref)
    print("farneback %dx%d L%d I%d: mean %.3g frac>1e-2 %.3g frac>1 %.3g identical %.5f" % (w, h, levels, iters, mean, f2, f0, same))
    assert mean <= STRICT_MEAN and f2 <= STRICT_FRAC_1E2 and f0 <= 2e-4


def test_farneback_1080p(ctx, pkg, oracle, synth):
    prev, nxt = synth.flow_pair(1080, 1920, seed=3)
    got = ctx.farneback(prev, nxt)
    ref = oracle.farneback(prev, nxt)
    mean, f2, f0, same = _stats(got, ref)
    print("farneback 1080p: mean %.3g frac>1e-2 %.3g frac>1 %.3g identical %.5f" % (mean, f2, f0, same))
    assert mean <= 1e-3 and f2 <= 1e-3 and f0 <= 2e-4
    # end-point error against the true translation, interior
    epe = np.hypot(got[100:-100, 100:-100, 0] - 2.5, got[100:-100, 100:-100, 1] + 1.5).mean()
    epe_ref = np.hypot(ref[100:-100, 100:-100, 0] - 2.5, ref[100:-100, 100:-100, 1] + 1.5).mean()
    assert epe <= epe_ref * 1.01 + 1e-6


def test_farneback_params_and_errors(ctx, pkg, synth):
    prev, nxt = synth.flow_pair(64, 80, seed=5)
    with pytest.raises(pkg.OfxcvError):
        ctx.farneback(prev, nxt, pkg.FbParams(winsize=5))
    with pytest.raises(pkg.OfxcvError):
        ctx.farneback(prev, nxt, pkg.FbParams(flags=4))
    # poly_n 7 / sigma 1.5 (the other setting OpenCV documents)
    import oracle
    got = ctx.farneback(prev, nxt, pkg.FbParams(poly_n=7, poly_sigma=1.5, levels=1, iterations=2))
    ref = oracle.farneback(prev, nxt, levels=1, iters=2, poly_n=7, poly_sigma=1.5)
    assert np.abs(got - ref).max(axis=2).mean() <= STRICT_MEAN


def test_sequence_equals_pairwise_and_builds_each_pyramid_once(ctx, pkg, synth):
    """ofxcv_farneback_sequence_u8[_host] must give the bits of the pair-by-pair call, and build each frame's pyramid once."""
    h, w, n = 135, 240, 5
    base = synth.gray(synth.texture(h, w, seed=11))
    frames = [synth.shift_bilinear(base, 1.5 * f, -0.75 * f) for f in range(n)]
    p = pkg.FbParams(levels=2, iterations=4)
    ref = [ctx.farneback(frames[t], frames[t + 1], p) for t in range(n - 1)]
    b0, h0 = ctx.farneback_cache_stats()
    got = ctx.farneback_sequence(frames, p)
    b1, h1 = ctx.farneback_cache_stats()
    assert b1 - b0 == n                            # one pyramid per frame (the host pipeline hands them on directly)
    for t in range(n - 1):
        assert np.array_equal(got[t], ref[t]), "host sequence flow %d differs from the pairwise call" % t
    # device-resident flavour
    d_frames = ctx.to_device(np.stack(frames))
    d_flows = ctx.alloc((n - 1) * w * h * 8)
    ctx.farneback_sequence_dev(d_frames.ptr, w, h, n, d_flows.ptr, p)
    dev = d_flows.download((n - 1, h, w, 2), np.float32)
    b2, h2 = ctx.farneback_cache_stats()
    assert b2 - b1 == n and h2 == h1               # again one pyramid per frame
    for t in range(n - 1):
        assert np.array_equal(dev[t], ref[t])
    # pinned host buffers take the overlapped path: same bits
    pf = [ctx.pinned_array((h, w), np.uint8) for _ in range(n)]
    for a, b in zip(pf, frames):
        a[...] = b
    po = [ctx.pinned_array((h, w, 2), np.float32) for _ in range(n - 1)]
    ctx.farneback_sequence(pf, p, out=po)
    for t in range(n - 1):
        assert np.array_equal(po[t], ref[t])


def test_keyed_forward_backward_shares_pyramids(ctx, pkg, synth):
    """A default VectorGenerator render needs t->t+1 and t->t-1: with keys, frame t is expanded once."""
    h, w = 96, 128
    base = synth.gray(synth.texture(h, w, seed=12))
    f = [synth.shift_bilinear(base, 2.0 * i, 1.0 * i) for i in range(3)]
    d = [ctx.to_device(x) for x in f]
    out_f, out_b = ctx.alloc(w * h * 8), ctx.alloc(w * h * 8)
    p = pkg.FbParams(levels=1, iterations=3)
    ref_f, ref_b = ctx.farneback(f[1], f[2], p), ctx.farneback(f[1], f[0], p)
    b0, h0 = ctx.farneback_cache_stats()
    ctx.farneback_keyed_dev(d[1].ptr, d[2].ptr, w, h, out_f.ptr, 1001, 1002, p)
    ctx.farneback_keyed_dev(d[1].ptr, d[0].ptr, w, h, out_b.ptr, 1001, 1000, p)
    ctx.synchronize()
    b1, h1 = ctx.farneback_cache_stats()
    assert (b1 - b0, h1 - h0) == (3, 1)
    assert np.array_equal(out_f.download((h, w, 2), np.float32), ref_f)
    assert np.array_equal(out_b.download((h, w, 2), np.float32), ref_b)
    # a different size under the same key must not hit
    small = synth.gray(synth.texture(64, 80, seed=13))
    s0, s1, sf = ctx.to_device(small), ctx.to_device(small), ctx.alloc(80 * 64 * 8)   # alive until the synchronize below
    ctx.farneback_keyed_dev(s0.ptr, s1.ptr, 80, 64, sf.ptr, 1001, 1002, p)
    ctx.synchronize()
    b2, h2 = ctx.farneback_cache_stats()
    assert b2 - b1 == 2 and h2 == h1


def test_content_key(ctx, synth):
    a = synth.gray(synth.texture(64, 80, seed=21))
    b = a.copy()
    b[10, 10] ^= 1
    c = np.ascontiguousarray(a[:, ::-1])          # same histogram, different positions
    bufs = [ctx.to_device(x) for x in (a, b, c, a.copy())]    # keep the device buffers alive while their keys are taken
    ka, kb, kc, ka2 = (ctx.content_key(d.ptr, 80, 64) for d in bufs)
    assert ka == ka2
    assert len({ka, kb, kc}) == 3 and 0 not in (ka, kb, kc)


@pytest.mark.parametrize("h,w,levels,iters,poly_n", [(7, 29, 0, 2, 5), (9, 8, 0, 3, 5), (40, 50, 3, 3, 5), (33, 31, 1, 2, 3), (64, 95, 2, 0, 5), (61, 62, 2, 1, 7)])
def test_farneback_edge_shapes(ctx, pkg, oracle, synth, h, w, levels, iters, poly_n):
    """Single strip / single band images, level clipping (< 32 px), zero and one iteration, the generic PolyExp path."""
    prev, nxt = synth.flow_pair(h, w, seed=7, dx=1.25, dy=0.5)
    sigma = 1.1 if poly_n == 5 else 1.5 if poly_n == 7 else 0.9
    got = ctx.farneback(prev, nxt, pkg.FbParams(levels=levels, iterations=iters, poly_n=poly_n, poly_sigma=sigma))
    ref = oracle.farneback(prev, nxt, levels=levels, iters=iters, poly_n=poly_n, poly_sigma=sigma)
    mean, f2, f0, same = _stats(got, ref)
    assert mean <= STRICT_MEAN and f2 <= STRICT_FRAC_1E2 and f0 <= 2e-4, (mean, f2, f0, same)


def test_farneback_strided_device_buffers(ctx, pkg, oracle, synth):
    """Device entry point with row padding on both frames and on the flow field (OFX rowBytes > width)."""
    h, w, pad, fpad = 70, 101, 27, 40
    prev, nxt = synth.flow_pair(h, w, seed=9)
    buf = np.full((2, h, w + pad), 123, np.uint8)
    buf[0, :, :w], buf[1, :, :w] = prev, nxt
    d = ctx.to_device(buf)
    fstride = w * 8 + fpad
    d_flow = ctx.alloc(h * fstride)
    p = pkg.FbParams(levels=2, iterations=3)
    ctx.farneback_dev(d.ptr, d.ptr + h * (w + pad), w, h, d_flow.ptr, p, stride=w + pad, flow_stride=fstride)
    ctx.synchronize()
    raw = d_flow.download((h, fstride), np.uint8)
    got = np.ascontiguousarray(raw[:, :w * 8]).view(np.float32).reshape(h, w, 2)
    ref = oracle.farneback(prev, nxt, levels=2, iters=3)
    assert np.abs(got - ref).max(axis=2).mean() <= STRICT_MEAN


def test_farneback_4k_properties(ctx, pkg, synth):
    """BASELINE config size (3840x2160, default parameters): size-independent properties instead of the (slow) oracle --
    the known global translation is recovered, two runs give identical bits, the clip call equals the pair call."""
    h, w = 2160, 3840
    base = synth.gray(synth.texture(h, w, seed=2000))
    f = [base, synth.shift_bilinear(base, 2.5, -1.5), synth.shift_bilinear(base, 5.0, -3.0)]
    a = ctx.farneback(f[0], f[1])
    b = ctx.farneback(f[0], f[1])
    assert np.array_equal(a, b)
    epe = np.hypot(a[200:-200, 200:-200, 0] - 2.5, a[200:-200, 200:-200, 1] + 1.5)
    assert epe.mean() < 0.5 and np.median(epe) < 0.3, (epe.mean(), np.median(epe))   # u8 texture, polyN 5: ~0.2 px
    seq = ctx.farneback_sequence(f)
    assert np.array_equal(seq[0], a) and np.array_equal(seq[1], ctx.farneback(f[1], f[2]))


def test_farneback_4k_vs_oracle(ctx, pkg, oracle, synth):
    """The bench's headline workload itself (3840x2160, default plugin parameters, the bench's frames: Texture(seed 2000)
    and its (2.5, -1.5) px translate) against the CPU oracle (about 8 s), SURVEY.md 8c tolerances."""
    h, w = 2160, 3840
    base = synth.gray(synth.texture(h, w, seed=2000))
    nxt = synth.shift_bilinear(base, 2.5, -1.5)
    got = ctx.farneback(base, nxt)
    ref = oracle.farneback(base, nxt)
    mean, f2, f0, same = _stats(got, ref)
    print("farneback 4K: mean %.3g frac>1e-2 %.3g frac>1 %.3g identical %.5f" % (mean, f2, f0, same))
    assert mean <= 1e-3 and f2 <= 1e-3 and f0 <= 2e-4, (mean, f2, f0, same)


def test_flow_clip_driver_single_rank(ctx, pkg, synth):
    """sequence.flow_clip (the multi-GPU clip driver) on one rank: loads each frame once, equals the pairwise call."""
    import importlib
    seq = importlib.import_module("openfx-opencv_b200.sequence")
    h, w, n = 90, 120, 4
    base = synth.gray(synth.texture(h, w, seed=31))
    loaded = []

    def load(t):
        loaded.append(t)
        return synth.shift_bilinear(base, 1.0 * t, 0.5 * t)

    p = pkg.FbParams(levels=1, iterations=3)
    first, flows, sums = seq.flow_clip(ctx, load, n, p)
    assert first == 0 and loaded == [0, 1, 2, 3] and len(flows) == n - 1
    for t in range(n - 1):
        ref = ctx.farneback(synth.shift_bilinear(base, 1.0 * t, 0.5 * t), synth.shift_bilinear(base, 1.0 * (t + 1), 0.5 * (t + 1)), p)
        assert np.array_equal(flows[t], ref) and sums[t] == seq.checksum64(ref)


def test_farneback_c5_8k_levels5_properties(ctx, pkg, synth):
    """BASELINE.json config 5's frame (7680x4320 Texture(seed 2000) shifted by (2.5, -1.5) px, levels = 5 -> 6 scales):
    the oracle would take minutes, so size-independent properties: the known translation is recovered as well as at 4K,
    two runs are bit-identical."""
    h, w = 4320, 7680
    par = pkg.FbParams(levels=5)
    assert pkg.lib().ofxcv_farneback_scales(w, h, par) == 6
    base = synth.gray(synth.texture(h, w, seed=2000))
    nxt = synth.shift_bilinear(base, 2.5, -1.5)
    a = ctx.farneback(base, nxt, par)
    b = ctx.farneback(base, nxt, par)
    assert np.array_equal(a, b) and np.isfinite(a).all()
    c = a[400:-400, 400:-400]
    epe = np.hypot(c[..., 0] - 2.5, c[..., 1] + 1.5)
    assert epe.mean() < 0.5 and np.median(epe) < 0.3, (epe.mean(), np.median(epe))


def test_farneback_c5_8k_levels5_vs_oracle(ctx, pkg, oracle, synth):
    """BASELINE.json config 5's frame pair (7680x4320, levels = 5 -> 6 scales) against the CPU oracle (about 35 s),
    SURVEY.md 8c tolerances."""
    h, w = 4320, 7680
    par = pkg.FbParams(levels=5)
    base = synth.gray(synth.texture(h, w, seed=2000))
    nxt = synth.shift_bilinear(base, 2.5, -1.5)
    got = ctx.farneback(base, nxt, par)
    ref = oracle.farneback(base, nxt, levels=5)
    mean, f2, f0, same = _stats(got, ref)
    print("farneback 8K L5: mean %.3g frac>1e-2 %.3g frac>1 %.3g identical %.5f" % (mean, f2, f0, same))
    assert mean <= 1e-3 and f2 <= 1e-3 and f0 <= 2e-4, (mean, f2, f0, same)


def test_farneback_720p_polyn7_levels4(ctx, pkg, oracle, synth):
    """The other PolyExp instantiation (N = 7, sigma 1.5) on a 5-scale pyramid."""
    prev, nxt = synth.flow_pair(720, 1280, seed=11, dx=3.25, dy=-2.5)
    par = pkg.FbParams(levels=4, iterations=5, poly_n=7, poly_sigma=1.5)
    got = ctx.farneback(prev, nxt, par)
    ref = oracle.farneback(prev, nxt, levels=4, iters=5, poly_n=7, poly_sigma=1.5)
    assert np.array_equal(got, ref)
