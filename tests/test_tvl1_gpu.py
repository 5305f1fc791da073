"""Dual TV-L1 (the VectorGenerator plugin's second method, VectorGenerator.cpp:436-492) through the C ABI against the
CPU oracle.  Parity of the METHOD is unpinned (no OpenCV build with DualTVL1 here); what these tests establish is that
the CUDA path and the oracle — whose OpenCV primitives are pinned to cv2 in test_oracle_golden.py — agree bit for bit,
including the data-dependent early exit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pair(synth, h, w, seed, dx=2.5, dy=-1.5):
    base = synth.gray(synth.texture(h, w, seed=seed))
    return base, synth.shift_bilinear(base, dx, dy)


@pytest.mark.parametrize("shape,kw", [
    ((48, 64), dict(nscales=3, warps=2, iterations=5, outer_iterations=3)),
    ((61, 97), dict(nscales=5, warps=3, iterations=7, outer_iterations=2)),
    ((61, 97), dict(nscales=4, warps=2, iterations=6, outer_iterations=2, median_filtering=1)),
    ((20, 18), dict(nscales=5, warps=2, iterations=4, outer_iterations=2)),          # scales below 16 px are dropped
    ((5, 7), dict(nscales=3, warps=2, iterations=3, outer_iterations=3)),            # smaller than every window
    ((1, 40), dict(nscales=2, warps=1, iterations=3, outer_iterations=1)),           # a single row
    ((37, 1), dict(nscales=2, warps=1, iterations=3, outer_iterations=1)),           # a single column
    ((33, 300), dict(nscales=2, warps=2, iterations=5, outer_iterations=2, tau=0.2, lambda_=0.3, theta=0.25)),
    ((120, 160), dict()),                                                            # the plugin's defaults
])
def test_tvl1_matches_oracle_bit_for_bit(pkg, ctx, oracle, synth, shape, kw):
    h, w = shape
    prev, nxt = _pair(synth, h, w, seed=h + w)
    got, it = ctx.tvl1(prev, nxt, pkg.Tvl1Params(**kw))
    okw = dict(kw)
    if "iterations" in okw: okw["inner"] = okw.pop("iterations")
    if "outer_iterations" in okw: okw["outer"] = okw.pop("outer_iterations")
    if "median_filtering" in okw: okw["median"] = okw.pop("median_filtering")
    ref, it_ref = oracle.tvl1(prev, nxt, **okw)
    assert it == it_ref                      # same early-exit decisions
    assert np.array_equal(got, ref)


def test_tvl1_early_exit_and_large_motion(pkg, ctx, oracle, synth):
    """A loose epsilon stops every warping after a few iterations; a large shift sends taps outside the image (the
    constant-0 border branch of the remap)."""
    prev, nxt = _pair(synth, 72, 88, seed=5, dx=9.0, dy=-7.0)
    par = pkg.Tvl1Params(epsilon=0.08, nscales=4)
    got, it = ctx.tvl1(prev, nxt, par)
    ref, it_ref = oracle.tvl1(prev, nxt, epsilon=0.08, nscales=4)
    assert it == it_ref and it < 4 * 5 * 10 * 15 // 4
    assert np.array_equal(got, ref)


def test_tvl1_strides_and_repeat(pkg, ctx, synth):
    """Row strides on both sides; a second call on the same context (workspace reuse) gives the same bits."""
    h, w = 50, 70
    prev, nxt = _pair(synth, h, w, seed=9)
    par = pkg.Tvl1Params(nscales=3, warps=2, iterations=5, outer_iterations=2)
    ref, _ = ctx.tvl1(prev, nxt, par)
    ps, fs = w + 13, (w + 3) * 8
    pa = np.zeros((h, ps), np.uint8); pb = np.zeros((h, ps), np.uint8)
    pa[:, :w], pb[:, :w] = prev, nxt
    a, b, f = ctx.to_device(pa), ctx.to_device(pb), ctx.alloc(h * fs)
    ctx.tvl1_dev(a.ptr, b.ptr, w, h, f.ptr, par, stride=ps, flow_stride=fs)
    out = f.download((h, w + 3, 2), np.float32)[:, :w]
    assert np.array_equal(out, ref)
    again, _ = ctx.tvl1(prev, nxt, par)
    assert np.array_equal(again, ref)


def test_tvl1_bad_arguments(pkg, ctx):
    a = ctx.alloc(64 * 64); f = ctx.alloc(64 * 64 * 8)
    with pytest.raises(pkg.OfxcvError):
        ctx.tvl1_dev(a.ptr, a.ptr, 64, 64, f.ptr, pkg.Tvl1Params(median_filtering=3))
    with pytest.raises(pkg.OfxcvError):
        ctx.tvl1_dev(a.ptr, a.ptr, 64, 64, f.ptr, pkg.Tvl1Params(scale_step=1.0))
    with pytest.raises(pkg.OfxcvError):
        ctx.tvl1_dev(a.ptr, a.ptr, 64, 64, f.ptr, pkg.Tvl1Params(), stride=32)
    with pytest.raises(pkg.OfxcvError):
        ctx.tvl1_dev(None, a.ptr, 64, 64, f.ptr, pkg.Tvl1Params())


def test_tvl1_1080p_recovers_translation(pkg, ctx, synth):
    """BASELINE config 2's frame size: the (2.5, -1.5) px shift is recovered; deterministic from run to run."""
    h, w = 1080, 1920
    prev, nxt = _pair(synth, h, w, seed=3)
    par = pkg.Tvl1Params(warps=3, outer_iterations=2, iterations=10)
    flow, it = ctx.tvl1(prev, nxt, par)
    inner = flow[32:-32, 32:-32]
    assert abs(np.median(inner[..., 0]) - 2.5) < 0.05 and abs(np.median(inner[..., 1]) + 1.5) < 0.05
    assert it > 0
    flow2, it2 = ctx.tvl1(prev, nxt, par)
    assert it2 == it and np.array_equal(flow, flow2)


def test_tvl1_clip_driver_single_rank(pkg, ctx, synth):
    """sequence.flow_clip with method="tvl1": every pair of the clip equals the pair call."""
    import importlib
    seq = importlib.import_module("openfx-opencv_b200.sequence")
    base = synth.gray(synth.texture(48, 64, seed=17))
    frames = [synth.shift_bilinear(base, 1.0 * t, -0.5 * t) for t in range(4)]
    par = pkg.Tvl1Params(nscales=3, warps=2, iterations=4, outer_iterations=2)
    first, flows, sums = seq.flow_clip(ctx, lambda t: frames[t], len(frames), par, method="tvl1")
    assert first == 0 and len(flows) == 3
    for t in range(3):
        ref, _ = ctx.tvl1(frames[t], frames[t + 1], par)
        assert np.array_equal(flows[t], ref) and sums[t] == seq.checksum64(ref)
