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
