"""Bit-exact parity of the CUDA FMM inpaint (Telea + Navier-Stokes) with the CPU oracle, through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TELEA, NS = 1, 0


def _masks(synth, h, w):
    border = np.zeros((h, w), np.uint8)
    border[0:3, :] = 255
    border[:, w - 2:] = 255
    border[h // 2, 0] = 255
    blob = np.zeros((h, w), np.uint8)
    blob[10:22, 12:28] = 255
    return {
        "iid05": synth.iid_mask(h, w, 2, 0.05),
        "iid10": synth.iid_mask(h, w, 3, 0.10),
        "iid30": synth.iid_mask(h, w, 4, 0.30),
        "blob": blob,
        "blobs": synth.blob_mask(h, w, 5, nblobs=5, rmax=7),
        "border": border,
    }


@pytest.mark.parametrize("method", [TELEA, NS])
@pytest.mark.parametrize("cn", [1, 3])
@pytest.mark.parametrize("radius", [1.4, 3, 5])
def test_inpaint_small_all_masks(ctx, oracle, synth, method, cn, radius):
    h, w = 40, 56
    img = synth.texture(h, w, 1)
    if cn == 1:
        img = synth.gray(img)
    for name, mask in _masks(synth, h, w).items():
        got = ctx.inpaint(img, mask, radius, method)
        ref = oracle.inpaint(img, mask, radius, method)
        nbad = int((got != ref).sum())
        assert nbad == 0, "%s: %d differing bytes (method %d, cn %d, r %s)" % (name, nbad, method, cn, radius)
        assert np.array_equal(got[mask == 0], img[mask == 0])


@pytest.mark.parametrize("method", [TELEA, NS])
def test_inpaint_c1_640x480(ctx, oracle, synth, method):
    """BASELINE.json config 1: 640x480, 5% random mask, radius 3."""
    img = synth.texture(480, 640, 1)
    mask = synth.iid_mask(480, 640, 2, 0.05)
    got = ctx.inpaint(img, mask, 3, method)
    ref = oracle.inpaint(img, mask, 3, method)
    assert int((got != ref).sum()) == 0
    st = ctx.inpaint_stats()
    assert st["hole_pixels"] == int((mask != 0).sum())


def test_inpaint_march_order_and_T(ctx, oracle, synth):
    """The marching stage alone: T map and fill order must equal the sequential CPU march."""
    h, w = 60, 84
    img = synth.texture(h, w, 2)
    for mask in (synth.iid_mask(h, w, 3, 0.3), synth.blob_mask(h, w, 5, nblobs=6, rmax=12)):
        for method in (TELEA, NS):
            got = ctx.inpaint(img, mask, 3, method)
            t, order = ctx.inpaint_debug_maps(w, h)
            ref, t_ref, seq_ref = oracle.inpaint(img, mask, 3, method, want_debug=True)
            assert np.array_equal(order, seq_ref)
            hole = np.pad(mask != 0, 1)
            assert np.array_equal(t[hole], t_ref[hole])
            if method == TELEA:
                assert np.array_equal(t, t_ref)
            assert np.array_equal(got, ref)


def test_inpaint_large_blob(ctx, oracle, synth):
    img = synth.texture(200, 260, 7)
    mask = synth.blob_mask(200, 260, 9, nblobs=4, rmax=30)
    for method in (TELEA, NS):
        got = ctx.inpaint(img, mask, 3, method)
        ref = oracle.inpaint(img, mask, 3, method)
        assert int((got != ref).sum()) == 0


def test_inpaint_edge_cases(ctx, oracle, synth, pkg):
    img = synth.texture(24, 31, 3)
    empty = np.zeros((24, 31), np.uint8)
    assert np.array_equal(ctx.inpaint(img, empty, 3, TELEA), img)
    full = np.full((24, 31), 255, np.uint8)   # nothing known: OpenCV leaves the image untouched
    for method in (TELEA, NS):
        assert np.array_equal(ctx.inpaint(img, full, 3, method), oracle.inpaint(img, full, 3, method))
    one = empty.copy(); one[0, 0] = 255; one[23, 30] = 255; one[5, 7] = 1
    for method in (TELEA, NS):
        assert np.array_equal(ctx.inpaint(img, one, 2, method), oracle.inpaint(img, one, 2, method))
    with pytest.raises(pkg.OfxcvError):
        ctx.inpaint(img, empty, 3, 7)


def test_inpaint_1080p_ns_10pct(ctx, oracle, synth):
    img = synth.texture(1080, 1920, 100)
    mask = synth.iid_mask(1080, 1920, 1000, 0.10)
    got = ctx.inpaint(img, mask, 3, NS)
    ref = oracle.inpaint(img, mask, 3, NS)
    assert int((got != ref).sum()) == 0


@pytest.mark.parametrize("sched", ["0", "1"])
@pytest.mark.parametrize("method", [TELEA, NS])
def test_inpaint_both_schedulers(ctx, oracle, synth, method, sched, monkeypatch):
    """The fill kernel's two schedulers (in-order tickets + flag spinning / dependency counters + ready queue) must give
    the same bytes on every kind of mask; the library picks one by itself, OFXCV_IP_READYQ forces it."""
    monkeypatch.setenv("OFXCV_IP_READYQ", sched)
    h, w = 96, 160
    img = synth.texture(h, w, 6)
    masks = _masks(synth, h, w)
    lines = np.zeros((h, w), np.uint8)
    lines[8:90:9, 5:150] = 255                     # scratches: every line is one long dependency chain
    lines[20:70, 77] = 255
    masks["lines"] = lines
    for name, mask in masks.items():
        for radius in (3, 6):
            got = ctx.inpaint(img, mask, radius, method)
            ref = oracle.inpaint(img, mask, radius, method)
            assert np.array_equal(got, ref), "%s r=%d scheduler %s: %d differing bytes" % (name, radius, sched, int((got != ref).sum()))
