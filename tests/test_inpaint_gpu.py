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


def test_inpaint_clip_equals_frame_by_frame(pkg, ctx, oracle, synth):
    """ofxcv_inpaint_sequence_u8[_host]: several frames in flight on worker sub-contexts give the bits of one frame
    at a time (and of the oracle), for every in-flight count incl. more workers than frames."""
    h, w = 72, 96
    imgs = [synth.texture(h, w, 30 + f) for f in range(5)]
    masks = [synth.iid_mask(h, w, 40 + f, 0.05 + 0.03 * f) for f in range(4)] + [np.zeros((h, w), np.uint8)]   # last: nothing to do
    for method in (pkg.INPAINT_NS, pkg.INPAINT_TELEA):
        ref = [oracle.inpaint(a, m, 3, method) for a, m in zip(imgs, masks)]
        for k in (0, 1, 3, 8):
            outs = ctx.inpaint_sequence(imgs, masks, 3, method, frames_in_flight=k)
            assert all(np.array_equal(o, r) for o, r in zip(outs, ref)), (method, k)
    # device-pointer flavour
    di = [ctx.to_device(a) for a in imgs]; dm = [ctx.to_device(m) for m in masks]; do = [ctx.alloc(h * w * 3) for _ in imgs]
    ctx.inpaint_sequence_dev([b.ptr for b in di], 3, [b.ptr for b in dm], [b.ptr for b in do], w, h, 3.0, pkg.INPAINT_TELEA, 2)
    ref = [oracle.inpaint(a, m, 3, pkg.INPAINT_TELEA) for a, m in zip(imgs, masks)]
    assert all(np.array_equal(b.download((h, w, 3), np.uint8), r) for b, r in zip(do, ref))
    assert ctx.inpaint_sequence([], [], 3, 0) == []
    assert ctx.inpaint_stats()["hole_pixels"] == sum(int((m != 0).sum()) for m in masks)


def test_inpaint_c4_4k_ns_10pct(ctx, oracle, synth):
    """BASELINE.json config 4's frame: 3840x2160 RGB8, 10 % iid mask, radius 3, Navier-Stokes — 0 differing bytes."""
    img = synth.texture(2160, 3840, 100)
    mask = synth.iid_mask(2160, 3840, 1000, 0.10)
    got = ctx.inpaint(img, mask, 3, NS)
    ref = oracle.inpaint(img, mask, 3, NS)
    assert int((got != ref).sum()) == 0
    assert np.array_equal(got[mask == 0], img[mask == 0])


def test_inpaint_gray_1080p_telea_radius5(ctx, oracle, synth):
    """One channel, a larger radius (81 -> 121 taps), blobs + iid holes together."""
    h, w = 1080, 1920
    img = synth.gray(synth.texture(h, w, 77))
    mask = np.maximum(synth.iid_mask(h, w, 78, 0.03), synth.blob_mask(h, w, 79, 40, 12))
    got = ctx.inpaint(img, mask, 5, TELEA)
    ref = oracle.inpaint(img, mask, 5, TELEA)
    assert int((got != ref).sum()) == 0


@pytest.mark.parametrize("method", [TELEA, NS])
@pytest.mark.parametrize("cn", [1, 3])
def test_inpaint_mask_shapes_both_schedulers(ctx, oracle, synth, method, cn):
    h, w = 240, 320
    img = synth.texture(h, w, 9)
    if cn == 1:
        img = synth.gray(img)
    for name, mask in synth.shape_masks(h, w).items():
        for radius in (3, 4):
            got = ctx.inpaint(img, mask, radius, method)
            ref = oracle.inpaint(img, mask, radius, method)
            nbad = int((got != ref).sum())
            assert nbad == 0, "%s: %d differing bytes (method %d, cn %d, r %d)" % (name, nbad, method, cn, radius)


@pytest.mark.parametrize("radius", [3, 5])
@pytest.mark.parametrize("method", [TELEA, NS])
def test_inpaint_in_place(ctx, oracle, synth, method, radius):
    """out == img: the fill must not depend on what a warp running ahead has already written into the image."""
    h, w = 200, 300
    img = synth.texture(h, w, 5)
    mask = synth.iid_mask(h, w, 6, 0.25)
    mask[0, :] = 255          # a whole border row of holes: taps on row 1 read it through clamped indices
    d_img, d_mask = ctx.to_device(img), ctx.to_device(mask)
    ctx.inpaint_dev(d_img.ptr, 3, d_mask.ptr, d_img.ptr, w, h, float(radius), method)
    ctx.synchronize()
    got = d_img.download((h, w, 3), np.uint8)
    assert int((got != oracle.inpaint(img, mask, radius, method)).sum()) == 0
