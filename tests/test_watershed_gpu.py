"""Bit-exact parity of the CUDA watershed with the CPU oracle (np.array_equal on the int32 label map)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("h,w,n,seed", [(120, 160, 9, 5), (480, 640, 64, 7), (33, 47, 3, 1)])
def test_watershed_vs_oracle(ctx, oracle, synth, h, w, n, seed):
    img = synth.texture(h, w, seed + 10)
    mk = synth.seed_markers(h, w, n, seed)
    got = ctx.watershed(img, mk)
    ref, pops = oracle.watershed(img, mk)
    assert np.array_equal(got, ref)
    assert ctx.watershed_stats()["pops"] == pops


def test_watershed_white_noise_and_negatives(ctx, oracle):
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (90, 121, 3), dtype=np.uint8)
    mk = np.zeros((90, 121), np.int32)
    mk[10:14, 10:14] = 1
    mk[60:64, 100:104] = 2
    mk[40, 50] = 3
    mk[5, 5] = -1      # stale ridge from an earlier run: must be treated as unlabelled
    mk[0, 30] = 7      # label on the border ring: overwritten by -1
    got = ctx.watershed(img, mk)
    ref, _ = oracle.watershed(img, mk)
    assert np.array_equal(got, ref)


def test_watershed_no_seeds(ctx, oracle):
    img = np.zeros((20, 30, 3), np.uint8)
    mk = np.zeros((20, 30), np.int32)
    got = ctx.watershed(img, mk)
    ref, _ = oracle.watershed(img, mk)
    assert np.array_equal(got, ref)


def test_watershed_c3_4k_256_seeds(ctx, oracle, synth):
    """BASELINE.json config 3 itself: 3840x2160, 256 seed markers — the whole int32 label map equals the oracle's
    (which equals cv2.watershed on this very input: tests/golden + SURVEY A.3), pop count included."""
    h, w = 2160, 3840
    img = synth.texture(h, w, 4)
    mk = synth.seed_markers(h, w, 256, 5)
    got = ctx.watershed(img, mk)
    ref, pops = oracle.watershed(img, mk)
    assert np.array_equal(got, ref)
    assert ctx.watershed_stats()["pops"] == pops
    assert (got[0] == -1).all() and (got[:, 0] == -1).all() and set(np.unique(got)) <= set(range(-1, 257)) and (got != 0).all()


def test_watershed_clip_equals_frame_by_frame(pkg, ctx, oracle, synth):
    """Frames in flight (ofxcv_watershed_u8c3_batch through the clip driver) give each frame's own label map."""
    import importlib
    seq = importlib.import_module("openfx-opencv_b200.sequence")
    h, w = 60, 84
    frames = [(synth.texture(h, w, 50 + f), synth.seed_markers(h, w, 4 + f, 60 + f)) for f in range(7)]
    first, labs, sums = seq.watershed_clip(ctx, lambda t: frames[t], len(frames), frames_in_flight=3)
    assert first == 0 and len(labs) == 7
    for (img, mk), lab, cs in zip(frames, labs, sums):
        ref, _ = oracle.watershed(img, mk)
        assert np.array_equal(lab, ref) and cs == seq.checksum64(ref)


def test_watershed_1080p_1000_seeds(ctx, oracle, synth):
    h, w = 1080, 1920
    img = synth.texture(h, w, 91)
    mk = synth.seed_markers(h, w, 1000, 92)
    got = ctx.watershed(img, mk)
    ref, pops = oracle.watershed(img, mk)
    assert np.array_equal(got, ref) and ctx.watershed_stats()["pops"] == pops


def _image_classes(h, w, seed):
    rng = np.random.default_rng(seed)
    yield "noise", rng.integers(0, 256, (h, w, 3), dtype=np.uint8)            # levels spread over 0..255: phases above 32
    yield "flat", np.full((h, w, 3), 77, np.uint8)                             # one phase, pure BFS generations
    g = np.zeros((h, w, 3), np.uint8)
    g[..., 0] = (np.arange(w) % 256)[None, :]
    g[..., 1] = (np.arange(h) % 256)[:, None]
    yield "gradient", g                                                        # level 1 everywhere, 255 at the wrap lines
    b = rng.integers(0, 256, (h // 16 + 1, w // 16 + 1, 3), dtype=np.uint8)
    yield "blocks", np.ascontiguousarray(np.kron(b, np.ones((16, 16, 1), np.uint8))[:h, :w])   # big sub-floods


@pytest.mark.parametrize("mode", ["par", "seq"])
def test_watershed_both_floods_on_image_classes(ctx, oracle, synth, monkeypatch, mode):
    """The exact intra-frame parallel flood (watershed_par.cu: what one frame uses) and the one-thread flood (what a
    long clip uses) both give the oracle's label map and pop count on every image class, whatever the scheduling."""
    monkeypatch.setenv("OFXCV_WS_MODE", mode)
    h, w = 270, 480
    mk = synth.seed_markers(h, w, 40, 5)
    for name, img in _image_classes(h, w, 7):
        ref, pops = oracle.watershed(img, mk)
        for rep in range(2 if mode == "par" else 1):      # the parallel flood converges to the same fixed point every time
            got = ctx.watershed(img, mk)
            assert np.array_equal(got, ref), (mode, name)
            st = ctx.watershed_stats()
            assert st["pops"] == pops, (mode, name, st)


def test_watershed_parallel_flood_is_the_single_frame_path(ctx, synth):
    """One frame goes through the round-synchronous parallel flood (rounds > 0 in the statistics), not the one-thread kernel."""
    h, w = 240, 320
    ctx.watershed(synth.texture(h, w, 3), synth.seed_markers(h, w, 12, 4))
    import ctypes as C
    s = (C.c_int64 * 4)()
    import importlib
    importlib.import_module("openfx-opencv_b200").lib().ofxcv_watershed_last_stats(ctx.h, s)
    assert s[2] > 0 and s[3] >= s[2]
