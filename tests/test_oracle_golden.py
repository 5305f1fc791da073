"""Pins the CPU oracle against the committed golden vectors (cv2 4.13 = the OpenCV arithmetic the reference
plugins call; the reference's own ofxsLut.cpp for the staging LUT).  Runs without a GPU, cv2 or /root/reference."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_inpaint_bit_exact(oracle, synth):
    z = np.load(os.path.join(G, "inpaint_cv2.npz"))
    rgb = z["rgb"]
    n = 0
    for key in z.files:
        if not key.startswith("out_"):
            continue
        _, mname, c, r, meth = key.split("_")
        img = rgb if c == "c3" else synth.gray(rgb)
        got = oracle.inpaint(img, z["mask_" + mname], float(r[1:]), 1 if meth == "telea" else 0)
        assert np.array_equal(got, z[key]), key
        n += 1
    assert n == 48


def test_inpaint_mask_shapes_bit_exact(oracle, synth):
    """cv2 outputs on masks whose fill order differs in kind (lines, a bar, diagonal scratches) and on holes on the image's
    border ring, where cv::inpaint reads neighbours through clamped indices (what the GPU fill's source-colour table serves)."""
    z = np.load(os.path.join(G, "inpaint_shapes_cv2.npz"))
    rgb = z["rgb"]
    h, w = rgb.shape[:2]
    masks = synth.shape_masks(h, w)
    n = 0
    for key in z.files:
        if not key.startswith("out_"):
            continue
        _, mname, c, r, meth = key.split("_")
        assert np.array_equal(masks[mname], z["mask_" + mname]), mname   # the GPU tests use the same generator
        img = rgb if c == "c3" else synth.gray(rgb)
        got = oracle.inpaint(img, z["mask_" + mname], float(r[1:]), 1 if meth == "telea" else 0)
        assert np.array_equal(got, z[key]), key
        n += 1
    assert n == 32


def test_watershed_bit_exact(oracle):
    z = np.load(os.path.join(G, "watershed_cv2.npz"))
    for name in ("a", "b", "noise"):
        got, _ = oracle.watershed(z[name + "_img"], z[name + "_markers"])
        assert np.array_equal(got, z[name + "_labels"]), name


def test_farneback_tolerance(oracle):
    """Tolerance of SURVEY.md 8c: mean |d| <= 1e-3 px, frac(|d|>1e-2) <= 1e-3, frac(|d|>1) <= 2e-4."""
    z = np.load(os.path.join(G, "farneback_cv2.npz"))
    for name in ("a", "b", "c"):
        levels, iters, n, sig = z[name + "_params"]
        got = oracle.farneback(z[name + "_prev"], z[name + "_next"], levels=int(levels), iters=int(iters), poly_n=int(n), poly_sigma=float(sig))
        d = np.abs(got - z[name + "_flow"]).max(axis=2)
        assert d.mean() <= 1e-3 and (d > 1e-2).mean() <= 1e-3 and (d > 1).mean() <= 2e-4, (name, d.mean(), d.max())
        assert d.mean() <= 5e-5   # what the restatement actually achieves (float-noise exact)


def test_lut_matches_reference_tables(oracle):
    z = np.load(os.path.join(G, "lut_srgb_ref.npz"))
    to, fr = oracle.srgb_tables()
    assert np.array_equal(((to.astype(np.int32) + 0x80) >> 8).astype(np.uint8), z["to_byte_by_hipart"])
    assert np.array_equal(fr, z["from_byte"])
    rgb = z["luma_rgb"]
    got = oracle.luma_srgb_gray8(rgb.reshape(64, 64, 3))
    assert np.array_equal(got.ravel(), z["luma_byte"])


def test_oracle_building_blocks(oracle):
    # getGaussianKernel fixed table and normalisation; pyramid level count rule
    assert np.array_equal(oracle.gaussian_kernel(3, 0), np.float32([0.25, 0.5, 0.25]))
    k = oracle.gaussian_kernel(9, 1.5)
    assert abs(float(k.sum()) - 1) < 1e-6 and np.array_equal(k, k[::-1])
    L = oracle.lib()
    import ctypes as C
    f = L.orc_farneback_levels
    f.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int]
    assert f(1920, 1080, 0.5, 3) == 3 and f(7680, 4320, 0.5, 5) == 5 and f(100, 100, 0.5, 3) == 1 and f(40, 40, 0.5, 3) == 0


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(G), "..", "oracle", "_ref", "libofxs_lut_ref.so")), reason="oracle/_ref not built")
def test_oracle_lut_vs_compiled_reference(oracle):
    import ctypes as C
    L = C.CDLL(os.path.join(os.path.dirname(G), "..", "oracle", "_ref", "libofxs_lut_ref.so"))
    L.ref_luma_to_byte.restype = C.c_ubyte
    L.ref_luma_to_byte.argtypes = [C.c_float] * 3
    rng = np.random.default_rng(5)
    rgb = rng.random((2000, 3), dtype=np.float32)
    ref = np.array([L.ref_luma_to_byte(*[C.c_float(v) for v in p]) for p in rgb], np.uint8)
    got = oracle.luma_srgb_gray8(rgb.reshape(40, 50, 3)).ravel()
    assert np.array_equal(got, ref)


def test_packed_conversions_match_reference_tables(oracle):
    """oracle to_byte_packed / from_byte_packed against values produced by the reference's own compiled ofxsLut code
    (tests/golden/lut_srgb_ref.npz: hipart table, from-byte table, floatToInt<256> / intToFloat<256> sweeps)."""
    z = np.load(os.path.join(G, "lut_srgb_ref.npz"))
    # colour channels: every hipart code once (mid-bucket float), through a 3-component image
    codes = np.arange(0x10000, dtype=np.uint32)
    f = ((codes << 16) | 0x8000).view(np.float32)
    img = np.stack([f, f, f], axis=-1).reshape(256, 256, 3)
    got = oracle.to_byte_packed(img, 3)
    assert np.array_equal(got[..., 0].ravel(), z["to_byte_by_hipart"]) and np.array_equal(got[..., 2].ravel(), z["to_byte_by_hipart"])
    # alpha: the dense sweep incl. rounding ties, as a 1-component image and as channel 3 of RGBA
    a = z["alpha_in"]
    assert np.array_equal(oracle.to_byte_packed(a.reshape(1, -1), 1)[0, :, 0], z["alpha_byte"])
    rgba = np.zeros((1, a.size, 4), np.float32); rgba[0, :, 3] = a
    out = oracle.to_byte_packed(rgba, 4)
    assert np.array_equal(out[0, :, 3], z["alpha_byte"]) and not out[..., :3].any()
    # RGB source -> RGBA destination leaves alpha 0; RGBA -> alpha-only keeps alpha
    assert not oracle.to_byte_packed(img, 4)[..., 3].any()
    assert np.array_equal(oracle.to_byte_packed(rgba, 1)[0, :, 0], z["alpha_byte"])
    # from_byte_packed: all 256 codes
    b = np.arange(256, dtype=np.uint8)
    back = oracle.from_byte_packed(np.stack([b, b, b, b], axis=-1).reshape(1, 256, 4))
    assert np.array_equal(back[0, :, 0], z["from_byte"]) and np.array_equal(back[0, :, 3], z["alpha_from"])
    assert np.array_equal(oracle.from_byte_packed(b.reshape(1, 256, 1))[0, :, 0], z["alpha_from"])
    # every 8-bit colour code survives the round trip (the reference patches its table for exactly this)
    assert np.array_equal(oracle.to_byte_packed(back, 4)[0, :, :3], np.stack([b, b, b], axis=-1))


def test_tvl1_primitives_against_cv2(oracle):
    """Dual TV-L1 itself cannot be pinned (no OpenCV build with it here: parity unpinned, see oracle/tvl1.c); the three
    OpenCV primitives it is assembled from can: bicubic remap and the 5x5 median bit-exactly, the bilinear resize to
    2 ulp (cv2's SIMD path contracts the two products differently)."""
    z = np.load(os.path.join(G, "tvl1_primitives_cv2.npz"))
    assert np.array_equal(oracle.remap_cubic(z["img"], z["mapx"], z["mapy"]), z["remap_cubic"])
    assert np.array_equal(oracle.median5(z["img"]), z["median5"])
    h, w = z["noise"].shape
    dh, dw = z["down08"].shape
    assert (dh, dw) == (round(h * 0.8), round(w * 0.8))
    down = oracle.tvl1_resize(z["noise"], dw, dh, 1 / 0.8, 1 / 0.8)
    assert np.abs(down - z["down08"]).max() <= 3.1e-5      # 2 ulp at 255
    up = oracle.tvl1_resize(z["down08"], w, h, 1 / (w / dw), 1 / (h / dh))
    assert np.abs(up - z["up"]).max() <= 3.1e-5


def test_tvl1_oracle_recovers_a_translation(oracle, synth):
    """Sanity of the restated method as a whole: a (2.5, -1.5) px shift of a smooth texture comes back to ~0.01 px, the
    early exit is taken (fewer inner iterations than the 5 x 5 x 10 x 15 maximum), deterministic."""
    base = synth.gray(synth.texture(120, 160, seed=3))
    nxt = synth.shift_bilinear(base, 2.5, -1.5)
    flow, iters = oracle.tvl1(base, nxt)
    inner = flow[12:-12, 12:-12]
    assert abs(np.median(inner[..., 0]) - 2.5) < 0.02 and abs(np.median(inner[..., 1]) + 1.5) < 0.02
    assert np.abs(inner - np.array([2.5, -1.5], np.float32)).mean() < 0.05
    assert 0 < iters < 5 * 5 * 10 * 15
    flow2, iters2 = oracle.tvl1(base, nxt)
    assert iters2 == iters and np.array_equal(flow, flow2)


def test_tvl1_oracle_properties(oracle, synth):
    """More sanity of the unpinned restatement: identical frames give exactly zero flow and stop at once; the backward
    flow of a translated pair is close to the negated forward flow; a larger epsilon never runs more iterations."""
    base = synth.gray(synth.texture(96, 128, seed=9))
    flow, iters = oracle.tvl1(base, base)
    assert not flow.any() and iters == 5 * 5          # one inner iteration per warping, then the early exit
    nxt = synth.shift_bilinear(base, 1.5, 0.75)
    fwd, it_f = oracle.tvl1(base, nxt)
    bwd, _ = oracle.tvl1(nxt, base)
    inner = (slice(12, -12), slice(12, -12))
    assert np.abs(fwd[inner] + bwd[inner]).mean() < 0.1
    assert abs(np.median(fwd[inner][..., 0]) - 1.5) < 0.05 and abs(np.median(fwd[inner][..., 1]) - 0.75) < 0.05
    _, it_loose = oracle.tvl1(base, nxt, epsilon=0.05)
    assert it_loose <= it_f
