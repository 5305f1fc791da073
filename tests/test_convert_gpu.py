"""Staging conversions either side of the filter bodies, against numpy restatements and the oracle LUT."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_luma_srgb_gray8(ctx, oracle):
    rng = np.random.default_rng(0)
    img = rng.random((37, 53, 4), dtype=np.float32) * 1.2 - 0.1   # includes <0 and >1
    img[0, 0] = [0, 0, 0, 1]; img[0, 1] = [1, 1, 1, 1]; img[0, 2] = [np.inf, 0, 0, 1]; img[0, 3] = [np.nan, 0, 0, 1]
    got = ctx.rgba32f_to_srgb_gray8(img)
    ref = oracle.luma_srgb_gray8(img)
    assert np.array_equal(got, ref)
    got3 = ctx.rgba32f_to_srgb_gray8(np.ascontiguousarray(img[..., :3]))
    assert np.array_equal(got3, ref)


def test_srgb_bytes_roundtrip(ctx, oracle):
    # every 8-bit code, converted to linear with the LUT's from-table, must come back unchanged
    lin = oracle.srgb_from_byte_table()
    img = np.zeros((1, 256, 4), np.float32)
    img[0, :, 0] = img[0, :, 1] = img[0, :, 2] = lin
    got = ctx.rgba32f_to_srgb_gray8(img)
    # luma of (v,v,v) in double = v*(0.2126+0.7152+0.0722) may differ from v by an ulp: allow +-1 code there
    assert np.abs(got[0].astype(int) - np.arange(256)).max() <= 1
    one = ctx.rgba32f_to_srgb_gray8(lin.reshape(1, 256))
    assert np.array_equal(one[0], np.arange(256, dtype=np.uint8))


def test_flow_to_rgba(ctx):
    rng = np.random.default_rng(1)
    flow = rng.standard_normal((19, 23, 2)).astype(np.float32) * 5
    dst = rng.random((19, 23, 4), dtype=np.float32)
    got = ctx.flow_to_rgba32f(flow, dst, [0, 1, -1, 0], 0.5, 0.25)
    ref = dst.copy()
    ref[..., 0] = (flow[..., 0].astype(np.float64) / 0.5).astype(np.float32)
    ref[..., 1] = (flow[..., 1].astype(np.float64) / 0.25).astype(np.float32)
    ref[..., 3] = ref[..., 0]
    assert np.array_equal(got, ref)


def test_rgba8_split_mask_dilate(ctx, synth):
    rng = np.random.default_rng(2)
    rgba = rng.integers(0, 256, (31, 45, 4), dtype=np.uint8)
    rgba[rng.random((31, 45)) < 0.05, :3] = 0
    rgba[3, 4, :3] = [1, 0, 3]   # gray rounds to 0 -> counts as hole, like cvCvtColor
    for n in (0, 1, 2):
        rgb, mask = ctx.rgba8_to_rgb8_mask(rgba, n)
        assert np.array_equal(rgb, rgba[..., :3])
        ref = (synth.gray(rgba[..., :3]) == 0)
        if n:
            pad = np.pad(ref, n)
            ref = np.zeros_like(ref)
            for dy in range(2 * n + 1):
                for dx in range(2 * n + 1):
                    ref |= pad[dy:dy + 31, dx:dx + 45]
        assert np.array_equal(mask, ref.astype(np.uint8) * 255)


def test_rgb_to_rgba_and_seed_grid_and_paint(ctx, oracle, synth):
    rgb = synth.texture(40, 50, 3)
    rgba = ctx.rgb8_to_rgba8(rgb)
    assert np.array_equal(rgba[..., :3], rgb) and (rgba[..., 3] == 255).all()
    mk = ctx.seed_grid(50, 40, 4, 3, 1)
    assert set(np.unique(mk)) == set(range(13))
    assert all((mk == l).sum() == 9 for l in range(1, 13))
    lab = ctx.watershed(rgb, mk)
    ref, _ = oracle.watershed(rgb, mk)
    assert np.array_equal(lab, ref)
    vis = ctx.labels_to_rgba8(rgb, lab, 12)
    for l in (1, 7, 12):
        sel = lab == l
        mean = np.floor(rgb[sel].astype(np.float64).mean(axis=0) + 0.5 + 1e-9)
        assert np.abs(vis[sel][0, :3].astype(int) - mean).max() <= 1
    assert (vis[lab == -1][:, :3] == 0).all()


def test_packed_srgb_conversions(ctx, oracle):
    """ofxcv_rgba32f_to_srgb8_packed / ofxcv_srgb8_packed_to_rgba32f against the oracle (itself pinned to the reference's
    compiled ofxsLut code): every component-count combination, out-of-range values, all 256 codes round trip."""
    rng = np.random.default_rng(3)
    img = (rng.random((29, 41, 4), dtype=np.float32) * 1.3 - 0.15).astype(np.float32)
    img[0, 0] = [0, 1, np.inf, 0.5]; img[0, 1] = [np.nan, -np.inf, 1e-8, 1.0]; img[0, 2] = [0.0031308, 0.04045, 0.5, 0.0019607844]
    for sn in (4, 3, 1):
        src = img[..., 0].copy() if sn == 1 else np.ascontiguousarray(img[..., :sn])
        for dn in (4, 3, 1):
            got = ctx.rgba32f_to_srgb8_packed(src, dn)
            assert np.array_equal(got, oracle.to_byte_packed(src, dn)), (sn, dn)
    codes = rng.integers(0, 256, (17, 23, 4), dtype=np.uint8)
    codes[0, :, :] = np.arange(23)[:, None] * 11
    for n in (4, 3, 1):
        src = np.ascontiguousarray(codes[..., :n])
        got = ctx.srgb8_packed_to_rgba32f(src)
        assert np.array_equal(got, oracle.from_byte_packed(src)), n
    b = np.arange(256, dtype=np.uint8)
    lin = ctx.srgb8_packed_to_rgba32f(np.stack([b, b, b, b], axis=-1).reshape(1, 256, 4))
    assert np.array_equal(ctx.rgba32f_to_srgb8_packed(lin, 4)[0], np.stack([b, b, b, b], axis=-1))


def _hash(x, y, s):
    m = 0xFFFFFFFF
    h = ((x * 0x9E3779B1) & m) ^ (((y + 0x7F4A7C15) & m) * 0x85EBCA77 & m) ^ (((s + 1) & m) * 0xC2B2AE3D & m)
    h ^= h >> 15; h = h * 0x2C1B3C6D & m; h ^= h >> 12; h = h * 0x297A2D39 & m; h ^= h >> 15
    return h


@pytest.mark.parametrize("shape", [(9, 1284), (5, 132), (3, 4), (6, 1023)])
def test_word_wide_staging_paths(ctx, oracle, synth, shape):
    """Widths that are multiples of 4 take the 4-pixels-per-thread kernels (several warps and blocks per row, ragged
    last warp); 1023 takes the byte kernels.  Both must give the numpy / oracle answer."""
    h, w = shape
    rng = np.random.default_rng(w)
    img = (rng.random((h, w, 4), dtype=np.float32) * 1.3 - 0.15).astype(np.float32)
    assert np.array_equal(ctx.rgba32f_to_srgb_gray8(img), oracle.luma_srgb_gray8(img))
    assert np.array_equal(ctx.rgba32f_to_srgb8_packed(img, 4), oracle.to_byte_packed(img, 4))
    codes = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    assert np.array_equal(ctx.srgb8_packed_to_rgba32f(codes), oracle.from_byte_packed(codes))
    rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    rgba[rng.random((h, w)) < 0.03, :3] = 0
    rgba[0, 0, :3] = 0; rgba[h - 1, w - 1, :3] = 0          # holes in the corners: dilation meets the border
    hole = synth.gray(rgba[..., :3]) == 0
    for n in (0, 1, 3, 4, 5):                                 # 5 is beyond the word-wide dilation
        rgb, mask = ctx.rgba8_to_rgb8_mask(rgba, n)
        assert np.array_equal(rgb, rgba[..., :3])
        ref = hole
        if n:
            pad = np.pad(hole, n)
            ref = np.zeros_like(hole)
            for dy in range(2 * n + 1):
                for dx in range(2 * n + 1):
                    ref |= pad[dy:dy + h, dx:dx + w]
        assert np.array_equal(mask, ref.astype(np.uint8) * 255), n
    rgb = np.ascontiguousarray(rgba[..., :3])
    out = ctx.rgb8_to_rgba8(rgb)
    assert np.array_equal(out[..., :3], rgb) and (out[..., 3] == 255).all()
    m8 = hole.astype(np.uint8) * 255
    for div in (0, 1, 2):
        got = ctx.rgb8_to_rgba8_noise(rgb, m8, div, 77)
        ref = np.concatenate([rgb, np.full((h, w, 1), 255, np.uint8)], axis=-1).astype(np.int64)
        if div:
            for y, x in zip(*np.nonzero(hole)):
                if x % 4 == 0:
                    d = int((_hash(int(x), int(y), 77) % 10 - 5) / div)   # C division truncates toward zero
                    ref[y, x, :3] = np.clip(ref[y, x, :3] + d, 0, 255)
        assert np.array_equal(got, ref.astype(np.uint8)), div
