"""Seeded synthetic frames for the parity tests and the bench (numpy/scipy only — no cv2, no /root/reference).

Shapes/seeds follow SURVEY.md section 8d; the blur here is scipy's (the survey used cv2.GaussianBlur), which only
changes the texture, not the workload.
"""
import numpy as np


def texture(h, w, seed, sigma=3.0):
    """HxWx3 uint8 blurred-noise texture, min-max normalised."""
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(seed)
    a = rng.random((h, w, 3), dtype=np.float32)
    a = gaussian_filter(a, sigma=(sigma, sigma, 0), mode="reflect")
    a = (a - a.min()) / (a.max() - a.min()) * 255.0
    return a.astype(np.uint8)


def gray(rgb):
    """cv2.cvtColor(RGB2GRAY) fixed-point formula (verified in SURVEY.md A.2)."""
    r, g, b = (rgb[..., i].astype(np.int32) for i in range(3))
    return ((9798 * r + 19235 * g + 3735 * b + 16384) >> 15).astype(np.uint8)


def shift_bilinear(img, dx, dy):
    """next(x, y) = img(x - dx, y - dy), bilinear, replicate border (a global translation by (dx, dy))."""
    h, w = img.shape
    xs = np.arange(w, dtype=np.float64) - dx
    ys = np.arange(h, dtype=np.float64) - dy
    x0 = np.floor(xs).astype(int); fx = (xs - x0).astype(np.float32)
    y0 = np.floor(ys).astype(int); fy = (ys - y0).astype(np.float32)
    x0c, x1c = np.clip(x0, 0, w - 1), np.clip(x0 + 1, 0, w - 1)
    y0c, y1c = np.clip(y0, 0, h - 1), np.clip(y0 + 1, 0, h - 1)
    f = img.astype(np.float32)
    top = f[y0c][:, x0c] * (1 - fx) + f[y0c][:, x1c] * fx
    bot = f[y1c][:, x0c] * (1 - fx) + f[y1c][:, x1c] * fx
    out = top * (1 - fy)[:, None] + bot * fy[:, None]
    return np.ascontiguousarray(np.clip(np.rint(out), 0, 255).astype(np.uint8))


def flow_pair(h, w, seed=3, dx=2.5, dy=-1.5):
    g = gray(texture(h, w, seed))
    return g, shift_bilinear(g, dx, dy)


def iid_mask(h, w, seed, frac):
    rng = np.random.default_rng(seed)
    return ((rng.random((h, w)) < frac) * 255).astype(np.uint8)


def blob_mask(h, w, seed, nblobs=6, rmax=9):
    rng = np.random.default_rng(seed)
    m = np.zeros((h, w), np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(nblobs):
        cy, cx = rng.integers(0, h), rng.integers(0, w)
        ry, rx = rng.integers(2, rmax + 1), rng.integers(2, rmax + 1)
        m[((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0] = 255
    return m


def seed_markers(h, w, n, seed, r=2):
    """n (2r+1)^2 squares labelled 1..n at seeded positions (later squares overwrite), int32."""
    rng = np.random.default_rng(seed)
    ys = rng.integers(8, h - 8, n)
    xs = rng.integers(8, w - 8, n)
    mk = np.zeros((h, w), np.int32)
    for i, (y, x) in enumerate(zip(ys, xs)):
        mk[y - r:y + r + 1, x - r:x + r + 1] = i + 1
    return mk


def shape_masks(h, w):
    """Masks whose fill order differs in kind: chains laid out back to back (horizontal lines: the ready-queue fill),
    chains interleaved by the fill order (diagonal scratches, a thick bar: the in-order incremental fill)."""
    lines = np.zeros((h, w), np.uint8)
    lines[20:h - 20:12, 15:w - 15] = 255
    thick = np.zeros((h, w), np.uint8)
    thick[h // 2 - 4:h // 2 + 5, 10:w - 10] = 255
    diag = np.zeros((h, w), np.uint8)
    rng = np.random.default_rng(11)
    for _ in range(14):
        x0, y0, n, sl = int(rng.integers(5, w // 2)), int(rng.integers(10, h - 10)), int(rng.integers(40, w // 2 - 10)), rng.uniform(-0.8, 0.8)
        for t in range(n):
            yy = min(max(int(y0 + sl * t), 0), h - 2)
            diag[yy:yy + 2, x0 + t] = 255
    edge = np.zeros((h, w), np.uint8)   # holes on the image's border ring and next to it: the clamped-index reads of the CPU code
    edge[0, ::3] = 255
    edge[1, 1::4] = 255
    edge[h - 1, ::2] = 255
    edge[:, 0] = 255
    edge[2:h - 2:5, w - 1] = 255
    edge[2:h - 2:7, w - 2] = 255
    return {"lines": lines, "thick": thick, "diag": diag, "edge": edge}
