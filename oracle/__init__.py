"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end of oracle/libofxcv_oracle.so (the plain-C CPU restatement of the OpenCV bodies the
reference plugins call).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module; the product package never does.

Reference call sites restated (see each .c header):
  farneback  /root/reference/VectorGenerator/VectorGenerator.cpp:403
  inpaint    /root/reference/opencv2fx/inpaint/inpaint.cpp:311-318   (+ NS per BASELINE.json)
  watershed  BASELINE.json config 3 (replaces /root/reference/opencv2fx/segment/segment.cpp:296-302)
  lut        /root/reference/SupportExt/ofxsLut.h:171-190,:220-223,:447-486
  tvl1       /root/reference/VectorGenerator/VectorGenerator.cpp:436-492   (parity unpinned: see tvl1.c)
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

INPAINT_NS, INPAINT_TELEA = 0, 1


def build(force=False):
    so = os.path.join(_HERE, "libofxcv_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("farneback.c", "watershed.c", "inpaint.c", "lut.c", "tvl1.c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "libofxcv_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_watershed.restype = C.c_long
    return _LIB


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def farneback(prev, nxt, levels=3, iters=15, poly_n=5, poly_sigma=1.1, winsize=3, pyr_scale=0.5):
    prev = np.ascontiguousarray(prev, np.uint8)
    nxt = np.ascontiguousarray(nxt, np.uint8)
    h, w = prev.shape
    flow = np.empty((h, w, 2), np.float32)
    lib().orc_farneback(_p(prev), _p(nxt), C.c_int(w), C.c_int(w), C.c_int(h), _p(flow), C.c_double(pyr_scale),
                        C.c_int(levels), C.c_int(winsize), C.c_int(iters), C.c_int(poly_n), C.c_double(poly_sigma))
    return flow


def gaussian_kernel(n, sigma):
    k = np.empty(n, np.float32)
    lib().orc_gaussian_kernel(C.c_int(n), C.c_double(sigma), _p(k))
    return k


def gaussian_blur(img, ksz, sigma):
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    out = np.empty_like(img)
    lib().orc_gaussian_blur_f32(_p(img), _p(out), C.c_int(w), C.c_int(h), C.c_int(ksz), C.c_double(sigma))
    return out


def resize_linear(img, dw, dh):
    img = np.ascontiguousarray(img, np.float32)
    cn = 1 if img.ndim == 2 else img.shape[2]
    h, w = img.shape[:2]
    out = np.empty((dh, dw) if img.ndim == 2 else (dh, dw, cn), np.float32)
    lib().orc_resize_linear_f32(_p(img), C.c_int(w), C.c_int(h), _p(out), C.c_int(dw), C.c_int(dh), C.c_int(cn))
    return out


def polyexp(img, n=5, sigma=1.1):
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    out = np.empty((h, w, 5), np.float32)
    lib().orc_polyexp(_p(img), C.c_int(w), C.c_int(h), C.c_int(n), C.c_double(sigma), _p(out))
    return out


def polyexp_setup(n=5, sigma=1.1):
    g = np.empty(2 * n + 1, np.float32); xg = np.empty_like(g); xxg = np.empty_like(g)
    ig = np.empty(4, np.float64)
    lib().orc_polyexp_setup(C.c_int(n), C.c_double(sigma), _p(g), _p(xg), _p(xxg), _p(ig))
    return g, xg, xxg, ig


def update_matrices(R0, R1, flow):
    h, w = flow.shape[:2]
    M = np.empty((h, w, 5), np.float32)
    lib().orc_update_matrices(_p(np.ascontiguousarray(R0, np.float32)), _p(np.ascontiguousarray(R1, np.float32)),
                              _p(np.ascontiguousarray(flow, np.float32)), _p(M), C.c_int(w), C.c_int(h), C.c_int(0), C.c_int(h))
    return M


def update_flow_blur(R0, R1, flow, M, winsize=3, update=True):
    """In-place on copies; returns (flow, M)."""
    flow = np.array(flow, np.float32, copy=True, order="C")
    M = np.array(M, np.float32, copy=True, order="C")
    h, w = flow.shape[:2]
    lib().orc_update_flow_blur(_p(np.ascontiguousarray(R0, np.float32)), _p(np.ascontiguousarray(R1, np.float32)), _p(flow), _p(M),
                               C.c_int(w), C.c_int(h), C.c_int(winsize), C.c_int(1 if update else 0))
    return flow, M


class Tvl1Params(C.Structure):
    """Defaults = the plugin's (VectorGenerator.cpp:874-929; iterations :814) + OpenCV's fixed ones."""
    _fields_ = [("tau", C.c_double), ("lambda_", C.c_double), ("theta", C.c_double), ("epsilon", C.c_double),
                ("scale_step", C.c_double), ("nscales", C.c_int), ("warps", C.c_int), ("inner", C.c_int),
                ("outer", C.c_int), ("median", C.c_int)]

    def __init__(self, tau=0.25, lambda_=0.15, theta=0.3, epsilon=0.01, scale_step=0.8, nscales=5, warps=5, inner=15,
                 outer=10, median=5):
        super().__init__(tau, lambda_, theta, epsilon, scale_step, nscales, warps, inner, outer, median)


def tvl1(prev, nxt, **kw):
    """Dual TV-L1 flow prev -> nxt (HxW u8).  Returns (flow HxWx2 f32, inner iterations run)."""
    prev = np.ascontiguousarray(prev, np.uint8); nxt = np.ascontiguousarray(nxt, np.uint8)
    h, w = prev.shape
    flow = np.empty((h, w, 2), np.float32)
    par = Tvl1Params(**kw)
    it = lib().orc_tvl1(_p(prev), _p(nxt), C.c_int(w), C.c_int(w), C.c_int(h), _p(flow), C.byref(par))
    return flow, it


def tvl1_resize(img, dw, dh, scale_x, scale_y):
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    out = np.empty((dh, dw), np.float32)
    lib().orc_tvl1_resize_f32(_p(img), C.c_int(w), C.c_int(h), _p(out), C.c_int(dw), C.c_int(dh), C.c_double(scale_x), C.c_double(scale_y))
    return out


def remap_cubic(img, mapx, mapy):
    img = np.ascontiguousarray(img, np.float32)
    mapx = np.ascontiguousarray(mapx, np.float32); mapy = np.ascontiguousarray(mapy, np.float32)
    h, w = img.shape
    dh, dw = mapx.shape
    out = np.empty((dh, dw), np.float32)
    src = (C.c_void_p * 1)(img.ctypes.data); dst = (C.c_void_p * 1)(out.ctypes.data)
    lib().orc_remap_cubic_f32(src, C.c_int(1), C.c_int(w), C.c_int(h), _p(mapx), _p(mapy), dst, C.c_int(dw), C.c_int(dh))
    return out


def median5(img):
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape
    out = np.empty_like(img)
    lib().orc_median5_f32(_p(img), _p(out), C.c_int(w), C.c_int(h))
    return out


def watershed(img, markers):
    img = np.ascontiguousarray(img, np.uint8)
    m = np.array(markers, np.int32, copy=True, order="C")
    h, w = m.shape
    pops = lib().orc_watershed(_p(img), C.c_int(w * 3), _p(m), C.c_int(w), C.c_int(w), C.c_int(h))
    return m, int(pops)


def inpaint(img, mask, radius, method, want_debug=False):
    img = np.ascontiguousarray(img, np.uint8)
    mask = np.ascontiguousarray(mask, np.uint8)
    h, w = mask.shape
    cn = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty_like(img)
    t = np.empty((h + 2, w + 2), np.float32) if want_debug else None
    seq = np.empty((h, w), np.int32) if want_debug else None
    lib().orc_inpaint(_p(img), C.c_int(w * cn), _p(mask), C.c_int(w), _p(out), C.c_int(w * cn), C.c_int(w), C.c_int(h), C.c_int(cn),
                      C.c_double(radius), C.c_int(method), _p(t) if want_debug else None, _p(seq) if want_debug else None)
    return (out, t, seq) if want_debug else out


def srgb_tables():
    to = np.empty(0x10000, np.uint16)
    fr = np.empty(256, np.float32)
    lib().orc_srgb_tables(_p(to), _p(fr))
    return to, fr


def srgb_from_byte_table():
    return srgb_tables()[1]


def to_byte_packed(img, dst_ncomp):
    """Lut::to_byte_packed_nodither over whole rows: float (h,w[,n]) -> uint8 (h,w[,dst_ncomp])."""
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape[:2]
    sn = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty((h, w, dst_ncomp), np.uint8)
    lib().orc_to_byte_packed(_p(img), C.c_int(sn), _p(out), C.c_int(dst_ncomp), C.c_int(w), C.c_int(h))
    return out


def from_byte_packed(img):
    """Lut::from_byte_packed: uint8 (h,w,n) -> float32 (h,w,n)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w, n = img.shape
    out = np.empty((h, w, n), np.float32)
    lib().orc_from_byte_packed(_p(img), _p(out), C.c_int(n), C.c_int(w), C.c_int(h))
    return out


def luma_srgb_gray8(img):
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape[:2]
    nc = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty((h, w), np.uint8)
    lib().orc_luma_srgb_gray8(_p(img), C.c_int(nc), _p(out), C.c_int(w), C.c_int(h))
    return out
