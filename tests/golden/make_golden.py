"""Generates the committed golden vectors in tests/golden/ from the REFERENCE arithmetic available in the build
container: OpenCV 4.13 (cv2, the library the reference plugins call) and the reference's own
SupportExt/ofxsLut.cpp (compiled by `make -C oracle ref` into oracle/_ref/).  Run here only; the GPU box and the
tests never need cv2 or /root/reference — they read the .npz files.

    python tests/golden/make_golden.py
"""
import ctypes as C
import importlib
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
synth = importlib.import_module("openfx-opencv_b200.synth")


def farneback():
    out = {}
    for name, (h, w, levels, iters, n, sig) in {
        "a": (96, 128, 1, 3, 5, 1.1),
        "b": (120, 160, 3, 15, 5, 1.1),
        "c": (64, 80, 1, 2, 7, 1.5),
    }.items():
        prev, nxt = synth.flow_pair(h, w, seed=3)
        flow = cv2.calcOpticalFlowFarneback(prev, nxt, None, 0.5, levels, 3, iters, n, sig, 0)
        out[name + "_prev"], out[name + "_next"], out[name + "_flow"] = prev, nxt, flow
        out[name + "_params"] = np.array([levels, iters, n, sig], np.float64)
    np.savez_compressed(os.path.join(HERE, "farneback_cv2.npz"), **out)


def inpaint():
    out = {}
    h, w = 40, 56
    rgb = synth.texture(h, w, 1)
    masks = {"iid10": synth.iid_mask(h, w, 3, 0.10), "iid30": synth.iid_mask(h, w, 4, 0.30), "blobs": synth.blob_mask(h, w, 5, 5, 7)}
    border = np.zeros((h, w), np.uint8); border[0:3, :] = 255; border[:, w - 2:] = 255
    masks["border"] = border
    out["rgb"] = rgb
    for mname, m in masks.items():
        out["mask_" + mname] = m
        for cn in (1, 3):
            img = rgb if cn == 3 else synth.gray(rgb)
            for r in (1.4, 3, 5):
                for meth, flag in (("telea", cv2.INPAINT_TELEA), ("ns", cv2.INPAINT_NS)):
                    out["out_%s_c%d_r%s_%s" % (mname, cn, r, meth)] = cv2.inpaint(img, m, r, flag)
    np.savez_compressed(os.path.join(HERE, "inpaint_cv2.npz"), **out)


def inpaint_shapes():
    """Masks whose fill order differs in kind (synth.shape_masks): chains laid out back to back, interleaved chains,
    holes on the image's border ring (the clamped-index reads of cv::inpaint)."""
    out = {}
    h, w = 96, 128
    rgb = synth.texture(h, w, 9)
    out["rgb"] = rgb
    for mname, m in synth.shape_masks(h, w).items():
        out["mask_" + mname] = m
        for cn in (1, 3):
            img = rgb if cn == 3 else synth.gray(rgb)
            for r in (3, 4):
                for meth, flag in (("telea", cv2.INPAINT_TELEA), ("ns", cv2.INPAINT_NS)):
                    out["out_%s_c%d_r%d_%s" % (mname, cn, r, meth)] = cv2.inpaint(img, m, r, flag)
    np.savez_compressed(os.path.join(HERE, "inpaint_shapes_cv2.npz"), **out)


def watershed():
    out = {}
    for name, (h, w, n, seed) in {"a": (120, 160, 9, 5), "b": (64, 64, 4, 2)}.items():
        img = synth.texture(h, w, seed + 10)
        mk = synth.seed_markers(h, w, n, seed)
        lab = mk.copy()
        cv2.watershed(img, lab)
        out[name + "_img"], out[name + "_markers"], out[name + "_labels"] = img, mk, lab
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (90, 121, 3), dtype=np.uint8)
    mk = np.zeros((90, 121), np.int32); mk[10:14, 10:14] = 1; mk[60:64, 100:104] = 2; mk[40, 50] = 3
    lab = mk.copy(); cv2.watershed(img, lab)
    out["noise_img"], out["noise_markers"], out["noise_labels"] = img, mk, lab
    np.savez_compressed(os.path.join(HERE, "watershed_cv2.npz"), **out)


def tvl1_primitives():
    """The three cv2 primitives OpenCV's Dual TV-L1 is assembled from, on float images (the method itself is not in
    cv2 4.13's main modules, so only its building blocks can be pinned)."""
    rng = np.random.default_rng(11)
    h, w = 61, 97
    img = cv2.GaussianBlur((rng.random((h, w), dtype=np.float32) * 255).astype(np.float32), (0, 0), 1.5)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    mx = (xx + rng.standard_normal((h, w)).astype(np.float32) * 3).astype(np.float32)
    my = (yy + rng.standard_normal((h, w)).astype(np.float32) * 3).astype(np.float32)
    mx[0, :5] = -10; my[1, :5] = 200; mx[2, 0] = -1.2; mx[3, 0] = 96.7; mx[4, 0] = -2.99; mx[5, 0] = -3.0; mx[6, 0] = 97.0; mx[7, 0] = 98.0
    out = {"img": img, "mapx": mx, "mapy": my,
           "remap_cubic": cv2.remap(img, mx, my, cv2.INTER_CUBIC),
           "median5": cv2.medianBlur(img, 5)}
    noise = (rng.random((h, w), dtype=np.float32) * 255).astype(np.float32)
    down = cv2.resize(noise, None, fx=0.8, fy=0.8, interpolation=cv2.INTER_LINEAR)
    out["noise"], out["down08"] = noise, down
    out["up"] = cv2.resize(down, (w, h), interpolation=cv2.INTER_LINEAR)
    np.savez_compressed(os.path.join(HERE, "tvl1_primitives_cv2.npz"), **out)


def lut():
    so = os.path.join(ROOT, "oracle", "_ref", "libofxs_lut_ref.so")
    L = C.CDLL(so)
    L.ref_srgb_to_byte.restype = C.c_ubyte; L.ref_srgb_to_byte.argtypes = [C.c_float]
    L.ref_srgb_from_byte.restype = C.c_float; L.ref_srgb_from_byte.argtypes = [C.c_ubyte]
    L.ref_luma_to_byte.restype = C.c_ubyte; L.ref_luma_to_byte.argtypes = [C.c_float] * 3
    to = np.empty(0x10000, np.uint8)
    for i in range(0x10000):
        f = np.array([(i << 16) | 0x8000], np.uint32).view(np.float32)[0]
        to[i] = L.ref_srgb_to_byte(C.c_float(f))
    fr = np.array([L.ref_srgb_from_byte(b) for b in range(256)], np.float32)
    rng = np.random.default_rng(0)
    rgb = (rng.random((4096, 3), dtype=np.float32) * 1.2 - 0.1).astype(np.float32)
    luma = np.array([L.ref_luma_to_byte(*[C.c_float(v) for v in p]) for p in rgb], np.uint8)
    # alpha quantisers of the packed conversions (floatToInt<256> / intToFloat<256>): dense sweep incl. the rounding ties
    L.ref_alpha_to_byte.restype = C.c_ubyte; L.ref_alpha_to_byte.argtypes = [C.c_float]
    L.ref_alpha_from_byte.restype = C.c_float; L.ref_alpha_from_byte.argtypes = [C.c_ubyte]
    alpha_in = np.concatenate([np.linspace(-0.25, 1.25, 6001, dtype=np.float32),
                               ((np.arange(256, dtype=np.float32) + 0.5) / 255).astype(np.float32),
                               np.nextafter(((np.arange(256, dtype=np.float32) + 0.5) / 255).astype(np.float32), np.float32(0))])
    alpha_byte = np.array([L.ref_alpha_to_byte(C.c_float(v)) for v in alpha_in], np.uint8)
    alpha_from = np.array([L.ref_alpha_from_byte(b) for b in range(256)], np.float32)
    np.savez_compressed(os.path.join(HERE, "lut_srgb_ref.npz"), to_byte_by_hipart=to, from_byte=fr, luma_rgb=rgb, luma_byte=luma,
                        alpha_in=alpha_in, alpha_byte=alpha_byte, alpha_from=alpha_from)


if __name__ == "__main__":
    print("cv2", cv2.__version__)
    farneback(); inpaint(); inpaint_shapes(); watershed(); lut(); tvl1_primitives()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
