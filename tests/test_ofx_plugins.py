"""The three .ofx bundles through the OFX boundary, driven by the mini-host.  The describe-time half runs on a
CPU-only box (identity, parameters, clips, exports: SURVEY.md section 8b); the render half needs the GPU and
compares the output clip with the CPU oracle composed exactly like the reference's render actions."""
import importlib
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mh():
    return importlib.import_module("openfx-opencv_b200.minihost")


IDENT = {
    "VectorGenerator": ("net.sf.openfx.VectorGenerator", (1, 0), "VectorGeneratorOFX", "Time"),
    "inpaint": ("uk.org.bratwurstandhaggis:cvInpaint", (0, 5), "openCV Inpaint", "Draw"),
    "segment": ("uk.org.bratwurstandhaggis:cvPyrSegmentation", (0, 5), "openCV Segment", "Draw"),
}


@pytest.mark.parametrize("name", sorted(IDENT))
def test_bundle_identity_and_exports(mh, name):
    path = mh.bundle_path(name)
    assert os.path.exists(path) and os.path.exists(os.path.join(os.path.dirname(os.path.dirname(path)), "Info.plist"))
    out = subprocess.check_output(["nm", "-D", "--defined-only", path]).decode()
    exported = sorted(l.split()[-1].split("@")[0] for l in out.splitlines() if " T " in l)
    assert exported == ["OfxGetNumberOfPlugins", "OfxGetPlugin"]
    p = mh.Plugin(name)
    ident, version, label, group = IDENT[name]
    assert p.load_status == 0
    assert (p.identifier, p.version, p.api) == (ident, version, "OfxImageEffectPluginAPI")
    assert p.prop_string("OfxPropLabel") == label and p.prop_string("OfxImageEffectPluginPropGrouping") == group
    assert "OfxImageEffectContextFilter" in p.prop_strings("OfxImageEffectPropSupportedContexts")
    assert sorted(p.clips()) == ["Output", "Source"]
    assert p.prop_int("OfxImageEffectPropSupportsTiles") == 0
    assert p.action("OfxImageEffectActionGetRegionOfDefinition") == mh.STAT_REPLY_DEFAULT   # unhandled action
    p.close()


def test_inpaint_and_segment_parameters(mh):
    p = mh.Plugin("inpaint")
    assert p.prop_strings("OfxImageEffectPropSupportedPixelDepths") == ["OfxBitDepthByte"]
    par = p.params()
    for n in ("threshold1", "threshold2", "inpaintnoise"):
        assert par[n] == "OfxParamTypeDouble"
    assert par["method"] == "OfxParamTypeChoice"
    # inpaint.cpp:424-452: Radius 3 [1,10], Dilation 1 [1,5], noise 0 [0,1]
    for n, (d, lo, hi, label) in {"threshold1": (3, 1, 10, "Radius"), "threshold2": (1, 1, 5, "Dilation"), "inpaintnoise": (0, 0, 1, "Inpaint noise")}.items():
        assert p.param_prop(n, "OfxParamPropDefault")["double"] == d
        assert p.param_prop(n, "OfxParamPropDisplayMin")["double"] == lo and p.param_prop(n, "OfxParamPropDisplayMax")["double"] == hi
        assert p.param_prop(n, "OfxPropLabel")["string"] == label
    assert [p.param_prop("method", "OfxParamPropChoiceOption", i)["string"] for i in range(2)] == ["Telea", "Navier-Stokes"]
    assert p.param_prop("method", "OfxParamPropDefault")["int"] == 0
    assert p.clip_components("Source") == ["OfxImageComponentRGBA"]
    p.close()
    s = mh.Plugin("segment")
    assert s.param_prop("threshold1", "OfxParamPropDefault")["double"] == 250 and s.param_prop("threshold2", "OfxParamPropDefault")["double"] == 30
    assert s.param_prop("seeds", "OfxParamPropDefault")["int"] == 256
    s.close()


def test_vectorgenerator_descriptor(mh):
    p = mh.Plugin("VectorGenerator")
    assert p.prop_strings("OfxImageEffectPropSupportedPixelDepths") == ["OfxBitDepthFloat"]
    assert p.prop_string("OfxImageEffectPluginRenderThreadSafety") == "OfxImageEffectRenderFullySafe"
    assert p.prop_int("OfxImageEffectPropTemporalClipAccess") == 1 and p.prop_int("OfxImageEffectPropSupportsMultiResolution") == 1
    assert p.prop_int("OfxImageEffectPluginPropHostFrameThreading") == 0
    assert p.prop_string("OfxImageEffectPropCudaRenderSupported") == "true"
    assert p.clip_components("Source") == ["OfxImageComponentRGBA", "OfxImageComponentRGB", "OfxImageComponentAlpha"]
    assert p.clip_components("Output") == ["OfxImageComponentRGBA"]
    par = p.params()
    for i, n in enumerate(("rChannel", "gChannel", "bChannel", "aChannel")):
        assert par[n] == "OfxParamTypeChoice" and p.param_prop(n, "OfxParamPropDefault")["int"] == i + 1     # VectorGenerator.cpp:739-779
        assert [p.param_prop(n, "OfxParamPropChoiceOption", k)["string"] for k in range(5)] == ["0", "forward.u", "forward.v", "backward.u", "backward.v"]
    assert p.param_prop("levels", "OfxParamPropDefault")["int"] == 3 and p.param_prop("iterations", "OfxParamPropDefault")["int"] == 15
    assert p.param_prop("neighborhood", "OfxParamPropDefault")["int"] == 5 and p.param_prop("sigma", "OfxParamPropDefault")["double"] == 1.1
    assert p.param_prop("method", "OfxParamPropChoiceOption", 0)["string"] == "Farneback"
    assert p.param_prop("tau", "OfxParamPropDefault")["double"] == 0.25 and p.param_prop("warps", "OfxParamPropDefault")["int"] == 5
    assert p.create_instance() == 0
    # getFramesNeeded (VectorGenerator.cpp:675-695): defaults need both directions -> [t-1, t+1]
    assert p.frames_needed(10.0) == (0, (9.0, 11.0))
    for n in ("bChannel", "aChannel"):
        p.set_param(n, 0)
    assert p.frames_needed(10.0) == (0, (10.0, 11.0))
    p.set_param("rChannel", 3); p.set_param("gChannel", 0)
    assert p.frames_needed(10.0) == (0, (9.0, 10.0))
    p.set_param("rChannel", 0)
    assert p.frames_needed(10.0)[0] == mh.STAT_REPLY_DEFAULT
    # method change hides the Farneback controls (updateVisibility, VectorGenerator.cpp:642-662)
    p.set_param("method", 1)
    assert p.instance_changed("method") == 0
    assert p.param_prop("levels", "OfxParamPropSecret", instance=True)["int"] == 1
    assert p.param_prop("iterations", "OfxParamPropSecret", instance=True)["int"] == 0   # shared by both methods
    assert p.param_prop("tau", "OfxParamPropSecret", instance=True)["int"] == 0
    p.set_param("method", 0); p.instance_changed("method")
    assert p.param_prop("levels", "OfxParamPropSecret", instance=True)["int"] == 0
    assert p.param_prop("warps", "OfxParamPropSecret", instance=True)["int"] == 1
    assert p.destroy_instance() == 0
    p.close()


@pytest.mark.skipif(not os.path.isdir("/root/reference/openfx/include"), reason="reference headers not present")
def test_ofx_min_header_layout_matches_reference(tmp_path):
    """ofx_min.h restates the OFX ABI; prove it is layout-identical to the headers the reference builds against."""
    src = tmp_path / "layout.cpp"
    mine = os.path.join(ROOT, "openfx-opencv_b200", "ofx", "ofx_min.h")
    suites = {
        "OfxPropertySuiteV1": ["propSetPointer", "propSetString", "propSetDouble", "propSetInt", "propSetPointerN", "propSetStringN", "propSetDoubleN", "propSetIntN", "propGetPointer", "propGetString", "propGetDouble", "propGetInt", "propGetPointerN", "propGetStringN", "propGetDoubleN", "propGetIntN", "propReset", "propGetDimension"],
        "OfxImageEffectSuiteV1": ["getPropertySet", "getParamSet", "clipDefine", "clipGetHandle", "clipGetPropertySet", "clipGetImage", "clipReleaseImage", "clipGetRegionOfDefinition", "abort", "imageMemoryAlloc", "imageMemoryFree", "imageMemoryLock", "imageMemoryUnlock"],
        "OfxParameterSuiteV1": ["paramDefine", "paramGetHandle", "paramSetGetPropertySet", "paramGetPropertySet", "paramGetValue", "paramGetValueAtTime", "paramGetDerivative", "paramGetIntegral", "paramSetValue", "paramSetValueAtTime", "paramGetNumKeys", "paramGetKeyTime", "paramGetKeyIndex", "paramDeleteKey", "paramDeleteAllKeys", "paramCopy", "paramEditBegin", "paramEditEnd"],
        "OfxPlugin": ["pluginApi", "apiVersion", "pluginIdentifier", "pluginVersionMajor", "pluginVersionMinor", "setHost", "mainEntry"],
        "OfxHost": ["host", "fetchSuite"],
        "OfxRectI": ["x1", "y1", "x2", "y2"],
    }
    consts = ["kOfxActionLoad", "kOfxActionDescribe", "kOfxImageEffectActionRender", "kOfxImageEffectActionGetFramesNeeded", "kOfxImagePropData",
              "kOfxImagePropBounds", "kOfxImagePropRowBytes", "kOfxImageEffectPropRenderWindow", "kOfxImageEffectPropCudaEnabled",
              "kOfxImageEffectPropCudaRenderSupported", "kOfxParamPropChoiceOption", "kOfxImageEffectPropFrameRange", "kOfxPropInstanceData",
              "kOfxImageEffectPluginRenderThreadSafety", "kOfxImageEffectRenderFullySafe", "kOfxBitDepthFloat", "kOfxImageComponentRGBA"]
    code = ['#include <stddef.h>', '#include <string.h>', '#include <limits.h>', '#include <stdint.h>', 'namespace ref {', '#include "ofxCore.h"', '#include "ofxProperty.h"', '#include "ofxParam.h"', '#include "ofxImageEffect.h"', '}']
    ref_consts = ["static const char* ref_%s = %s;" % (c, c) for c in consts] + ["static const int ref_statfmt = kOfxStatErrImageFormat, ref_default = kOfxStatReplyDefault, ref_mem = kOfxStatErrMemory;"]
    undef = ["#undef " + c for c in consts]
    # the reference headers use include guards + macros: read the constants first, then drop every kOfx macro
    code += ref_consts
    text = open(mine).read()
    import re
    macros = sorted(set(re.findall(r"#define (kOfx\w+|OfxExport)", text)))
    code += ["#undef " + m for m in macros]
    code += ['namespace mine {', '#include "%s"' % mine, '}']
    for s, members in suites.items():
        code.append("static_assert(sizeof(ref::%s) == sizeof(mine::%s), \"size %s\");" % (s, s, s))
        for m in members:
            code.append("static_assert(offsetof(ref::%s, %s) == offsetof(mine::%s, %s), \"offset %s.%s\");" % (s, m, s, m, s, m))
    code.append("int main() { int bad = 0;")
    for c in consts:
        code.append("bad += strcmp(ref_%s, %s) != 0;" % (c, c))
    code.append("bad += ref_statfmt != kOfxStatErrImageFormat; bad += ref_default != kOfxStatReplyDefault; bad += ref_mem != kOfxStatErrMemory; return bad; }")
    src.write_text("\n".join(code))
    exe = tmp_path / "layout"
    subprocess.check_call(["g++", "-std=gnu++14", "-w", "-DOFX_EXTENSIONS_RESOLVE", "-I/root/reference/openfx/include", str(src), "-o", str(exe)])
    assert subprocess.call([str(exe)]) == 0


# ------------------------------------------------------------------------------------------------ GPU half
def _rgba_with_holes(synth, h, w, seed, frac):
    rgb = np.maximum(synth.texture(h, w, seed), 1)   # keep every known pixel's gray >= 1 (mask rule inpaint.cpp:305-307)
    hole = synth.iid_mask(h, w, seed + 1, frac) > 0
    rgba = np.dstack([rgb, np.full((h, w), 200, np.uint8)])
    rgba[hole, :3] = 0
    return np.ascontiguousarray(rgba), hole


@pytest.mark.gpu
@pytest.mark.parametrize("method", [0, 1])
@pytest.mark.parametrize("flip", [False, True])
def test_inpaint_plugin_render(mh, oracle, synth, method, flip):
    h, w = 72, 96
    rgba, hole = _rgba_with_holes(synth, h, w, 3, 0.04)
    p = mh.Plugin("inpaint")
    assert p.create_instance() == 0
    p.set_param("method", method)
    dst = np.zeros_like(rgba)
    p.set_image("Source", 0, rgba, flip_rows=flip)
    p.set_image("Output", 0, dst, flip_rows=flip)
    assert p.render(0, (0, 0, w, h)) == 0
    assert p.images_outstanding() == 0
    # reference composition: mask = gray==0, one 3x3 dilation (Dilation default 1), inpaint radius 3, alpha 255
    mask = (synth.gray(rgba[..., :3]) == 0)
    pad = np.pad(mask, 1)
    dil = np.zeros_like(mask)
    for dy in range(3):
        for dx in range(3):
            dil |= pad[dy:dy + h, dx:dx + w]
    src_rgb = rgba[..., :3] if not flip else rgba[::-1, :, :3]
    dil_m = dil if not flip else dil[::-1]
    ref = oracle.inpaint(np.ascontiguousarray(src_rgb), (dil_m * 255).astype(np.uint8), 3, 1 if method == 0 else 0)
    got = dst if not flip else dst[::-1]
    assert np.array_equal(got[..., :3], ref) and (got[..., 3] == 255).all()
    # noise parameter: only hole pixels with x % 4 == 0 may change, by at most 5 levels
    p.set_param("inpaintnoise", 1.0)
    dst2 = np.zeros_like(rgba)
    p.clear_images("Output"); p.set_image("Output", 0, dst2, flip_rows=flip)
    assert p.render(0, (0, 0, w, h)) == 0
    diff = np.abs(dst2.astype(int) - dst.astype(int)).max(axis=2)
    assert diff.max() <= 5 and not diff[~dil].any() and not diff[:, np.arange(w) % 4 != 0].any()
    # missing source image -> kOfxStatFailed, and nothing leaks
    p.clear_images("Source")
    assert p.render(0, (0, 0, w, h)) == mh.STAT_FAILED and p.images_outstanding() == 0
    # wrong pixel depth -> kOfxStatErrImageFormat
    p.set_image("Source", 0, rgba.astype(np.float32))
    assert p.render(0, (0, 0, w, h)) == mh.STAT_ERR_IMAGE_FORMAT
    p.close()


@pytest.mark.gpu
def test_segment_plugin_render(mh, oracle, synth, ctx):
    h, w = 90, 120
    rgb = synth.texture(h, w, 8)
    rgba = np.ascontiguousarray(np.dstack([rgb, np.full((h, w), 255, np.uint8)]))
    p = mh.Plugin("segment")
    assert p.create_instance() == 0
    p.set_param("seeds", 12)
    dst = np.zeros_like(rgba)
    p.set_image("Source", 0, rgba); p.set_image("Output", 0, dst)
    assert p.render(0, (0, 0, w, h)) == 0 and p.images_outstanding() == 0
    gx = 3; gy = 4   # round(sqrt(12)) = 3 columns, ceil(12/3) = 4 rows
    mk = ctx.seed_grid(w, h, gx, gy, 2)
    lab, _ = oracle.watershed(rgb, mk)
    assert (dst[..., 3] == 255).all() and (dst[lab == -1][:, :3] == 0).all()
    for l in range(1, gx * gy + 1):
        sel = lab == l
        cols = np.unique(dst[sel][:, :3], axis=0)
        assert len(cols) == 1                                     # one flat colour per watershed basin
        assert np.abs(cols[0].astype(float) - rgb[sel].mean(axis=0)).max() <= 0.51
    p.close()


def _float_rgba(synth, oracle, gray):
    lin = oracle.srgb_from_byte_table()[gray]
    return np.ascontiguousarray(np.dstack([lin, lin, lin, np.ones_like(lin)]).astype(np.float32))


@pytest.mark.gpu
def test_vectorgenerator_plugin_render(mh, oracle, synth):
    h, w = 120, 160
    base = synth.gray(synth.texture(h, w, 21))
    frames = {t: _float_rgba(synth, oracle, synth.shift_bilinear(base, 1.5 * t, -0.75 * t)) for t in (4, 5, 6)}
    g = {t: oracle.luma_srgb_gray8(f) for t, f in frames.items()}
    p = mh.Plugin("VectorGenerator")
    assert p.create_instance() == 0
    for t, f in frames.items():
        p.set_image("Source", t, f)
    dst = np.full((h, w, 4), 7.0, np.float32)
    p.set_image("Output", 5, dst)
    p.set_image_props("Output", 5, scale=(0.5, 0.25))     # the host renders at this scale: the output image carries it
    assert p.render(5, (0, 0, w, h), scale=(0.5, 0.25)) == 0 and p.images_outstanding() == 0
    fwd = oracle.farneback(g[5], g[6])
    bwd = oracle.farneback(g[5], g[4])
    ref = np.dstack([fwd[..., 0].astype(np.float64) / 0.5, fwd[..., 1].astype(np.float64) / 0.25,
                     bwd[..., 0].astype(np.float64) / 0.5, bwd[..., 1].astype(np.float64) / 0.25]).astype(np.float32)
    d = np.abs(dst - ref).max(axis=2)
    assert d.mean() <= 1e-3 and (d > 1e-2).mean() <= 1e-3
    # channel routing + "0" channels + RGB and Alpha sources
    p.set_param("rChannel", 2); p.set_param("gChannel", 0); p.set_param("bChannel", 1); p.set_param("aChannel", 0)
    p.set_param("levels", 2); p.set_param("iterations", 4)
    dst2 = np.full((h, w, 4), 7.0, np.float32)
    p.clear_images("Output"); p.set_image("Output", 5, dst2)
    p.clear_images("Source")
    p.set_image("Source", 5, np.ascontiguousarray(frames[5][..., :3]))
    p.set_image("Source", 6, np.ascontiguousarray(frames[6][..., :3]))
    assert p.render(5, (0, 0, w, h)) == 0
    f2 = oracle.farneback(g[5], g[6], levels=2, iters=4)
    assert np.abs(dst2[..., 0] - f2[..., 1]).mean() <= 1e-3 and np.abs(dst2[..., 2] - f2[..., 0]).mean() <= 1e-3
    assert not dst2[..., 1].any() and not dst2[..., 3].any()
    # frame t+1 missing -> kOfxStatFailed (VectorGenerator.cpp:563-566)
    p.clear_images("Source"); p.set_image("Source", 5, frames[5])
    assert p.render(5, (0, 0, w, h)) == mh.STAT_FAILED and p.images_outstanding() == 0
    # Dual TV-L1 (VectorGenerator.cpp:436-492) with the plugin's controls: forward flow into R,G
    p.set_image("Source", 6, frames[6]); p.set_param("method", 1)
    p.set_param("rChannel", 1); p.set_param("gChannel", 2); p.set_param("bChannel", 0); p.set_param("aChannel", 0)
    p.set_param("iterations", 6); p.set_param("warps", 2); p.set_param("nScales", 3); p.set_param("epsilon", 0.02)
    dst3 = np.full((h, w, 4), 7.0, np.float32)
    p.clear_images("Output"); p.set_image("Output", 5, dst3)
    assert p.render(5, (0, 0, w, h)) == 0 and p.images_outstanding() == 0
    tv, _ = oracle.tvl1(g[5], g[6], inner=6, warps=2, nscales=3, epsilon=0.02)
    assert np.array_equal(dst3[..., 0], tv[..., 0]) and np.array_equal(dst3[..., 1], tv[..., 1])
    assert not dst3[..., 2].any() and not dst3[..., 3].any()
    p.close()


@pytest.mark.gpu
def test_vectorgenerator_cuda_render_handoff(mh, oracle, synth, ctx):
    """OfxImageEffectPropCudaEnabled=1: the clip images are device pointers, no staging copy."""
    h, w = 96, 128
    base = synth.gray(synth.texture(h, w, 33))
    f0 = _float_rgba(synth, oracle, base)
    f1 = _float_rgba(synth, oracle, synth.shift_bilinear(base, 2.0, 1.0))
    d0, d1, dout = ctx.to_device(f0), ctx.to_device(f1), ctx.alloc(w * h * 16)
    p = mh.Plugin("VectorGenerator")
    assert p.create_instance() == 0
    p.set_param("bChannel", 0); p.set_param("aChannel", 0)
    p.set_device_image("Source", 1, d0.ptr, w, h, 4, np.float32)
    p.set_device_image("Source", 2, d1.ptr, w, h, 4, np.float32)
    p.set_device_image("Output", 1, dout.ptr, w, h, 4, np.float32)
    assert p.render(1, (0, 0, w, h), cuda_enabled=1) == 0
    got = dout.download((h, w, 4), np.float32)
    ref = oracle.farneback(oracle.luma_srgb_gray8(f0), oracle.luma_srgb_gray8(f1))
    d = np.abs(got[..., :2] - ref).max(axis=2)
    assert d.mean() <= 1e-3 and not got[..., 2:].any()
    p.close()


@pytest.mark.gpu
@pytest.mark.parametrize("flip", [False, True])
def test_vectorgenerator_plugin_large_host_images(mh, pkg, ctx, oracle, synth, flip):
    """1080p host-memory clips take the chunked staging path of the glue (worker threads copy row chunks into pinned
    memory while the H2D copies run; the result is scattered by the workers): the render must equal the C ABI called
    directly on the same frames, with bottom-up and top-down (negative rowBytes) host images."""
    h, w = 1080, 1920
    base = synth.gray(synth.texture(h, w, 31))
    gray = {t: synth.shift_bilinear(base, 1.5 * t, -0.75 * t) for t in (0, 1)}
    frames = {t: _float_rgba(synth, oracle, g) for t, g in gray.items()}
    p = mh.Plugin("VectorGenerator")
    assert p.create_instance() == 0
    for n, v in (("rChannel", 1), ("gChannel", 2), ("bChannel", 0), ("aChannel", 0), ("levels", 2), ("iterations", 3)):
        p.set_param(n, v)
    for t, f in frames.items():
        p.set_image("Source", t, f, flip_rows=flip)
    dst = np.full((h, w, 4), 7.0, np.float32)
    p.set_image("Output", 0, dst, flip_rows=flip)
    assert p.render(0, (0, 0, w, h)) == 0 and p.images_outstanding() == 0
    g0 = ctx.rgba32f_to_srgb_gray8(frames[0] if not flip else frames[0][::-1])
    g1 = ctx.rgba32f_to_srgb_gray8(frames[1] if not flip else frames[1][::-1])
    assert np.array_equal(g0, gray[0] if not flip else gray[0][::-1])          # the staging conversion round-trips the bytes
    ref = ctx.farneback(g0, g1, pkg.FbParams(levels=2, iterations=3))
    got = dst if not flip else dst[::-1]
    assert np.array_equal(got[..., 0], ref[..., 0]) and np.array_equal(got[..., 1], ref[..., 1])
    assert not got[..., 2].any() and not got[..., 3].any()
    p.close()


@pytest.mark.gpu
@pytest.mark.parametrize("flip", [False, True])
def test_inpaint_plugin_large_host_images(mh, pkg, ctx, synth, flip):
    """A 1500x1600 RGBA8 clip (9.6 MB) goes through the chunked staging path; sub-window render with an offset origin.
    The render equals the C ABI composition called directly (split + dilate, inpaint, merge)."""
    h, w = 1500, 1600
    rgba, hole = _rgba_with_holes(synth, h, w, 5, 0.01)
    p = mh.Plugin("inpaint")
    assert p.create_instance() == 0
    dst = np.zeros_like(rgba)
    p.set_image("Source", 0, rgba, flip_rows=flip)
    p.set_image("Output", 0, dst, flip_rows=flip)
    assert p.render(0, (0, 0, w, h)) == 0 and p.images_outstanding() == 0
    src = rgba if not flip else rgba[::-1]
    rgb, mask = ctx.rgba8_to_rgb8_mask(np.ascontiguousarray(src), 1)
    ref = ctx.inpaint(rgb, mask, 3, pkg.INPAINT_TELEA)
    got = dst if not flip else dst[::-1]
    assert np.array_equal(got[..., :3], ref) and (got[..., 3] == 255).all()
    p.close()


def _bundle_lib(mh, name):
    """The private libofxcv_b200.so of a bundle (already loaded by the plugin; CDLL of the same path is the same instance)."""
    import ctypes as C
    import os
    L = C.CDLL(os.path.join(os.path.dirname(mh.bundle_path(name)), "libofxcv_b200.so"))
    L.ofxcv_transfer_stats.restype = None
    L.ofxcv_transfer_stats.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    return L


@pytest.mark.gpu
def test_vectorgenerator_render_argument_checks(mh, oracle, synth):
    """VectorGenerator.cpp:531-536: an output image whose render scale or field differs from the render arguments fails the
    render; a CUDA-enabled render handed host pointers is refused; abort leaves with kOfxStatOK and no image outstanding."""
    h, w = 64, 80
    base = synth.gray(synth.texture(h, w, 5))
    frames = {t: _float_rgba(synth, oracle, synth.shift_bilinear(base, 1.0 * t, 0.5 * t)) for t in (1, 2, 3)}
    p = mh.Plugin("VectorGenerator")
    assert p.create_instance() == 0
    for t, f in frames.items():
        p.set_image("Source", t, f)
    dst = np.zeros((h, w, 4), np.float32)
    p.set_image("Output", 2, dst)
    assert p.render(2, (0, 0, w, h)) == 0
    good = dst.copy()
    p.set_image_props("Output", 2, scale=(0.5, 0.5))
    assert p.render(2, (0, 0, w, h)) == mh.STAT_FAILED and p.images_outstanding() == 0      # scale (1,1) rendered into a half-scale image
    assert p.render(2, (0, 0, w, h), scale=(0.5, 0.5)) == 0
    p.set_image_props("Output", 2, scale=(1.0, 1.0), field="OfxFieldLower")
    assert p.render(2, (0, 0, w, h)) == mh.STAT_FAILED and p.images_outstanding() == 0      # field None rendered into a lower-field image
    p.set_image_props("Output", 2)
    assert p.render(2, (0, 0, w, h), cuda_enabled=1) == mh.STAT_ERR_UNSUPPORTED and p.images_outstanding() == 0   # host pointers are not device images
    p.set_abort(1)
    dst[...] = -5.0
    assert p.render(2, (0, 0, w, h)) == 0 and p.images_outstanding() == 0
    p.set_abort(0)
    assert p.render(2, (0, 0, w, h)) == 0 and np.array_equal(dst, good)
    p.close()


@pytest.mark.gpu
def test_vectorgenerator_keeps_staged_frames_between_renders(mh, oracle, synth):
    """getFramesNeeded (VectorGenerator.cpp:675-695) makes render t+1 refetch two of render t's three frames: with images
    labelled by the host (kOfxImagePropUniqueIdentifier) the second render uploads ONE new frame, not three, and gives the
    same pixels as a host that does not label its images."""
    h, w = 90, 120
    base = synth.gray(synth.texture(h, w, 9))
    frames = {t: _float_rgba(synth, oracle, synth.shift_bilinear(base, 1.25 * t, -0.5 * t)) for t in range(1, 6)}
    L = _bundle_lib(mh, "VectorGenerator")
    frame_bytes = w * h * 16

    def h2d():
        import ctypes as C
        a, b = C.c_uint64(0), C.c_uint64(0)
        L.ofxcv_transfer_stats(C.byref(a), C.byref(b))
        return int(a.value)

    outs = {}
    for labelled in (True, False):
        mh.Plugin.provide_unique_identifiers(labelled)
        p = mh.Plugin("VectorGenerator")
        assert p.create_instance() == 0
        for t, f in frames.items():
            p.set_image("Source", t, f)
        per_render = []
        for t in (2, 3, 4):
            dst = np.zeros((h, w, 4), np.float32)
            p.clear_images("Output"); p.set_image("Output", t, dst)
            before = h2d()
            assert p.render(t, (0, 0, w, h)) == 0 and p.images_outstanding() == 0
            per_render.append(h2d() - before)
            outs[(labelled, t)] = dst
        p.close()
        if labelled:
            assert per_render == [3 * frame_bytes, frame_bytes, frame_bytes], per_render
        else:
            assert per_render == [3 * frame_bytes] * 3, per_render
    mh.Plugin.provide_unique_identifiers(True)
    for t in (2, 3, 4):
        assert np.array_equal(outs[(True, t)], outs[(False, t)])
