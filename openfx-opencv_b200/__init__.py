"""openfx-opencv_b200 — B200-native filter bodies behind openfx-opencv's OFX render actions.

This package is the thin Python host side above the C ABI of ``include/ofxcv_abi.h`` (``libofxcv_b200.so``,
hand-written sm_100a CUDA).  It exists for the parity tests, the bench and the sequence driver; the
reference-facing product is the set of ``.ofx`` bundles built from ``ofx/`` which call the same C ABI.

There is NO CPU fallback: importing works anywhere (so that the symbol table can be checked on a CPU box), but
every compute call raises unless the CUDA library is built and a GPU is present.

The package directory name contains a hyphen (it is the name the build contract asks for), so import it with
``importlib.import_module("openfx-opencv_b200")`` — ``tests/conftest.py`` and ``bench.py`` do exactly that.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libofxcv_b200.so")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include", "ofxcv_abi.h")
_LIB = None

OK = 0
INPAINT_NS, INPAINT_TELEA = 0, 1


class OfxcvError(RuntimeError):
    def __init__(self, status, where, detail=""):
        self.status = status
        msg = "%s failed: %s (%d)" % (where, _status_string(status), status)
        if detail:
            msg += " — " + detail
        super().__init__(msg)


class FbParams(C.Structure):
    """ofxcv_fb_params (defaults = the reference plugin's: VectorGenerator.cpp:391-399,:804-834)."""
    _fields_ = [("pyr_scale", C.c_double), ("levels", C.c_int), ("winsize", C.c_int), ("iterations", C.c_int),
                ("poly_n", C.c_int), ("poly_sigma", C.c_double), ("flags", C.c_int)]

    def __init__(self, pyr_scale=0.5, levels=3, winsize=3, iterations=15, poly_n=5, poly_sigma=1.1, flags=0):
        super().__init__(pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags)


class Tvl1Params(C.Structure):
    """ofxcv_tvl1_params (defaults = the reference plugin's, VectorGenerator.cpp:814,:874-929, + OpenCV's fixed ones)."""
    _fields_ = [("tau", C.c_double), ("lambda_", C.c_double), ("theta", C.c_double), ("epsilon", C.c_double),
                ("nscales", C.c_int), ("warps", C.c_int), ("iterations", C.c_int), ("outer_iterations", C.c_int),
                ("scale_step", C.c_double), ("median_filtering", C.c_int)]

    def __init__(self, tau=0.25, lambda_=0.15, theta=0.3, epsilon=0.01, nscales=5, warps=5, iterations=15,
                 outer_iterations=10, scale_step=0.8, median_filtering=5):
        super().__init__(tau, lambda_, theta, epsilon, nscales, warps, iterations, outer_iterations, scale_step,
                         median_filtering)


def build(force=False):
    """Compile libofxcv_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    import subprocess
    csrc = os.path.join(_HERE, "csrc")
    if force:
        subprocess.check_call(["make", "-C", csrc, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", csrc, "-j4"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    """The loaded C-ABI library.  Fails loudly when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libofxcv_b200.so is missing (%s): run `python -c 'import __graft_entry__ as g; g.build()'`; "
                           "there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i, sz, pd, d = C.c_void_p, C.c_int, C.c_size_t, C.c_ssize_t, C.c_double
    fbp = C.POINTER(FbParams)
    tvp = C.POINTER(Tvl1Params)
    sigs = {
        "ofxcv_abi_version": (i, []),
        "ofxcv_status_string": (C.c_char_p, [i]),
        "ofxcv_device_count": (i, []),
        "ofxcv_create": (vp, [i]),
        "ofxcv_destroy": (None, [vp]),
        "ofxcv_device": (i, [vp]),
        "ofxcv_ctx_stream": (vp, [vp]),
        "ofxcv_synchronize": (i, [vp]),
        "ofxcv_last_error": (C.c_char_p, [vp]),
        "ofxcv_launch_count": (C.c_uint64, [vp]),
        "ofxcv_kernel_time_ms": (C.c_uint64, [vp, i, C.POINTER(C.c_double)]),
        "ofxcv_kernel_time_enable": (None, [vp, i]),
        "ofxcv_prof_enable": (None, [vp, i]),
        "ofxcv_prof_report": (sz, [vp, C.c_char_p, sz]),
        "ofxcv_device_alloc": (vp, [vp, sz]),
        "ofxcv_device_free": (None, [vp, vp]),
        "ofxcv_pinned_alloc": (vp, [vp, sz]),
        "ofxcv_pinned_free": (None, [vp, vp]),
        "ofxcv_scratch_device": (vp, [vp, i, sz]),
        "ofxcv_scratch_pinned": (vp, [vp, i, sz]),
        "ofxcv_upload": (i, [vp, vp, vp, vp, sz]),
        "ofxcv_download": (i, [vp, vp, vp, vp, sz]),
        "ofxcv_device_copy": (i, [vp, vp, vp, vp, sz]),
        "ofxcv_memset": (i, [vp, vp, vp, i, sz]),
        "ofxcv_aux_stream": (vp, [vp]),
        "ofxcv_stream_wait": (i, [vp, vp, vp]),
        "ofxcv_stream_synchronize": (i, [vp, vp]),
        "ofxcv_current_device": (i, []),
        "ofxcv_pointer_device": (i, [vp]),
        "ofxcv_upload_rows": (i, [vp, vp, vp, vp, pd, sz, i]),
        "ofxcv_download_rows": (i, [vp, vp, vp, pd, vp, sz, i]),
        "ofxcv_transfer_stats": (None, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
        "ofxcv_set_abort_callback": (None, [vp, vp, vp]),
        "ofxcv_fb_default_params": (None, [fbp]),
        "ofxcv_farneback_scales": (i, [i, i, fbp]),
        "ofxcv_farneback_algorithmic_bytes": (d, [i, i, fbp]),
        "ofxcv_farneback_iter_bytes": (d, [i, i, fbp]),
        "ofxcv_farneback_workspace_bytes": (sz, [i, i, fbp]),
        "ofxcv_farneback_u8": (i, [vp, vp, vp, vp, pd, i, i, vp, pd, fbp]),
        "ofxcv_tvl1_default_params": (None, [tvp]),
        "ofxcv_tvl1_scales": (i, [i, i, tvp]),
        "ofxcv_tvl1_workspace_bytes": (sz, [i, i, tvp]),
        "ofxcv_tvl1_iter_bytes": (d, [i, i]),
        "ofxcv_tvl1_u8": (i, [vp, vp, vp, vp, pd, i, i, vp, pd, tvp]),
        "ofxcv_tvl1_iterations_run": (C.c_int64, [vp]),
        "ofxcv_farneback_u8_host": (i, [vp, vp, vp, pd, i, i, vp, pd, fbp]),
        "ofxcv_farneback_u8_keyed": (i, [vp, vp, vp, vp, pd, i, i, vp, pd, fbp, C.c_uint64, C.c_uint64]),
        "ofxcv_content_key_u8": (i, [vp, vp, vp, pd, i, i, C.POINTER(C.c_uint64)]),
        "ofxcv_farneback_set_lanes": (None, [vp, i]),
        "ofxcv_farneback_cache_clear": (None, [vp]),
        "ofxcv_farneback_cache_stats": (i, [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
        "ofxcv_farneback_sequence_u8": (i, [vp, vp, vp, pd, sz, i, i, i, vp, pd, sz, fbp]),
        "ofxcv_farneback_sequence_u8_host": (i, [vp, C.POINTER(vp), pd, i, i, i, C.POINTER(vp), pd, fbp]),
        "ofxcv_inpaint_u8": (i, [vp, vp, vp, pd, i, vp, pd, vp, pd, i, i, d, i]),
        "ofxcv_inpaint_u8_host": (i, [vp, vp, pd, i, vp, pd, vp, pd, i, i, d, i]),
        "ofxcv_inpaint_sequence_u8": (i, [vp, vp, vp, pd, i, vp, pd, vp, pd, i, i, i, d, i, i]),
        "ofxcv_inpaint_sequence_u8_host": (i, [vp, vp, pd, i, vp, pd, vp, pd, i, i, i, d, i, i]),
        "ofxcv_inpaint_workspace_bytes": (sz, [i, i, i]),
        "ofxcv_inpaint_set_fill_blocks": (None, [vp, i]),
        "ofxcv_inpaint_last_stats": (i, [vp, C.POINTER(C.c_int64)]),
        "ofxcv_inpaint_debug_maps": (i, [vp, i, i, vp, vp]),
        "ofxcv_watershed_u8c3": (i, [vp, vp, vp, pd, vp, pd, i, i]),
        "ofxcv_watershed_u8c3_host": (i, [vp, vp, pd, vp, pd, i, i]),
        "ofxcv_watershed_u8c3_batch": (i, [vp, vp, vp, pd, sz, vp, pd, sz, i, i, i]),
        "ofxcv_watershed_workspace_bytes": (sz, [i, i, i]),
        "ofxcv_watershed_last_stats": (i, [vp, C.POINTER(C.c_int64)]),
        "ofxcv_rgba32f_to_srgb_gray8": (i, [vp, vp, vp, pd, i, vp, pd, i, i]),
        "ofxcv_rgba32f_to_srgb8_packed": (i, [vp, vp, vp, pd, i, vp, pd, i, i, i]),
        "ofxcv_srgb8_packed_to_rgba32f": (i, [vp, vp, vp, pd, vp, pd, i, i, i]),
        "ofxcv_flow_to_rgba32f": (i, [vp, vp, vp, pd, vp, pd, i, i, C.POINTER(C.c_int), d, d]),
        "ofxcv_rgba8_to_rgb8_mask": (i, [vp, vp, vp, pd, vp, pd, vp, pd, i, i, i]),
        "ofxcv_rgb8_to_rgba8": (i, [vp, vp, vp, pd, vp, pd, i, i]),
        "ofxcv_rgb8_to_rgba8_noise": (i, [vp, vp, vp, pd, vp, pd, vp, pd, i, i, i, C.c_uint]),
        "ofxcv_seed_grid": (i, [vp, vp, vp, pd, i, i, i, i, i]),
        "ofxcv_labels_to_rgba8": (i, [vp, vp, vp, pd, vp, pd, vp, pd, i, i, i]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._ofxcv_sigs = sigs
    _LIB = L
    return L


def declared_symbols():
    """Every OFXCV_API function name declared in include/ofxcv_abi.h."""
    import re
    text = open(INCLUDE).read()
    return sorted(set(re.findall(r"OFXCV_API[^;(]*?\b(ofxcv_[a-z0-9_]+)\s*\(", text)))


def _status_string(st):
    try:
        return lib().ofxcv_status_string(st).decode()
    except Exception:
        return "status"


def _hp(a):
    return a.ctypes.data_as(C.c_void_p)


class DeviceBuffer:
    """A device allocation owned through the C ABI (no torch needed)."""

    def __init__(self, ctx, nbytes):
        self.ctx, self.nbytes = ctx, int(nbytes)
        self.ptr = lib().ofxcv_device_alloc(ctx.h, self.nbytes)
        if not self.ptr:
            raise OfxcvError(-3, "ofxcv_device_alloc", ctx.last_error())

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        ctx = self.ctx
        ctx._check(lib().ofxcv_upload(ctx.h, None, self.ptr, _hp(arr), arr.nbytes), "ofxcv_upload")
        ctx.synchronize()
        return self

    def download(self, shape, dtype):
        out = np.empty(shape, dtype)
        assert out.nbytes <= self.nbytes
        ctx = self.ctx
        ctx._check(lib().ofxcv_download(ctx.h, None, _hp(out), self.ptr, out.nbytes), "ofxcv_download")
        ctx.synchronize()
        return out

    def free(self):
        if self.ptr:
            lib().ofxcv_device_free(self.ctx.h, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """ofxcv_ctx: one per host thread; owns a stream, the workspaces and the pinned staging buffers."""

    def __init__(self, device=-1):
        L = lib()
        if L.ofxcv_device_count() <= 0:
            raise OfxcvError(-2, "ofxcv_create", "no CUDA device visible; this library has no CPU fallback")
        self.h = L.ofxcv_create(int(device))
        if not self.h:
            raise OfxcvError(-2, "ofxcv_create")

    def close(self):
        if getattr(self, "h", None):
            for ptr in getattr(self, "_pinned", []):
                lib().ofxcv_pinned_free(self.h, ptr)
            self._pinned = []
            lib().ofxcv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def last_error(self):
        return lib().ofxcv_last_error(self.h).decode()

    def _check(self, st, where):
        if st != OK:
            raise OfxcvError(st, where, self.last_error())

    def device(self):
        return int(lib().ofxcv_device(self.h))

    def synchronize(self):
        self._check(lib().ofxcv_synchronize(self.h), "ofxcv_synchronize")

    def stream(self):
        return lib().ofxcv_ctx_stream(self.h)

    def set_abort_callback(self, fn):
        """fn() -> truthy aborts the flow calls of this context between pyramid scales / pairs (None clears it)."""
        if fn is None:
            self._abort_cb = None
            lib().ofxcv_set_abort_callback(self.h, None, None)
            return
        self._abort_cb = C.CFUNCTYPE(C.c_int, C.c_void_p)(lambda _u: 1 if fn() else 0)
        lib().ofxcv_set_abort_callback(self.h, C.cast(self._abort_cb, C.c_void_p), None)

    def upload_rows(self, dst_d, arr2d_bytes_view, row_bytes, rows, src_stride):
        self._check(lib().ofxcv_upload_rows(self.h, None, dst_d, arr2d_bytes_view, src_stride, row_bytes, rows), "ofxcv_upload_rows")

    def download_rows(self, dst_ptr, dst_stride, src_d, row_bytes, rows):
        self._check(lib().ofxcv_download_rows(self.h, None, dst_ptr, dst_stride, src_d, row_bytes, rows), "ofxcv_download_rows")

    def launch_count(self):
        return int(lib().ofxcv_launch_count(self.h))

    def alloc(self, nbytes):
        return DeviceBuffer(self, nbytes)

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        return DeviceBuffer(self, max(arr.nbytes, 1)).upload(arr)

    def timing(self, enable):
        lib().ofxcv_kernel_time_enable(self.h, 1 if enable else 0)

    def kernel_time_ms(self, family):
        ms = C.c_double(0)
        n = lib().ofxcv_kernel_time_ms(self.h, family, C.byref(ms))
        return int(n), float(ms.value)

    def prof(self, enable):
        lib().ofxcv_prof_enable(self.h, 1 if enable else 0)

    def prof_report(self):
        """[(name, tag, launches, total_ms)] of the labelled launches since the last report."""
        buf = C.create_string_buffer(1 << 16)
        lib().ofxcv_prof_report(self.h, buf, len(buf))
        rows = []
        for line in buf.value.decode().splitlines():
            name, tag, n, ms = line.split()
            rows.append((name, int(tag), int(n), float(ms)))
        return rows

    # ---- host-buffer entry points (what the OFX glue calls for host-memory clips) ----------------------
    def pinned_array(self, shape, dtype):
        """A numpy array over page-locked host memory (ofxcv_pinned_alloc); freed with the context."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        ptr = lib().ofxcv_pinned_alloc(self.h, max(n, 1))
        if not ptr:
            raise OfxcvError(-3, "ofxcv_pinned_alloc", self.last_error())
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(ptr)
        buf = (C.c_uint8 * max(n, 1)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def farneback(self, prev, nxt, params=None, out=None):
        """prev, nxt: HxW uint8 (numpy, host).  Returns HxWx2 float32 flow (cv2.calcOpticalFlowFarneback layout)."""
        params = params or FbParams()
        prev = np.ascontiguousarray(prev, np.uint8)
        nxt = np.ascontiguousarray(nxt, np.uint8)
        if prev.ndim != 2 or prev.shape != nxt.shape:
            raise ValueError("prev/next must be equal-shape HxW uint8")
        h, w = prev.shape
        flow = out if out is not None else np.empty((h, w, 2), np.float32)
        assert flow.shape == (h, w, 2) and flow.dtype == np.float32 and flow.flags.c_contiguous
        st = lib().ofxcv_farneback_u8_host(self.h, _hp(prev), _hp(nxt), w, w, h, _hp(flow), w * 8, C.byref(params))
        self._check(st, "ofxcv_farneback_u8_host")
        return flow

    def inpaint(self, img, mask, radius, method):
        img = np.ascontiguousarray(img, np.uint8)
        mask = np.ascontiguousarray(mask, np.uint8)
        cn = 1 if img.ndim == 2 else img.shape[2]
        h, w = mask.shape
        if img.shape[:2] != (h, w) or cn not in (1, 3):
            raise ValueError("img must be HxW or HxWx3 uint8 matching the mask")
        out = np.empty_like(img)
        st = lib().ofxcv_inpaint_u8_host(self.h, _hp(img), w * cn, cn, _hp(mask), w, _hp(out), w * cn, w, h, float(radius), int(method))
        self._check(st, "ofxcv_inpaint_u8_host")
        return out

    def inpaint_sequence(self, imgs, masks, radius, method, frames_in_flight=0):
        """A clip of independent frames (host arrays), several in flight (ofxcv_inpaint_sequence_u8_host)."""
        imgs = [np.ascontiguousarray(a, np.uint8) for a in imgs]
        masks = [np.ascontiguousarray(m, np.uint8) for m in masks]
        if not imgs:
            return []
        if len(imgs) != len(masks):
            raise ValueError("one mask per frame")
        h, w = masks[0].shape
        cn = 1 if imgs[0].ndim == 2 else imgs[0].shape[2]
        if any(a.shape != imgs[0].shape for a in imgs) or any(m.shape != (h, w) for m in masks) or imgs[0].shape[:2] != (h, w):
            raise ValueError("all frames and masks of a clip must have the same size")
        outs = [np.empty_like(a) for a in imgs]
        n = len(imgs)
        arr = lambda xs: (C.c_void_p * n)(*[x.ctypes.data for x in xs])
        st = lib().ofxcv_inpaint_sequence_u8_host(self.h, arr(imgs), w * cn, cn, arr(masks), w, arr(outs), w * cn, w, h, n, float(radius),
                                                  int(method), int(frames_in_flight))
        self._check(st, "ofxcv_inpaint_sequence_u8_host")
        return outs

    def inpaint_sequence_dev(self, img_ptrs, cn, mask_ptrs, out_ptrs, w, h, radius, method, frames_in_flight=0, stream=None):
        n = len(img_ptrs)
        arr = lambda xs: (C.c_void_p * n)(*xs)
        st = lib().ofxcv_inpaint_sequence_u8(self.h, stream, arr(img_ptrs), w * cn, cn, arr(mask_ptrs), w, arr(out_ptrs), w * cn, w, h, n,
                                             float(radius), int(method), int(frames_in_flight))
        self._check(st, "ofxcv_inpaint_sequence_u8")

    def inpaint_set_fill_blocks(self, blocks_per_sm):
        lib().ofxcv_inpaint_set_fill_blocks(self.h, int(blocks_per_sm))

    def inpaint_stats(self):
        s = (C.c_int64 * 4)()
        self._check(lib().ofxcv_inpaint_last_stats(self.h, s), "ofxcv_inpaint_last_stats")
        return dict(hole_pixels=s[0], batches=s[1], rounds=s[2], launches=s[3])

    def inpaint_debug_maps(self, w, h):
        """(T map (h+2)x(w+2) float32, fill order hxw int32) of the last inpaint call — test hook."""
        t = np.empty((h + 2, w + 2), np.float32)
        order = np.empty((h, w), np.int32)
        self._check(lib().ofxcv_inpaint_debug_maps(self.h, w, h, _hp(t), _hp(order)), "ofxcv_inpaint_debug_maps")
        return t, order

    def watershed(self, rgb, markers):
        """rgb HxWx3 uint8, markers HxW int32 -> new label map (cv2.watershed semantics; input not modified)."""
        rgb = np.ascontiguousarray(rgb, np.uint8)
        m = np.array(markers, np.int32, copy=True, order="C")
        h, w = m.shape
        if rgb.shape != (h, w, 3):
            raise ValueError("rgb must be HxWx3 uint8 matching the markers")
        st = lib().ofxcv_watershed_u8c3_host(self.h, _hp(rgb), w * 3, _hp(m), w * 4, w, h)
        self._check(st, "ofxcv_watershed_u8c3_host")
        return m

    def watershed_sequence(self, rgbs, markers):
        """A clip of independent frames flooded concurrently (ofxcv_watershed_u8c3_batch): lists of HxWx3 uint8 frames and
        HxW int32 marker maps -> list of label maps."""
        n = len(rgbs)
        if n == 0:
            return []
        if len(markers) != n:
            raise ValueError("one marker map per frame")
        h, w = np.asarray(markers[0]).shape
        d_rgb, d_mk = self.alloc(w * h * 3 * n), self.alloc(w * h * 4 * n)
        try:
            for f in range(n):
                a = np.ascontiguousarray(rgbs[f], np.uint8)
                m = np.ascontiguousarray(markers[f], np.int32)
                if a.shape != (h, w, 3) or m.shape != (h, w):
                    raise ValueError("all frames and marker maps of a clip must have the same size")
                self._check(lib().ofxcv_upload(self.h, None, d_rgb.ptr + f * w * h * 3, _hp(a), w * h * 3), "ofxcv_upload")
                self._check(lib().ofxcv_upload(self.h, None, d_mk.ptr + f * w * h * 4, _hp(m), w * h * 4), "ofxcv_upload")
                self.synchronize()   # `a` / `m` may be temporaries
            self.watershed_dev(d_rgb.ptr, d_mk.ptr, w, h, n)
            out = d_mk.download((n, h, w), np.int32)
        finally:
            d_rgb.free(); d_mk.free()
        return [out[f] for f in range(n)]

    def watershed_stats(self):
        s = (C.c_int64 * 4)()
        self._check(lib().ofxcv_watershed_last_stats(self.h, s), "ofxcv_watershed_last_stats")
        return dict(pops=s[0], frames=s[1])

    # ---- device-pointer entry points (frames already resident in HBM) ------------------------------------
    def farneback_dev(self, prev_d, next_d, w, h, flow_d, params=None, stride=None, flow_stride=None, stream=None):
        params = params or FbParams()
        st = lib().ofxcv_farneback_u8(self.h, stream, prev_d, next_d, stride or w, w, h, flow_d, flow_stride or w * 8, C.byref(params))
        self._check(st, "ofxcv_farneback_u8")

    def tvl1_dev(self, prev_d, next_d, w, h, flow_d, params=None, stride=None, flow_stride=None, stream=None):
        params = params or Tvl1Params()
        st = lib().ofxcv_tvl1_u8(self.h, stream, prev_d, next_d, stride or w, w, h, flow_d, flow_stride or w * 8, C.byref(params))
        self._check(st, "ofxcv_tvl1_u8")

    def tvl1(self, prev, nxt, params=None):
        """Dual TV-L1 flow prev -> nxt (HxW uint8, host).  Returns (HxWx2 float32 flow, inner iterations run)."""
        prev = np.ascontiguousarray(prev, np.uint8)
        nxt = np.ascontiguousarray(nxt, np.uint8)
        if prev.ndim != 2 or prev.shape != nxt.shape:
            raise ValueError("prev/next must be equal-shape HxW uint8")
        h, w = prev.shape
        a, b, f = self.to_device(prev), self.to_device(nxt), self.alloc(w * h * 8)
        self.tvl1_dev(a.ptr, b.ptr, w, h, f.ptr, params)
        flow = f.download((h, w, 2), np.float32)
        return flow, int(lib().ofxcv_tvl1_iterations_run(self.h))

    def farneback_keyed_dev(self, prev_d, next_d, w, h, flow_d, key_prev, key_next, params=None, stream=None):
        params = params or FbParams()
        st = lib().ofxcv_farneback_u8_keyed(self.h, stream, prev_d, next_d, w, w, h, flow_d, w * 8, C.byref(params), key_prev, key_next)
        self._check(st, "ofxcv_farneback_u8_keyed")

    def farneback_set_lanes(self, lanes):
        lib().ofxcv_farneback_set_lanes(self.h, int(lanes))

    def content_key(self, img_d, w, h):
        k = C.c_uint64(0)
        self._check(lib().ofxcv_content_key_u8(self.h, None, img_d, w, w, h, C.byref(k)), "ofxcv_content_key_u8")
        return int(k.value)

    def farneback_cache_stats(self):
        b, h = C.c_uint64(0), C.c_uint64(0)
        self._check(lib().ofxcv_farneback_cache_stats(self.h, C.byref(b), C.byref(h)), "ofxcv_farneback_cache_stats")
        return int(b.value), int(h.value)

    def farneback_sequence_dev(self, frames_d, w, h, nframes, flows_d, params=None, stream=None):
        """frames_d: nframes contiguous HxW u8 frames on the device; flows_d: nframes-1 contiguous HxWx2 f32 fields."""
        params = params or FbParams()
        st = lib().ofxcv_farneback_sequence_u8(self.h, stream, frames_d, w, w * h, w, h, nframes, flows_d, w * 8, w * h * 8, C.byref(params))
        self._check(st, "ofxcv_farneback_sequence_u8")

    def farneback_sequence(self, frames, params=None, out=None):
        """frames: list of HxW uint8 host arrays (page-locked ones overlap copies with compute).  Returns the list of
        len(frames)-1 HxWx2 float32 flow fields (written into `out` when given)."""
        params = params or FbParams()
        frames = [np.ascontiguousarray(f, np.uint8) for f in frames]   # no copy for C-ordered (e.g. page-locked) arrays
        h, w = frames[0].shape
        n = len(frames)
        if any(f.shape != (h, w) for f in frames):
            raise ValueError("all frames must be HxW uint8 of one size")
        flows = out if out is not None else [np.empty((h, w, 2), np.float32) for _ in range(n - 1)]
        fp = (C.c_void_p * n)(*[f.ctypes.data for f in frames])
        assert all(f.shape == (h, w, 2) and f.dtype == np.float32 and f.flags.c_contiguous for f in flows)
        op = (C.c_void_p * (n - 1))(*[f.ctypes.data for f in flows])
        st = lib().ofxcv_farneback_sequence_u8_host(self.h, fp, w, w, h, n, op, w * 8, C.byref(params))
        self._check(st, "ofxcv_farneback_sequence_u8_host")
        return flows

    def inpaint_dev(self, img_d, cn, mask_d, out_d, w, h, radius, method, stream=None):
        st = lib().ofxcv_inpaint_u8(self.h, stream, img_d, w * cn, cn, mask_d, w, out_d, w * cn, w, h, float(radius), int(method))
        self._check(st, "ofxcv_inpaint_u8")

    def watershed_dev(self, rgb_d, markers_d, w, h, nframes=1, stream=None):
        st = lib().ofxcv_watershed_u8c3_batch(self.h, stream, rgb_d, w * 3, w * h * 3, markers_d, w * 4, w * h * 4, w, h, nframes)
        self._check(st, "ofxcv_watershed_u8c3_batch")

    # ---- staging conversions ---------------------------------------------------------------------------
    def rgba32f_to_srgb_gray8(self, img):
        img = np.ascontiguousarray(img, np.float32)
        h, w = img.shape[:2]
        nc = 1 if img.ndim == 2 else img.shape[2]
        src = self.to_device(img)
        dst = self.alloc(w * h)
        self._check(lib().ofxcv_rgba32f_to_srgb_gray8(self.h, None, src.ptr, w * nc * 4, nc, dst.ptr, w, w, h), "ofxcv_rgba32f_to_srgb_gray8")
        return dst.download((h, w), np.uint8)

    def rgba32f_to_srgb8_packed(self, img, dst_ncomp):
        img = np.ascontiguousarray(img, np.float32)
        h, w = img.shape[:2]
        sn = 1 if img.ndim == 2 else img.shape[2]
        src, dst = self.to_device(img), self.alloc(w * h * dst_ncomp)
        self._check(lib().ofxcv_rgba32f_to_srgb8_packed(self.h, None, src.ptr, w * sn * 4, sn, dst.ptr, w * dst_ncomp, dst_ncomp, w, h),
                    "ofxcv_rgba32f_to_srgb8_packed")
        return dst.download((h, w, dst_ncomp), np.uint8)

    def srgb8_packed_to_rgba32f(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        h, w, n = img.shape
        src, dst = self.to_device(img), self.alloc(w * h * n * 4)
        self._check(lib().ofxcv_srgb8_packed_to_rgba32f(self.h, None, src.ptr, w * n, dst.ptr, w * n * 4, n, w, h), "ofxcv_srgb8_packed_to_rgba32f")
        return dst.download((h, w, n), np.float32)

    def flow_to_rgba32f(self, flow, dst, chan_sel, scale_x=1.0, scale_y=1.0):
        flow = np.ascontiguousarray(flow, np.float32)
        dst = np.ascontiguousarray(dst, np.float32)
        h, w = flow.shape[:2]
        fd, dd = self.to_device(flow), self.to_device(dst)
        sel = (C.c_int * 4)(*chan_sel)
        self._check(lib().ofxcv_flow_to_rgba32f(self.h, None, fd.ptr, w * 8, dd.ptr, w * 16, w, h, sel, scale_x, scale_y), "ofxcv_flow_to_rgba32f")
        return dd.download((h, w, 4), np.float32)

    def rgba8_to_rgb8_mask(self, rgba, dilate_iterations=0):
        rgba = np.ascontiguousarray(rgba, np.uint8)
        h, w = rgba.shape[:2]
        src = self.to_device(rgba)
        rgb, mask = self.alloc(w * h * 3), self.alloc(w * h)
        self._check(lib().ofxcv_rgba8_to_rgb8_mask(self.h, None, src.ptr, w * 4, rgb.ptr, w * 3, mask.ptr, w, w, h, int(dilate_iterations)),
                    "ofxcv_rgba8_to_rgb8_mask")
        return rgb.download((h, w, 3), np.uint8), mask.download((h, w), np.uint8)

    def rgb8_to_rgba8(self, rgb):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        h, w = rgb.shape[:2]
        src, dst = self.to_device(rgb), self.alloc(w * h * 4)
        self._check(lib().ofxcv_rgb8_to_rgba8(self.h, None, src.ptr, w * 3, dst.ptr, w * 4, w, h), "ofxcv_rgb8_to_rgba8")
        return dst.download((h, w, 4), np.uint8)

    def rgb8_to_rgba8_noise(self, rgb, mask, noise_div, seed):
        """RGB -> RGBA with the inpaint plugin's optional noise on hole pixels whose x is a multiple of 4."""
        rgb = np.ascontiguousarray(rgb, np.uint8)
        mask = np.ascontiguousarray(mask, np.uint8)
        h, w = rgb.shape[:2]
        src, m, dst = self.to_device(rgb), self.to_device(mask), self.alloc(w * h * 4)
        self._check(lib().ofxcv_rgb8_to_rgba8_noise(self.h, None, src.ptr, w * 3, m.ptr, w, dst.ptr, w * 4, w, h, int(noise_div), int(seed)),
                    "ofxcv_rgb8_to_rgba8_noise")
        return dst.download((h, w, 4), np.uint8)

    def seed_grid(self, w, h, gx, gy, half):
        dst = self.alloc(w * h * 4)
        self._check(lib().ofxcv_seed_grid(self.h, None, dst.ptr, w * 4, w, h, gx, gy, half), "ofxcv_seed_grid")
        return dst.download((h, w), np.int32)

    def labels_to_rgba8(self, rgb, labels, nlabels):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        labels = np.ascontiguousarray(labels, np.int32)
        h, w = labels.shape
        r, l, o = self.to_device(rgb), self.to_device(labels), self.alloc(w * h * 4)
        self._check(lib().ofxcv_labels_to_rgba8(self.h, None, r.ptr, w * 3, l.ptr, w * 4, o.ptr, w * 4, w, h, int(nlabels)), "ofxcv_labels_to_rgba8")
        return o.download((h, w, 4), np.uint8)


def transfer_stats(library=None):
    """(host->device bytes, device->host bytes) moved by the staging helpers of a loaded libofxcv_b200.so so far."""
    L = library or lib()
    a, b = C.c_uint64(0), C.c_uint64(0)
    L.ofxcv_transfer_stats(C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


def farneback_algorithmic_bytes(w, h, params=None):
    params = params or FbParams()
    return float(lib().ofxcv_farneback_algorithmic_bytes(w, h, C.byref(params)))


def farneback_iter_bytes(w, h, params=None):
    params = params or FbParams()
    return float(lib().ofxcv_farneback_iter_bytes(w, h, C.byref(params)))


def farneback_scales(w, h, params=None):
    params = params or FbParams()
    return int(lib().ofxcv_farneback_scales(w, h, C.byref(params)))
