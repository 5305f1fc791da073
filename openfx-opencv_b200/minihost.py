"""ctypes front-end of ofx/libofx_minihost.so — a tiny OFX host that loads one of the drop-in bundles and drives
Load / Describe / DescribeInContext / CreateInstance / Render the way a real host (Natron, Nuke, Resolve) does.
Used by the integration tests and by tools that render a sequence through the plugin boundary."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "ofx", "libofx_minihost.so")
BUNDLES = os.path.join(_HERE, "ofx", "bundles")
_LIB = None

STAT_OK, STAT_FAILED, STAT_ERR_UNSUPPORTED, STAT_REPLY_DEFAULT, STAT_ERR_IMAGE_FORMAT = 0, 1, 5, 14, 1000
DEPTH_BYTE, DEPTH_FLOAT = "OfxBitDepthByte", "OfxBitDepthFloat"
RGBA, RGB, ALPHA = "OfxImageComponentRGBA", "OfxImageComponentRGB", "OfxImageComponentAlpha"


def bundle_path(name):
    return os.path.join(BUNDLES, name + ".ofx.bundle", "Contents", "Linux-x86-64", name + ".ofx")


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libofx_minihost.so missing: run __graft_entry__.build()")
        L = C.CDLL(LIB_PATH)
        vp, i, d, s = C.c_void_p, C.c_int, C.c_double, C.c_char_p
        L.mh_load.restype = vp; L.mh_load.argtypes = [s, s, C.POINTER(i)]
        for f in ("mh_plugin_identifier", "mh_plugin_api"):
            getattr(L, f).restype = s; getattr(L, f).argtypes = [vp]
        L.mh_plugin_version.restype = i; L.mh_plugin_version.argtypes = [vp, i]
        for f in ("mh_create_instance", "mh_destroy_instance", "mh_param_count", "mh_clip_count", "mh_images_outstanding"):
            getattr(L, f).restype = i; getattr(L, f).argtypes = [vp]
        L.mh_unload.restype = None; L.mh_unload.argtypes = [vp]
        L.mh_effect_prop_string.restype = s; L.mh_effect_prop_string.argtypes = [vp, i, s, i]
        L.mh_effect_prop_int.restype = i; L.mh_effect_prop_int.argtypes = [vp, i, s, i, C.POINTER(i)]
        L.mh_param_name.restype = s; L.mh_param_name.argtypes = [vp, i]
        L.mh_param_type.restype = s; L.mh_param_type.argtypes = [vp, s]
        L.mh_param_prop_double.restype = i
        L.mh_param_prop_double.argtypes = [vp, i, s, s, i, C.POINTER(d), C.POINTER(i), C.POINTER(s)]
        L.mh_set_param_double.restype = i; L.mh_set_param_double.argtypes = [vp, s, d]
        L.mh_clip_name.restype = s; L.mh_clip_name.argtypes = [vp, i]
        L.mh_clip_prop_string.restype = s; L.mh_clip_prop_string.argtypes = [vp, s, s, i]
        L.mh_set_clip_image.restype = i; L.mh_set_clip_image.argtypes = [vp, s, d, vp, i, i, i, s, s, i, i]
        L.mh_clear_clip_images.restype = i; L.mh_clear_clip_images.argtypes = [vp, s]
        L.mh_set_image_props.restype = i; L.mh_set_image_props.argtypes = [vp, s, d, d, d, s]
        L.mh_provide_unique_identifiers.restype = None; L.mh_provide_unique_identifiers.argtypes = [i]
        L.mh_render.restype = i; L.mh_render.argtypes = [vp, d, i, i, i, i, d, d, i]
        L.mh_frames_needed.restype = i; L.mh_frames_needed.argtypes = [vp, d, s, C.POINTER(d)]
        L.mh_instance_changed.restype = i; L.mh_instance_changed.argtypes = [vp, s]
        L.mh_action.restype = i; L.mh_action.argtypes = [vp, s]
        L.mh_set_abort.restype = None; L.mh_set_abort.argtypes = [vp, i]
        _LIB = L
    return _LIB


class Plugin:
    def __init__(self, name, context="OfxImageEffectContextFilter"):
        path = bundle_path(name)
        if not os.path.exists(path):
            raise RuntimeError("bundle missing: " + path)
        st = C.c_int(-1)
        self.h = lib().mh_load(path.encode(), context.encode(), C.byref(st))
        self.load_status = st.value
        if not self.h:
            raise RuntimeError("could not load " + path)
        self._keep = []

    # -- descriptor ------------------------------------------------------------------------------------
    @property
    def identifier(self):
        return lib().mh_plugin_identifier(self.h).decode()

    @property
    def version(self):
        return lib().mh_plugin_version(self.h, 0), lib().mh_plugin_version(self.h, 1)

    @property
    def api(self):
        return lib().mh_plugin_api(self.h).decode()

    def prop_string(self, name, idx=0, instance=False):
        v = lib().mh_effect_prop_string(self.h, int(instance), name.encode(), idx)
        return v.decode() if v is not None else None

    def prop_strings(self, name):
        out = []
        while True:
            v = self.prop_string(name, len(out))
            if v is None:
                return out
            out.append(v)

    def prop_int(self, name, idx=0):
        v = C.c_int(0)
        return v.value if lib().mh_effect_prop_int(self.h, 0, name.encode(), idx, C.byref(v)) == 0 else None

    def params(self):
        return {lib().mh_param_name(self.h, i).decode(): lib().mh_param_type(self.h, lib().mh_param_name(self.h, i)).decode()
                for i in range(lib().mh_param_count(self.h))}

    def param_prop(self, name, prop, idx=0, instance=False):
        d, i, s = C.c_double(0), C.c_int(0), C.c_char_p()
        st = lib().mh_param_prop_double(self.h, int(instance), name.encode(), prop.encode(), idx, C.byref(d), C.byref(i), C.byref(s))
        if st != 0:
            return None
        return dict(double=d.value, int=i.value, string=(s.value or b"").decode())

    def clips(self):
        return [lib().mh_clip_name(self.h, i).decode() for i in range(lib().mh_clip_count(self.h))]

    def clip_components(self, clip):
        out = []
        while True:
            v = lib().mh_clip_prop_string(self.h, clip.encode(), b"OfxImageEffectPropSupportedComponents", len(out))
            if v is None:
                return out
            out.append(v.decode())

    # -- instance ----------------------------------------------------------------------------------------
    def create_instance(self):
        return lib().mh_create_instance(self.h)

    def destroy_instance(self):
        return lib().mh_destroy_instance(self.h)

    def set_param(self, name, value):
        st = lib().mh_set_param_double(self.h, name.encode(), float(value))
        if st != 0:
            raise KeyError(name)

    def set_image(self, clip, time, arr, x1=0, y1=0, flip_rows=False):
        """arr: HxWxC numpy array (uint8 or float32), row 0 = OFX bottom row.  flip_rows stores the image top-down
        in memory and hands the host a NEGATIVE rowBytes (legal: ofxImageEffect.h:909-921)."""
        assert arr.flags.c_contiguous and arr.ndim == 3
        h, w, c = arr.shape
        depth = DEPTH_BYTE if arr.dtype == np.uint8 else DEPTH_FLOAT
        comps = {4: RGBA, 3: RGB, 1: ALPHA}[c]
        row_bytes = w * c * arr.itemsize
        self._keep.append(arr)
        if flip_rows:
            ptr = arr.ctypes.data + (h - 1) * row_bytes
            row_bytes = -row_bytes
        else:
            ptr = arr.ctypes.data
        st = lib().mh_set_clip_image(self.h, clip.encode(), float(time), C.c_void_p(ptr), w, h, row_bytes, depth.encode(), comps.encode(), x1, y1)
        assert st == 0, st

    def set_device_image(self, clip, time, dptr, w, h, c, dtype, x1=0, y1=0):
        depth = DEPTH_BYTE if np.dtype(dtype) == np.uint8 else DEPTH_FLOAT
        comps = {4: RGBA, 3: RGB, 1: ALPHA}[c]
        st = lib().mh_set_clip_image(self.h, clip.encode(), float(time), C.c_void_p(dptr), w, h, w * c * np.dtype(dtype).itemsize,
                                     depth.encode(), comps.encode(), x1, y1)
        assert st == 0, st

    def set_image_props(self, clip, time, scale=(1.0, 1.0), field="OfxFieldNone"):
        """What the host writes on an already set image: its render scale and field (a mismatch with the render arguments
        must fail the render: VectorGenerator.cpp:531-536)."""
        st = lib().mh_set_image_props(self.h, clip.encode(), float(time), float(scale[0]), float(scale[1]), field.encode())
        assert st == 0, st

    @staticmethod
    def provide_unique_identifiers(on):
        """Hosts that do not set kOfxImagePropUniqueIdentifier (the staged-frame cache of the plugin must then stay off)."""
        lib().mh_provide_unique_identifiers(1 if on else 0)

    def clear_images(self, clip):
        lib().mh_clear_clip_images(self.h, clip.encode())

    def render(self, time, window, scale=(1.0, 1.0), cuda_enabled=-1):
        x1, y1, x2, y2 = window
        return lib().mh_render(self.h, float(time), x1, y1, x2, y2, float(scale[0]), float(scale[1]), int(cuda_enabled))

    def frames_needed(self, time, clip="Source"):
        r = (C.c_double * 2)()
        st = lib().mh_frames_needed(self.h, float(time), clip.encode(), r)
        return st, (r[0], r[1])

    def instance_changed(self, param):
        return lib().mh_instance_changed(self.h, param.encode())

    def action(self, name):
        return lib().mh_action(self.h, name.encode())

    def set_abort(self, v):
        lib().mh_set_abort(self.h, int(v))

    def images_outstanding(self):
        return lib().mh_images_outstanding(self.h)

    def close(self):
        if self.h:
            lib().mh_unload(self.h)
            self.h = None
