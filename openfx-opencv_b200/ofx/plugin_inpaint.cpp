// "openCV Inpaint" — drop-in replacement of /root/reference/opencv2fx/inpaint/inpaint.cpp: same plugin identifier,
// version, label, grouping, clips, parameters and action set; the body (RGBA8 -> RGB8 + mask, dilate, cvInpaint,
// write-back: inpaint.cpp:292-358) runs as sm_100a CUDA through the C ABI (include/ofxcv_abi.h).  Raw OFX C API,
// like the reference plugin.  New parameter (BASELINE.json config 4): `method` {Telea, Navier-Stokes}, default Telea
// (= the value the reference hard-wires at inpaint.cpp:311).
#include <math.h>
#include <stdlib.h>

#include "ofxcv_glue.h"

using namespace ofxcv;

namespace {
Host gHost;
ContextPool gPool;

const char* kRadius = "threshold1";   // opencv2fx.h: INPAINT_RADIUS
const char* kDilation = "threshold2";  // DILATION
const char* kNoise = "inpaintnoise";   // INPAINT_NOISE
const char* kMethod = "method";

struct Instance {
    OfxImageClipHandle src = nullptr, dst = nullptr;
    OfxParamHandle radius = nullptr, dilation = nullptr, noise = nullptr, method = nullptr;
};

OfxStatus describe(OfxImageEffectHandle effect)
{
    OfxPropertySetHandle p = nullptr;
    check(gHost.effect->getPropertySet(effect, &p));
    const OfxPropertySuiteV1* P = gHost.prop;
    check(P->propSetString(p, kOfxImageEffectPropSupportedPixelDepths, 0, kOfxBitDepthByte));
    check(P->propSetString(p, kOfxPropLabel, 0, "openCV Inpaint"));
    check(P->propSetString(p, kOfxImageEffectPluginPropGrouping, 0, "Draw"));
    check(P->propSetString(p, kOfxPropPluginDescription, 0,
                           "Fast-marching inpainting (Telea / Navier-Stokes) of the black areas of the source, computed on the GPU."));
    check(P->propSetString(p, kOfxImageEffectPropSupportedContexts, 0, kOfxImageEffectContextFilter));
    check(P->propSetInt(p, kOfxImageEffectPluginPropHostFrameThreading, 0, 0));
    check(P->propSetInt(p, kOfxImageEffectPropSupportsMultiResolution, 0, 0));
    check(P->propSetInt(p, kOfxImageEffectPropSupportsTiles, 0, 0));
    check(P->propSetInt(p, kOfxImageEffectPropTemporalClipAccess, 0, 0));
    return kOfxStatOK;
}

OfxStatus describe_in_context(OfxImageEffectHandle effect)
{
    const OfxPropertySuiteV1* P = gHost.prop;
    OfxPropertySetHandle p = nullptr;
    check(gHost.effect->clipDefine(effect, kOfxImageEffectOutputClipName, &p));
    check(P->propSetString(p, kOfxImageEffectPropSupportedComponents, 0, kOfxImageComponentRGBA));
    check(gHost.effect->clipDefine(effect, kOfxImageEffectSimpleSourceClipName, &p));
    check(P->propSetString(p, kOfxImageEffectPropSupportedComponents, 0, kOfxImageComponentRGBA));
    OfxParamSetHandle ps = nullptr;
    check(gHost.effect->getParamSet(effect, &ps));
    define_double(gHost, ps, kRadius, "Radius", "Sets the inpaint radius", 1, 10, 3);
    define_double(gHost, ps, kDilation, "Dilation", "Sets the size of the boundary of intact pixels taken for the inpainting", 1, 5, 1);
    define_double(gHost, ps, kNoise, "Inpaint noise", "Sets additional noise to fake camera noise", 0, 1, 0);
    define_choice(gHost, ps, kMethod, "Method", "Inpainting algorithm", {"Telea", "Navier-Stokes"}, 0);
    check(gHost.param->paramDefine(ps, kOfxParamTypePage, "Main", &p));
    check(P->propSetString(p, kOfxParamPropPageChild, 0, kRadius));
    check(P->propSetString(p, kOfxParamPropPageChild, 1, kDilation));
    check(P->propSetString(p, kOfxParamPropPageChild, 2, kNoise));
    check(P->propSetString(p, kOfxParamPropPageChild, 3, kMethod));
    return kOfxStatOK;
}

OfxStatus create_instance(OfxImageEffectHandle effect)
{
    OfxPropertySetHandle p = nullptr;
    check(gHost.effect->getPropertySet(effect, &p));
    OfxParamSetHandle ps = nullptr;
    check(gHost.effect->getParamSet(effect, &ps));
    Instance* d = new Instance;
    try {
        d->radius = param_handle(gHost, ps, kRadius);
        d->dilation = param_handle(gHost, ps, kDilation);
        d->noise = param_handle(gHost, ps, kNoise);
        d->method = param_handle(gHost, ps, kMethod);
        check(gHost.effect->clipGetHandle(effect, kOfxImageEffectSimpleSourceClipName, &d->src, nullptr));
        check(gHost.effect->clipGetHandle(effect, kOfxImageEffectOutputClipName, &d->dst, nullptr));
        check(gHost.prop->propSetPointer(p, kOfxPropInstanceData, 0, d));
    } catch (...) {
        delete d;
        throw;
    }
    return kOfxStatOK;
}

Instance* instance_data(OfxImageEffectHandle effect)
{
    OfxPropertySetHandle p = nullptr;
    check(gHost.effect->getPropertySet(effect, &p));
    void* d = nullptr;
    check(gHost.prop->propGetPointer(p, kOfxPropInstanceData, 0, &d));
    if (!d) throw StatusException{kOfxStatErrBadHandle};
    return (Instance*)d;
}

OfxStatus destroy_instance(OfxImageEffectHandle effect)
{
    delete instance_data(effect);
    return kOfxStatOK;
}

OfxStatus render(OfxImageEffectHandle effect, OfxPropertySetHandle inArgs)
{
    Instance* d = instance_data(effect);
    RenderArgs a = render_args(gHost, inArgs);
    ImageGuard dst(gHost, d->dst, a.time), src(gHost, d->src, a.time);
    if (src.img.depth != kOfxBitDepthByte || dst.img.depth != kOfxBitDepthByte || src.img.components != kOfxImageComponentRGBA ||
        dst.img.components != kOfxImageComponentRGBA)
        return kOfxStatErrImageFormat;
    // the reference filters the whole source image and writes the render window (no tiles: window == bounds)
    const OfxRectI win = a.window;
    if (!window_inside(win, src.img.bounds) || !window_inside(win, dst.img.bounds)) return kOfxStatFailed;
    const double t1 = param_double(gHost, d->radius, a.time), t2 = param_double(gHost, d->dilation, a.time);
    const double ng = param_double(gHost, d->noise, a.time);
    const int method = param_int(gHost, d->method, a.time) == 1 ? OFXCV_INPAINT_NS : OFXCV_INPAINT_TELEA;
    const int W = win.x2 - win.x1, H = win.y2 - win.y1;
    if (W < 2 || H < 2) return kOfxStatFailed;

    ContextLease lease(gPool);
    ofxcv_ctx* ctx = lease.ctx;
    const size_t n = (size_t)W * H;
    DevBuf d_rgba(ctx, 0, n * 4), d_rgb(ctx, 1, n * 3), d_out(ctx, 2, n * 3), d_mask(ctx, 3, n);
    upload_window(ctx, src.img, win, 4, d_rgba.p);
    check_cv(ofxcv_rgba8_to_rgb8_mask(ctx, nullptr, (const uint8_t*)d_rgba.p, (ptrdiff_t)W * 4, (uint8_t*)d_rgb.p, (ptrdiff_t)W * 3,
                                      (uint8_t*)d_mask.p, W, W, H, t2 > 0 ? (int)t2 : 0));
    if (gHost.effect->abort(effect)) return kOfxStatOK;
    check_cv(ofxcv_inpaint_u8(ctx, nullptr, (const uint8_t*)d_rgb.p, (ptrdiff_t)W * 3, 3, (const uint8_t*)d_mask.p, W, (uint8_t*)d_out.p,
                              (ptrdiff_t)W * 3, W, H, t1, method));
    int noise_div = 0;
    if (ng > 0) noise_div = (int)(1 / ng);  // inpaint.cpp:320-324
    check_cv(ofxcv_rgb8_to_rgba8_noise(ctx, nullptr, (const uint8_t*)d_out.p, (ptrdiff_t)W * 3, (const uint8_t*)d_mask.p, W,
                                       (uint8_t*)d_rgba.p, (ptrdiff_t)W * 4, W, H, noise_div, (unsigned)(long)a.time));
    if (gHost.effect->abort(effect)) return kOfxStatOK;
    download_window(ctx, dst.img, win, 4, d_rgba.p);
    return kOfxStatOK;
}

OfxStatus plugin_main(const char* action, const void* handle, OfxPropertySetHandle inArgs, OfxPropertySetHandle /*outArgs*/)
{
    return guarded([&]() -> OfxStatus {
        OfxImageEffectHandle effect = (OfxImageEffectHandle)handle;
        if (!strcmp(action, kOfxActionLoad)) return gHost.fetch();
        if (!strcmp(action, kOfxActionUnload)) {
            gPool.clear();
            return kOfxStatOK;
        }
        if (!gHost.effect) return kOfxStatErrMissingHostFeature;
        if (!strcmp(action, kOfxActionDescribe)) return describe(effect);
        if (!strcmp(action, kOfxImageEffectActionDescribeInContext)) return describe_in_context(effect);
        if (!strcmp(action, kOfxActionCreateInstance)) return create_instance(effect);
        if (!strcmp(action, kOfxActionDestroyInstance)) return destroy_instance(effect);
        if (!strcmp(action, kOfxImageEffectActionRender)) return render(effect, inArgs);
        return kOfxStatReplyDefault;
    });
}

void set_host(OfxHost* h) { gHost.host = h; }

OfxPlugin gPlugin = {kOfxImageEffectPluginApi, 1, "uk.org.bratwurstandhaggis:cvInpaint", 0, 5, set_host, plugin_main};
}  // namespace

OfxExport int OfxGetNumberOfPlugins(void) { return 1; }
OfxExport OfxPlugin* OfxGetPlugin(int nth) { return nth == 0 ? &gPlugin : nullptr; }
