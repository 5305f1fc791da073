// "openCV Segment" — drop-in replacement of /root/reference/opencv2fx/segment/segment.cpp: same plugin identifier,
// version, label, grouping, clips, `threshold1`/`threshold2` parameters and action set.  The reference body
// (cvPyrSegmentation, segment.cpp:296-302) no longer exists in OpenCV >= 3; as BASELINE.json config 3 requires, the
// body is cv::watershed(rgb8, int32 markers) on the GPU.  The plugin has no marker
// input, so seeds come from a deterministic grid: new int parameter `seeds` (default 256 -> 16x16 squares of 5x5
// pixels).  Output: every segment painted with its mean colour (what cvPyrSegmentation's output looked like),
// watershed ridges black, alpha 255.  threshold1/2 are kept for host-project compatibility and are unused.
#include <math.h>
#include <stdlib.h>

#include "ofxcv_glue.h"

using namespace ofxcv;

namespace {
Host gHost;
ContextPool gPool;

struct Instance {
    OfxImageClipHandle src = nullptr, dst = nullptr;
    OfxParamHandle t1 = nullptr, t2 = nullptr, seeds = nullptr;
};

OfxStatus describe(OfxImageEffectHandle effect)
{
    OfxPropertySetHandle p = nullptr;
    check(gHost.effect->getPropertySet(effect, &p));
    const OfxPropertySuiteV1* P = gHost.prop;
    check(P->propSetString(p, kOfxImageEffectPropSupportedPixelDepths, 0, kOfxBitDepthByte));
    check(P->propSetString(p, kOfxPropLabel, 0, "openCV Segment"));
    check(P->propSetString(p, kOfxImageEffectPluginPropGrouping, 0, "Draw"));
    check(P->propSetString(p, kOfxPropPluginDescription, 0, "Marker-based watershed segmentation of the source, computed on the GPU."));
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
    define_double(gHost, ps, "threshold1", "threshold 1", "Sets the threshold #1", 1, 255, 250);  // segment.cpp:385-393
    define_double(gHost, ps, "threshold2", "threshold 2", "Sets the threshold #2", 1, 255, 30);   // segment.cpp:395-403
    define_int(gHost, ps, "seeds", "Seeds", "Number of watershed seed markers, laid out on a regular grid", 256, 1, 4096);
    check(gHost.param->paramDefine(ps, kOfxParamTypePage, "Main", &p));
    check(P->propSetString(p, kOfxParamPropPageChild, 0, "threshold1"));
    check(P->propSetString(p, kOfxParamPropPageChild, 1, "threshold2"));
    check(P->propSetString(p, kOfxParamPropPageChild, 2, "seeds"));
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
        d->t1 = param_handle(gHost, ps, "threshold1");
        d->t2 = param_handle(gHost, ps, "threshold2");
        d->seeds = param_handle(gHost, ps, "seeds");
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

OfxStatus render(OfxImageEffectHandle effect, OfxPropertySetHandle inArgs)
{
    Instance* d = instance_data(effect);
    RenderArgs a = render_args(gHost, inArgs);
    ImageGuard dst(gHost, d->dst, a.time), src(gHost, d->src, a.time);
    if (src.img.depth != kOfxBitDepthByte || dst.img.depth != kOfxBitDepthByte || src.img.components != kOfxImageComponentRGBA ||
        dst.img.components != kOfxImageComponentRGBA)
        return kOfxStatErrImageFormat;
    const OfxRectI win = a.window;
    if (!window_inside(win, src.img.bounds) || !window_inside(win, dst.img.bounds)) return kOfxStatFailed;
    const int W = win.x2 - win.x1, H = win.y2 - win.y1;
    if (W < 3 || H < 3) return kOfxStatFailed;
    int seeds = param_int(gHost, d->seeds, a.time);
    seeds = seeds < 1 ? 1 : seeds;
    int gx = (int)lround(sqrt((double)seeds));
    gx = gx < 1 ? 1 : gx;
    int gy = (seeds + gx - 1) / gx;

    ContextLease lease(gPool);
    ofxcv_ctx* ctx = lease.ctx;
    const size_t n = (size_t)W * H;
    DevBuf d_rgba(ctx, 0, n * 4), d_rgb(ctx, 1, n * 3), d_mask(ctx, 2, n), d_lab(ctx, 3, n * 4);
    upload_window(ctx, src.img, win, 4, d_rgba.p);
    check_cv(ofxcv_rgba8_to_rgb8_mask(ctx, nullptr, (const uint8_t*)d_rgba.p, (ptrdiff_t)W * 4, (uint8_t*)d_rgb.p, (ptrdiff_t)W * 3,
                                      (uint8_t*)d_mask.p, W, W, H, 0));
    check_cv(ofxcv_seed_grid(ctx, nullptr, (int32_t*)d_lab.p, (ptrdiff_t)W * 4, W, H, gx, gy, 2));
    if (gHost.effect->abort(effect)) return kOfxStatOK;
    check_cv(ofxcv_watershed_u8c3(ctx, nullptr, (const uint8_t*)d_rgb.p, (ptrdiff_t)W * 3, (int32_t*)d_lab.p, (ptrdiff_t)W * 4, W, H));
    check_cv(ofxcv_labels_to_rgba8(ctx, nullptr, (const uint8_t*)d_rgb.p, (ptrdiff_t)W * 3, (const int32_t*)d_lab.p, (ptrdiff_t)W * 4,
                                   (uint8_t*)d_rgba.p, (ptrdiff_t)W * 4, W, H, gx * gy));
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
        if (!strcmp(action, kOfxActionDestroyInstance)) {
            delete instance_data(effect);
            return kOfxStatOK;
        }
        if (!strcmp(action, kOfxImageEffectActionRender)) return render(effect, inArgs);
        return kOfxStatReplyDefault;
    });
}

void set_host(OfxHost* h) { gHost.host = h; }

OfxPlugin gPlugin = {kOfxImageEffectPluginApi, 1, "uk.org.bratwurstandhaggis:cvPyrSegmentation", 0, 5, set_host, plugin_main};
}  // namespace

OfxExport int OfxGetNumberOfPlugins(void) { return 1; }
OfxExport OfxPlugin* OfxGetPlugin(int nth) { return nth == 0 ? &gPlugin : nullptr; }
