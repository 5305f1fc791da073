// "VectorGeneratorOFX" — drop-in replacement of /root/reference/VectorGenerator/VectorGenerator.cpp (+ the staging of
// /root/reference/OpenCV/GenericOpenCVPlugin.cpp): same plugin identifier/version/label/grouping, clips, channel and
// Farneback parameters, render thread safety and action set (Describe, DescribeInContext, CreateInstance,
// DestroyInstance, Render, GetFramesNeeded, InstanceChanged).  The render body
//   float RGBA -> Rec.709 luma -> sRGB 8-bit (fetchCVImage8UGrayscale, GenericOpenCVPlugin.cpp:223-265)
//   -> calcOpticalFlowFarneback(prev, next, flow, 0.5, levels, 3, iters, polyN, polySigma, 0) (VectorGenerator.cpp:403)
//   -> dst[x*4+ch] = flow[x*2+c] / renderScale (VectorGenerator.cpp:494-519)
// runs as sm_100a CUDA through the C ABI: frames are staged into HBM once per render (or used in place when the
// host enables OFX CUDA render, ofxImageEffect.h:1013-1049).  Written against the raw OFX C API (no Support library).
// The second method, Dual TV-L1 (VectorGenerator.cpp:436-492), runs through ofxcv_tvl1_u8 with the plugin's tau / lambda /
// theta / nScales / warps / epsilon / iterations controls (parity of that method is unpinned: see the ABI header).
// Deliberate differences, all documented in DESIGN.md: channels set to "0" are written as 0.0 (the reference leaves
// them uninitialised); the whole row is converted (the pinned reference converts a quarter: SURVEY.md B1).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <chrono>

#include "ofxcv_glue.h"

using namespace ofxcv;

namespace {
Host gHost;
ContextPool gPool;
GrayCache gGray;

struct Instance {
    OfxImageClipHandle src = nullptr, dst = nullptr;
    OfxParamHandle chan[4] = {nullptr, nullptr, nullptr, nullptr};
    OfxParamHandle method = nullptr, levels = nullptr, iterations = nullptr, neighborhood = nullptr, sigma = nullptr;
    OfxParamHandle tvl1[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};
const char* kChanNames[4] = {"rChannel", "gChannel", "bChannel", "aChannel"};
const char* kChanLabels[4] = {"R channel", "G channel", "B channel", "A channel"};
const char* kTvl1Names[6] = {"tau", "lambda", "theta", "nScales", "warps", "epsilon"};

OfxStatus describe(OfxImageEffectHandle effect)
{
    // genericCVDescribe (GenericOpenCVPlugin.cpp:327-358)
    OfxPropertySetHandle p = nullptr;
    check(gHost.effect->getPropertySet(effect, &p));
    const OfxPropertySuiteV1* P = gHost.prop;
    check(P->propSetString(p, kOfxPropLabel, 0, "VectorGeneratorOFX"));
    check(P->propSetString(p, kOfxImageEffectPluginPropGrouping, 0, "Time"));
    check(P->propSetString(p, kOfxPropPluginDescription, 0, "Compute optical flow for the input sequence, using OpenCV."));
    check(P->propSetString(p, kOfxImageEffectPropSupportedContexts, 0, kOfxImageEffectContextFilter));
    check(P->propSetString(p, kOfxImageEffectPropSupportedContexts, 1, kOfxImageEffectContextGeneral));
    check(P->propSetString(p, kOfxImageEffectPropSupportedPixelDepths, 0, kOfxBitDepthFloat));
    check(P->propSetInt(p, kOfxImageEffectPluginPropHostFrameThreading, 0, 0));
    check(P->propSetInt(p, kOfxImageEffectPropSupportsMultiResolution, 0, 1));
    check(P->propSetInt(p, kOfxImageEffectPropSupportsTiles, 0, 0));
    check(P->propSetInt(p, kOfxImageEffectPropTemporalClipAccess, 0, 1));
    check(P->propSetString(p, kOfxImageEffectPluginRenderThreadSafety, 0, kOfxImageEffectRenderFullySafe));
    // Resolve-style CUDA hand-off: images may arrive as device pointers (ignored by hosts that do not know it)
    P->propSetString(p, kOfxImageEffectPropCudaRenderSupported, 0, "true");
    return kOfxStatOK;
}

OfxStatus describe_in_context(OfxImageEffectHandle effect)
{
    const OfxPropertySuiteV1* P = gHost.prop;
    OfxPropertySetHandle p = nullptr;
    check(gHost.effect->clipDefine(effect, kOfxImageEffectSimpleSourceClipName, &p));
    check(P->propSetString(p, kOfxImageEffectPropSupportedComponents, 0, kOfxImageComponentRGBA));
    check(P->propSetString(p, kOfxImageEffectPropSupportedComponents, 1, kOfxImageComponentRGB));
    check(P->propSetString(p, kOfxImageEffectPropSupportedComponents, 2, kOfxImageComponentAlpha));
    check(P->propSetInt(p, kOfxImageEffectPropTemporalClipAccess, 0, 1));
    check(P->propSetInt(p, kOfxImageEffectPropSupportsTiles, 0, 0));
    check(gHost.effect->clipDefine(effect, kOfxImageEffectOutputClipName, &p));
    check(P->propSetString(p, kOfxImageEffectPropSupportedComponents, 0, kOfxImageComponentRGBA));
    check(P->propSetInt(p, kOfxImageEffectPropSupportsTiles, 0, 0));

    OfxParamSetHandle ps = nullptr;
    check(gHost.effect->getParamSet(effect, &ps));
    const std::vector<const char*> chans = {"0", "forward.u", "forward.v", "backward.u", "backward.v"};
    for (int c = 0; c < 4; c++)  // VectorGenerator.cpp:731-782, defaults 1,2,3,4
        define_choice(gHost, ps, kChanNames[c], kChanLabels[c],
                      "Selects which component of the motion vectors to set in this channel of the output image", chans, c + 1);
    define_choice(gHost, ps, "method", "Method", "", {"Farneback", "Dual TV L1"}, 0);
    define_int(gHost, ps, "levels", "Levels",
               "Number of pyramid levels including initial image. If 1 that means no extra layer will be created and only the original images are used.",
               3, 1, 10);
    define_int(gHost, ps, "iterations", "Iterations", "Number of iterations the algorithm uses at each pyramid level", 15, 1, 100);
    define_int(gHost, ps, "neighborhood", "Neighborhood",
               "Size of the pixel neighborhood used to find the polynomial expansion in each pixel. Typically 5 or 7.", 5, 1, 16);
    define_plain_double(gHost, ps, "sigma", "Sigma",
                        "Standard deviation of the Gaussian used to smooth derivatives used as a basis of the polynomial expansion. "
                        "For a Neighborhood of 5 you can set Sigma to 1.1, for 7 a good value would be 1.5.",
                        1.1, 0.1, 5);
    // Dual TV-L1 controls (VectorGenerator.cpp:874-929), hidden while the method is Farneback (updateVisibility, :642-662)
    const double tvd[6] = {0.25, 0.15, 0.3, 5, 5, 0.01};
    const char* tvl[6] = {"Tau", "Lambda", "Theta", "N. Scales", "Warps", "Epsilon"};
    for (int i = 0; i < 6; i++) {
        if (i == 3 || i == 4) define_int(gHost, ps, kTvl1Names[i], tvl[i], "Dual TV-L1 parameter", (int)tvd[i], 1, 20);
        else define_plain_double(gHost, ps, kTvl1Names[i], tvl[i], "Dual TV-L1 parameter", tvd[i], 0, 1);
        OfxParamHandle ph = nullptr;
        OfxPropertySetHandle pp = nullptr;
        if (gHost.param->paramGetHandle(ps, kTvl1Names[i], &ph, &pp) == kOfxStatOK && pp) P->propSetInt(pp, kOfxParamPropSecret, 0, 1);
    }
    check(gHost.param->paramDefine(ps, kOfxParamTypePage, "Controls", &p));
    int k = 0;
    for (int c = 0; c < 4; c++) check(P->propSetString(p, kOfxParamPropPageChild, k++, kChanNames[c]));
    for (const char* n : {"method", "levels", "iterations", "neighborhood", "sigma"}) check(P->propSetString(p, kOfxParamPropPageChild, k++, n));
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
        for (int c = 0; c < 4; c++) d->chan[c] = param_handle(gHost, ps, kChanNames[c]);
        d->method = param_handle(gHost, ps, "method");
        d->levels = param_handle(gHost, ps, "levels");
        d->iterations = param_handle(gHost, ps, "iterations");
        d->neighborhood = param_handle(gHost, ps, "neighborhood");
        d->sigma = param_handle(gHost, ps, "sigma");
        for (int i = 0; i < 6; i++) d->tvl1[i] = param_handle(gHost, ps, kTvl1Names[i]);
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

void read_channels(Instance* d, OfxTime t, int ch[4], bool& fwd, bool& bwd)
{
    fwd = bwd = false;
    for (int c = 0; c < 4; c++) {
        ch[c] = param_int(gHost, d->chan[c], t);
        fwd |= ch[c] == 1 || ch[c] == 2;
        bwd |= ch[c] == 3 || ch[c] == 4;
    }
}

// one frame -> 8-bit sRGB gray on the device.  Host images: the render window goes to the device through the row
// pipeline of the C ABI (or, when it sticks out of the image bounds, is gathered replicate-clamped first =
// copyMakeBorder(BORDER_REPLICATE) to the union bounds, VectorGenerator.cpp:387-388), then one conversion kernel.
void stage_gray(ofxcv_ctx* ctx, ofxcv_stream s, const Image& img, const OfxRectI& win, bool device_ptrs, DevBuf& d_float, uint8_t* d_gray)
{
    const int W = win.x2 - win.x1, H = win.y2 - win.y1, nc = img.ncomp();
    if (img.depth != kOfxBitDepthFloat) throw StatusException{kOfxStatErrImageFormat};
    if (device_ptrs) {
        if (!window_inside(win, img.bounds)) throw StatusException{kOfxStatErrUnsupported};
        // the host's device image must live on the GPU this context runs on (multi-GPU hosts render on several)
        if (ofxcv_pointer_device(img.data) != ofxcv_device(ctx)) throw StatusException{kOfxStatErrUnsupported};
        check_cv(ofxcv_rgba32f_to_srgb_gray8(ctx, s, (const float*)img.pixel(win.x1, win.y1), img.rowBytes, nc, d_gray, W, W, H));
        return;
    }
    if (window_inside(win, img.bounds)) {
        check_cv(ofxcv_upload_rows(ctx, s, d_float.p, img.pixel(win.x1, win.y1), img.rowBytes, (size_t)W * nc * 4, H));
    } else {
        PinBuf stage(ctx, 0, (size_t)W * H * nc * 4);
        check_cv(ofxcv_stream_synchronize(ctx, s));  // an earlier upload out of this pinned slot must have left it
        float* sp = (float*)stage.p;
        const OfxRectI& b = img.bounds;
        for (int y = win.y1; y < win.y2; y++) {
            const int yy = y < b.y1 ? b.y1 : y >= b.y2 ? b.y2 - 1 : y;
            float* out = sp + (size_t)(y - win.y1) * W * nc;
            const int xa = win.x1 > b.x1 ? win.x1 : b.x1, xb = win.x2 < b.x2 ? win.x2 : b.x2;  // overlap [xa, xb)
            if (xb > xa) memcpy(out + (size_t)(xa - win.x1) * nc, img.pixel(xa, yy), (size_t)(xb - xa) * nc * 4);
            for (int x = win.x1; x < win.x2; x++) {
                if (x >= xa && x < xb) { x = xb - 1; continue; }
                const int xx = x < b.x1 ? b.x1 : x >= b.x2 ? b.x2 - 1 : x;
                memcpy(out + (size_t)(x - win.x1) * nc, img.pixel(xx, yy), (size_t)nc * 4);
            }
        }
        check_cv(ofxcv_upload(ctx, s, d_float.p, stage.p, (size_t)W * H * nc * 4));
    }
    check_cv(ofxcv_rgba32f_to_srgb_gray8(ctx, s, (const float*)d_float.p, (ptrdiff_t)W * nc * 4, nc, d_gray, W, W, H));
}

// the staged gray frame of an image + its content key: from the cache of staged frames when the host labels its images
// (kOfxImagePropUniqueIdentifier), else staged into `own` (a scratch plane of the render)
struct StagedFrame {
    GrayCache::Entry* entry = nullptr;
    const uint8_t* gray = nullptr;
    uint64_t key = 0;
    ~StagedFrame() { gGray.release(entry); }
};
// staging (upload, conversion, key) runs on the context's staging stream, so that it overlaps whatever flow is already
// queued on the main stream; the main stream is made to wait for it before the caller enqueues anything that reads the plane
void get_gray(ofxcv_ctx* ctx, const Image& img, const OfxRectI& win, bool dev, DevBuf& d_float, uint8_t* own, StagedFrame& out)
{
    ofxcv_stream aux = ofxcv_aux_stream(ctx);
    const int W = win.x2 - win.x1, H = win.y2 - win.y1;
    const int device = ofxcv_device(ctx);
    if ((out.entry = gGray.find(img.uid, win, device))) {
        out.gray = (const uint8_t*)out.entry->gray;
        out.key = out.entry->key;
        return;
    }
    out.entry = gGray.claim(ctx, img.uid, win, (size_t)W * H);
    uint8_t* dstp = out.entry ? (uint8_t*)out.entry->gray : own;
    stage_gray(ctx, aux, img, win, dev, d_float, dstp);
    check_cv(ofxcv_content_key_u8(ctx, aux, dstp, W, W, H, &out.key));  // synchronises the staging stream: the plane is complete
    if (out.entry) out.entry->key = out.key;
    out.gray = dstp;
}

// OFXCV_TRACE=1: wall-clock of the phases of a render on stderr (what tools/plugin_render_time.py reads)
struct Trace {
    bool on;
    std::chrono::steady_clock::time_point t0;
    Trace() : on(getenv("OFXCV_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* what)
    {
        if (!on) return;
        const auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[vg] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t - t0).count());
        t0 = t;
    }
};

int abort_cb(void* effect) { return gHost.effect->abort((OfxImageEffectHandle)effect); }

OfxStatus render(OfxImageEffectHandle effect, OfxPropertySetHandle inArgs)
{
    Instance* d = instance_data(effect);
    RenderArgs a = render_args(gHost, inArgs);
    ImageGuard dst(gHost, d->dst, a.time);
    // VectorGenerator.cpp:531-536: an output image whose scale or field is not the one being rendered fails the render
    if (dst.img.scale.x != a.scale.x || dst.img.scale.y != a.scale.y || (!a.field.empty() && !dst.img.field.empty() && dst.img.field != a.field))
        return kOfxStatFailed;
    if (dst.img.depth != kOfxBitDepthFloat || dst.img.components != kOfxImageComponentRGBA) return kOfxStatErrImageFormat;
    ImageGuard ref(gHost, d->src, a.time);  // missing image -> kOfxStatFailed (VectorGenerator.cpp:539-544)
    int ch[4];
    bool fwd, bwd;
    read_channels(d, a.time, ch, fwd, bwd);
    const int method = param_int(gHost, d->method, a.time);
    if (method != 0 && method != 1) return kOfxStatErrUnsupported;
    ofxcv_tvl1_params tv;
    ofxcv_tvl1_default_params(&tv);
    if (method == 1) {  // VectorGenerator.cpp:459-486
        tv.tau = param_double(gHost, d->tvl1[0], a.time);
        tv.lambda = param_double(gHost, d->tvl1[1], a.time);
        tv.theta = param_double(gHost, d->tvl1[2], a.time);
        tv.nscales = param_int(gHost, d->tvl1[3], a.time);
        tv.warps = param_int(gHost, d->tvl1[4], a.time);
        tv.epsilon = param_double(gHost, d->tvl1[5], a.time);
        tv.iterations = param_int(gHost, d->iterations, a.time);
    }
    ofxcv_fb_params par;
    ofxcv_fb_default_params(&par);
    par.levels = param_int(gHost, d->levels, a.time);
    par.iterations = param_int(gHost, d->iterations, a.time);
    par.poly_n = param_int(gHost, d->neighborhood, a.time);
    par.poly_sigma = param_double(gHost, d->sigma, a.time);

    const OfxRectI win = a.window;
    if (!window_inside(win, dst.img.bounds)) return kOfxStatFailed;
    const int W = win.x2 - win.x1, H = win.y2 - win.y1;
    const size_t n = (size_t)W * H;
    const bool dev = a.cudaEnabled != 0;
    // CUDA render: the context of the GPU that owns the host's images (a host pointer here is a host bug -> unsupported)
    int device = -1;
    if (dev) {
        device = ofxcv_pointer_device(dst.img.data);
        if (device < 0) return kOfxStatErrUnsupported;
    }

    ContextLease lease(gPool, device);
    ofxcv_ctx* ctx = lease.ctx;
    SyncOnExit sync(dev ? ctx : nullptr);  // nothing may still write into (or read from) the host's device images when we leave
    ofxcv_set_abort_callback(ctx, abort_cb, effect);  // polled between pyramid scales inside the flow calls
    DevBuf d_float(ctx, 0, dev ? 16 : n * 16), d_gray0(ctx, 1, n), d_gray1(ctx, 2, n), d_flow(ctx, 3, n * 8), d_dst(ctx, 4, dev ? 16 : n * 16);
    DevBuf d_gray2(ctx, 5, n);  // one plane per neighbour frame: t-1 may still be read by the backward flow while t+1 is staged
    float* out_dev = dev ? (float*)dst.img.pixel(win.x1, win.y1) : (float*)d_dst.p;
    const ptrdiff_t out_stride = dev ? dst.img.rowBytes : (ptrdiff_t)W * 16;
    // content keys: the pyramid of a gray frame is shared by the forward and the backward flow of this render and
    // by the neighbouring renders of the clip (frame t+1 here is frame t of the next render)
    Trace tr;
    StagedFrame f0;
    get_gray(ctx, ref.img, win, dev, d_float, (uint8_t*)d_gray0.p, f0);
    tr.mark("frame t staged");
    // channels set to "0" must read 0: scatter a zero flow into all four channels first
    check_cv(ofxcv_memset(ctx, nullptr, d_flow.p, 0, n * 8));
    {
        const int all[4] = {0, 0, 0, 0};
        check_cv(ofxcv_flow_to_rgba32f(ctx, nullptr, (const float*)d_flow.p, (ptrdiff_t)W * 8, out_dev, out_stride, W, H, all, 1.0, 1.0));
    }
    bool aborted = false;
    // backward first: frames t and t-1 are usually staged already (render t-1 needed them), so its solve is queued at once
    // and runs while frame t+1 is uploaded and converted on the staging stream
    for (int pass = 0; pass < 2 && !aborted; pass++) {
        const int dir = 1 - pass;
        if (!(dir == 0 ? fwd : bwd)) continue;
        if (gHost.effect->abort(effect)) break;
        ImageGuard other(gHost, d->src, dir == 0 ? a.time + 1 : a.time - 1);
        SyncOnExit sync_other(dev ? ctx : nullptr);
        StagedFrame f1;
        get_gray(ctx, other.img, win, dev, d_float, (uint8_t*)(dir == 0 ? d_gray1.p : d_gray2.p), f1);
        tr.mark(dir == 0 ? "frame t+1 staged" : "frame t-1 staged");
        int st;
        if (method == 1)
            st = ofxcv_tvl1_u8(ctx, nullptr, f0.gray, f1.gray, W, W, H, (float*)d_flow.p, (ptrdiff_t)W * 8, &tv);
        else
            st = ofxcv_farneback_u8_keyed(ctx, nullptr, f0.gray, f1.gray, W, W, H, (float*)d_flow.p, (ptrdiff_t)W * 8, &par, f0.key, f1.key);
        if (st == OFXCV_ABORTED) {
            aborted = true;
            break;
        }
        check_cv(st);
        int sel[4];
        const int u = dir == 0 ? 1 : 3, v = dir == 0 ? 2 : 4;
        for (int c = 0; c < 4; c++) sel[c] = ch[c] == u ? 0 : ch[c] == v ? 1 : -1;
        check_cv(ofxcv_flow_to_rgba32f(ctx, nullptr, (const float*)d_flow.p, (ptrdiff_t)W * 8, out_dev, out_stride, W, H, sel, a.scale.x,
                                       a.scale.y));
        tr.mark("flow enqueued");
    }
    if (aborted || gHost.effect->abort(effect)) {  // like the reference's `if (abort()) return;` -- after the queue has drained
        check_cv(ofxcv_synchronize(ctx));
        return kOfxStatOK;
    }
    if (!dev) download_window(ctx, dst.img, win, 16, d_dst.p);
    else check_cv(ofxcv_synchronize(ctx));
    tr.mark(dev ? "synchronised" : "result downloaded");
    return kOfxStatOK;
}

OfxStatus frames_needed(OfxImageEffectHandle effect, OfxPropertySetHandle inArgs, OfxPropertySetHandle outArgs)
{
    // VectorGenerator.cpp:675-695
    Instance* d = instance_data(effect);
    double time = 0;
    check(gHost.prop->propGetDouble(inArgs, kOfxPropTime, 0, &time));
    int ch[4];
    bool fwd, bwd;
    read_channels(d, time, ch, fwd, bwd);
    if (!fwd && !bwd) return kOfxStatReplyDefault;
    double range[2] = {time - (int)bwd, time + (int)fwd};
    check(gHost.prop->propSetDoubleN(outArgs, kOfxImageEffectPropFrameRange "_" kOfxImageEffectSimpleSourceClipName, 2, range));
    return kOfxStatOK;
}

OfxStatus instance_changed(OfxImageEffectHandle effect, OfxPropertySetHandle inArgs)
{
    // VectorGenerator.cpp:642-673: Farneback controls for Farneback, `iterations` for both methods, TV-L1 controls for TV-L1
    char* name = nullptr;
    if (gHost.prop->propGetString(inArgs, kOfxPropName, 0, &name) != kOfxStatOK || !name || strcmp(name, "method")) return kOfxStatReplyDefault;
    Instance* d = instance_data(effect);
    const int method = param_int(gHost, d->method, 0);
    auto secret = [&](OfxParamHandle h, bool hide) {
        OfxPropertySetHandle pp = nullptr;
        if (gHost.param->paramGetPropertySet(h, &pp) == kOfxStatOK && pp) gHost.prop->propSetInt(pp, kOfxParamPropSecret, 0, hide ? 1 : 0);
    };
    for (OfxParamHandle h : {d->levels, d->neighborhood, d->sigma}) secret(h, method != 0);
    secret(d->iterations, method != 0 && method != 1);
    for (OfxParamHandle h : d->tvl1) secret(h, method != 1);
    return kOfxStatOK;
}

OfxStatus plugin_main(const char* action, const void* handle, OfxPropertySetHandle inArgs, OfxPropertySetHandle outArgs)
{
    return guarded([&]() -> OfxStatus {
        OfxImageEffectHandle effect = (OfxImageEffectHandle)handle;
        if (!strcmp(action, kOfxActionLoad)) return gHost.fetch();
        if (!strcmp(action, kOfxActionUnload)) {
            if (ofxcv_device_count() > 0) {
                ContextLease lease(gPool);
                gGray.clear(lease.ctx);
            }
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
        if (!strcmp(action, kOfxImageEffectActionGetFramesNeeded)) return frames_needed(effect, inArgs, outArgs);
        if (!strcmp(action, kOfxActionInstanceChanged)) return instance_changed(effect, inArgs);
        return kOfxStatReplyDefault;
    });
}

void set_host(OfxHost* h) { gHost.host = h; }

OfxPlugin gPlugin = {kOfxImageEffectPluginApi, 1, "net.sf.openfx.VectorGenerator", 1, 0, set_host, plugin_main};
}  // namespace

OfxExport int OfxGetNumberOfPlugins(void) { return 1; }
OfxExport OfxPlugin* OfxGetPlugin(int nth) { return nth == 0 ? &gPlugin : nullptr; }
