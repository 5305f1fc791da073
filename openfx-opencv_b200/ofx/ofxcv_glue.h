// ofxcv_glue.h — what the three plugin bundles share: host/suite bookkeeping, exception -> OfxStatus mapping,
// parameter definition helpers, image access and a small pool of C-ABI contexts.
// Mirrors the helper layer of the reference's raw-C-API plugins (/root/reference/opencv2fx/opencv2fx.{h,cpp}:
// throwSuiteStatusException, defineDoubleParam, clamp) and of its Support-library plugin
// (/root/reference/OpenCV/GenericOpenCVPlugin.cpp:327-358 genericCVDescribe).
#pragma once
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ofxcv_abi.h"
#include "ofx_min.h"

namespace ofxcv {

struct StatusException {
    OfxStatus status;
};
inline void check(OfxStatus st)
{
    // opencv2fx.cpp:6-58: kOfxStatOK / ReplyYes / ReplyNo / ReplyDefault pass, everything else is thrown
    switch (st) {
        case kOfxStatOK:
        case kOfxStatReplyYes:
        case kOfxStatReplyNo:
        case kOfxStatReplyDefault: return;
        default: throw StatusException{st};
    }
}

struct Host {
    OfxHost* host = nullptr;
    const OfxImageEffectSuiteV1* effect = nullptr;
    const OfxPropertySuiteV1* prop = nullptr;
    const OfxParameterSuiteV1* param = nullptr;
    OfxStatus fetch()
    {
        if (!host) return kOfxStatErrMissingHostFeature;
        effect = (const OfxImageEffectSuiteV1*)host->fetchSuite(host->host, kOfxImageEffectSuite, 1);
        prop = (const OfxPropertySuiteV1*)host->fetchSuite(host->host, kOfxPropertySuite, 1);
        param = (const OfxParameterSuiteV1*)host->fetchSuite(host->host, kOfxParameterSuite, 1);
        if (!effect || !prop || !param) return kOfxStatErrMissingHostFeature;
        return kOfxStatOK;
    }
};

// maps a C-ABI status onto the OFX error convention (SURVEY.md 8b "Error convention")
inline OfxStatus to_ofx(int st)
{
    switch (st) {
        case OFXCV_OK: return kOfxStatOK;
        case OFXCV_ERR_MEMORY: return kOfxStatErrMemory;
        case OFXCV_ERR_UNSUPPORTED: return kOfxStatErrUnsupported;
        case OFXCV_ERR_NO_DEVICE: return kOfxStatErrMissingHostFeature;  // no GPU: there is no CPU fallback
        default: return kOfxStatFailed;
    }
}
inline void check_cv(int st)
{
    if (st != OFXCV_OK) throw StatusException{to_ofx(st)};
}

// runs an action body, catching everything at the entry point like inpaint.cpp:554-569 / ofxsImageEffect.cpp:5241-5296
template <class F>
OfxStatus guarded(F&& f)
{
    try {
        return f();
    } catch (const StatusException& e) {
        return e.status;
    } catch (const std::bad_alloc&) {
        return kOfxStatErrMemory;
    } catch (...) {
        return kOfxStatFailed;
    }
}

// ---- contexts: one per concurrent render (FullySafe), created lazily on the calling thread's current device ----
class ContextPool {
public:
    ofxcv_ctx* acquire()
    {
        {
            std::lock_guard<std::mutex> l(m_);
            if (!free_.empty()) {
                ofxcv_ctx* c = free_.back();
                free_.pop_back();
                return c;
            }
        }
        int dev = -1;
        if (const char* e = getenv("OFXCV_DEVICE")) dev = atoi(e);
        ofxcv_ctx* c = ofxcv_create(dev);
        if (!c) throw StatusException{kOfxStatErrMissingHostFeature};
        return c;
    }
    void release(ofxcv_ctx* c)
    {
        std::lock_guard<std::mutex> l(m_);
        free_.push_back(c);
    }
    void clear()
    {
        std::lock_guard<std::mutex> l(m_);
        for (ofxcv_ctx* c : free_) ofxcv_destroy(c);
        free_.clear();
    }

private:
    std::mutex m_;
    std::vector<ofxcv_ctx*> free_;
};
struct ContextLease {
    ContextPool& pool;
    ofxcv_ctx* ctx;
    explicit ContextLease(ContextPool& p) : pool(p), ctx(p.acquire()) {}
    ~ContextLease() { pool.release(ctx); }
};

// ---- images --------------------------------------------------------------------------------------------
struct Image {
    OfxPropertySetHandle h = nullptr;
    char* data = nullptr;  // address of pixel (bounds.x1, bounds.y1); rows go UP with +rowBytes
    OfxRectI bounds{0, 0, 0, 0};
    int rowBytes = 0;
    std::string depth, components;
    int ncomp() const { return components == kOfxImageComponentRGBA ? 4 : components == kOfxImageComponentRGB ? 3 : 1; }
    int bytes_per_comp() const { return depth == kOfxBitDepthFloat ? 4 : depth == kOfxBitDepthShort ? 2 : 1; }
    char* row(int y) const { return data + (ptrdiff_t)(y - bounds.y1) * rowBytes; }
    char* pixel(int x, int y) const { return row(y) + (ptrdiff_t)(x - bounds.x1) * ncomp() * bytes_per_comp(); }
};

// RAII clipGetImage / clipReleaseImage (handles are only valid inside the action: ofxImageEffect.h:1253-1258)
class ImageGuard {
public:
    ImageGuard(const Host& h, OfxImageClipHandle clip, OfxTime time) : host_(h)
    {
        OfxPropertySetHandle p = nullptr;
        OfxStatus st = h.effect->clipGetImage(clip, time, nullptr, &p);
        if (st != kOfxStatOK || !p) throw StatusException{kOfxStatFailed};  // missing image -> kOfxStatFailed
        img.h = p;
        void* d = nullptr;
        char* s = nullptr;
        check(h.prop->propGetPointer(p, kOfxImagePropData, 0, &d));
        check(h.prop->propGetIntN(p, kOfxImagePropBounds, 4, &img.bounds.x1));
        check(h.prop->propGetInt(p, kOfxImagePropRowBytes, 0, &img.rowBytes));
        check(h.prop->propGetString(p, kOfxImageEffectPropPixelDepth, 0, &s));
        img.depth = s ? s : "";
        check(h.prop->propGetString(p, kOfxImageEffectPropComponents, 0, &s));
        img.components = s ? s : "";
        img.data = (char*)d;
        if (!img.data) throw StatusException{kOfxStatFailed};
    }
    ~ImageGuard()
    {
        if (img.h) host_.effect->clipReleaseImage(img.h);
    }
    ImageGuard(const ImageGuard&) = delete;
    Image img;

private:
    const Host& host_;
};

struct RenderArgs {
    OfxTime time = 0;
    OfxRectI window{0, 0, 0, 0};
    OfxPointD scale{1, 1};
    int cudaEnabled = 0;
};
inline RenderArgs render_args(const Host& h, OfxPropertySetHandle inArgs)
{
    RenderArgs a;
    check(h.prop->propGetDouble(inArgs, kOfxPropTime, 0, &a.time));
    check(h.prop->propGetIntN(inArgs, kOfxImageEffectPropRenderWindow, 4, &a.window.x1));
    if (h.prop->propGetDoubleN(inArgs, kOfxImageEffectPropRenderScale, 2, &a.scale.x) != kOfxStatOK) a.scale = {1, 1};
    if (h.prop->propGetInt(inArgs, kOfxImageEffectPropCudaEnabled, 0, &a.cudaEnabled) != kOfxStatOK) a.cudaEnabled = 0;
    return a;
}

// ---- parameter definition ---------------------------------------------------------------------------------
// opencv2fx.cpp:60-94
inline void define_double(const Host& h, OfxParamSetHandle ps, const char* name, const char* label, const char* hint, double dmin,
                          double dmax, double def)
{
    OfxPropertySetHandle p = nullptr;
    check(h.param->paramDefine(ps, kOfxParamTypeDouble, name, &p));
    check(h.prop->propSetString(p, kOfxParamPropDoubleType, 0, kOfxParamDoubleTypeScale));
    check(h.prop->propSetDouble(p, kOfxParamPropDefault, 0, def));
    check(h.prop->propSetDouble(p, kOfxParamPropMin, 0, 0.0));
    check(h.prop->propSetDouble(p, kOfxParamPropDisplayMin, 0, dmin));
    check(h.prop->propSetDouble(p, kOfxParamPropDisplayMax, 0, dmax));
    check(h.prop->propSetString(p, kOfxParamPropHint, 0, hint));
    check(h.prop->propSetString(p, kOfxParamPropScriptName, 0, name));
    check(h.prop->propSetString(p, kOfxPropLabel, 0, label));
}
inline void define_plain_double(const Host& h, OfxParamSetHandle ps, const char* name, const char* label, const char* hint, double def,
                                double dmin, double dmax)
{
    OfxPropertySetHandle p = nullptr;
    check(h.param->paramDefine(ps, kOfxParamTypeDouble, name, &p));
    check(h.prop->propSetDouble(p, kOfxParamPropDefault, 0, def));
    check(h.prop->propSetDouble(p, kOfxParamPropDisplayMin, 0, dmin));
    check(h.prop->propSetDouble(p, kOfxParamPropDisplayMax, 0, dmax));
    check(h.prop->propSetString(p, kOfxParamPropHint, 0, hint));
    check(h.prop->propSetString(p, kOfxParamPropScriptName, 0, name));
    check(h.prop->propSetString(p, kOfxPropLabel, 0, label));
}
inline void define_int(const Host& h, OfxParamSetHandle ps, const char* name, const char* label, const char* hint, int def, int dmin,
                       int dmax)
{
    OfxPropertySetHandle p = nullptr;
    check(h.param->paramDefine(ps, kOfxParamTypeInteger, name, &p));
    check(h.prop->propSetInt(p, kOfxParamPropDefault, 0, def));
    check(h.prop->propSetInt(p, kOfxParamPropDisplayMin, 0, dmin));
    check(h.prop->propSetInt(p, kOfxParamPropDisplayMax, 0, dmax));
    check(h.prop->propSetString(p, kOfxParamPropHint, 0, hint));
    check(h.prop->propSetString(p, kOfxParamPropScriptName, 0, name));
    check(h.prop->propSetString(p, kOfxPropLabel, 0, label));
}
inline void define_choice(const Host& h, OfxParamSetHandle ps, const char* name, const char* label, const char* hint,
                          const std::vector<const char*>& options, int def)
{
    OfxPropertySetHandle p = nullptr;
    check(h.param->paramDefine(ps, kOfxParamTypeChoice, name, &p));
    for (size_t i = 0; i < options.size(); i++) check(h.prop->propSetString(p, kOfxParamPropChoiceOption, (int)i, options[i]));
    check(h.prop->propSetInt(p, kOfxParamPropDefault, 0, def));
    check(h.prop->propSetInt(p, kOfxParamPropAnimates, 0, 0));
    check(h.prop->propSetString(p, kOfxParamPropHint, 0, hint));
    check(h.prop->propSetString(p, kOfxParamPropScriptName, 0, name));
    check(h.prop->propSetString(p, kOfxPropLabel, 0, label));
}
inline OfxParamHandle param_handle(const Host& h, OfxParamSetHandle ps, const char* name)
{
    OfxParamHandle p = nullptr;
    check(h.param->paramGetHandle(ps, name, &p, nullptr));
    return p;
}
inline double param_double(const Host& h, OfxParamHandle p, OfxTime t)
{
    double v = 0;
    check(h.param->paramGetValueAtTime(p, t, &v));
    return v;
}
inline int param_int(const Host& h, OfxParamHandle p, OfxTime t)
{
    int v = 0;
    check(h.param->paramGetValueAtTime(p, t, &v));
    return v;
}

// a device buffer owned through the C ABI
// Staging buffers of a render: grow-only scratch slots of the leased context (no cudaMalloc / cudaMallocHost per render).
// One slot per buffer of a render action; a context is used by one render at a time (ContextLease).
struct DevBuf {
    void* p;
    DevBuf(ofxcv_ctx* c, int slot, size_t bytes) : p(ofxcv_scratch_device(c, slot, bytes))
    {
        if (!p) throw StatusException{kOfxStatErrMemory};
    }
    DevBuf(const DevBuf&) = delete;
};
struct PinBuf {
    void* p;
    PinBuf(ofxcv_ctx* c, int slot, size_t bytes) : p(ofxcv_scratch_pinned(c, slot, bytes))
    {
        if (!p) throw StatusException{kOfxStatErrMemory};
    }
    PinBuf(const PinBuf&) = delete;
};

// copy the render window of a host image into a tight (pitch = w*bpp) staging buffer and back; rowBytes may be
// negative (ofxImageEffect.h:909-921), rows are addressed through Image::row
inline void gather_rows(const Image& img, const OfxRectI& win, int bpp, char* tight)
{
    const int w = win.x2 - win.x1;
    for (int y = win.y1; y < win.y2; y++)
        memcpy(tight + (size_t)(y - win.y1) * w * bpp, img.row(y) + (ptrdiff_t)(win.x1 - img.bounds.x1) * bpp, (size_t)w * bpp);
}
inline void scatter_rows(const Image& img, const OfxRectI& win, int bpp, const char* tight)
{
    const int w = win.x2 - win.x1;
    for (int y = win.y1; y < win.y2; y++)
        memcpy(img.row(y) + (ptrdiff_t)(win.x1 - img.bounds.x1) * bpp, tight + (size_t)(y - win.y1) * w * bpp, (size_t)w * bpp);
}
inline bool window_inside(const OfxRectI& win, const OfxRectI& b)
{
    return win.x1 >= b.x1 && win.y1 >= b.y1 && win.x2 <= b.x2 && win.y2 <= b.y2 && win.x2 > win.x1 && win.y2 > win.y1;
}

// Large host images (a 4K float RGBA frame is 133 MB): one thread copying rows into the pinned staging buffer runs at
// ~8 GB/s and would take three times as long as the PCIe transfer it feeds.  The window is cut into row chunks; a few
// workers copy chunks into (out of) the pinned buffer while the calling thread enqueues the H2D copy of every chunk as
// soon as it is complete, so the row copies and the DMA overlap.  `win` must lie inside the image bounds.
inline int staging_workers()
{
    const unsigned hc = std::thread::hardware_concurrency();
    return hc >= 8 ? 4 : hc >= 4 ? 2 : 1;
}
inline void upload_window(ofxcv_ctx* ctx, const Image& img, const OfxRectI& win, int bpp, char* pinned, void* dev)
{
    const int w = win.x2 - win.x1, h = win.y2 - win.y1;
    const size_t row = (size_t)w * bpp;
    const int nch = h >= 256 && row * h >= (8u << 20) ? 16 : 1;
    if (nch == 1) {
        gather_rows(img, win, bpp, pinned);
        check_cv(ofxcv_upload(ctx, nullptr, dev, pinned, row * h));
        return;
    }
    std::vector<std::atomic<int>> done(nch);
    for (auto& d : done) d.store(0, std::memory_order_relaxed);
    std::atomic<int> next{0};
    auto rows_of = [&](int k, int& y0, int& y1) {
        y0 = (int)((long long)h * k / nch);
        y1 = (int)((long long)h * (k + 1) / nch);
    };
    auto work = [&]() {
        for (;;) {
            const int k = next.fetch_add(1);
            if (k >= nch) return;
            int y0, y1;
            rows_of(k, y0, y1);
            for (int y = y0; y < y1; y++)
                memcpy(pinned + (size_t)y * row, img.row(win.y1 + y) + (ptrdiff_t)(win.x1 - img.bounds.x1) * bpp, row);
            done[k].store(1, std::memory_order_release);
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < staging_workers(); t++) th.emplace_back(work);
    int st = OFXCV_OK;
    for (int k = 0; k < nch; k++) {
        while (!done[k].load(std::memory_order_acquire)) std::this_thread::yield();
        int y0, y1;
        rows_of(k, y0, y1);
        if (st == OFXCV_OK) st = ofxcv_upload(ctx, nullptr, (char*)dev + (size_t)y0 * row, pinned + (size_t)y0 * row, (size_t)(y1 - y0) * row);
    }
    for (auto& t : th) t.join();
    check_cv(st);
}
// device (tight rows) -> host image window: one D2H copy, then the rows are scattered by the workers
inline void download_window(ofxcv_ctx* ctx, const Image& img, const OfxRectI& win, int bpp, char* pinned, const void* dev)
{
    const int w = win.x2 - win.x1, h = win.y2 - win.y1;
    const size_t row = (size_t)w * bpp;
    check_cv(ofxcv_download(ctx, nullptr, pinned, dev, row * h));
    check_cv(ofxcv_synchronize(ctx));
    const int nt = h >= 256 && row * h >= (8u << 20) ? staging_workers() : 1;
    if (nt == 1) {
        scatter_rows(img, win, bpp, pinned);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
        th.emplace_back([&, t]() {
            for (int y = (int)((long long)h * t / nt); y < (int)((long long)h * (t + 1) / nt); y++)
                memcpy(img.row(win.y1 + y) + (ptrdiff_t)(win.x1 - img.bounds.x1) * bpp, pinned + (size_t)y * row, row);
        });
    for (auto& t : th) t.join();
}

}  // namespace ofxcv
