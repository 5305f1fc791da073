// ofxcv_glue.h — what the three plugin bundles share: host/suite bookkeeping, exception -> OfxStatus mapping,
// parameter definition helpers, image access and a small pool of C-ABI contexts.
// Mirrors the helper layer of the reference's raw-C-API plugins (/root/reference/opencv2fx/opencv2fx.{h,cpp}:
// throwSuiteStatusException, defineDoubleParam, clamp) and of its Support-library plugin
// (/root/reference/OpenCV/GenericOpenCVPlugin.cpp:327-358 genericCVDescribe).
#pragma once
#include <stdint.h>
#include <string.h>

#include <stdlib.h>

#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
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

// ---- contexts: one per concurrent render (FullySafe), pooled PER DEVICE ------------------------------------------------
// A context lives on one GPU (its stream, workspaces, cached pyramids).  The device of a render is OFXCV_DEVICE when set,
// else the device that owns the host's device images (CUDA render), else the calling thread's current device; a pooled
// context is only handed out again for the device it was created on.
class ContextPool {
public:
    static int default_device()
    {
        if (const char* e = getenv("OFXCV_DEVICE")) return atoi(e);
        return ofxcv_current_device();
    }
    ofxcv_ctx* acquire(int device)
    {
        if (device < 0) device = default_device();
        {
            std::lock_guard<std::mutex> l(m_);
            for (size_t i = free_.size(); i-- > 0;)
                if (ofxcv_device(free_[i]) == device) {
                    ofxcv_ctx* c = free_[i];
                    free_.erase(free_.begin() + (ptrdiff_t)i);
                    return c;
                }
        }
        ofxcv_ctx* c = ofxcv_create(device);
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
    explicit ContextLease(ContextPool& p, int device = -1) : pool(p), ctx(p.acquire(device)) {}
    ~ContextLease()
    {
        ofxcv_set_abort_callback(ctx, nullptr, nullptr);
        pool.release(ctx);
    }
};
// declared AFTER the image guards of a render that hands device pointers to the library: whatever way the action is left
// (abort, a missing neighbour frame, an error), the context's queued work is finished before the host gets its images back
struct SyncOnExit {
    ofxcv_ctx* ctx;
    explicit SyncOnExit(ofxcv_ctx* c) : ctx(c) {}
    ~SyncOnExit()
    {
        if (!ctx) return;
        ofxcv_stream_synchronize(ctx, ofxcv_aux_stream(ctx));  // staging stream (conversions read the host's device images)
        ofxcv_synchronize(ctx);
    }
    SyncOnExit(const SyncOnExit&) = delete;
};

// ---- images --------------------------------------------------------------------------------------------
struct Image {
    OfxPropertySetHandle h = nullptr;
    char* data = nullptr;  // address of pixel (bounds.x1, bounds.y1); rows go UP with +rowBytes
    OfxRectI bounds{0, 0, 0, 0};
    int rowBytes = 0;
    std::string depth, components, uid, field;  // uid: kOfxImagePropUniqueIdentifier ("" when the host gives none)
    OfxPointD scale{1, 1};
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
        s = nullptr;
        if (h.prop->propGetString(p, kOfxImagePropUniqueIdentifier, 0, &s) == kOfxStatOK && s) img.uid = s;
        s = nullptr;
        if (h.prop->propGetString(p, kOfxImagePropField, 0, &s) == kOfxStatOK && s) img.field = s;
        if (h.prop->propGetDoubleN(p, kOfxImageEffectPropRenderScale, 2, &img.scale.x) != kOfxStatOK) img.scale = {1, 1};
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
    std::string field;
};
inline RenderArgs render_args(const Host& h, OfxPropertySetHandle inArgs)
{
    RenderArgs a;
    check(h.prop->propGetDouble(inArgs, kOfxPropTime, 0, &a.time));
    check(h.prop->propGetIntN(inArgs, kOfxImageEffectPropRenderWindow, 4, &a.window.x1));
    if (h.prop->propGetDoubleN(inArgs, kOfxImageEffectPropRenderScale, 2, &a.scale.x) != kOfxStatOK) a.scale = {1, 1};
    if (h.prop->propGetInt(inArgs, kOfxImageEffectPropCudaEnabled, 0, &a.cudaEnabled) != kOfxStatOK) a.cudaEnabled = 0;
    char* f = nullptr;
    if (h.prop->propGetString(inArgs, kOfxImageEffectPropFieldToRender, 0, &f) == kOfxStatOK && f) a.field = f;
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

// host image window <-> tight device rows: the row-chunk pipeline lives behind the C ABI (ofxcv_upload_rows /
// ofxcv_download_rows: worker threads + the context's pinned staging + chunk events); `win` must lie inside the bounds
inline void upload_window(ofxcv_ctx* ctx, const Image& img, const OfxRectI& win, int bpp, void* dev)
{
    check_cv(ofxcv_upload_rows(ctx, nullptr, dev, img.pixel(win.x1, win.y1), img.rowBytes, (size_t)(win.x2 - win.x1) * bpp, win.y2 - win.y1));
}
inline void download_window(ofxcv_ctx* ctx, const Image& img, const OfxRectI& win, int bpp, const void* dev)
{
    check_cv(ofxcv_download_rows(ctx, nullptr, img.pixel(win.x1, win.y1), img.rowBytes, dev, (size_t)(win.x2 - win.x1) * bpp, win.y2 - win.y1));
}

// ---- staged frames kept across renders -------------------------------------------------------------------------------
// getFramesNeeded (VectorGenerator.cpp:675-695) makes render t+1 ask for two of the three frames render t staged.  A host
// that labels its images (kOfxImagePropUniqueIdentifier: "changes whenever the image changes") lets us keep the staged
// 8-bit gray frame + its content key on the device: the next render uploads and converts ONE new frame instead of three.
class GrayCache {
public:
    struct Entry {
        std::string uid;
        OfxRectI win{0, 0, 0, 0};
        int device = -1;
        void* gray = nullptr;  // W*H bytes on `device` (owned)
        size_t cap = 0;
        uint64_t key = 0;      // ofxcv_content_key_u8 of the plane
        uint64_t tick = 0;
        int users = 0;
    };
    // a hit pins the entry (release() when the render is over); a miss returns nullptr
    Entry* find(const std::string& uid, const OfxRectI& win, int device)
    {
        if (uid.empty()) return nullptr;
        std::lock_guard<std::mutex> l(m_);
        for (Entry& e : e_)
            if (e.gray && e.device == device && e.uid == uid && !memcmp(&e.win, &win, sizeof(win))) {
                e.tick = ++tick_;
                e.users++;
                hits_++;
                return &e;
            }
        return nullptr;
    }
    // a pinned slot of at least `bytes` on the context's device for a frame that is about to be staged (nullptr: every
    // slot is in use by concurrent renders -> the caller stages into its own scratch)
    Entry* claim(ofxcv_ctx* ctx, const std::string& uid, const OfxRectI& win, size_t bytes)
    {
        if (uid.empty()) return nullptr;
        std::lock_guard<std::mutex> l(m_);
        Entry* v = nullptr;
        for (Entry& e : e_)
            if (e.users == 0 && (!v || e.tick < v->tick)) v = &e;
        if (!v) return nullptr;
        const int device = ofxcv_device(ctx);
        if (v->gray && (v->device != device || v->cap < bytes)) {
            ofxcv_device_free(ctx, v->gray);  // cudaFree synchronises: nobody is reading it (users == 0)
            v->gray = nullptr;
        }
        if (!v->gray) {
            v->gray = ofxcv_device_alloc(ctx, bytes);
            if (!v->gray) return nullptr;
            v->cap = bytes;
            v->device = device;
        }
        v->uid = uid;
        v->win = win;
        v->key = 0;
        v->tick = ++tick_;
        v->users = 1;
        return v;
    }
    void release(Entry* e)
    {
        if (!e) return;
        std::lock_guard<std::mutex> l(m_);
        e->users--;
        if (e->key == 0) e->uid.clear();  // staging did not complete: never match it
    }
    void clear(ofxcv_ctx* any_ctx_or_null)
    {
        std::lock_guard<std::mutex> l(m_);
        for (Entry& e : e_) {
            if (e.gray && any_ctx_or_null) ofxcv_device_free(any_ctx_or_null, e.gray);
            e = Entry();
        }
    }
    uint64_t hits() const { return hits_; }

private:
    std::mutex m_;
    Entry e_[8];
    uint64_t tick_ = 0, hits_ = 0;
};

}  // namespace ofxcv
