// minihost — a small OFX image-effect HOST for tests and the integration bench: it loads one .ofx bundle, runs
// Load / Describe / DescribeInContext / CreateInstance, lets the caller set parameters and clip images (any size,
// byte or float, host or CUDA device memory), and calls the Render / GetFramesNeeded / InstanceChanged actions.
// Modelled on what /root/reference/openfx/HostSupport/examples/hostDemo.cpp:98-303 does with the HostSupport
// library (hostDemoClipInstance.cpp:152-224 for the image property set), but written directly against the C ABI
// in ofx_min.h: property sets, a parameter suite and the image-effect suite in ~400 lines.
// Exposed to Python (ctypes) through the mh_* C functions at the bottom.
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "ofx_min.h"

namespace {

struct Value {
    int i = 0;
    double d = 0;
    void* p = nullptr;
    std::string s;
};
struct PropSet {
    std::map<std::string, std::vector<Value>> v;
    Value& at(const char* name, int idx)
    {
        auto& vec = v[name];
        if ((int)vec.size() <= idx) vec.resize(idx + 1);
        return vec[idx];
    }
    const Value* find(const char* name, int idx) const
    {
        auto it = v.find(name);
        if (it == v.end() || idx < 0 || idx >= (int)it->second.size()) return nullptr;
        return &it->second[idx];
    }
};
PropSet* PS(OfxPropertySetHandle h) { return (PropSet*)h; }

#define SETTER(NAME, T, FIELD)                                                                      \
    OfxStatus NAME(OfxPropertySetHandle h, const char* n, int idx, T val)                           \
    {                                                                                               \
        if (!h || !n || idx < 0) return kOfxStatErrBadHandle;                                       \
        PS(h)->at(n, idx).FIELD = val;                                                              \
        return kOfxStatOK;                                                                          \
    }
SETTER(propSetPointer, void*, p)
SETTER(propSetDouble, double, d)
SETTER(propSetInt, int, i)
OfxStatus propSetString(OfxPropertySetHandle h, const char* n, int idx, const char* val)
{
    if (!h || !n || idx < 0) return kOfxStatErrBadHandle;
    PS(h)->at(n, idx).s = val ? val : "";
    return kOfxStatOK;
}
OfxStatus propSetPointerN(OfxPropertySetHandle h, const char* n, int c, void* const* v) { for (int i = 0; i < c; i++) propSetPointer(h, n, i, v[i]); return kOfxStatOK; }
OfxStatus propSetStringN(OfxPropertySetHandle h, const char* n, int c, const char* const* v) { for (int i = 0; i < c; i++) propSetString(h, n, i, v[i]); return kOfxStatOK; }
OfxStatus propSetDoubleN(OfxPropertySetHandle h, const char* n, int c, const double* v) { for (int i = 0; i < c; i++) propSetDouble(h, n, i, v[i]); return kOfxStatOK; }
OfxStatus propSetIntN(OfxPropertySetHandle h, const char* n, int c, const int* v) { for (int i = 0; i < c; i++) propSetInt(h, n, i, v[i]); return kOfxStatOK; }
#define GETTER(NAME, T, FIELD)                                                                      \
    OfxStatus NAME(OfxPropertySetHandle h, const char* n, int idx, T* val)                          \
    {                                                                                               \
        if (!h || !n) return kOfxStatErrBadHandle;                                                  \
        const Value* v = PS(h)->find(n, idx);                                                       \
        if (!v) return PS(h)->v.count(n) ? kOfxStatErrBadIndex : kOfxStatErrUnknown;                \
        *val = v->FIELD;                                                                            \
        return kOfxStatOK;                                                                          \
    }
GETTER(propGetPointer, void*, p)
GETTER(propGetDouble, double, d)
GETTER(propGetInt, int, i)
OfxStatus propGetString(OfxPropertySetHandle h, const char* n, int idx, char** val)
{
    if (!h || !n) return kOfxStatErrBadHandle;
    const Value* v = PS(h)->find(n, idx);
    if (!v) return PS(h)->v.count(n) ? kOfxStatErrBadIndex : kOfxStatErrUnknown;
    *val = const_cast<char*>(v->s.c_str());
    return kOfxStatOK;
}
OfxStatus propGetPointerN(OfxPropertySetHandle h, const char* n, int c, void** v) { for (int i = 0; i < c; i++) { OfxStatus s = propGetPointer(h, n, i, v + i); if (s) return s; } return kOfxStatOK; }
OfxStatus propGetStringN(OfxPropertySetHandle h, const char* n, int c, char** v) { for (int i = 0; i < c; i++) { OfxStatus s = propGetString(h, n, i, v + i); if (s) return s; } return kOfxStatOK; }
OfxStatus propGetDoubleN(OfxPropertySetHandle h, const char* n, int c, double* v) { for (int i = 0; i < c; i++) { OfxStatus s = propGetDouble(h, n, i, v + i); if (s) return s; } return kOfxStatOK; }
OfxStatus propGetIntN(OfxPropertySetHandle h, const char* n, int c, int* v) { for (int i = 0; i < c; i++) { OfxStatus s = propGetInt(h, n, i, v + i); if (s) return s; } return kOfxStatOK; }
OfxStatus propReset(OfxPropertySetHandle h, const char* n) { if (!h) return kOfxStatErrBadHandle; PS(h)->v.erase(n); return kOfxStatOK; }
OfxStatus propGetDimension(OfxPropertySetHandle h, const char* n, int* c)
{
    if (!h) return kOfxStatErrBadHandle;
    auto it = PS(h)->v.find(n);
    if (it == PS(h)->v.end()) return kOfxStatErrUnknown;
    *c = (int)it->second.size();
    return kOfxStatOK;
}
OfxPropertySuiteV1 gPropSuite = {propSetPointer, propSetString, propSetDouble, propSetInt, propSetPointerN, propSetStringN, propSetDoubleN,
                                 propSetIntN, propGetPointer, propGetString, propGetDouble, propGetInt, propGetPointerN, propGetStringN,
                                 propGetDoubleN, propGetIntN, propReset, propGetDimension};

// ---- parameters ------------------------------------------------------------------------------------------
struct Param {
    std::string type, name;
    PropSet props;
    bool has_value = false;
    double d = 0;
    int i = 0;
};
struct ParamSet {
    std::map<std::string, std::unique_ptr<Param>> params;
    PropSet props;
};
bool is_int_type(const std::string& t) { return t == kOfxParamTypeInteger || t == kOfxParamTypeChoice || t == "OfxParamTypeBoolean"; }

OfxStatus paramDefine(OfxParamSetHandle ps, const char* type, const char* name, OfxPropertySetHandle* props)
{
    ParamSet* s = (ParamSet*)ps;
    if (!s) return kOfxStatErrBadHandle;
    if (s->params.count(name)) return kOfxStatErrExists;
    auto p = std::make_unique<Param>();
    p->type = type;
    p->name = name;
    if (props) *props = (OfxPropertySetHandle)&p->props;
    s->params[name] = std::move(p);
    return kOfxStatOK;
}
OfxStatus paramGetHandle(OfxParamSetHandle ps, const char* name, OfxParamHandle* h, OfxPropertySetHandle* props)
{
    ParamSet* s = (ParamSet*)ps;
    if (!s) return kOfxStatErrBadHandle;
    auto it = s->params.find(name);
    if (it == s->params.end()) return kOfxStatErrUnknown;
    if (h) *h = (OfxParamHandle)it->second.get();
    if (props) *props = (OfxPropertySetHandle)&it->second->props;
    return kOfxStatOK;
}
OfxStatus paramSetGetPropertySet(OfxParamSetHandle ps, OfxPropertySetHandle* props) { *props = (OfxPropertySetHandle) & ((ParamSet*)ps)->props; return kOfxStatOK; }
OfxStatus paramGetPropertySet(OfxParamHandle p, OfxPropertySetHandle* props) { *props = (OfxPropertySetHandle) & ((Param*)p)->props; return kOfxStatOK; }
void param_read(Param* p, va_list ap)
{
    if (is_int_type(p->type)) {
        int* out = va_arg(ap, int*);
        if (p->has_value) *out = p->i;
        else { const Value* v = p->props.find(kOfxParamPropDefault, 0); *out = v ? v->i : 0; }
    } else {
        double* out = va_arg(ap, double*);
        if (p->has_value) *out = p->d;
        else { const Value* v = p->props.find(kOfxParamPropDefault, 0); *out = v ? v->d : 0; }
    }
}
OfxStatus paramGetValue(OfxParamHandle h, ...)
{
    if (!h) return kOfxStatErrBadHandle;
    va_list ap;
    va_start(ap, h);
    param_read((Param*)h, ap);
    va_end(ap);
    return kOfxStatOK;
}
OfxStatus paramGetValueAtTimeReal(OfxParamHandle h, OfxTime t, ...)
{
    (void)t;
    if (!h) return kOfxStatErrBadHandle;
    va_list ap;
    va_start(ap, t);
    param_read((Param*)h, ap);
    va_end(ap);
    return kOfxStatOK;
}
OfxStatus paramSetValue(OfxParamHandle h, ...)
{
    if (!h) return kOfxStatErrBadHandle;
    Param* p = (Param*)h;
    va_list ap;
    va_start(ap, h);
    if (is_int_type(p->type)) p->i = va_arg(ap, int);
    else p->d = va_arg(ap, double);
    p->has_value = true;
    va_end(ap);
    return kOfxStatOK;
}
OfxStatus paramUnsupportedT(OfxParamHandle, OfxTime, ...) { return kOfxStatErrUnsupported; }
OfxStatus paramUnsupportedTT(OfxParamHandle, OfxTime, OfxTime, ...) { return kOfxStatErrUnsupported; }
OfxStatus paramGetNumKeys(OfxParamHandle, unsigned int* n) { *n = 0; return kOfxStatOK; }
OfxStatus paramGetKeyTime(OfxParamHandle, unsigned int, OfxTime*) { return kOfxStatErrBadIndex; }
OfxStatus paramGetKeyIndex(OfxParamHandle, OfxTime, int, int*) { return kOfxStatFailed; }
OfxStatus paramDeleteKey(OfxParamHandle, OfxTime) { return kOfxStatErrBadIndex; }
OfxStatus paramDeleteAllKeys(OfxParamHandle) { return kOfxStatOK; }
OfxStatus paramCopy(OfxParamHandle, OfxParamHandle, OfxTime, const OfxRangeD*) { return kOfxStatErrUnsupported; }
OfxStatus paramEditBegin(OfxParamSetHandle, const char*) { return kOfxStatOK; }
OfxStatus paramEditEnd(OfxParamSetHandle) { return kOfxStatOK; }
OfxParameterSuiteV1 gParamSuite = {paramDefine, paramGetHandle, paramSetGetPropertySet, paramGetPropertySet, paramGetValue,
                                   paramGetValueAtTimeReal, paramUnsupportedT, paramUnsupportedTT, paramSetValue, paramUnsupportedT,
                                   paramGetNumKeys, paramGetKeyTime, paramGetKeyIndex, paramDeleteKey, paramDeleteAllKeys, paramCopy,
                                   paramEditBegin, paramEditEnd};

// ---- image effect ------------------------------------------------------------------------------------------
struct ImageDesc {
    void* data;
    int w, h, rowBytes;
    std::string depth, comps;
    int x1, y1;
    std::string uid;                // kOfxImagePropUniqueIdentifier: changes whenever the host replaces the image ("" = not provided)
    double sx = 1, sy = 1;          // kOfxImageEffectPropRenderScale of the image
    std::string field = "OfxFieldNone";
};
long gImageSerial = 0;
bool gProvideUid = true;
struct Clip {
    std::string name;
    PropSet props;
    std::map<long, ImageDesc> images;  // by time*1000
    int outstanding = 0;
};
struct Effect {
    PropSet props;
    ParamSet params;
    std::map<std::string, std::unique_ptr<Clip>> clips;
    int abort_flag = 0;
    int images_fetched = 0, images_released = 0;
};
struct ImageHandle {
    PropSet props;
    Clip* clip;
};

OfxStatus getPropertySet(OfxImageEffectHandle e, OfxPropertySetHandle* p) { if (!e) return kOfxStatErrBadHandle; *p = (OfxPropertySetHandle) & ((Effect*)e)->props; return kOfxStatOK; }
OfxStatus getParamSet(OfxImageEffectHandle e, OfxParamSetHandle* p) { if (!e) return kOfxStatErrBadHandle; *p = (OfxParamSetHandle) & ((Effect*)e)->params; return kOfxStatOK; }
OfxStatus clipDefine(OfxImageEffectHandle e, const char* name, OfxPropertySetHandle* p)
{
    Effect* ef = (Effect*)e;
    if (!ef) return kOfxStatErrBadHandle;
    auto& c = ef->clips[name];
    if (!c) { c = std::make_unique<Clip>(); c->name = name; }
    if (p) *p = (OfxPropertySetHandle)&c->props;
    return kOfxStatOK;
}
OfxStatus clipGetHandle(OfxImageEffectHandle e, const char* name, OfxImageClipHandle* c, OfxPropertySetHandle* p)
{
    Effect* ef = (Effect*)e;
    if (!ef) return kOfxStatErrBadHandle;
    auto it = ef->clips.find(name);
    if (it == ef->clips.end()) return kOfxStatErrBadHandle;
    if (c) *c = (OfxImageClipHandle)it->second.get();
    if (p) *p = (OfxPropertySetHandle)&it->second->props;
    return kOfxStatOK;
}
OfxStatus clipGetPropertySet(OfxImageClipHandle c, OfxPropertySetHandle* p) { *p = (OfxPropertySetHandle) & ((Clip*)c)->props; return kOfxStatOK; }
Effect* gCurrent = nullptr;
OfxStatus clipGetImage(OfxImageClipHandle ch, OfxTime t, const OfxRectD*, OfxPropertySetHandle* out)
{
    Clip* c = (Clip*)ch;
    if (!c) return kOfxStatErrBadHandle;
    auto it = c->images.find(lround(t * 1000));
    if (it == c->images.end()) return kOfxStatFailed;  // no image at that time
    const ImageDesc& d = it->second;
    ImageHandle* ih = new ImageHandle;
    ih->clip = c;
    OfxPropertySetHandle p = (OfxPropertySetHandle)&ih->props;
    propSetPointer(p, kOfxImagePropData, 0, d.data);
    int b[4] = {d.x1, d.y1, d.x1 + d.w, d.y1 + d.h};
    propSetIntN(p, kOfxImagePropBounds, 4, b);
    propSetIntN(p, "OfxImagePropRegionOfDefinition", 4, b);
    propSetInt(p, kOfxImagePropRowBytes, 0, d.rowBytes);
    propSetString(p, kOfxImageEffectPropPixelDepth, 0, d.depth.c_str());
    propSetString(p, kOfxImageEffectPropComponents, 0, d.comps.c_str());
    double sc[2] = {d.sx, d.sy};
    propSetDoubleN(p, kOfxImageEffectPropRenderScale, 2, sc);
    propSetString(p, "OfxImagePropField", 0, d.field.c_str());
    if (gProvideUid && !d.uid.empty()) propSetString(p, "OfxImagePropUniqueIdentifier", 0, d.uid.c_str());
    c->outstanding++;
    if (gCurrent) gCurrent->images_fetched++;
    *out = p;
    return kOfxStatOK;
}
OfxStatus clipReleaseImage(OfxPropertySetHandle h)
{
    if (!h) return kOfxStatErrBadHandle;
    ImageHandle* ih = (ImageHandle*)h;  // props is the first member
    ih->clip->outstanding--;
    if (gCurrent) gCurrent->images_released++;
    delete ih;
    return kOfxStatOK;
}
OfxStatus clipGetRegionOfDefinition(OfxImageClipHandle ch, OfxTime t, OfxRectD* b)
{
    Clip* c = (Clip*)ch;
    auto it = c->images.find(lround(t * 1000));
    if (it == c->images.end()) return kOfxStatFailed;
    b->x1 = it->second.x1; b->y1 = it->second.y1; b->x2 = it->second.x1 + it->second.w; b->y2 = it->second.y1 + it->second.h;
    return kOfxStatOK;
}
int effectAbort(OfxImageEffectHandle e) { return e ? ((Effect*)e)->abort_flag : 0; }
OfxStatus imageMemoryAlloc(OfxImageEffectHandle, size_t n, OfxImageMemoryHandle* h) { void* p = malloc(n ? n : 1); if (!p) return kOfxStatErrMemory; *h = (OfxImageMemoryHandle)p; return kOfxStatOK; }
OfxStatus imageMemoryFree(OfxImageMemoryHandle h) { free(h); return kOfxStatOK; }
OfxStatus imageMemoryLock(OfxImageMemoryHandle h, void** p) { *p = h; return kOfxStatOK; }
OfxStatus imageMemoryUnlock(OfxImageMemoryHandle) { return kOfxStatOK; }
OfxImageEffectSuiteV1 gEffectSuite = {getPropertySet, getParamSet, clipDefine, clipGetHandle, clipGetPropertySet, clipGetImage, clipReleaseImage,
                                      clipGetRegionOfDefinition, effectAbort, imageMemoryAlloc, imageMemoryFree, imageMemoryLock, imageMemoryUnlock};

PropSet gHostProps;
const void* fetchSuite(OfxPropertySetHandle, const char* name, int version)
{
    if (version != 1) return nullptr;
    if (!strcmp(name, kOfxPropertySuite)) return &gPropSuite;
    if (!strcmp(name, kOfxParameterSuite)) return &gParamSuite;
    if (!strcmp(name, kOfxImageEffectSuite)) return &gEffectSuite;
    return nullptr;
}
OfxHost gHost = {(OfxPropertySetHandle)&gHostProps, fetchSuite};

struct Loaded {
    void* dl = nullptr;
    OfxPlugin* plugin = nullptr;
    Effect descriptor;                 // describe + describeInContext target
    std::unique_ptr<Effect> instance;  // createInstance target
    std::string context;
};

void copy_descriptor(const Effect& d, Effect& inst)
{
    inst.props = d.props;
    for (auto& kv : d.params.params) {
        auto p = std::make_unique<Param>();
        p->type = kv.second->type; p->name = kv.second->name; p->props = kv.second->props;
        inst.params.params[kv.first] = std::move(p);
    }
    for (auto& kv : d.clips) {
        auto c = std::make_unique<Clip>();
        c->name = kv.second->name; c->props = kv.second->props;
        inst.clips[kv.first] = std::move(c);
    }
}

}  // namespace

#define MH extern "C" __attribute__((visibility("default")))

MH void* mh_load(const char* ofx_path, const char* context, int* status)
{
    Loaded* L = new Loaded;
    *status = kOfxStatFailed;
    L->dl = dlopen(ofx_path, RTLD_LAZY | RTLD_LOCAL);  // ofxhBinary.cpp:65
    if (!L->dl) { fprintf(stderr, "minihost: dlopen failed: %s\n", dlerror()); delete L; return nullptr; }
    auto getN = (int (*)(void))dlsym(L->dl, "OfxGetNumberOfPlugins");
    auto getP = (OfxPlugin * (*)(int)) dlsym(L->dl, "OfxGetPlugin");
    if (!getN || !getP || getN() < 1) { delete L; return nullptr; }
    L->plugin = getP(0);
    L->context = context ? context : kOfxImageEffectContextFilter;
    propSetString(gHost.host, kOfxPropName, 0, "ofxcv.minihost");
    L->plugin->setHost(&gHost);
    *status = L->plugin->mainEntry(kOfxActionLoad, nullptr, nullptr, nullptr);
    if (*status != kOfxStatOK && *status != kOfxStatReplyDefault) return L;
    *status = L->plugin->mainEntry(kOfxActionDescribe, &L->descriptor, nullptr, nullptr);
    if (*status != kOfxStatOK && *status != kOfxStatReplyDefault) return L;
    PropSet in;
    propSetString((OfxPropertySetHandle)&in, kOfxImageEffectPropContext, 0, L->context.c_str());
    propSetString((OfxPropertySetHandle)&L->descriptor.props, kOfxImageEffectPropContext, 0, L->context.c_str());
    *status = L->plugin->mainEntry(kOfxImageEffectActionDescribeInContext, &L->descriptor, (OfxPropertySetHandle)&in, nullptr);
    return L;
}
MH const char* mh_plugin_identifier(void* h) { return ((Loaded*)h)->plugin->pluginIdentifier; }
MH int mh_plugin_version(void* h, int minor) { Loaded* L = (Loaded*)h; return minor ? (int)L->plugin->pluginVersionMinor : (int)L->plugin->pluginVersionMajor; }
MH const char* mh_plugin_api(void* h) { return ((Loaded*)h)->plugin->pluginApi; }
MH int mh_create_instance(void* h)
{
    Loaded* L = (Loaded*)h;
    L->instance = std::make_unique<Effect>();
    copy_descriptor(L->descriptor, *L->instance);
    return L->plugin->mainEntry(kOfxActionCreateInstance, L->instance.get(), nullptr, nullptr);
}
MH int mh_destroy_instance(void* h)
{
    Loaded* L = (Loaded*)h;
    if (!L->instance) return kOfxStatErrBadHandle;
    int st = L->plugin->mainEntry(kOfxActionDestroyInstance, L->instance.get(), nullptr, nullptr);
    L->instance.reset();
    return st;
}
MH void mh_unload(void* h)
{
    Loaded* L = (Loaded*)h;
    if (L->instance) mh_destroy_instance(h);
    if (L->plugin) L->plugin->mainEntry(kOfxActionUnload, nullptr, nullptr, nullptr);
    if (L->dl) dlclose(L->dl);
    delete L;
}
static Effect* target(Loaded* L, int instance) { return instance && L->instance ? L->instance.get() : &L->descriptor; }
MH const char* mh_effect_prop_string(void* h, int instance, const char* name, int idx)
{
    const Value* v = target((Loaded*)h, instance)->props.find(name, idx);
    return v ? v->s.c_str() : nullptr;
}
MH int mh_effect_prop_int(void* h, int instance, const char* name, int idx, int* out)
{
    const Value* v = target((Loaded*)h, instance)->props.find(name, idx);
    if (!v) return kOfxStatErrUnknown;
    *out = v->i;
    return kOfxStatOK;
}
MH int mh_param_count(void* h) { return (int)((Loaded*)h)->descriptor.params.params.size(); }
MH const char* mh_param_name(void* h, int i)
{
    auto& m = ((Loaded*)h)->descriptor.params.params;
    auto it = m.begin();
    std::advance(it, i);
    return it->first.c_str();
}
MH const char* mh_param_type(void* h, const char* name)
{
    auto& m = ((Loaded*)h)->descriptor.params.params;
    auto it = m.find(name);
    return it == m.end() ? nullptr : it->second->type.c_str();
}
MH int mh_param_prop_double(void* h, int instance, const char* name, const char* prop, int idx, double* out, int* iout, const char** sout)
{
    auto& m = target((Loaded*)h, instance)->params.params;
    auto it = m.find(name);
    if (it == m.end()) return kOfxStatErrUnknown;
    const Value* v = it->second->props.find(prop, idx);
    if (!v) return kOfxStatErrBadIndex;
    if (out) *out = v->d;
    if (iout) *iout = v->i;
    if (sout) *sout = v->s.c_str();
    return kOfxStatOK;
}
MH int mh_set_param_double(void* h, const char* name, double v)
{
    Loaded* L = (Loaded*)h;
    if (!L->instance) return kOfxStatErrBadHandle;
    auto it = L->instance->params.params.find(name);
    if (it == L->instance->params.params.end()) return kOfxStatErrUnknown;
    it->second->has_value = true;
    it->second->d = v;
    it->second->i = (int)v;
    return kOfxStatOK;
}
MH int mh_clip_count(void* h) { return (int)((Loaded*)h)->descriptor.clips.size(); }
MH const char* mh_clip_name(void* h, int i)
{
    auto& m = ((Loaded*)h)->descriptor.clips;
    auto it = m.begin();
    std::advance(it, i);
    return it->first.c_str();
}
MH const char* mh_clip_prop_string(void* h, const char* clip, const char* prop, int idx)
{
    auto& m = ((Loaded*)h)->descriptor.clips;
    auto it = m.find(clip);
    if (it == m.end()) return nullptr;
    const Value* v = it->second->props.find(prop, idx);
    return v ? v->s.c_str() : nullptr;
}
// data = address of the pixel at (x1, y1) (the lower-left one); rowBytes may be negative
MH int mh_set_clip_image(void* h, const char* clip, double time, void* data, int w, int hgt, int rowBytes, const char* depth,
                         const char* comps, int x1, int y1)
{
    Loaded* L = (Loaded*)h;
    if (!L->instance) return kOfxStatErrBadHandle;
    auto it = L->instance->clips.find(clip);
    if (it == L->instance->clips.end()) return kOfxStatErrUnknown;
    ImageDesc desc{data, w, hgt, rowBytes, depth, comps, x1, y1};
    desc.uid = std::string(clip) + ":" + std::to_string(lround(time * 1000)) + ":" + std::to_string(++gImageSerial);
    it->second->images[lround(time * 1000)] = desc;
    propSetInt((OfxPropertySetHandle)&it->second->props, kOfxImageClipPropConnected, 0, 1);
    return kOfxStatOK;
}
// test hooks: what a (mis)behaving host may put on an image, and hosts that do not label their images
MH int mh_set_image_props(void* h, const char* clip, double time, double sx, double sy, const char* field)
{
    Loaded* L = (Loaded*)h;
    auto it = L->instance->clips.find(clip);
    if (it == L->instance->clips.end()) return kOfxStatErrUnknown;
    auto im = it->second->images.find(lround(time * 1000));
    if (im == it->second->images.end()) return kOfxStatFailed;
    im->second.sx = sx;
    im->second.sy = sy;
    if (field) im->second.field = field;
    return kOfxStatOK;
}
MH void mh_provide_unique_identifiers(int on) { gProvideUid = on != 0; }
MH int mh_clear_clip_images(void* h, const char* clip)
{
    Loaded* L = (Loaded*)h;
    auto it = L->instance->clips.find(clip);
    if (it == L->instance->clips.end()) return kOfxStatErrUnknown;
    it->second->images.clear();
    return kOfxStatOK;
}
MH int mh_render(void* h, double time, int x1, int y1, int x2, int y2, double sx, double sy, int cuda_enabled)
{
    Loaded* L = (Loaded*)h;
    if (!L->instance) return kOfxStatErrBadHandle;
    PropSet in;  // ofxhImageEffect.cpp:1392-1436
    OfxPropertySetHandle p = (OfxPropertySetHandle)&in;
    propSetDouble(p, kOfxPropTime, 0, time);
    int win[4] = {x1, y1, x2, y2};
    propSetIntN(p, kOfxImageEffectPropRenderWindow, 4, win);
    double sc[2] = {sx, sy};
    propSetDoubleN(p, kOfxImageEffectPropRenderScale, 2, sc);
    propSetString(p, kOfxImageEffectPropFieldToRender, 0, "OfxFieldNone");
    if (cuda_enabled >= 0) propSetInt(p, kOfxImageEffectPropCudaEnabled, 0, cuda_enabled);
    gCurrent = L->instance.get();
    int st = L->plugin->mainEntry(kOfxImageEffectActionRender, L->instance.get(), p, nullptr);
    gCurrent = nullptr;
    return st;
}
MH int mh_frames_needed(void* h, double time, const char* clip, double* range)
{
    Loaded* L = (Loaded*)h;
    PropSet in, out;
    propSetDouble((OfxPropertySetHandle)&in, kOfxPropTime, 0, time);
    int st = L->plugin->mainEntry(kOfxImageEffectActionGetFramesNeeded, L->instance.get(), (OfxPropertySetHandle)&in, (OfxPropertySetHandle)&out);
    std::string key = std::string(kOfxImageEffectPropFrameRange) + "_" + clip;
    const Value* a = out.find(key.c_str(), 0);
    const Value* b = out.find(key.c_str(), 1);
    if (a && b) { range[0] = a->d; range[1] = b->d; } else { range[0] = range[1] = time; }
    return st;
}
MH int mh_instance_changed(void* h, const char* param)
{
    Loaded* L = (Loaded*)h;
    PropSet in;
    propSetString((OfxPropertySetHandle)&in, kOfxPropName, 0, param);
    propSetString((OfxPropertySetHandle)&in, "OfxPropType", 0, "OfxTypeParameter");
    propSetString((OfxPropertySetHandle)&in, "OfxPropChangeReason", 0, "OfxChangeUserEdited");
    return L->plugin->mainEntry(kOfxActionInstanceChanged, L->instance.get(), (OfxPropertySetHandle)&in, nullptr);
}
MH int mh_action(void* h, const char* action) { Loaded* L = (Loaded*)h; return L->plugin->mainEntry(action, L->instance ? (void*)L->instance.get() : (void*)&L->descriptor, nullptr, nullptr); }
MH void mh_set_abort(void* h, int v) { Loaded* L = (Loaded*)h; if (L->instance) L->instance->abort_flag = v; }
MH int mh_images_outstanding(void* h)
{
    Loaded* L = (Loaded*)h;
    int n = 0;
    for (auto& kv : L->instance->clips) n += kv.second->outstanding;
    return n;
}
