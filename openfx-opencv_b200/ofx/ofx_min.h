/*
 * ofx_min.h — the part of the OpenFX 1.4 image-effect C ABI these plugins use, restated from the public standard
 * so that the bundles build hermetically (the GPU box has no /root/reference).  Layout-identical to the headers the
 * reference compiles against (/root/reference/openfx/include/ofxCore.h:61-229,:550-598, ofxProperty.h:49-328,
 * ofxImageEffect.h:1145-1414, ofxParam.h:879-1250); tests/test_ofx_plugins.py
 * (test_ofx_min_header_layout_matches_reference) compiles a translation unit that
 * includes BOTH and static_asserts every struct size / member offset whenever the reference tree is present.
 */
#ifndef OFX_MIN_H
#define OFX_MIN_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int OfxStatus;
typedef double OfxTime;
typedef struct OfxPropertySetStruct* OfxPropertySetHandle;
typedef struct OfxImageEffectStruct* OfxImageEffectHandle;
typedef struct OfxImageClipStruct* OfxImageClipHandle;
typedef struct OfxImageMemoryStruct* OfxImageMemoryHandle;
typedef struct OfxParamStruct* OfxParamHandle;
typedef struct OfxParamSetStruct* OfxParamSetHandle;

typedef struct OfxRectI { int x1, y1, x2, y2; } OfxRectI;
typedef struct OfxRectD { double x1, y1, x2, y2; } OfxRectD;
typedef struct OfxRangeD { double min, max; } OfxRangeD;
typedef struct OfxPointD { double x, y; } OfxPointD;

typedef struct OfxHost {
    OfxPropertySetHandle host;
    const void* (*fetchSuite)(OfxPropertySetHandle host, const char* suiteName, int suiteVersion);
} OfxHost;

typedef OfxStatus(OfxPluginEntryPoint)(const char* action, const void* handle, OfxPropertySetHandle inArgs, OfxPropertySetHandle outArgs);

typedef struct OfxPlugin {
    const char* pluginApi;
    int apiVersion;
    const char* pluginIdentifier;
    unsigned int pluginVersionMajor;
    unsigned int pluginVersionMinor;
    void (*setHost)(OfxHost* host);
    OfxPluginEntryPoint* mainEntry;
} OfxPlugin;

#define kOfxStatOK 0
#define kOfxStatFailed 1
#define kOfxStatErrFatal 2
#define kOfxStatErrUnknown 3
#define kOfxStatErrMissingHostFeature 4
#define kOfxStatErrUnsupported 5
#define kOfxStatErrExists 6
#define kOfxStatErrFormat 7
#define kOfxStatErrMemory 8
#define kOfxStatErrBadHandle 9
#define kOfxStatErrBadIndex 10
#define kOfxStatErrValue 11
#define kOfxStatReplyYes 12
#define kOfxStatReplyNo 13
#define kOfxStatReplyDefault 14
#define kOfxStatErrImageFormat 1000

typedef struct OfxPropertySuiteV1 {
    OfxStatus (*propSetPointer)(OfxPropertySetHandle, const char*, int, void*);
    OfxStatus (*propSetString)(OfxPropertySetHandle, const char*, int, const char*);
    OfxStatus (*propSetDouble)(OfxPropertySetHandle, const char*, int, double);
    OfxStatus (*propSetInt)(OfxPropertySetHandle, const char*, int, int);
    OfxStatus (*propSetPointerN)(OfxPropertySetHandle, const char*, int, void* const*);
    OfxStatus (*propSetStringN)(OfxPropertySetHandle, const char*, int, const char* const*);
    OfxStatus (*propSetDoubleN)(OfxPropertySetHandle, const char*, int, const double*);
    OfxStatus (*propSetIntN)(OfxPropertySetHandle, const char*, int, const int*);
    OfxStatus (*propGetPointer)(OfxPropertySetHandle, const char*, int, void**);
    OfxStatus (*propGetString)(OfxPropertySetHandle, const char*, int, char**);
    OfxStatus (*propGetDouble)(OfxPropertySetHandle, const char*, int, double*);
    OfxStatus (*propGetInt)(OfxPropertySetHandle, const char*, int, int*);
    OfxStatus (*propGetPointerN)(OfxPropertySetHandle, const char*, int, void**);
    OfxStatus (*propGetStringN)(OfxPropertySetHandle, const char*, int, char**);
    OfxStatus (*propGetDoubleN)(OfxPropertySetHandle, const char*, int, double*);
    OfxStatus (*propGetIntN)(OfxPropertySetHandle, const char*, int, int*);
    OfxStatus (*propReset)(OfxPropertySetHandle, const char*);
    OfxStatus (*propGetDimension)(OfxPropertySetHandle, const char*, int*);
} OfxPropertySuiteV1;

typedef struct OfxImageEffectSuiteV1 {
    OfxStatus (*getPropertySet)(OfxImageEffectHandle, OfxPropertySetHandle*);
    OfxStatus (*getParamSet)(OfxImageEffectHandle, OfxParamSetHandle*);
    OfxStatus (*clipDefine)(OfxImageEffectHandle, const char* name, OfxPropertySetHandle*);
    OfxStatus (*clipGetHandle)(OfxImageEffectHandle, const char* name, OfxImageClipHandle*, OfxPropertySetHandle*);
    OfxStatus (*clipGetPropertySet)(OfxImageClipHandle, OfxPropertySetHandle*);
    OfxStatus (*clipGetImage)(OfxImageClipHandle, OfxTime, const OfxRectD* region, OfxPropertySetHandle* image);
    OfxStatus (*clipReleaseImage)(OfxPropertySetHandle image);
    OfxStatus (*clipGetRegionOfDefinition)(OfxImageClipHandle, OfxTime, OfxRectD* bounds);
    int (*abort)(OfxImageEffectHandle);
    OfxStatus (*imageMemoryAlloc)(OfxImageEffectHandle, size_t nBytes, OfxImageMemoryHandle*);
    OfxStatus (*imageMemoryFree)(OfxImageMemoryHandle);
    OfxStatus (*imageMemoryLock)(OfxImageMemoryHandle, void** returnedPtr);
    OfxStatus (*imageMemoryUnlock)(OfxImageMemoryHandle);
} OfxImageEffectSuiteV1;

typedef struct OfxParameterSuiteV1 {
    OfxStatus (*paramDefine)(OfxParamSetHandle, const char* paramType, const char* name, OfxPropertySetHandle*);
    OfxStatus (*paramGetHandle)(OfxParamSetHandle, const char* name, OfxParamHandle*, OfxPropertySetHandle*);
    OfxStatus (*paramSetGetPropertySet)(OfxParamSetHandle, OfxPropertySetHandle*);
    OfxStatus (*paramGetPropertySet)(OfxParamHandle, OfxPropertySetHandle*);
    OfxStatus (*paramGetValue)(OfxParamHandle, ...);
    OfxStatus (*paramGetValueAtTime)(OfxParamHandle, OfxTime, ...);
    OfxStatus (*paramGetDerivative)(OfxParamHandle, OfxTime, ...);
    OfxStatus (*paramGetIntegral)(OfxParamHandle, OfxTime, OfxTime, ...);
    OfxStatus (*paramSetValue)(OfxParamHandle, ...);
    OfxStatus (*paramSetValueAtTime)(OfxParamHandle, OfxTime, ...);
    OfxStatus (*paramGetNumKeys)(OfxParamHandle, unsigned int*);
    OfxStatus (*paramGetKeyTime)(OfxParamHandle, unsigned int, OfxTime*);
    OfxStatus (*paramGetKeyIndex)(OfxParamHandle, OfxTime, int, int*);
    OfxStatus (*paramDeleteKey)(OfxParamHandle, OfxTime);
    OfxStatus (*paramDeleteAllKeys)(OfxParamHandle);
    OfxStatus (*paramCopy)(OfxParamHandle, OfxParamHandle, OfxTime, const OfxRangeD*);
    OfxStatus (*paramEditBegin)(OfxParamSetHandle, const char*);
    OfxStatus (*paramEditEnd)(OfxParamSetHandle);
} OfxParameterSuiteV1;

#define kOfxImageEffectPluginApi "OfxImageEffectPluginAPI"
#define kOfxImageEffectPluginApiVersion 1
#define kOfxPropertySuite "OfxPropertySuite"
#define kOfxImageEffectSuite "OfxImageEffectSuite"
#define kOfxParameterSuite "OfxParameterSuite"

#define kOfxActionLoad "OfxActionLoad"
#define kOfxActionUnload "OfxActionUnload"
#define kOfxActionDescribe "OfxActionDescribe"
#define kOfxActionCreateInstance "OfxActionCreateInstance"
#define kOfxActionDestroyInstance "OfxActionDestroyInstance"
#define kOfxActionInstanceChanged "OfxActionInstanceChanged"
#define kOfxImageEffectActionDescribeInContext "OfxImageEffectActionDescribeInContext"
#define kOfxImageEffectActionRender "OfxImageEffectActionRender"
#define kOfxImageEffectActionGetFramesNeeded "OfxImageEffectActionGetFramesNeeded"

#define kOfxPropName "OfxPropName"
#define kOfxPropLabel "OfxPropLabel"
#define kOfxPropShortLabel "OfxPropShortLabel"
#define kOfxPropLongLabel "OfxPropLongLabel"
#define kOfxPropPluginDescription "OfxPropPluginDescription"
#define kOfxPropInstanceData "OfxPropInstanceData"
#define kOfxPropTime "OfxPropTime"
#define kOfxImageEffectPluginPropGrouping "OfxImageEffectPluginPropGrouping"
#define kOfxImageEffectPropSupportedPixelDepths "OfxImageEffectPropSupportedPixelDepths"
#define kOfxImageEffectPropSupportedContexts "OfxImageEffectPropSupportedContexts"
#define kOfxImageEffectPropSupportedComponents "OfxImageEffectPropSupportedComponents"
#define kOfxImageEffectPropSupportsTiles "OfxImageEffectPropSupportsTiles"
#define kOfxImageEffectPropSupportsMultiResolution "OfxImageEffectPropSupportsMultiResolution"
#define kOfxImageEffectPropTemporalClipAccess "OfxImageEffectPropTemporalClipAccess"
#define kOfxImageEffectPluginRenderThreadSafety "OfxImageEffectPluginRenderThreadSafety"
#define kOfxImageEffectPluginPropHostFrameThreading "OfxImageEffectPluginPropHostFrameThreading"
#define kOfxImageEffectRenderFullySafe "OfxImageEffectRenderFullySafe"
#define kOfxImageEffectRenderInstanceSafe "OfxImageEffectRenderInstanceSafe"
#define kOfxImageEffectPropContext "OfxImageEffectPropContext"
#define kOfxImageEffectContextFilter "OfxImageEffectContextFilter"
#define kOfxImageEffectContextGeneral "OfxImageEffectContextGeneral"
#define kOfxImageEffectPropRenderWindow "OfxImageEffectPropRenderWindow"
#define kOfxImageEffectPropRenderScale "OfxImageEffectPropRenderScale"
#define kOfxImagePropUniqueIdentifier "OfxImagePropUniqueIdentifier"
#define kOfxImagePropField "OfxImagePropField"
#define kOfxImageEffectPropFieldToRender "OfxImageEffectPropFieldToRender"
#define kOfxImageEffectPropFrameRange "OfxImageEffectPropFrameRange"
#define kOfxImageEffectPropPixelDepth "OfxImageEffectPropPixelDepth"
#define kOfxImageEffectPropComponents "OfxImageEffectPropComponents"
#define kOfxImageEffectPropCudaRenderSupported "OfxImageEffectPropCudaRenderSupported"
#define kOfxImageEffectPropCudaEnabled "OfxImageEffectPropCudaEnabled"
#define kOfxImageClipPropOptional "OfxImageClipPropOptional"
#define kOfxImageClipPropConnected "OfxImageClipPropConnected"
#define kOfxImagePropData "OfxImagePropData"
#define kOfxImagePropBounds "OfxImagePropBounds"
#define kOfxImagePropRowBytes "OfxImagePropRowBytes"
#define kOfxImageEffectSimpleSourceClipName "Source"
#define kOfxImageEffectOutputClipName "Output"
#define kOfxBitDepthByte "OfxBitDepthByte"
#define kOfxBitDepthShort "OfxBitDepthShort"
#define kOfxBitDepthFloat "OfxBitDepthFloat"
#define kOfxImageComponentRGBA "OfxImageComponentRGBA"
#define kOfxImageComponentRGB "OfxImageComponentRGB"
#define kOfxImageComponentAlpha "OfxImageComponentAlpha"

#define kOfxParamTypeInteger "OfxParamTypeInteger"
#define kOfxParamTypeDouble "OfxParamTypeDouble"
#define kOfxParamTypeChoice "OfxParamTypeChoice"
#define kOfxParamTypePage "OfxParamTypePage"
#define kOfxParamPropDefault "OfxParamPropDefault"
#define kOfxParamPropHint "OfxParamPropHint"
#define kOfxParamPropScriptName "OfxParamPropScriptName"
#define kOfxParamPropMin "OfxParamPropMin"
#define kOfxParamPropMax "OfxParamPropMax"
#define kOfxParamPropDisplayMin "OfxParamPropDisplayMin"
#define kOfxParamPropDisplayMax "OfxParamPropDisplayMax"
#define kOfxParamPropChoiceOption "OfxParamPropChoiceOption"
#define kOfxParamPropPageChild "OfxParamPropPageChild"
#define kOfxParamPropAnimates "OfxParamPropAnimates"
#define kOfxParamPropSecret "OfxParamPropSecret"
#define kOfxParamPropDoubleType "OfxParamPropDoubleType"
#define kOfxParamDoubleTypeScale "OfxParamDoubleTypeScale"
#define kOfxParamDoubleTypePlain "OfxParamDoubleTypePlain"

#if defined(__GNUC__)
#define OfxExport extern "C" __attribute__((visibility("default")))
#else
#define OfxExport extern "C"
#endif

#ifdef __cplusplus
}
#endif
#endif
