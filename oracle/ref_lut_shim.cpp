// ORACLE — TEST INFRASTRUCTURE ONLY.  C shim over the REFERENCE's own SupportExt/ofxsLut.{h,cpp}, compiled from
// where they lie under /root/reference by `make ref` (outputs only into oracle/_ref/).  Used to pin oracle/lut.c
// and to generate tests/golden/lut_srgb.npz.  No reference source is copied into this repository.
#include "ofxsLut.h"

namespace {
struct NoMutex {
    void lock() {}
    void unlock() {}
};
OFX::Color::LutManager<NoMutex> g_manager;
}  // namespace

extern "C" {
unsigned char ref_srgb_to_byte(float linear) { return g_manager.sRGBLut()->toColorSpaceUint8FromLinearFloatFast(linear); }
float ref_srgb_from_byte(unsigned char b) { return g_manager.sRGBLut()->fromColorSpaceUint8ToLinearFloatFast(b); }
// Rec.709 luma exactly as Lut::to_byte_grayscale_nodither spells it (ofxsLut.h:479), then the table lookup
unsigned char ref_luma_to_byte(float r, float g, float b)
{
    float l = 0.2126 * r + 0.7152 * g + 0.0722 * b;
    return g_manager.sRGBLut()->toColorSpaceUint8FromLinearFloatFast(l);
}
// alpha channel quantisers of the packed conversions (ofxsLut.h:47-68)
unsigned char ref_alpha_to_byte(float a) { return (unsigned char)OFX::Color::floatToInt<256>(a); }
float ref_alpha_from_byte(unsigned char b) { return OFX::Color::intToFloat<256>(b); }
}
