/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/farneback.c header for the rules).
 *
 * CPU restatement of the float -> 8-bit sRGB staging the reference does before every OpenCV call:
 *   GenericOpenCVPlugin::fetchCVImage8UGrayscale   /root/reference/OpenCV/GenericOpenCVPlugin.cpp:223-265
 *   Lut::to_byte_grayscale_nodither                /root/reference/SupportExt/ofxsLut.h:447-486
 *   Lut::toColorSpaceUint8FromLinearFloatFast      /root/reference/SupportExt/ofxsLut.h:220-223
 *   Lut::fillTables                                /root/reference/SupportExt/ofxsLut.h:171-190
 *   Lut::hipart / Lut::index_to_float              /root/reference/SupportExt/ofxsLut.cpp:62-119
 *   to_func_srgb / from_func_srgb                  /root/reference/SupportExt/ofxsLut.h:660-679
 * over the WHOLE row (the pinned reference converts only a quarter of each RGBA row: SURVEY.md Appendix B1).
 * Parity pin: oracle/_ref/libofxs_lut_ref.so (the reference's own ofxsLut.cpp compiled by `make ref`) compared
 * table-for-table in tests/test_oracle_lut.py, and tests/golden/lut_srgb.npz generated from it.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static float orc_to_srgb(float v)
{
    if (v < 0.0031308f) return (v < 0.0f) ? 0.0f : v * 12.92f;
    return 1.055f * powf(v, 1.0f / 2.4f) - 0.055f;
}
static float orc_from_srgb(float v)
{
    if (v < 0.04045f) return (v < 0.0f) ? 0.0f : v * (1.0f / 12.92f);
    return powf((v + 0.055f) * (1.0f / 1.055f), 2.4f);
}
static float orc_index_to_float(uint16_t i)
{
    if ((i < 0x80) || ((i >= 0x8000) && (i < 0x8080))) return 0;
    if ((i >= 0x7f80) && (i < 0x8000)) return 3.402823466e+38f;
    if (i >= 0xff80) return -3.402823466e+38f;
    uint32_t bits = ((uint32_t)i << 16) | 0x8000u;
    float f;
    memcpy(&f, &bits, 4);
    return f;
}
static int orc_float_to_ff01(float value)
{
    if (value <= 0) return 0;
    if (value >= 1.) return 0xff00;
    return (int)(value * 0xff00 + 0.5);
}
static uint16_t orc_hipart(float f)
{
    uint32_t bits;
    memcpy(&bits, &f, 4);
    return (uint16_t)(bits >> 16);
}

/* to_table: 65536 x u16 (hipart -> uint8xx); from_table: 256 x f32 (byte -> linear) */
void orc_srgb_tables(uint16_t* to_table, float* from_table)
{
    for (int i = 0; i < 0x10000; ++i) to_table[i] = (uint16_t)orc_float_to_ff01(orc_to_srgb(orc_index_to_float((uint16_t)i)));
    for (int b = 0; b < 256; ++b) {
        float f = orc_from_srgb(b / (float)255);
        from_table[b] = f;
        to_table[orc_hipart(f)] = (uint16_t)(b << 8);
    }
}

/* src: h x w x ncomp f32 (ncomp 4, 3 or 1; 1 = already luminance), dst: h x w u8 */
void orc_luma_srgb_gray8(const float* src, int ncomp, uint8_t* dst, int w, int h)
{
    static uint16_t to_table[0x10000];
    static float from_table[256];
    static int init = 0;
    if (!init) { orc_srgb_tables(to_table, from_table); init = 1; }
    for (long p = 0; p < (long)w * h; p++) {
        const float* s = src + p * ncomp;
        float l = ncomp == 1 ? s[0] : (float)(0.2126 * s[0] + 0.7152 * s[1] + 0.0722 * s[2]);
        dst[p] = (uint8_t)((to_table[orc_hipart(l)] + 0x80) >> 8);
    }
}

/* ---- packed conversions: Lut::to_byte_packed_nodither (ofxsLut.h:389-444), Lut::from_byte_packed (:536-581),
 * floatToInt<256> / intToFloat<256> (:47-68), over whole rows ------------------------------------------------ */
static uint8_t orc_alpha_to_byte(float v)
{
    if (v <= 0) return 0;
    if (v >= 1.) return 255;
    return (uint8_t)(int)(v * 255 + 0.5);
}

void orc_to_byte_packed(const float* src, int sn, uint8_t* dst, int dn, int w, int h)
{
    static uint16_t to_table[0x10000];
    static float from_table[256];
    static int init = 0;
    if (!init) { orc_srgb_tables(to_table, from_table); init = 1; }
    for (long p = 0; p < (long)w * h; p++) {
        const float* s = src + p * sn;
        uint8_t t[4] = {0, 0, 0, 0};
        if (sn == 1) t[3] = orc_alpha_to_byte(s[0]);
        else {
            for (int k = 0; k < 3; k++) t[k] = (uint8_t)((to_table[orc_hipart(s[k])] + 0x80) >> 8);
            if (sn == 4) t[3] = orc_alpha_to_byte(s[3]);
        }
        uint8_t* d = dst + p * dn;
        if (dn == 1) d[0] = t[3];
        else for (int k = 0; k < dn; k++) d[k] = t[k];
    }
}

void orc_from_byte_packed(const uint8_t* src, float* dst, int n, int w, int h)
{
    static uint16_t to_table[0x10000];
    static float from_table[256];
    static int init = 0;
    if (!init) { orc_srgb_tables(to_table, from_table); init = 1; }
    for (long p = 0; p < (long)w * h; p++) {
        const uint8_t* s = src + p * n;
        float* d = dst + p * n;
        if (n == 1) d[0] = s[0] / (float)255;
        else {
            for (int k = 0; k < 3; k++) d[k] = from_table[s[k]];
            if (n == 4) d[3] = s[3] / (float)255;
        }
    }
}
