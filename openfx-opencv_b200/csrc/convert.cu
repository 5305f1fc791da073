// Staging conversions either side of the filter bodies.
// All are single-pass HBM-bound element-wise kernels.
#include <math.h>

#include <mutex>

#include "common.cuh"

namespace {

// ---- the sRGB 16-bit-hipart -> uint8xx table of SupportExt's Lut (built on the host exactly like
// Lut::fillTables, /root/reference/SupportExt/ofxsLut.h:171-190 with to_func_srgb :671-679, from_func_srgb
// :660-668, index_to_float ofxsLut.cpp:88-119, floatToInt<0xff01> ofxsLut.h:57-68) ---------------------------
float to_srgb(float v)
{
    if (v < 0.0031308f) return (v < 0.0f) ? 0.0f : v * 12.92f;
    return 1.055f * powf(v, 1.0f / 2.4f) - 0.055f;
}
float from_srgb(float v)
{
    if (v < 0.04045f) return (v < 0.0f) ? 0.0f : v * (1.0f / 12.92f);
    return powf((v + 0.055f) * (1.0f / 1.055f), 2.4f);
}
float index_to_float(unsigned short i)
{
    if ((i < 0x80) || ((i >= 0x8000) && (i < 0x8080))) return 0;
    if ((i >= 0x7f80) && (i < 0x8000)) return 3.402823466e+38f;
    if (i >= 0xff80) return -3.402823466e+38f;
    uint32_t bits = ((uint32_t)i << 16) | 0x8000u;
    float f;
    memcpy(&f, &bits, 4);
    return f;
}
int float_to_ff01(float value)
{
    if (value <= 0) return 0;
    if (value >= 1.) return 0xff00;
    return (int)(value * 0xff00 + 0.5);  // float product, double add, truncation (as the reference's template)
}

inline bool cv_aligned(const void* p, ptrdiff_t stride, int a) { return ((uintptr_t)p % a) == 0 && (stride % a) == 0; }

std::once_flag g_lut_once;
uint16_t g_lut[0x10000];
float g_from_lut[256];  // byte -> linear float (Lut::fromFunc_uint8_to_float, ofxsLut.h:183-189)

void build_lut()
{
    for (int i = 0; i < 0x10000; ++i) g_lut[i] = (uint16_t)float_to_ff01(to_srgb(index_to_float((unsigned short)i)));
    for (int b = 0; b < 256; ++b) {
        float f = from_srgb(b / (float)255);
        uint32_t bits;
        memcpy(&bits, &f, 4);
        g_lut[bits >> 16] = (uint16_t)(b << 8);
        g_from_lut[b] = f;
    }
}

// floatToInt<256> of the reference (ofxsLut.h:57-68): float product, double +0.5, truncation
__device__ __forceinline__ uint8_t alpha_to_byte(float v)
{
    if (v <= 0.f) return 0;
    if (v >= 1.f) return 255;
    return (uint8_t)(int)((double)(v * 255.f) + 0.5);
}

__global__ void __launch_bounds__(256) cv_luma_srgb8(const char* __restrict__ src, ptrdiff_t src_stride, int ncomp,
                                                     uint8_t* __restrict__ dst, ptrdiff_t dst_stride, int W, int H,
                                                     const uint16_t* __restrict__ lut)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W) return;
    const float* row = (const float*)(src + (ptrdiff_t)y * src_stride);
    float l;
    if (ncomp == 4) {
        float4 p = reinterpret_cast<const float4*>(row)[x];
        l = (float)__dadd_rn(__dadd_rn(__dmul_rn(0.2126, (double)p.x), __dmul_rn(0.7152, (double)p.y)), __dmul_rn(0.0722, (double)p.z));
    } else if (ncomp == 3) {
        const float* p = row + (size_t)x * 3;
        l = (float)__dadd_rn(__dadd_rn(__dmul_rn(0.2126, (double)p[0]), __dmul_rn(0.7152, (double)p[1])), __dmul_rn(0.0722, (double)p[2]));
    } else {
        l = row[x];
    }
    unsigned hi = __float_as_uint(l) >> 16;
    dst[(size_t)y * dst_stride + x] = (uint8_t)((lut[hi] + 0x80) >> 8);
}

// RGBA fast path of cv_luma_srgb8: a warp covers 128 pixels of a row, every lane loads pixels lane+32k (coalesced 16-byte loads,
// four independent table gathers in flight) and the bytes are transposed through shuffles so that lane L stores pixels
// 4L..4L+3 as one word (W % 4 == 0, dst rows 4-byte aligned, src rows 16-byte aligned)
__global__ void __launch_bounds__(256) cv_luma_srgb8_rgba4(const char* __restrict__ src, ptrdiff_t src_stride, uint8_t* __restrict__ dst,
                                                           ptrdiff_t dst_stride, int W, int H, const uint16_t* __restrict__ lut)
{
    const int lane = threadIdx.x & 31;
    const int x0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 128;
    const int y = blockIdx.y;
    if (x0 >= W) return;
    const float4* row = reinterpret_cast<const float4*>(src + (ptrdiff_t)y * src_stride);
    float4 p[4];
#pragma unroll
    for (int k = 0; k < 4; k++) p[k] = x0 + lane + 32 * k < W ? row[x0 + lane + 32 * k] : make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float l = (float)__dadd_rn(__dadd_rn(__dmul_rn(0.2126, (double)p[k].x), __dmul_rn(0.7152, (double)p[k].y)),
                                         __dmul_rn(0.0722, (double)p[k].z));
        w |= (uint32_t)((lut[__float_as_uint(l) >> 16] + 0x80) >> 8) << (8 * k);
    }
    uint32_t o = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) o |= ((__shfl_sync(0xffffffffu, w, (4 * lane + j) & 31) >> (8 * (lane >> 3))) & 0xffu) << (8 * j);
    if (x0 + 4 * lane < W) reinterpret_cast<uint32_t*>(dst + (size_t)y * dst_stride)[(x0 >> 2) + lane] = o;
}

// RGBA -> RGBA fast paths of the two packed conversions: 4 pixels per thread, 256 apart (coalesced, 12 gathers in flight)
__global__ void __launch_bounds__(256) cv_to_byte_packed44(const char* __restrict__ src, ptrdiff_t src_stride, uint8_t* __restrict__ dst,
                                                           ptrdiff_t dst_stride, int W, int H, const uint16_t* __restrict__ lut)
{
    const int x0 = blockIdx.x * 1024 + threadIdx.x;
    const int y = blockIdx.y;
    const float4* row = reinterpret_cast<const float4*>(src + (ptrdiff_t)y * src_stride);
    uchar4* out = reinterpret_cast<uchar4*>(dst + (size_t)y * dst_stride);
    float4 p[4];
#pragma unroll
    for (int k = 0; k < 4; k++) p[k] = x0 + 256 * k < W ? row[x0 + 256 * k] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uchar4 t;
        t.x = (uint8_t)((lut[__float_as_uint(p[k].x) >> 16] + 0x80) >> 8);
        t.y = (uint8_t)((lut[__float_as_uint(p[k].y) >> 16] + 0x80) >> 8);
        t.z = (uint8_t)((lut[__float_as_uint(p[k].z) >> 16] + 0x80) >> 8);
        t.w = alpha_to_byte(p[k].w);
        if (x0 + 256 * k < W) out[x0 + 256 * k] = t;
    }
}
__global__ void __launch_bounds__(256) cv_from_byte_packed44(const uint8_t* __restrict__ src, ptrdiff_t src_stride, char* __restrict__ dst,
                                                             ptrdiff_t dst_stride, int W, int H, const float* __restrict__ from)
{
    const int x0 = blockIdx.x * 1024 + threadIdx.x;
    const int y = blockIdx.y;
    const uchar4* row = reinterpret_cast<const uchar4*>(src + (size_t)y * src_stride);
    float4* out = reinterpret_cast<float4*>(dst + (ptrdiff_t)y * dst_stride);
    uchar4 p[4];
#pragma unroll
    for (int k = 0; k < 4; k++) p[k] = x0 + 256 * k < W ? row[x0 + 256 * k] : make_uchar4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (x0 + 256 * k < W) out[x0 + 256 * k] = make_float4(from[p[k].x], from[p[k].y], from[p[k].z], (float)p[k].w / 255.f);
}

// Lut::to_byte_packed_nodither (ofxsLut.h:389-444) over whole rows: colour channels through the hipart table,
// alpha through floatToInt<256>; a 3-component source leaves alpha 0, a 1-component source is alpha only
__global__ void __launch_bounds__(256) cv_to_byte_packed(const char* __restrict__ src, ptrdiff_t src_stride, int sn,
                                                         uint8_t* __restrict__ dst, ptrdiff_t dst_stride, int dn, int W, int H,
                                                         const uint16_t* __restrict__ lut)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W) return;
    const float* p = (const float*)(src + (ptrdiff_t)y * src_stride) + (size_t)x * sn;
    uint8_t t[4] = {0, 0, 0, 0};
    if (sn == 1) t[3] = alpha_to_byte(p[0]);
    else {
#pragma unroll
        for (int k = 0; k < 3; k++) t[k] = (uint8_t)((lut[__float_as_uint(p[k]) >> 16] + 0x80) >> 8);
        if (sn == 4) t[3] = alpha_to_byte(p[3]);
    }
    uint8_t* q = dst + (size_t)y * dst_stride + (size_t)x * dn;
    if (dn == 1) q[0] = t[3];
    else
        for (int k = 0; k < dn; k++) q[k] = t[k];
}

// Lut::from_byte_packed (ofxsLut.h:536-581): colour bytes through the 256-entry table, alpha = b / 255.f
__global__ void __launch_bounds__(256) cv_from_byte_packed(const uint8_t* __restrict__ src, ptrdiff_t src_stride,
                                                           char* __restrict__ dst, ptrdiff_t dst_stride, int n, int W, int H,
                                                           const float* __restrict__ from)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W) return;
    const uint8_t* p = src + (size_t)y * src_stride + (size_t)x * n;
    float* q = (float*)(dst + (ptrdiff_t)y * dst_stride) + (size_t)x * n;
    if (n == 1) q[0] = (float)p[0] / 255.f;
    else {
#pragma unroll
        for (int k = 0; k < 3; k++) q[k] = from[p[k]];
        if (n == 4) q[3] = (float)p[3] / 255.f;
    }
}

__global__ void __launch_bounds__(256) cv_flow_to_rgba(const char* __restrict__ flow, ptrdiff_t flow_stride,
                                                       char* __restrict__ dst, ptrdiff_t dst_stride, int W, int H, int s0,
                                                       int s1, int s2, int s3, double sx, double sy)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W) return;
    const float* f = (const float*)(flow + (ptrdiff_t)y * flow_stride) + 2 * (size_t)x;
    float* d = (float*)(dst + (ptrdiff_t)y * dst_stride) + 4 * (size_t)x;
    // dstPix[x*4+c] = flow[x*2+coord] / renderScale.{x|y}: float / double -> double division, stored as float
    float vx = (float)__ddiv_rn((double)f[0], sx), vy = (float)__ddiv_rn((double)f[1], sy);
    if (s0 >= 0) d[0] = s0 ? vy : vx;
    if (s1 >= 0) d[1] = s1 ? vy : vx;
    if (s2 >= 0) d[2] = s2 ? vy : vx;
    if (s3 >= 0) d[3] = s3 ? vy : vx;
}

// RGBA8 -> RGB8 and the un-dilated hole mask (cvCvtColor RGBA2GRAY fixed point, then THRESH_BINARY_INV at 0)
__global__ void __launch_bounds__(256) cv_rgba_split(const uint8_t* __restrict__ rgba, ptrdiff_t rgba_stride,
                                                     uint8_t* __restrict__ rgb, ptrdiff_t rgb_stride,
                                                     uint8_t* __restrict__ mask, ptrdiff_t mask_stride, int W, int H)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W) return;
    const uint8_t* p = rgba + (size_t)y * rgba_stride + 4 * (size_t)x;
    int r = p[0], g = p[1], b = p[2];
    uint8_t* q = rgb + (size_t)y * rgb_stride + 3 * (size_t)x;
    q[0] = (uint8_t)r; q[1] = (uint8_t)g; q[2] = (uint8_t)b;
    int gray = (9798 * r + 19235 * g + 3735 * b + 16384) >> 15;
    mask[(size_t)y * mask_stride + x] = gray == 0 ? 255 : 0;
}

// 4 pixels per thread: one 16-byte RGBA load, three 4-byte RGB stores, one 4-byte mask store (rows 4-byte aligned, W % 4 == 0)
__global__ void __launch_bounds__(256) cv_rgba_split4(const uint8_t* __restrict__ rgba, ptrdiff_t rgba_stride,
                                                      uint8_t* __restrict__ rgb, ptrdiff_t rgb_stride,
                                                      uint8_t* __restrict__ mask, ptrdiff_t mask_stride, int W4, int H)
{
    const int x4 = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x4 >= W4) return;
    const uint4 p = reinterpret_cast<const uint4*>(rgba + (size_t)y * rgba_stride)[x4];
    const uint32_t px[4] = {p.x, p.y, p.z, p.w};
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int r = px[j] & 0xff, g = (px[j] >> 8) & 0xff, b = (px[j] >> 16) & 0xff;
        const int gray = (9798 * r + 19235 * g + 3735 * b + 16384) >> 15;
        m |= (gray == 0 ? 255u : 0u) << (8 * j);
    }
    // r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
    uint32_t* q = reinterpret_cast<uint32_t*>(rgb + (size_t)y * rgb_stride) + 3 * (size_t)x4;
    q[0] = (px[0] & 0xffffffu) | (px[1] << 24);
    q[1] = ((px[1] >> 8) & 0xffffu) | (px[2] << 16);
    q[2] = ((px[2] >> 16) & 0xffu) | (px[3] << 8);
    reinterpret_cast<uint32_t*>(mask + (size_t)y * mask_stride)[x4] = m;
}

// 4 pixels per thread, bytes as lanes of 32-bit words (rows 4-byte aligned, W % 4 == 0, n <= 4)
__global__ void __launch_bounds__(256) cv_dilate_rows4(const uint8_t* __restrict__ in, ptrdiff_t is, uint8_t* __restrict__ out,
                                                       ptrdiff_t os, int W4, int H, int n)
{
    const int x4 = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x4 >= W4) return;
    const uint32_t* row = reinterpret_cast<const uint32_t*>(in + (size_t)y * is);
    const uint32_t w0 = x4 > 0 ? row[x4 - 1] : 0u, w1 = row[x4], w2 = x4 + 1 < W4 ? row[x4 + 1] : 0u;  // outside = not set
    uint32_t v = w1;
    for (int l = 1; l <= n; l++) {
        v |= __funnelshift_rc(w1, w2, 8 * l);      // bytes x+l (clamped shift: l = 4 gives w2)
        v |= __funnelshift_rc(w0, w1, 32 - 8 * l); // bytes x-l
    }
    reinterpret_cast<uint32_t*>(out + (size_t)y * os)[x4] = v;
}
__global__ void __launch_bounds__(256) cv_dilate_cols4(const uint8_t* __restrict__ in, ptrdiff_t is, uint8_t* __restrict__ out,
                                                       ptrdiff_t os, int W4, int H, int n)
{
    const int x4 = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x4 >= W4) return;
    uint32_t v = 0;
    for (int k = max(y - n, 0); k <= min(y + n, H - 1); k++) v |= reinterpret_cast<const uint32_t*>(in + (size_t)k * is)[x4];
    reinterpret_cast<uint32_t*>(out + (size_t)y * os)[x4] = v;
}

// 4 pixels per thread: three 4-byte RGB loads, one 16-byte RGBA store, optional noise (same rule as cv_rgb_to_rgba_noise)
__device__ __forceinline__ unsigned cv_hash(unsigned x, unsigned y, unsigned s);
__global__ void __launch_bounds__(256) cv_rgb_to_rgba4(const uint8_t* __restrict__ rgb, ptrdiff_t rgb_stride,
                                                       const uint8_t* __restrict__ mask, ptrdiff_t mask_stride,
                                                       uint8_t* __restrict__ rgba, ptrdiff_t rgba_stride, int W4, int H, int noise_div,
                                                       unsigned seed)
{
    const int x4 = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x4 >= W4) return;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(rgb + (size_t)y * rgb_stride) + 3 * (size_t)x4;
    const uint32_t a = p[0], b = p[1], c = p[2];
    uint32_t px[4] = {a & 0xffffffu, (a >> 24) | ((b & 0xffffu) << 8), (b >> 16) | ((c & 0xffu) << 16), c >> 8};
    if (noise_div > 0 && mask[(size_t)y * mask_stride + 4 * (size_t)x4]) {  // only pixels whose x is a multiple of 4 get noise
        const int d = ((int)(cv_hash(4u * x4, y, seed) % 10u) - 5) / noise_div;
        uint32_t o = 0;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) o |= (uint32_t)min(max((int)((px[0] >> (8 * ch)) & 0xff) + d, 0), 255) << (8 * ch);
        px[0] = o;
    }
    reinterpret_cast<uint4*>(rgba + (size_t)y * rgba_stride)[x4] =
        make_uint4(px[0] | 0xff000000u, px[1] | 0xff000000u, px[2] | 0xff000000u, px[3] | 0xff000000u);
}

// n successive 3x3 rect dilations == one (2n+1)^2 rect dilation (constant border = not set)
__global__ void __launch_bounds__(256) cv_dilate_rows(const uint8_t* __restrict__ in, ptrdiff_t is, uint8_t* __restrict__ out,
                                                      ptrdiff_t os, int W, int H, int n)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W) return;
    const uint8_t* row = in + (size_t)y * is;
    uint8_t v = 0;
    for (int l = max(x - n, 0); l <= min(x + n, W - 1); l++) v |= row[l];
    out[(size_t)y * os + x] = v;
}
__global__ void __launch_bounds__(256) cv_dilate_cols(const uint8_t* __restrict__ in, ptrdiff_t is, uint8_t* __restrict__ out,
                                                      ptrdiff_t os, int W, int H, int n)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W) return;
    uint8_t v = 0;
    for (int k = max(y - n, 0); k <= min(y + n, H - 1); k++) v |= in[(size_t)k * is + x];
    out[(size_t)y * os + x] = v;
}

__global__ void __launch_bounds__(256) cv_rgb_to_rgba(const uint8_t* __restrict__ rgb, ptrdiff_t rgb_stride,
                                                      uint8_t* __restrict__ rgba, ptrdiff_t rgba_stride, int W, int H)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W) return;
    const uint8_t* p = rgb + (size_t)y * rgb_stride + 3 * (size_t)x;
    uint8_t* q = rgba + (size_t)y * rgba_stride + 4 * (size_t)x;
    q[0] = p[0]; q[1] = p[1]; q[2] = p[2]; q[3] = 255;
}

// write-back with the optional "camera noise" of inpaint.cpp:320-347: on hole pixels whose x is a multiple of 4,
// a = ((r % 10) - 5) / noise_div is added to the three colour bytes.  The reference draws r from libc rand()
// reseeded from pixel data (not reproducible across libcs: SURVEY.md B5); here r is a counter-based hash of
// (x, y, seed) -- same amplitude and spatial pattern, documented as outside the parity contract.
__device__ __forceinline__ unsigned cv_hash(unsigned x, unsigned y, unsigned s)
{
    unsigned h = x * 0x9E3779B1u ^ (y + 0x7F4A7C15u) * 0x85EBCA77u ^ (s + 1u) * 0xC2B2AE3Du;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}
__global__ void __launch_bounds__(256) cv_rgb_to_rgba_noise(const uint8_t* __restrict__ rgb, ptrdiff_t rgb_stride,
                                                            const uint8_t* __restrict__ mask, ptrdiff_t mask_stride,
                                                            uint8_t* __restrict__ rgba, ptrdiff_t rgba_stride, int W, int H,
                                                            int noise_div, unsigned seed)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W) return;
    const uint8_t* p = rgb + (size_t)y * rgb_stride + 3 * (size_t)x;
    uint8_t* q = rgba + (size_t)y * rgba_stride + 4 * (size_t)x;
    int a = 0;
    if (noise_div > 0 && (x & 3) == 0 && mask[(size_t)y * mask_stride + x]) a = ((int)(cv_hash(x, y, seed) % 10u) - 5) / noise_div;
    q[0] = (uint8_t)min(max((int)p[0] + a, 0), 255);
    q[1] = (uint8_t)min(max((int)p[1] + a, 0), 255);
    q[2] = (uint8_t)min(max((int)p[2] + a, 0), 255);
    q[3] = 255;
}

// position-sensitive content hash of a u8 plane, two independent 64-bit lanes: lane k = sum over pixels of
// mix_k(value + golden * (index+1)) (splitmix64 / murmur3 finalisers with different constants)
__global__ void __launch_bounds__(256) cv_content_key(const uint8_t* __restrict__ img, ptrdiff_t stride, int W, int H,
                                                      unsigned long long* __restrict__ acc)
{
    unsigned long long sum = 0, sum2 = 0;
    const int y = blockIdx.y;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < W; x += gridDim.x * blockDim.x) {
        const unsigned long long v = (unsigned long long)img[(size_t)y * stride + x];
        const unsigned long long idx = (unsigned long long)y * W + x + 1;
        unsigned long long z = v + 0x9E3779B97F4A7C15ull * idx;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        sum += z ^ (z >> 31);
        unsigned long long u = (v << 56) ^ (0xD6E8FEB86659FD93ull * idx);
        u = (u ^ (u >> 33)) * 0xFF51AFD7ED558CCDull;
        u = (u ^ (u >> 33)) * 0xC4CEB9FE1A85EC53ull;
        sum2 += u ^ (u >> 33);
    }
    for (int d = 16; d > 0; d >>= 1) {
        sum += __shfl_down_sync(0xffffffffu, sum, d);
        sum2 += __shfl_down_sync(0xffffffffu, sum2, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (sum) atomicAdd(acc, sum);
        if (sum2) atomicAdd(acc + 1, sum2);
    }
}

__global__ void __launch_bounds__(256) cv_seed_grid(int32_t* __restrict__ m, ptrdiff_t ms, int W, int H, int gx, int gy, int half)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W) return;
    // seed (i,j) is centred at ((2i+1)W/(2gx), (2j+1)H/(2gy))
    int i = min((int)(((long)x * gx) / W), gx - 1), j = min((int)(((long)y * gy) / H), gy - 1);
    int cx = (int)(((long)(2 * i + 1) * W) / (2 * gx)), cy = (int)(((long)(2 * j + 1) * H) / (2 * gy));
    int lab = 0;
    if (abs(x - cx) <= half && abs(y - cy) <= half) lab = j * gx + i + 1;
    m[(size_t)y * ms + x] = lab;
}

// per-label colour sums (u32 x3 + count) by atomics, then paint
__global__ void __launch_bounds__(256) cv_label_sums(const uint8_t* __restrict__ rgb, ptrdiff_t rgb_stride,
                                                     const int32_t* __restrict__ lab, ptrdiff_t ls, int W, int H, int nlabels,
                                                     unsigned long long* __restrict__ sums)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W) return;
    int l = lab[(size_t)y * ls + x];
    if (l <= 0 || l > nlabels) return;
    const uint8_t* p = rgb + (size_t)y * rgb_stride + 3 * (size_t)x;
    unsigned long long* s = sums + (size_t)l * 4;
    atomicAdd(s + 0, (unsigned long long)p[0]);
    atomicAdd(s + 1, (unsigned long long)p[1]);
    atomicAdd(s + 2, (unsigned long long)p[2]);
    atomicAdd(s + 3, 1ull);
}
__global__ void __launch_bounds__(256) cv_label_paint(const int32_t* __restrict__ lab, ptrdiff_t ls, uint8_t* __restrict__ rgba,
                                                      ptrdiff_t rgba_stride, int W, int H, int nlabels,
                                                      const unsigned long long* __restrict__ sums)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W) return;
    int l = lab[(size_t)y * ls + x];
    uint8_t* q = rgba + (size_t)y * rgba_stride + 4 * (size_t)x;
    uint8_t r = 0, g = 0, b = 0;
    if (l > 0 && l <= nlabels) {
        const unsigned long long* s = sums + (size_t)l * 4;
        unsigned long long n = s[3];
        if (n) {
            r = (uint8_t)((s[0] + n / 2) / n);
            g = (uint8_t)((s[1] + n / 2) / n);
            b = (uint8_t)((s[2] + n / 2) / n);
        }
    }
    q[0] = r; q[1] = g; q[2] = b; q[3] = 255;
}

}  // namespace

extern "C" {

static cudaStream_t pick(ofxcv_ctx* ctx, ofxcv_stream s) { return s ? (cudaStream_t)s : ctx->stream; }

// device copies of the two sRGB tables, uploaded once per context: [65536 x u16 hipart table | 256 x f32 from-table]
static int srgb_tables(ofxcv_ctx* ctx, cudaStream_t s, const uint16_t** to, const float** from)
{
    std::call_once(g_lut_once, build_lut);
    char* dl = (char*)ctx->ws[WS_LUT].p;
    if (!dl) {
        dl = (char*)ofxcv_ws(ctx, WS_LUT, sizeof(g_lut) + sizeof(g_from_lut));
        if (!dl) return OFXCV_ERR_MEMORY;
        OFXCV_CUDA(ctx, cudaMemcpyAsync(dl, g_lut, sizeof(g_lut), cudaMemcpyHostToDevice, s));
        OFXCV_CUDA(ctx, cudaMemcpyAsync(dl + sizeof(g_lut), g_from_lut, sizeof(g_from_lut), cudaMemcpyHostToDevice, s));
    }
    *to = (const uint16_t*)dl;
    *from = (const float*)(dl + sizeof(g_lut));
    return OFXCV_OK;
}

int ofxcv_rgba32f_to_srgb_gray8(ofxcv_ctx* ctx, ofxcv_stream stream, const float* src, ptrdiff_t src_stride, int ncomp,
                                uint8_t* dst, ptrdiff_t dst_stride, int W, int H)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    if (!src || !dst || W <= 0 || H <= 0 || (ncomp != 1 && ncomp != 3 && ncomp != 4) || dst_stride < W) return OFXCV_ERR_BAD_ARG;
    if (ncomp == 4 && (((uintptr_t)src | (size_t)(src_stride < 0 ? -src_stride : src_stride)) & 15)) return OFXCV_ERR_BAD_ARG;
    ofxcv_device_guard guard(ctx->device);
    cudaStream_t s = pick(ctx, stream);
    const uint16_t* dl;
    const float* from;
    int st = srgb_tables(ctx, s, &dl, &from);
    if (st < 0) return st;
    if (ncomp == 4 && (W & 3) == 0 && cv_aligned(dst, dst_stride, 4))
        cv_luma_srgb8_rgba4<<<dim3(ofxcv_div_up(W, 1024), H), 256, 0, s>>>((const char*)src, src_stride, dst, dst_stride, W, H, dl);
    else
        cv_luma_srgb8<<<dim3(ofxcv_div_up(W, 256), H), 256, 0, s>>>((const char*)src, src_stride, ncomp, dst, dst_stride, W, H, dl);
    OFXCV_LAUNCH_CHECK(ctx);
    return OFXCV_OK;
}

int ofxcv_rgba32f_to_srgb8_packed(ofxcv_ctx* ctx, ofxcv_stream stream, const float* src, ptrdiff_t src_stride, int src_ncomp,
                                  uint8_t* dst, ptrdiff_t dst_stride, int dst_ncomp, int W, int H)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    auto okn = [](int n) { return n == 1 || n == 3 || n == 4; };
    if (!src || !dst || W <= 0 || H <= 0 || !okn(src_ncomp) || !okn(dst_ncomp) || dst_stride < (ptrdiff_t)W * dst_ncomp || (src_stride & 3))
        return OFXCV_ERR_BAD_ARG;
    ofxcv_device_guard guard(ctx->device);
    cudaStream_t s = pick(ctx, stream);
    const uint16_t* to;
    const float* from;
    int st = srgb_tables(ctx, s, &to, &from);
    if (st < 0) return st;
    if (src_ncomp == 4 && dst_ncomp == 4 && cv_aligned(src, src_stride, 16) && cv_aligned(dst, dst_stride, 4))
        cv_to_byte_packed44<<<dim3(ofxcv_div_up(W, 1024), H), 256, 0, s>>>((const char*)src, src_stride, dst, dst_stride, W, H, to);
    else
        cv_to_byte_packed<<<dim3(ofxcv_div_up(W, 256), H), 256, 0, s>>>((const char*)src, src_stride, src_ncomp, dst, dst_stride, dst_ncomp, W, H, to);
    OFXCV_LAUNCH_CHECK(ctx);
    return OFXCV_OK;
}

int ofxcv_srgb8_packed_to_rgba32f(ofxcv_ctx* ctx, ofxcv_stream stream, const uint8_t* src, ptrdiff_t src_stride, float* dst,
                                  ptrdiff_t dst_stride, int ncomp, int W, int H)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    if (!src || !dst || W <= 0 || H <= 0 || (ncomp != 1 && ncomp != 3 && ncomp != 4) || src_stride < (ptrdiff_t)W * ncomp || (dst_stride & 3))
        return OFXCV_ERR_BAD_ARG;
    ofxcv_device_guard guard(ctx->device);
    cudaStream_t s = pick(ctx, stream);
    const uint16_t* to;
    const float* from;
    int st = srgb_tables(ctx, s, &to, &from);
    if (st < 0) return st;
    if (ncomp == 4 && cv_aligned(src, src_stride, 4) && cv_aligned(dst, dst_stride, 16))
        cv_from_byte_packed44<<<dim3(ofxcv_div_up(W, 1024), H), 256, 0, s>>>(src, src_stride, (char*)dst, dst_stride, W, H, from);
    else
        cv_from_byte_packed<<<dim3(ofxcv_div_up(W, 256), H), 256, 0, s>>>(src, src_stride, (char*)dst, dst_stride, ncomp, W, H, from);
    OFXCV_LAUNCH_CHECK(ctx);
    return OFXCV_OK;
}

int ofxcv_flow_to_rgba32f(ofxcv_ctx* ctx, ofxcv_stream stream, const float* flow, ptrdiff_t flow_stride, float* dst,
                          ptrdiff_t dst_stride, int W, int H, const int chan_sel[4], double scale_x, double scale_y)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    if (!flow || !dst || !chan_sel || W <= 0 || H <= 0) return OFXCV_ERR_BAD_ARG;
    ofxcv_device_guard guard(ctx->device);
    cv_flow_to_rgba<<<dim3(ofxcv_div_up(W, 256), H), 256, 0, pick(ctx, stream)>>>((const char*)flow, flow_stride, (char*)dst, dst_stride, W,
                                                                                  H, chan_sel[0], chan_sel[1], chan_sel[2],
                                                                                  chan_sel[3], scale_x, scale_y);
    OFXCV_LAUNCH_CHECK(ctx);
    return OFXCV_OK;
}

int ofxcv_rgba8_to_rgb8_mask(ofxcv_ctx* ctx, ofxcv_stream stream, const uint8_t* rgba, ptrdiff_t rgba_stride, uint8_t* rgb,
                             ptrdiff_t rgb_stride, uint8_t* mask, ptrdiff_t mask_stride, int W, int H, int dilate_iterations)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    if (!rgba || !rgb || !mask || W <= 0 || H <= 0) return OFXCV_ERR_BAD_ARG;
    ofxcv_device_guard guard(ctx->device);
    cudaStream_t s = pick(ctx, stream);
    dim3 grid(ofxcv_div_up(W, 256), H);
    // 4-pixel kernels need W % 4 == 0 and rows that start on the natural boundary of the word they use
    const bool w4 = (W & 3) == 0;
    const bool mask4 = w4 && cv_aligned(mask, mask_stride, 4);
    dim3 grid4(ofxcv_div_up(W / 4 > 0 ? W / 4 : 1, 256), H);
    if (mask4 && cv_aligned(rgba, rgba_stride, 16) && cv_aligned(rgb, rgb_stride, 4))
        cv_rgba_split4<<<grid4, 256, 0, s>>>(rgba, rgba_stride, rgb, rgb_stride, mask, mask_stride, W / 4, H);
    else
        cv_rgba_split<<<grid, 256, 0, s>>>(rgba, rgba_stride, rgb, rgb_stride, mask, mask_stride, W, H);
    OFXCV_LAUNCH_CHECK(ctx);
    if (dilate_iterations > 0) {
        uint8_t* tmp = (uint8_t*)ofxcv_ws(ctx, WS_MISC0, (size_t)W * H);
        if (!tmp) return OFXCV_ERR_MEMORY;
        if (mask4 && dilate_iterations <= 4) {
            cv_dilate_rows4<<<grid4, 256, 0, s>>>(mask, mask_stride, tmp, W, W / 4, H, dilate_iterations);
            OFXCV_LAUNCH_CHECK(ctx);
            cv_dilate_cols4<<<grid4, 256, 0, s>>>(tmp, W, mask, mask_stride, W / 4, H, dilate_iterations);
        } else {
            cv_dilate_rows<<<grid, 256, 0, s>>>(mask, mask_stride, tmp, W, W, H, dilate_iterations);
            OFXCV_LAUNCH_CHECK(ctx);
            cv_dilate_cols<<<grid, 256, 0, s>>>(tmp, W, mask, mask_stride, W, H, dilate_iterations);
        }
        OFXCV_LAUNCH_CHECK(ctx);
    }
    return OFXCV_OK;
}

int ofxcv_rgb8_to_rgba8(ofxcv_ctx* ctx, ofxcv_stream stream, const uint8_t* rgb, ptrdiff_t rgb_stride, uint8_t* rgba,
                        ptrdiff_t rgba_stride, int W, int H)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    if (!rgb || !rgba || W <= 0 || H <= 0) return OFXCV_ERR_BAD_ARG;
    ofxcv_device_guard guard(ctx->device);
    if ((W & 3) == 0 && cv_aligned(rgb, rgb_stride, 4) && cv_aligned(rgba, rgba_stride, 16))
        cv_rgb_to_rgba4<<<dim3(ofxcv_div_up(W / 4, 256), H), 256, 0, pick(ctx, stream)>>>(rgb, rgb_stride, nullptr, 0, rgba, rgba_stride,
                                                                                         W / 4, H, 0, 0u);
    else
        cv_rgb_to_rgba<<<dim3(ofxcv_div_up(W, 256), H), 256, 0, pick(ctx, stream)>>>(rgb, rgb_stride, rgba, rgba_stride, W, H);
    OFXCV_LAUNCH_CHECK(ctx);
    return OFXCV_OK;
}

int ofxcv_rgb8_to_rgba8_noise(ofxcv_ctx* ctx, ofxcv_stream stream, const uint8_t* rgb, ptrdiff_t rgb_stride, const uint8_t* mask,
                              ptrdiff_t mask_stride, uint8_t* rgba, ptrdiff_t rgba_stride, int W, int H, int noise_div, unsigned seed)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    if (!rgb || !rgba || W <= 0 || H <= 0 || (noise_div > 0 && !mask)) return OFXCV_ERR_BAD_ARG;
    ofxcv_device_guard guard(ctx->device);
    if ((W & 3) == 0 && cv_aligned(rgb, rgb_stride, 4) && cv_aligned(rgba, rgba_stride, 16))
        cv_rgb_to_rgba4<<<dim3(ofxcv_div_up(W / 4, 256), H), 256, 0, pick(ctx, stream)>>>(rgb, rgb_stride, mask, mask_stride, rgba,
                                                                                         rgba_stride, W / 4, H, noise_div, seed);
    else
        cv_rgb_to_rgba_noise<<<dim3(ofxcv_div_up(W, 256), H), 256, 0, pick(ctx, stream)>>>(rgb, rgb_stride, mask, mask_stride, rgba,
                                                                                           rgba_stride, W, H, noise_div, seed);
    OFXCV_LAUNCH_CHECK(ctx);
    return OFXCV_OK;
}

int ofxcv_content_key_u8(ofxcv_ctx* ctx, ofxcv_stream stream, const uint8_t* img, ptrdiff_t stride, int W, int H, uint64_t* key)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    if (!img || !key || W <= 0 || H <= 0 || stride < W) return OFXCV_ERR_BAD_ARG;
    ofxcv_device_guard guard(ctx->device);
    cudaStream_t s = pick(ctx, stream);
    unsigned long long* acc = (unsigned long long*)ofxcv_ws(ctx, WS_KEY, 16);   // own slots: keys are taken on the staging stream too
    unsigned long long* host = (unsigned long long*)ofxcv_pin(ctx, 14, 16);
    if (!acc || !host) return OFXCV_ERR_MEMORY;
    OFXCV_CUDA(ctx, cudaMemsetAsync(acc, 0, 16, s));
    cv_content_key<<<dim3(ofxcv_div_up(W, 1024) > 0 ? ofxcv_div_up(W, 1024) : 1, H), 256, 0, s>>>(img, stride, W, H, acc);
    OFXCV_LAUNCH_CHECK(ctx);
    OFXCV_CUDA(ctx, cudaMemcpyAsync(host, acc, 16, cudaMemcpyDeviceToHost, s));
    OFXCV_CUDA(ctx, cudaStreamSynchronize(s));
    // fold the two lanes and the geometry; a collision needs both independent 64-bit sums to agree
    uint64_t a = (uint64_t)host[0], b = (uint64_t)host[1] + 0x9E3779B97F4A7C15ull * (((uint64_t)W << 32) | (uint64_t)H);
    b = (b ^ (b >> 29)) * 0xBF58476D1CE4E5B9ull;
    uint64_t k = a ^ (b ^ (b >> 32));
    *key = k ? k : 1;
    return OFXCV_OK;
}

int ofxcv_seed_grid(ofxcv_ctx* ctx, ofxcv_stream stream, int32_t* markers, ptrdiff_t markers_stride, int W, int H, int gx, int gy,
                    int half)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    if (!markers || W <= 0 || H <= 0 || gx < 1 || gy < 1 || half < 0 || (markers_stride & 3)) return OFXCV_ERR_BAD_ARG;
    ofxcv_device_guard guard(ctx->device);
    cv_seed_grid<<<dim3(ofxcv_div_up(W, 256), H), 256, 0, pick(ctx, stream)>>>(markers, markers_stride / 4, W, H, gx, gy, half);
    OFXCV_LAUNCH_CHECK(ctx);
    return OFXCV_OK;
}

int ofxcv_labels_to_rgba8(ofxcv_ctx* ctx, ofxcv_stream stream, const uint8_t* rgb, ptrdiff_t rgb_stride, const int32_t* labels,
                          ptrdiff_t labels_stride, uint8_t* rgba, ptrdiff_t rgba_stride, int W, int H, int nlabels)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    if (!rgb || !labels || !rgba || W <= 0 || H <= 0 || nlabels < 0 || (labels_stride & 3)) return OFXCV_ERR_BAD_ARG;
    ofxcv_device_guard guard(ctx->device);
    cudaStream_t s = pick(ctx, stream);
    size_t sb = ((size_t)nlabels + 1) * 4 * sizeof(unsigned long long);
    unsigned long long* sums = (unsigned long long*)ofxcv_ws(ctx, WS_MISC1, sb);
    if (!sums) return OFXCV_ERR_MEMORY;
    OFXCV_CUDA(ctx, cudaMemsetAsync(sums, 0, sb, s));
    dim3 grid(ofxcv_div_up(W, 256), H);
    cv_label_sums<<<grid, 256, 0, s>>>(rgb, rgb_stride, labels, labels_stride / 4, W, H, nlabels, sums);
    OFXCV_LAUNCH_CHECK(ctx);
    cv_label_paint<<<grid, 256, 0, s>>>(labels, labels_stride / 4, rgba, rgba_stride, W, H, nlabels, sums);
    OFXCV_LAUNCH_CHECK(ctx);
    return OFXCV_OK;
}

}  // extern "C"
