// Dual TV-L1 optical flow — the VectorGenerator plugin's second method
// (/root/reference/VectorGenerator/VectorGenerator.cpp:436-492: createOptFlow_DualTVL1(), setTau/Lambda/Theta/
//  ScalesNumber/WarpingsNumber/Epsilon/InnerIterations, calc(prev, next, flow); parameter defaults :874-929).
//
// The arithmetic is OpenCV's (not vendored by the reference, not in this image's cv2): PARITY UNPINNED at the level of
// the whole method.  the CPU test oracle (tvl1.c) restates it (its header says from what); this file follows the oracle operation for
// operation (compiled with -fmad=false) and the two agree bit for bit.  The three OpenCV primitives it is built from
// (bicubic remap on the 1/32-pixel grid, 5x5 median, bilinear resize) ARE pinned against cv2.
//
// HBM layout (per scale, row-major, no padding): I0, I1 float planes of the pyramid; J = float4 (I1, dI1/dx, dI1/dy, 0)
// so that one bicubic tap of all three warped images is one LDG.128; A = float4 (I1wx, I1wy, |grad|^2, rho_c) written
// once per warping and read once per inner iteration; U = float2 flow; P = float4 dual variable (p11, p12, p21, p22).
// One inner iteration = ONE launch of tv_iter (64 B/px: A 16 + U 8 + P 16 in, U 8 + P 16 out): a warp walks a 31-column
// strip (+1 halo lane) down a band of rows; the primal step of row y (threshold step + divergence of p + update of u +
// squared-update sum) and the dual step of row y-1 (forward gradient of the NEW u + update of p) share the trip, the
// vertical neighbours stay in registers, the horizontal ones come through warp shuffles, the next row is prefetched
// while the current one is computed.  U and P are ping-pong pairs (a band reads only old values, so bands and strips
// are independent); which half is current is a flag on the device, flipped by the last block of every launch that
// actually ran.  The convergence test (error <= epsilon^2 * area stops the warping) never comes back to the host: the
// last block adds the block partials in index order (f64, deterministic) and publishes the iteration at which to
// stop; later launches of the same warping read it and return at once.
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace {

struct TvCtrl {
    int stop_at;        // inner-iteration index (within the current warping) after which nothing runs; INT_MAX = none
    unsigned ticket;    // blocks of the running launch that have finished (the last one publishes / flips)
    int ucur, pcur;     // which half of the U / P ping-pong pair is current
    unsigned long long iters;  // inner iterations actually run (statistics)
    double err;         // squared update of the last iteration
};

__constant__ float c_cubic[32][4];  // OpenCV's bicubic coefficients (A = -0.75) at k/32

__global__ void __launch_bounds__(256) tv_u8_to_f32(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, ptrdiff_t stride,
                                                    float* __restrict__ fa, float* __restrict__ fb, int w, int h)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    fa[(size_t)y * w + x] = (float)a[(size_t)y * stride + x];
    fb[(size_t)y * w + x] = (float)b[(size_t)y * stride + x];
}

// source column/row and fraction of destination index d (cv::resize, INTER_LINEAR: coordinate kept in double until
// the fraction is taken; the fraction is zeroed where the 2-tap window leaves the image)
__device__ __forceinline__ void tv_lin(int d, double scale, int sn, int& s0, int& s1, float& f)
{
    const double fd = __dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
    int s = (int)floor(fd);
    f = (float)(fd - (double)s);
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= sn - 1) { f = 0.f; s = sn - 1; }
    s0 = s;
    s1 = s + 1 < sn ? s + 1 : sn - 1;
}

// bilinear resize of an NC-channel float plane, result multiplied by `mul` (1 for images, 1/scaleStep for the flow)
template <int NC>
__global__ void __launch_bounds__(256) tv_resize(const float* __restrict__ src, int sw, int sh, float* __restrict__ dst, int dw, int dh,
                                                 double scale_x, double scale_y, float mul, int do_mul)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw) return;
    int x0, x1, y0, y1;
    float a1, b1;
    tv_lin(x, scale_x, sw, x0, x1, a1);
    tv_lin(y, scale_y, sh, y0, y1, b1);
    const float a0 = 1.f - a1, b0 = 1.f - b1;
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const float r0 = src[((size_t)y0 * sw + x0) * NC + c] * a0 + src[((size_t)y0 * sw + x1) * NC + c] * a1;
        const float r1 = src[((size_t)y1 * sw + x0) * NC + c] * a0 + src[((size_t)y1 * sw + x1) * NC + c] * a1;
        float v = r0 * b0 + r1 * b1;
        if (do_mul) v *= mul;
        dst[((size_t)y * dw + x) * NC + c] = v;
    }
}

// J = (I1, centred dI1/dx, centred dI1/dy, 0), replicated border
__global__ void __launch_bounds__(256) tv_grad_pack(const float* __restrict__ I, float4* __restrict__ J, int w, int h)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const int xm = x > 0 ? x - 1 : 0, xp = x + 1 < w ? x + 1 : w - 1, ym = y > 0 ? y - 1 : 0, yp = y + 1 < h ? y + 1 : h - 1;
    const float* r = I + (size_t)y * w;
    J[(size_t)y * w + x] = make_float4(r[x], 0.5f * (r[xp] - r[xm]), 0.5f * (I[(size_t)yp * w + x] - I[(size_t)ym * w + x]), 0.f);
}

// one warping: bicubic remap of (I1, I1x, I1y) at (x + u1, y + u2) on the 1/32-pixel grid, constant border 0, then
// grad = I1wx^2 + I1wy^2 and rho_c = I1w - I1wx u1 - I1wy u2 - I0.  Also re-arms the convergence control block.
__global__ void __launch_bounds__(256) tv_warp(const float4* __restrict__ J, const float2* __restrict__ U0, const float2* __restrict__ U1,
                                               const float* __restrict__ I0, float4* __restrict__ A, int w, int h, TvCtrl* __restrict__ ctrl)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const float2* __restrict__ U = ctrl->ucur ? U1 : U0;
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        ctrl->stop_at = INT_MAX;
        ctrl->ticket = 0;
    }
    if (x >= w) return;
    const size_t o = (size_t)y * w + x;
    const float2 u = U[o];
    const int fx = __float2int_rn(((float)x + u.x) * 32.f), fy = __float2int_rn(((float)y + u.y) * 32.f);
    int ix = fx >> 5, iy = fy >> 5;
    ix = min(max(ix, -32768), 32767);  // OpenCV keeps the integer part as a saturated short
    iy = min(max(iy, -32768), 32767);
    const int sx = ix - 1, sy = iy - 1;
    float cx[4], cy[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        cx[k] = c_cubic[fx & 31][k];
        cy[k] = c_cubic[fy & 31][k];
    }
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (sx >= 0 && sx + 3 < w && sy >= 0 && sy + 3 < h) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float4* r = J + (size_t)(sy + i) * w + sx;
            const float4 t0 = r[0], t1 = r[1], t2 = r[2], t3 = r[3];
            const float w0 = cy[i] * cx[0], w1 = cy[i] * cx[1], w2 = cy[i] * cx[2], w3 = cy[i] * cx[3];
            s0 += t0.x * w0 + t1.x * w1 + t2.x * w2 + t3.x * w3;
            s1 += t0.y * w0 + t1.y * w1 + t2.y * w2 + t3.y * w3;
            s2 += t0.z * w0 + t1.z * w1 + t2.z * w2 + t3.z * w3;
        }
    } else if (!(sx >= w || sx + 4 <= 0 || sy >= h || sy + 4 <= 0)) {
        for (int i = 0; i < 4; i++) {
            const int yi = sy + i;
            if (yi < 0 || yi >= h) continue;
            for (int j = 0; j < 4; j++) {
                const int xj = sx + j;
                if (xj < 0 || xj >= w) continue;
                const float4 t = J[(size_t)yi * w + xj];
                const float wt = cy[i] * cx[j];
                s0 += t.x * wt;
                s1 += t.y * wt;
                s2 += t.z * wt;
            }
        }
    }
    A[o] = make_float4(s1, s2, s1 * s1 + s2 * s2, s0 - s1 * u.x - s2 * u.y - I0[o]);
}

// 5x5 median of both flow components (replicated border) through a 99-exchange selection network (the classic
// median-of-25 network; tests/test_abi.py re-verifies this very list on all 2^25 zero-one inputs, which by the zero-one
// principle proves it for every input; the median of a multiset does not depend on how it is found).  The launch
// does nothing once the warping has stopped.
__device__ __forceinline__ void tv_cswap(float& a, float& b)
{
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    a = lo;
    b = hi;
}
#define TV_MED25_NET \
    X(0, 1) X(3, 4) X(2, 4) X(2, 3) X(6, 7) X(5, 7) X(5, 6) X(9, 10) X(8, 10) \
    X(8, 9) X(12, 13) X(11, 13) X(11, 12) X(15, 16) X(14, 16) X(14, 15) X(18, 19) X(17, 19) \
    X(17, 18) X(21, 22) X(20, 22) X(20, 21) X(23, 24) X(2, 5) X(3, 6) X(0, 6) X(0, 3) \
    X(4, 7) X(1, 7) X(1, 4) X(11, 14) X(8, 14) X(8, 11) X(12, 15) X(9, 15) X(9, 12) \
    X(13, 16) X(10, 16) X(10, 13) X(20, 23) X(17, 23) X(17, 20) X(21, 24) X(18, 24) X(18, 21) \
    X(19, 22) X(8, 17) X(9, 18) X(0, 18) X(0, 9) X(10, 19) X(1, 19) X(1, 10) X(11, 20) \
    X(2, 20) X(2, 11) X(12, 21) X(3, 21) X(3, 12) X(13, 22) X(4, 22) X(4, 13) X(14, 23) \
    X(5, 23) X(5, 14) X(15, 24) X(6, 24) X(6, 15) X(7, 16) X(7, 19) X(13, 21) X(15, 23) \
    X(7, 13) X(7, 15) X(1, 9) X(3, 11) X(5, 17) X(11, 17) X(9, 17) X(4, 10) X(6, 12) \
    X(7, 14) X(4, 6) X(4, 7) X(12, 14) X(10, 14) X(6, 7) X(10, 12) X(6, 10) X(6, 17) \
    X(12, 17) X(7, 17) X(7, 10) X(12, 18) X(7, 12) X(10, 18) X(12, 20) X(10, 20) X(10, 12)
__device__ __forceinline__ float tv_median25(float (&v)[25])
{
#define X(i, j) tv_cswap(v[i], v[j]);
    TV_MED25_NET
#undef X
    return v[12];
}

__global__ void __launch_bounds__(256) tv_median5(float2* __restrict__ U0, float2* __restrict__ U1, int w, int h, int iter, TvCtrl* __restrict__ ctrl)
{
    if (iter > ctrl->stop_at) return;
    const int cur = ctrl->ucur;
    const float2* __restrict__ U = cur ? U1 : U0;
    float2* __restrict__ out = cur ? U0 : U1;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x < w) {
        float a[25], b[25];
#pragma unroll
        for (int dy = -2; dy <= 2; dy++)
#pragma unroll
            for (int dx = -2; dx <= 2; dx++) {
                const int yy = min(max(y + dy, 0), h - 1), xx = min(max(x + dx, 0), w - 1);
                const float2 v = U[(size_t)yy * w + xx];
                a[(dy + 2) * 5 + dx + 2] = v.x;
                b[(dy + 2) * 5 + dx + 2] = v.y;
            }
        out[(size_t)y * w + x] = make_float2(tv_median25(a), tv_median25(b));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&ctrl->ticket, 1u) == gridDim.x * gridDim.y - 1) {  // every block has read `ucur` by now
            ctrl->ticket = 0;
            ctrl->ucur = cur ^ 1;
        }
    }
}

// one inner iteration (see the header of this file).  Geometry: warp = strip of 31 columns (lane 31 = right halo: it
// computes the new u of the next strip's first column, which the dual step of column 30 needs), block = 8 adjacent
// strips x one band of `rb` rows (+ the new u of the row below the band, not stored).
constexpr int TV_STRIP = 31;
struct TvRow {
    float4 a, p;   // (I1wx, I1wy, grad, rho_c), (p11, p12, p21, p22) of this lane's pixel
    float2 u, pl;  // flow; (p11, p21) of the pixel left of the strip (lane 0 only)
};
__device__ __forceinline__ void tv_load_row(TvRow& r, const float4* __restrict__ A, const float4* __restrict__ P, const float2* __restrict__ U,
                                            int x, int y, int w, bool inx, int lane)
{
    r.a = r.p = make_float4(0.f, 0.f, 0.f, 0.f);
    r.u = r.pl = make_float2(0.f, 0.f);
    if (inx) {
        const size_t i = (size_t)y * w + x;
        r.a = __ldcs(&A[i]);
        r.p = P[i];
        r.u = U[i];
        if (lane == 0 && x > 0) {
            const float4 t = P[i - 1];
            r.pl = make_float2(t.x, t.z);
        }
    }
}
template <int MINB>
__global__ void __launch_bounds__(256, MINB) tv_iter(const float4* __restrict__ A, float4* __restrict__ P0, float4* __restrict__ P1, float2* __restrict__ U0,
                                               float2* __restrict__ U1, int w, int h, int rb, float l_t, float theta, float taut,
                                               float scaled_eps, int iter, TvCtrl* __restrict__ ctrl, double* __restrict__ partials)
{
    if (iter > ctrl->stop_at) return;
    const int ucur = ctrl->ucur, pcur = ctrl->pcur;
    const float2* __restrict__ Ui = ucur ? U1 : U0;
    float2* __restrict__ Uo = ucur ? U0 : U1;
    const float4* __restrict__ Pi = pcur ? P1 : P0;
    float4* __restrict__ Po = pcur ? P0 : P1;
    const int lane = threadIdx.x & 31, strip = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int x = strip * TV_STRIP + lane, y0 = blockIdx.y * rb, y1 = min(y0 + rb, h);
    const bool inx = x < w, owner = inx && lane < TV_STRIP;
    double e = 0.;
    if (strip * TV_STRIP < w) {
        float pu_y = 0.f, pu_w = 0.f;  // p12, p22 of the row above
        if (y0 > 0 && inx) {
            const float4 t = Pi[(size_t)(y0 - 1) * w + x];
            pu_y = t.y;
            pu_w = t.w;
        }
        TvRow nxt;
        tv_load_row(nxt, A, Pi, Ui, x, y0, w, inx, lane);
        float4 p_prev = make_float4(0.f, 0.f, 0.f, 0.f);
        float2 n_prev = make_float2(0.f, 0.f), nr_prev = make_float2(0.f, 0.f);
        for (int y = y0; y <= y1; y++) {
            const bool have = y < h;  // y == y1 is the row below the band (or below the image)
            const TvRow c = nxt;
            if (y + 1 <= y1 && y + 1 < h) tv_load_row(nxt, A, Pi, Ui, x, y + 1, w, inx, lane);
            float2 n = make_float2(0.f, 0.f);
            if (have) {
                // ---- primal step of row y: estimateV + divergence + estimateU
                float plx = __shfl_up_sync(0xffffffffu, c.p.x, 1), plz = __shfl_up_sync(0xffffffffu, c.p.z, 1);
                if (lane == 0) { plx = c.pl.x; plz = c.pl.y; }
                const float rho = c.a.w + (c.a.x * c.u.x + c.a.y * c.u.y);
                float d1 = 0.f, d2 = 0.f;
                const float lg = l_t * c.a.z;
                if (rho < -lg) { d1 = l_t * c.a.x; d2 = l_t * c.a.y; }
                else if (rho > lg) { d1 = -l_t * c.a.x; d2 = -l_t * c.a.y; }
                else if (c.a.z > FLT_EPSILON) { const float fi = -rho / c.a.z; d1 = fi * c.a.x; d2 = fi * c.a.y; }
                const float v1 = c.u.x + d1, v2 = c.u.y + d2;
                const float a1 = x > 0 ? c.p.x - plx : c.p.x, a2 = x > 0 ? c.p.z - plz : c.p.z;
                const float b1 = y > 0 ? c.p.y - pu_y : c.p.y, b2 = y > 0 ? c.p.w - pu_w : c.p.w;
                n = make_float2(v1 + theta * (a1 + b1), v2 + theta * (a2 + b2));
                if (owner && y < y1) {
                    const float e1 = n.x - c.u.x, e2 = n.y - c.u.y;
                    e += (double)(e1 * e1 + e2 * e2);
                    Uo[(size_t)y * w + x] = n;
                }
            }
            if (y > y0) {
                // ---- dual step of row y-1: forward gradient of the new u + estimateDualVariables
                float u1x = 0.f, u2x = 0.f, u1y = 0.f, u2y = 0.f;
                if (x + 1 < w) { u1x = nr_prev.x - n_prev.x; u2x = nr_prev.y - n_prev.y; }
                if (have) { u1y = n.x - n_prev.x; u2y = n.y - n_prev.y; }
                const float g1 = (float)sqrt((double)u1x * (double)u1x + (double)u1y * (double)u1y);
                const float g2 = (float)sqrt((double)u2x * (double)u2x + (double)u2y * (double)u2y);
                const float ng1 = 1.f + taut * g1, ng2 = 1.f + taut * g2;
                if (owner)
                    Po[(size_t)(y - 1) * w + x] = make_float4((p_prev.x + taut * u1x) / ng1, (p_prev.y + taut * u1y) / ng1,
                                                              (p_prev.z + taut * u2x) / ng2, (p_prev.w + taut * u2y) / ng2);
            }
            p_prev = c.p;
            pu_y = c.p.y;
            pu_w = c.p.w;
            n_prev = n;
            nr_prev = make_float2(__shfl_down_sync(0xffffffffu, n.x, 1), __shfl_down_sync(0xffffffffu, n.y, 1));
        }
    }
    // block sum -> partials[block]; the last block to arrive adds the partials in index order, publishes, flips
    __shared__ double wsum[8];
    __shared__ bool last;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) e += __shfl_down_sync(0xffffffffu, e, d);
    if (lane == 0) wsum[threadIdx.x >> 5] = e;
    __syncthreads();
    const unsigned nblocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
    if (threadIdx.x == 0) {
        double s = 0.;
#pragma unroll
        for (int k = 0; k < 8; k++) s += wsum[k];
        partials[bid] = s;
        __threadfence();
        last = atomicAdd(&ctrl->ticket, 1u) == nblocks - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    double s = 0.;
    for (unsigned k = threadIdx.x; k < nblocks; k += blockDim.x) s += __ldcg(&partials[k]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
    if (lane == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.;
#pragma unroll
        for (int k = 0; k < 8; k++) t += wsum[k];
        ctrl->err = t;
        ctrl->iters += 1;
        ctrl->ticket = 0;
        ctrl->ucur = ucur ^ 1;  // every block has read the flags by now
        ctrl->pcur = pcur ^ 1;
        if (!(t > (double)scaled_eps)) ctrl->stop_at = iter;
    }
}

// out = current half of the U pair (the final flow, or the source of the resize to the next finer scale)
__global__ void __launch_bounds__(256) tv_store_flow(const float2* __restrict__ U0, const float2* __restrict__ U1, const TvCtrl* __restrict__ ctrl,
                                                     char* __restrict__ flow, ptrdiff_t flow_stride, int w, int h)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const float2* __restrict__ U = ctrl->ucur ? U1 : U0;
    reinterpret_cast<float2*>(flow + (ptrdiff_t)y * flow_stride)[x] = U[(size_t)y * w + x];
}

cudaStream_t pick(ofxcv_ctx* ctx, ofxcv_stream s) { return s ? (cudaStream_t)s : ctx->stream; }

constexpr int TV_MAXS = 32;
struct TvPlan {
    int ns, w[TV_MAXS], h[TV_MAXS];
    size_t off_i0[TV_MAXS], off_i1[TV_MAXS];  // float offsets into the arena
    size_t off_j, off_a, off_u, off_u2, off_p, off_p2, off_uc, off_part, off_ctrl, total;  // byte offsets
};

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

int tv_plan(int W, int H, const ofxcv_tvl1_params* p, TvPlan* pl)
{
    int ns = p->nscales < 1 ? 1 : p->nscales > TV_MAXS ? TV_MAXS : p->nscales;
    pl->w[0] = W;
    pl->h[0] = H;
    int built = 1;
    for (int s = 1; s < ns; s++) {
        const int cw = (int)lrint(pl->w[s - 1] * p->scale_step), ch = (int)lrint(pl->h[s - 1] * p->scale_step);
        if (cw < 16 || ch < 16) break;
        pl->w[s] = cw;
        pl->h[s] = ch;
        built = s + 1;
    }
    pl->ns = built;
    size_t off = 0;
    for (int s = 0; s < built; s++) {
        const size_t n = (size_t)pl->w[s] * pl->h[s];
        pl->off_i0[s] = off; off = align256(off + n * 4);
        pl->off_i1[s] = off; off = align256(off + n * 4);
    }
    const size_t n0 = (size_t)W * H;
    pl->off_j = off; off = align256(off + n0 * 16);
    pl->off_a = off; off = align256(off + n0 * 16);
    pl->off_p = off; off = align256(off + n0 * 16);
    pl->off_p2 = off; off = align256(off + n0 * 16);
    pl->off_u = off; off = align256(off + n0 * 8);
    pl->off_u2 = off; off = align256(off + n0 * 8);
    pl->off_uc = off; off = align256(off + n0 * 8);  // flow of the coarser scale (resize source)
    pl->off_part = off; off = align256(off + (size_t)(ofxcv_div_up(W, 8 * TV_STRIP) + 1) * (H / 4 + 2) * 8);  // one per tv_iter block
    pl->off_ctrl = off; off = align256(off + sizeof(TvCtrl));
    pl->total = off;
    return built;
}

}  // namespace

extern "C" {

void ofxcv_tvl1_default_params(ofxcv_tvl1_params* p)
{
    if (!p) return;
    p->tau = 0.25;
    p->lambda = 0.15;
    p->theta = 0.3;
    p->epsilon = 0.01;
    p->nscales = 5;
    p->warps = 5;
    p->iterations = 15;
    p->outer_iterations = 10;
    p->scale_step = 0.8;
    p->median_filtering = 5;
}

int ofxcv_tvl1_scales(int W, int H, const ofxcv_tvl1_params* params)
{
    ofxcv_tvl1_params d;
    ofxcv_tvl1_default_params(&d);
    TvPlan pl;
    return W > 0 && H > 0 ? tv_plan(W, H, params ? params : &d, &pl) : 0;
}

size_t ofxcv_tvl1_workspace_bytes(int W, int H, const ofxcv_tvl1_params* params)
{
    ofxcv_tvl1_params d;
    ofxcv_tvl1_default_params(&d);
    TvPlan pl;
    if (W <= 0 || H <= 0) return 0;
    tv_plan(W, H, params ? params : &d, &pl);
    return pl.total;
}

double ofxcv_tvl1_iter_bytes(int W, int H) { return 64.0 * (double)W * (double)H; }

int ofxcv_tvl1_u8(ofxcv_ctx* ctx, ofxcv_stream stream, const uint8_t* prev, const uint8_t* next, ptrdiff_t stride, int W, int H,
                  float* flow, ptrdiff_t flow_stride, const ofxcv_tvl1_params* params)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    ofxcv_tvl1_params P;
    ofxcv_tvl1_default_params(&P);
    if (params) P = *params;
    if (!prev || !next || !flow || W <= 0 || H <= 0 || stride < W || (flow_stride & 7) || ((uintptr_t)flow & 7)) return OFXCV_ERR_BAD_ARG;
    if (!(P.tau > 0) || !(P.theta > 0) || !(P.lambda > 0) || !(P.scale_step > 0 && P.scale_step < 1) || P.warps < 1 || P.iterations < 1 ||
        P.outer_iterations < 1 || P.nscales < 1 || !(P.median_filtering <= 1 || P.median_filtering == 5))
        return OFXCV_ERR_BAD_ARG;
    ofxcv_device_guard guard(ctx->device);
    cudaStream_t s = pick(ctx, stream);
    static bool tab_ready[64] = {false};  // per device; written once with the same bytes, so a race is benign
    if (ctx->device < 0 || ctx->device >= 64 || !tab_ready[ctx->device]) {
        float tab[32][4];
        const float A = -0.75f;
        for (int i = 0; i < 32; i++) {  // cv::interpolateCubic at i/32
            const float x = i * (1.f / 32);
            tab[i][0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
            tab[i][1] = ((A + 2) * x - (A + 3)) * x * x + 1;
            tab[i][2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
            tab[i][3] = 1.f - tab[i][0] - tab[i][1] - tab[i][2];
        }
        OFXCV_CUDA(ctx, cudaMemcpyToSymbolAsync(c_cubic, tab, sizeof(tab), 0, cudaMemcpyHostToDevice, s));
        OFXCV_CUDA(ctx, cudaStreamSynchronize(s));  // `tab` is on the stack
        if (ctx->device >= 0 && ctx->device < 64) tab_ready[ctx->device] = true;
    }
    TvPlan pl;
    tv_plan(W, H, &P, &pl);
    char* base = (char*)ofxcv_ws(ctx, WS_TV_ARENA, pl.total);
    if (!base) return OFXCV_ERR_MEMORY;
    ctx->tv_ctrl_off = pl.off_ctrl;
    float4* J = (float4*)(base + pl.off_j);
    float4* A = (float4*)(base + pl.off_a);
    float4* Pa = (float4*)(base + pl.off_p);
    float4* Pb = (float4*)(base + pl.off_p2);
    float2* Ua = (float2*)(base + pl.off_u);
    float2* Ub = (float2*)(base + pl.off_u2);
    float2* Uc = (float2*)(base + pl.off_uc);
    double* partials = (double*)(base + pl.off_part);
    TvCtrl* ctrl = (TvCtrl*)(base + pl.off_ctrl);
    auto I0 = [&](int k) { return (float*)(base + pl.off_i0[k]); };
    auto I1 = [&](int k) { return (float*)(base + pl.off_i1[k]); };
    auto grid = [](int w, int h) { return dim3(ofxcv_div_up(w, 256), h); };

    // The convergence flag lives on the device and the launches of a stopped warping return at once, but thousands of
    // no-op launches still cost microseconds each.  The host therefore paces itself: after every outer iteration the
    // flag is copied to pinned memory behind an event, and before enqueuing outer iteration n it waits for the copy
    // made after n-2.  The GPU always has a full outer iteration queued while the host waits, and a stopped warping
    // wastes at most the rest of that outer iteration and the next one.
    constexpr int TV_NEV = 16;
    int* h_stop = (int*)ofxcv_pin(ctx, 12, TV_NEV * sizeof(int));
    if (!h_stop) return OFXCV_ERR_MEMORY;
    for (auto& e : ctx->tv_ev)
        if (!e) OFXCV_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ofxcv_prof_scope ps_all(ctx, s, "tv_total", 0);
    OFXCV_CUDA(ctx, cudaMemsetAsync(ctrl, 0, sizeof(TvCtrl), s));
    tv_u8_to_f32<<<grid(W, H), 256, 0, s>>>(prev, next, stride, I0(0), I1(0), W, H);
    OFXCV_LAUNCH_CHECK(ctx);
    const double inv = 1. / P.scale_step;
    for (int k = 1; k < pl.ns; k++) {
        tv_resize<1><<<grid(pl.w[k], pl.h[k]), 256, 0, s>>>(I0(k - 1), pl.w[k - 1], pl.h[k - 1], I0(k), pl.w[k], pl.h[k], inv, inv, 1.f, 0);
        OFXCV_LAUNCH_CHECK(ctx);
        tv_resize<1><<<grid(pl.w[k], pl.h[k]), 256, 0, s>>>(I1(k - 1), pl.w[k - 1], pl.h[k - 1], I1(k), pl.w[k], pl.h[k], inv, inv, 1.f, 0);
        OFXCV_LAUNCH_CHECK(ctx);
    }
    const float l_t = (float)(P.lambda * P.theta), taut = (float)(P.tau / P.theta), theta = (float)P.theta;
    const float up = (float)(1. / P.scale_step);
    for (int k = pl.ns - 1; k >= 0; k--) {
        const int w = pl.w[k], h = pl.h[k];
        const size_t n = (size_t)w * h;
        const dim3 g = grid(w, h);
        const float scaled_eps = (float)(P.epsilon * P.epsilon * (double)n);
        // rows per band of tv_iter: short enough for ~48 warps per SM (a band costs one redundant halo row), 4..16
        // (measured at 4K: 121 us for 12-16 rows, 126-128 us for 8 or 24-32 rows)
        const long long strips = ofxcv_div_up(w, TV_STRIP);
        int rb = (int)std::min<long long>(16, std::max<long long>(4, (long long)h * strips / (48LL * ctx->num_sms)));
        static const int env_rb = getenv("OFXCV_TV_RB") ? atoi(getenv("OFXCV_TV_RB")) : 0;       // experiments
        static const int env_minb = getenv("OFXCV_TV_MINB") ? atoi(getenv("OFXCV_TV_MINB")) : 4;
        if (env_rb > 0) rb = env_rb;
        const dim3 gi(ofxcv_div_up(ofxcv_div_up(w, TV_STRIP), 8), ofxcv_div_up(h, rb));
        if (k == pl.ns - 1) {
            OFXCV_CUDA(ctx, cudaMemsetAsync(Ua, 0, n * 8, s));
        } else {  // flow of the coarser scale (current half -> Uc), resized and multiplied by 1 / scaleStep
            const int cw = pl.w[k + 1], ch = pl.h[k + 1];
            tv_store_flow<<<grid(cw, ch), 256, 0, s>>>(Ua, Ub, ctrl, (char*)Uc, (ptrdiff_t)cw * 8, cw, ch);
            OFXCV_LAUNCH_CHECK(ctx);
            tv_resize<2><<<g, 256, 0, s>>>((const float*)Uc, cw, ch, (float*)Ua, w, h, 1. / ((double)w / cw), 1. / ((double)h / ch), up, 1);
            OFXCV_LAUNCH_CHECK(ctx);
        }
        tv_grad_pack<<<g, 256, 0, s>>>(I1(k), J, w, h);
        OFXCV_LAUNCH_CHECK(ctx);
        OFXCV_CUDA(ctx, cudaMemsetAsync(Pa, 0, n * 16, s));
        OFXCV_CUDA(ctx, cudaMemsetAsync(&ctrl->ucur, 0, 2 * sizeof(int), s));  // current halves: Ua, Pa
        for (int wi = 0; wi < P.warps; wi++) {
            {
                ofxcv_prof_scope ps(ctx, s, "tv_warp", k);
                tv_warp<<<g, 256, 0, s>>>(J, Ua, Ub, I0(k), A, w, h, ctrl);
                OFXCV_LAUNCH_CHECK(ctx);
            }
            int it = 0;
            for (int no = 0; no < P.outer_iterations; no++) {
                if (no >= 2) {
                    OFXCV_CUDA(ctx, cudaEventSynchronize(ctx->tv_ev[(no - 2) % TV_NEV]));
                    if (h_stop[(no - 2) % TV_NEV] != INT_MAX) break;  // everything still queued returns at once
                }
                if (P.median_filtering > 1) {
                    ofxcv_prof_scope ps(ctx, s, "tv_median5", k);
                    tv_median5<<<g, 256, 0, s>>>(Ua, Ub, w, h, it, ctrl);
                    OFXCV_LAUNCH_CHECK(ctx);
                }
                ofxcv_prof_scope ps(ctx, s, "tv_iter", k);
                for (int ni = 0; ni < P.iterations; ni++, it++) {
                    const bool timed = ctx->timing && k == 0;
                    if (timed) ofxcv_time_begin(ctx, 1, s);
                    if (env_minb == 5)
                        tv_iter<5><<<gi, 256, 0, s>>>(A, Pa, Pb, Ua, Ub, w, h, rb, l_t, theta, taut, scaled_eps, it, ctrl, partials);
                    else
                        tv_iter<4><<<gi, 256, 0, s>>>(A, Pa, Pb, Ua, Ub, w, h, rb, l_t, theta, taut, scaled_eps, it, ctrl, partials);
                    OFXCV_LAUNCH_CHECK(ctx);
                    if (timed) ofxcv_time_end(ctx, 1, s);
                }
                if (no + 2 < P.outer_iterations) {
                    const int slot = no % TV_NEV;
                    OFXCV_CUDA(ctx, cudaMemcpyAsync(&h_stop[slot], &ctrl->stop_at, sizeof(int), cudaMemcpyDeviceToHost, s));
                    OFXCV_CUDA(ctx, cudaEventRecord(ctx->tv_ev[slot], s));
                }
            }
        }
    }
    tv_store_flow<<<grid(W, H), 256, 0, s>>>(Ua, Ub, ctrl, (char*)flow, flow_stride, W, H);
    OFXCV_LAUNCH_CHECK(ctx);
    return OFXCV_OK;
}

// inner iterations run by the last ofxcv_tvl1_u8 on this context (synchronises the context's stream)
int64_t ofxcv_tvl1_iterations_run(ofxcv_ctx* ctx)
{
    if (!ctx || !ctx->ws[WS_TV_ARENA].p) return -1;
    ofxcv_device_guard guard(ctx->device);
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    TvCtrl c;
    if (cudaMemcpy(&c, (char*)ctx->ws[WS_TV_ARENA].p + ctx->tv_ctrl_off, sizeof(c), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int64_t)c.iters;
}

}  // extern "C"
