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
// One inner iteration = two element-wise stencil kernels (88 B/px): tv_iter_u (threshold step + divergence + primal
// update + squared-update partial sums) and tv_iter_p (forward gradient + dual update).  The convergence test
// (error <= epsilon^2 * area stops the warping) never comes back to the host: the last block of tv_iter_u adds the
// block partials in index order (f64, deterministic) and publishes the iteration at which to stop; later launches of
// the same warping read it and return at once.
#include <float.h>
#include <limits.h>
#include <math.h>

#include "common.cuh"

namespace {

struct TvCtrl {
    int stop_at;        // inner-iteration index (within the current warping) after which nothing runs; INT_MAX = none
    unsigned ticket;    // blocks of tv_iter_u that have delivered their partial sum
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
__global__ void __launch_bounds__(256) tv_warp(const float4* __restrict__ J, const float2* __restrict__ U, const float* __restrict__ I0,
                                               float4* __restrict__ A, int w, int h, TvCtrl* __restrict__ ctrl)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
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

// 5x5 median of both flow components (replicated border).  rank selection: the median is the value with exactly 12
// predecessors in the order (value, window index).  Passes the flow through unchanged once the warping has stopped,
// so that the host can swap the two buffers unconditionally.
__global__ void __launch_bounds__(256) tv_median5(const float2* __restrict__ U, float2* __restrict__ out, int w, int h, int iter,
                                                  const TvCtrl* __restrict__ ctrl)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const size_t o = (size_t)y * w + x;
    if (iter > ctrl->stop_at) {
        out[o] = U[o];
        return;
    }
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
    float ma = a[12], mb = b[12];
#pragma unroll
    for (int i = 0; i < 25; i++) {
        int ra = 0, rb = 0;
#pragma unroll
        for (int j = 0; j < 25; j++) {
            ra += (a[j] < a[i]) || (a[j] == a[i] && j < i);
            rb += (b[j] < b[i]) || (b[j] == b[i] && j < i);
        }
        if (ra == 12) ma = a[i];
        if (rb == 12) mb = b[i];
    }
    out[o] = make_float2(ma, mb);
}

// threshold step (estimateV) + divergence of p + primal update (estimateU), in place on U; squared update summed in f64
__global__ void __launch_bounds__(256) tv_iter_u(const float4* __restrict__ A, const float4* __restrict__ P, float2* __restrict__ U, int w, int h,
                                                 float l_t, float theta, float scaled_eps, int iter, TvCtrl* __restrict__ ctrl,
                                                 double* __restrict__ partials)
{
    if (iter > ctrl->stop_at) return;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    double e = 0.;
    if (x < w) {
        const size_t i = (size_t)y * w + x;
        const float4 a = A[i];  // I1wx, I1wy, grad, rho_c
        const float2 u = U[i];
        const float4 p = P[i];
        const float rho = a.w + (a.x * u.x + a.y * u.y);
        float d1 = 0.f, d2 = 0.f;
        const float lg = l_t * a.z;
        if (rho < -lg) { d1 = l_t * a.x; d2 = l_t * a.y; }
        else if (rho > lg) { d1 = -l_t * a.x; d2 = -l_t * a.y; }
        else if (a.z > FLT_EPSILON) { const float fi = -rho / a.z; d1 = fi * a.x; d2 = fi * a.y; }
        const float v1 = u.x + d1, v2 = u.y + d2;
        float a1 = p.x, b1 = p.y, a2 = p.z, b2 = p.w;
        if (x > 0) { const float4 pl = P[i - 1]; a1 = p.x - pl.x; a2 = p.z - pl.z; }
        if (y > 0) { const float4 pu = P[i - w]; b1 = p.y - pu.y; b2 = p.w - pu.w; }
        const float n1 = v1 + theta * (a1 + b1), n2 = v2 + theta * (a2 + b2);
        const float e1 = n1 - u.x, e2 = n2 - u.y;
        e = (double)(e1 * e1 + e2 * e2);
        U[i] = make_float2(n1, n2);
    }
    // block sum -> partials[block]; the last block to arrive adds the partials in index order
    __shared__ double wsum[8];
    __shared__ bool last;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) e += __shfl_down_sync(0xffffffffu, e, d);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = e;
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
    // fixed-shape tree over the 256 strided sums
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.;
#pragma unroll
        for (int k = 0; k < 8; k++) t += wsum[k];
        ctrl->err = t;
        ctrl->iters += 1;
        ctrl->ticket = 0;
        if (!(t > (double)scaled_eps)) ctrl->stop_at = iter;  // this iteration's dual update still runs
    }
}

// forward gradient of u + dual update, in place on P
__global__ void __launch_bounds__(256) tv_iter_p(const float2* __restrict__ U, float4* __restrict__ P, int w, int h, float taut, int iter,
                                                 const TvCtrl* __restrict__ ctrl)
{
    if (iter > ctrl->stop_at) return;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const size_t i = (size_t)y * w + x;
    const float2 u = U[i];
    float u1x = 0.f, u2x = 0.f, u1y = 0.f, u2y = 0.f;
    if (x + 1 < w) { const float2 r = U[i + 1]; u1x = r.x - u.x; u2x = r.y - u.y; }
    if (y + 1 < h) { const float2 d = U[i + w]; u1y = d.x - u.x; u2y = d.y - u.y; }
    const float g1 = (float)sqrt((double)u1x * (double)u1x + (double)u1y * (double)u1y);
    const float g2 = (float)sqrt((double)u2x * (double)u2x + (double)u2y * (double)u2y);
    const float ng1 = 1.f + taut * g1, ng2 = 1.f + taut * g2;
    float4 p = P[i];
    p.x = (p.x + taut * u1x) / ng1;
    p.y = (p.y + taut * u1y) / ng1;
    p.z = (p.z + taut * u2x) / ng2;
    p.w = (p.w + taut * u2y) / ng2;
    P[i] = p;
}

__global__ void __launch_bounds__(256) tv_store_flow(const float2* __restrict__ U, char* __restrict__ flow, ptrdiff_t flow_stride, int w, int h)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    reinterpret_cast<float2*>(flow + (ptrdiff_t)y * flow_stride)[x] = U[(size_t)y * w + x];
}

cudaStream_t pick(ofxcv_ctx* ctx, ofxcv_stream s) { return s ? (cudaStream_t)s : ctx->stream; }

constexpr int TV_MAXS = 32;
struct TvPlan {
    int ns, w[TV_MAXS], h[TV_MAXS];
    size_t off_i0[TV_MAXS], off_i1[TV_MAXS];  // float offsets into the arena
    size_t off_j, off_a, off_u, off_u2, off_p, off_uc, off_part, off_ctrl, total;  // byte offsets
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
    pl->off_u = off; off = align256(off + n0 * 8);
    pl->off_u2 = off; off = align256(off + n0 * 8);
    pl->off_uc = off; off = align256(off + n0 * 8);  // flow of the coarser scale (resize source)
    pl->off_part = off; off = align256(off + (size_t)ofxcv_div_up(W, 256) * H * 8);
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

double ofxcv_tvl1_iter_bytes(int W, int H) { return 88.0 * (double)W * (double)H; }

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
    float4* Pd = (float4*)(base + pl.off_p);
    float2* U = (float2*)(base + pl.off_u);
    float2* U2 = (float2*)(base + pl.off_u2);
    float2* Uc = (float2*)(base + pl.off_uc);
    double* partials = (double*)(base + pl.off_part);
    TvCtrl* ctrl = (TvCtrl*)(base + pl.off_ctrl);
    auto I0 = [&](int k) { return (float*)(base + pl.off_i0[k]); };
    auto I1 = [&](int k) { return (float*)(base + pl.off_i1[k]); };
    auto grid = [](int w, int h) { return dim3(ofxcv_div_up(w, 256), h); };

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
        if (k == pl.ns - 1) {
            OFXCV_CUDA(ctx, cudaMemsetAsync(U, 0, n * 8, s));
        } else {  // flow of the coarser scale, resized and multiplied by 1 / scaleStep
            const int cw = pl.w[k + 1], ch = pl.h[k + 1];
            OFXCV_CUDA(ctx, cudaMemcpyAsync(Uc, U, (size_t)cw * ch * 8, cudaMemcpyDeviceToDevice, s));
            tv_resize<2><<<g, 256, 0, s>>>((const float*)Uc, cw, ch, (float*)U, w, h, 1. / ((double)w / cw), 1. / ((double)h / ch), up, 1);
            OFXCV_LAUNCH_CHECK(ctx);
        }
        tv_grad_pack<<<g, 256, 0, s>>>(I1(k), J, w, h);
        OFXCV_LAUNCH_CHECK(ctx);
        OFXCV_CUDA(ctx, cudaMemsetAsync(Pd, 0, n * 16, s));
        for (int wi = 0; wi < P.warps; wi++) {
            {
                ofxcv_prof_scope ps(ctx, s, "tv_warp", k);
                tv_warp<<<g, 256, 0, s>>>(J, U, I0(k), A, w, h, ctrl);
                OFXCV_LAUNCH_CHECK(ctx);
            }
            int it = 0;
            for (int no = 0; no < P.outer_iterations; no++) {
                if (P.median_filtering > 1) {
                    ofxcv_prof_scope ps(ctx, s, "tv_median5", k);
                    tv_median5<<<g, 256, 0, s>>>(U, U2, w, h, it, ctrl);
                    OFXCV_LAUNCH_CHECK(ctx);
                    float2* t = U; U = U2; U2 = t;
                }
                ofxcv_prof_scope ps(ctx, s, "tv_iter", k);
                for (int ni = 0; ni < P.iterations; ni++, it++) {
                    const bool timed = ctx->timing && k == 0;
                    if (timed) ofxcv_time_begin(ctx, 1, s);
                    tv_iter_u<<<g, 256, 0, s>>>(A, Pd, U, w, h, l_t, theta, scaled_eps, it, ctrl, partials);
                    OFXCV_LAUNCH_CHECK(ctx);
                    tv_iter_p<<<g, 256, 0, s>>>(U, Pd, w, h, taut, it, ctrl);
                    OFXCV_LAUNCH_CHECK(ctx);
                    if (timed) ofxcv_time_end(ctx, 1, s);
                }
            }
        }
    }
    tv_store_flow<<<grid(W, H), 256, 0, s>>>(U, (char*)flow, flow_stride, W, H);
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
