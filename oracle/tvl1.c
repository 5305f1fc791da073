/* CPU ORACLE (test infrastructure only - never linked into the product library, never timed as the product).
 *
 * Dual TV-L1 optical flow, the second method the VectorGenerator plugin exposes
 * (/root/reference/VectorGenerator/VectorGenerator.cpp:436-492: createOptFlow_DualTVL1(), setTau/Lambda/Theta/
 * ScalesNumber/WarpingsNumber/Epsilon/InnerIterations, calc(prev, next, flow); parameter defaults :874-929).
 *
 * PARITY UNPINNED.  The arithmetic lives in OpenCV (`video` in 2.4/3.x, contrib `optflow` in 4.x), which is neither
 * vendored by the reference nor present in this image (cv2 4.13 main modules have no DualTVL1).  This file restates
 * the published algorithm - Zach, Pock, Bischof, "A duality based approach for realtime TV-L1 optical flow" (DAGM
 * 2007) in the formulation of Sanchez, Meinhardt-Llopis, Facciolo, "TV-L1 Optical Flow Estimation" (IPOL 2013) - with
 * the structure of OpenCV's implementation: pyramid by bilinear resize with scale step 0.8, `warps` warpings per scale
 * (bicubic remap of I1 and of its centred gradient on a 1/32-pixel grid, constant-0 border), `outer` x `inner`
 * iterations per warping with a 5x5 median filter of the flow per outer iteration, early exit when the squared
 * update falls below epsilon^2 * area, gamma = 0 (no illumination term).  What IS pinned: the three OpenCV primitives
 * it is assembled from - cv2.remap(INTER_CUBIC), cv2.medianBlur(5) and cv2.resize(INTER_LINEAR) on float images -
 * are compared with orc_remap_cubic_f32 / orc_median5_f32 / orc_tvl1_resize_f32 in tests/test_oracle_golden.py.
 *
 * Every expression below fixes an evaluation order; the CUDA path (csrc/tvl1.cu) follows it operation for operation
 * (no FMA contraction on either side), so the two agree bit for bit.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    double tau, lambda, theta, epsilon, scale_step;
    int nscales, warps, inner, outer, median; /* median: 5 -> 5x5 median per outer iteration, <= 1 -> none */
} orc_tvl1_params;

/* ---- bilinear resize, OpenCV convention: src coordinate = (d + 0.5) * scale - 0.5, scale = 1 / (dst/src ratio) ---- */
static void lin_coeffs(int sn, int dn, double scale, int* ofs, float* a)
{
    for (int d = 0; d < dn; d++) {
        double fd = (d + 0.5) * scale - 0.5; /* cv2 4.13 keeps the coordinate in double until the fraction is taken */
        int s = (int)floor(fd);
        float f = (float)(fd - s);
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= sn - 1) { f = 0.f; s = sn - 1; }
        ofs[d] = s; a[d] = f;
    }
}
void orc_tvl1_resize_f32(const float* src, int sw, int sh, float* dst, int dw, int dh, double scale_x, double scale_y)
{
    int* xo = (int*)malloc(dw * sizeof(int)); int* yo = (int*)malloc(dh * sizeof(int));
    float* xa = (float*)malloc(dw * sizeof(float)); float* ya = (float*)malloc(dh * sizeof(float));
    lin_coeffs(sw, dw, scale_x, xo, xa);
    lin_coeffs(sh, dh, scale_y, yo, ya);
    for (int y = 0; y < dh; y++) {
        int y0 = yo[y], y1 = y0 + 1 < sh ? y0 + 1 : sh - 1;
        float b1 = ya[y], b0 = 1.f - b1;
        for (int x = 0; x < dw; x++) {
            int x0 = xo[x], x1 = x0 + 1 < sw ? x0 + 1 : sw - 1;
            float a1 = xa[x], a0 = 1.f - a1;
            float r0 = src[(size_t)y0 * sw + x0] * a0 + src[(size_t)y0 * sw + x1] * a1;
            float r1 = src[(size_t)y1 * sw + x0] * a0 + src[(size_t)y1 * sw + x1] * a1;
            dst[(size_t)y * dw + x] = r0 * b0 + r1 * b1;
        }
    }
    free(xo); free(yo); free(xa); free(ya);
}

/* ---- bicubic remap on OpenCV's 1/32-pixel grid (A = -0.75), constant border 0 ---- */
void orc_cubic_tab(float tab[32][4])
{
    const float A = -0.75f;
    for (int i = 0; i < 32; i++) {
        float x = i * (1.f / 32);
        tab[i][0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
        tab[i][1] = ((A + 2) * x - (A + 3)) * x * x + 1;
        tab[i][2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
        tab[i][3] = 1.f - tab[i][0] - tab[i][1] - tab[i][2];
    }
}
static inline int round_half_even(float v) { return (int)lrintf(v); } /* default rounding mode = cvRound */

/* dst(y,x) = bicubic(src at (mapx, mapy)); nsrc planes share the maps (I1, I1x, I1y) */
void orc_remap_cubic_f32(const float* const* src, int nsrc, int w, int h, const float* mapx, const float* mapy, float* const* dst,
                         int dw, int dh)
{
    float tab[32][4];
    orc_cubic_tab(tab);
    for (int y = 0; y < dh; y++)
        for (int x = 0; x < dw; x++) {
            size_t o = (size_t)y * dw + x;
            int fx = round_half_even(mapx[o] * 32.f), fy = round_half_even(mapy[o] * 32.f);
            int ix = fx >> 5, iy = fy >> 5; /* OpenCV keeps the integer part as a saturated short */
            ix = ix < -32768 ? -32768 : ix > 32767 ? 32767 : ix;
            iy = iy < -32768 ? -32768 : iy > 32767 ? 32767 : iy;
            const int sx = ix - 1, sy = iy - 1;
            const float* cx = tab[fx & 31]; const float* cy = tab[fy & 31];
            float wgt[16];
            for (int i = 0; i < 4; i++)
                for (int j = 0; j < 4; j++) wgt[i * 4 + j] = cy[i] * cx[j];
            for (int p = 0; p < nsrc; p++) {
                const float* S = src[p];
                float sum = 0.f;
                if (sx >= 0 && sx + 3 < w && sy >= 0 && sy + 3 < h) { /* (unsigned)sx < w-3 && (unsigned)sy < h-3 */
                    for (int i = 0; i < 4; i++) {
                        const float* r = S + (size_t)(sy + i) * w + sx;
                        sum += r[0] * wgt[i * 4] + r[1] * wgt[i * 4 + 1] + r[2] * wgt[i * 4 + 2] + r[3] * wgt[i * 4 + 3];
                    }
                } else if (sx >= w || sx + 4 <= 0 || sy >= h || sy + 4 <= 0) {
                    sum = 0.f;
                } else {
                    for (int i = 0; i < 4; i++) {
                        int yi = sy + i;
                        if (yi < 0 || yi >= h) continue;
                        for (int j = 0; j < 4; j++) {
                            int xj = sx + j;
                            if (xj >= 0 && xj < w) sum += S[(size_t)yi * w + xj] * wgt[i * 4 + j];
                        }
                    }
                }
                dst[p][o] = sum;
            }
        }
}

/* ---- 5x5 median, replicated border (cv::medianBlur on CV_32F) ---- */
void orc_median5_f32(const float* src, float* dst, int w, int h)
{
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float v[25];
            int n = 0;
            for (int dy = -2; dy <= 2; dy++)
                for (int dx = -2; dx <= 2; dx++) {
                    int yy = y + dy, xx = x + dx;
                    yy = yy < 0 ? 0 : yy >= h ? h - 1 : yy;
                    xx = xx < 0 ? 0 : xx >= w ? w - 1 : xx;
                    v[n++] = src[(size_t)yy * w + xx];
                }
            for (int i = 0; i <= 12; i++) { /* partial selection sort up to the median */
                int m = i;
                for (int j = i + 1; j < 25; j++) if (v[j] < v[m]) m = j;
                float t = v[i]; v[i] = v[m]; v[m] = t;
            }
            dst[(size_t)y * w + x] = v[12];
        }
}

/* centred differences with replicated border */
static void centered_gradient(const float* I, float* Ix, float* Iy, int w, int h)
{
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int xm = x > 0 ? x - 1 : 0, xp = x + 1 < w ? x + 1 : w - 1, ym = y > 0 ? y - 1 : 0, yp = y + 1 < h ? y + 1 : h - 1;
            Ix[(size_t)y * w + x] = 0.5f * (I[(size_t)y * w + xp] - I[(size_t)y * w + xm]);
            Iy[(size_t)y * w + x] = 0.5f * (I[(size_t)yp * w + x] - I[(size_t)ym * w + x]);
        }
}

/* one scale; u1, u2 in/out.  Returns the number of inner iterations actually run (all warpings). */
static int proc_one_scale(const float* I0, const float* I1, float* u1, float* u2, int w, int h, const orc_tvl1_params* P)
{
    const size_t n = (size_t)w * h;
    const float scaled_eps = (float)(P->epsilon * P->epsilon * (double)n);
    const float l_t = (float)(P->lambda * P->theta), taut = (float)(P->tau / P->theta), theta = (float)P->theta;
    float* buf = (float*)calloc(n * 14, sizeof(float));
    float *I1x = buf, *I1y = buf + n, *I1w = buf + 2 * n, *I1wx = buf + 3 * n, *I1wy = buf + 4 * n, *grad = buf + 5 * n, *rho_c = buf + 6 * n;
    float *p11 = buf + 7 * n, *p12 = buf + 8 * n, *p21 = buf + 9 * n, *p22 = buf + 10 * n, *mx = buf + 11 * n, *my = buf + 12 * n, *tmp = buf + 13 * n;
    int iters = 0;
    centered_gradient(I1, I1x, I1y, w, h);
    for (int wi = 0; wi < P->warps; wi++) {
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                mx[(size_t)y * w + x] = (float)x + u1[(size_t)y * w + x];
                my[(size_t)y * w + x] = (float)y + u2[(size_t)y * w + x];
            }
        const float* srcs[3] = {I1, I1x, I1y};
        float* dsts[3] = {I1w, I1wx, I1wy};
        orc_remap_cubic_f32(srcs, 3, w, h, mx, my, dsts, w, h);
        for (size_t i = 0; i < n; i++) {
            grad[i] = I1wx[i] * I1wx[i] + I1wy[i] * I1wy[i];
            rho_c[i] = I1w[i] - I1wx[i] * u1[i] - I1wy[i] * u2[i] - I0[i];
        }
        int stopped = 0;
        for (int no = 0; no < P->outer && !stopped; no++) {
            if (P->median > 1) {
                orc_median5_f32(u1, tmp, w, h); memcpy(u1, tmp, n * sizeof(float));
                orc_median5_f32(u2, tmp, w, h); memcpy(u2, tmp, n * sizeof(float));
            }
            for (int ni = 0; ni < P->inner && !stopped; ni++) {
                double err = 0.;
                iters++;
                /* estimateV + divergence + estimateU (p is read-only here, u is updated in place) */
                for (int y = 0; y < h; y++)
                    for (int x = 0; x < w; x++) {
                        size_t i = (size_t)y * w + x;
                        float rho = rho_c[i] + (I1wx[i] * u1[i] + I1wy[i] * u2[i]);
                        float d1 = 0.f, d2 = 0.f, lg = l_t * grad[i];
                        if (rho < -lg) { d1 = l_t * I1wx[i]; d2 = l_t * I1wy[i]; }
                        else if (rho > lg) { d1 = -l_t * I1wx[i]; d2 = -l_t * I1wy[i]; }
                        else if (grad[i] > FLT_EPSILON) { float fi = -rho / grad[i]; d1 = fi * I1wx[i]; d2 = fi * I1wy[i]; }
                        float v1 = u1[i] + d1, v2 = u2[i] + d2;
                        float a1 = x > 0 ? p11[i] - p11[i - 1] : p11[i], b1 = y > 0 ? p12[i] - p12[i - w] : p12[i];
                        float a2 = x > 0 ? p21[i] - p21[i - 1] : p21[i], b2 = y > 0 ? p22[i] - p22[i - w] : p22[i];
                        float n1 = v1 + theta * (a1 + b1), n2 = v2 + theta * (a2 + b2);
                        float e1 = n1 - u1[i], e2 = n2 - u2[i];
                        err += (double)(e1 * e1 + e2 * e2);
                        u1[i] = n1; u2[i] = n2;
                    }
                /* forward gradient of u + dual update */
                for (int y = 0; y < h; y++)
                    for (int x = 0; x < w; x++) {
                        size_t i = (size_t)y * w + x;
                        float u1x = x + 1 < w ? u1[i + 1] - u1[i] : 0.f, u1y = y + 1 < h ? u1[i + w] - u1[i] : 0.f;
                        float u2x = x + 1 < w ? u2[i + 1] - u2[i] : 0.f, u2y = y + 1 < h ? u2[i + w] - u2[i] : 0.f;
                        float g1 = (float)sqrt((double)u1x * u1x + (double)u1y * u1y), g2 = (float)sqrt((double)u2x * u2x + (double)u2y * u2y);
                        float ng1 = 1.f + taut * g1, ng2 = 1.f + taut * g2;
                        p11[i] = (p11[i] + taut * u1x) / ng1; p12[i] = (p12[i] + taut * u1y) / ng1;
                        p21[i] = (p21[i] + taut * u2x) / ng2; p22[i] = (p22[i] + taut * u2y) / ng2;
                    }
                if (!(err > (double)scaled_eps)) stopped = 1;
            }
        }
    }
    free(buf);
    return iters;
}

/* prev/next: u8 gray, row stride `stride`; flow: w*h*2 floats (u, v interleaved).  Returns total inner iterations. */
int orc_tvl1(const uint8_t* prev, const uint8_t* next, int stride, int w, int h, float* flow, const orc_tvl1_params* P)
{
    enum { MAXS = 32 };
    int ns = P->nscales < 1 ? 1 : P->nscales > MAXS ? MAXS : P->nscales;
    float *I0s[MAXS], *I1s[MAXS], *u1s[MAXS], *u2s[MAXS];
    int ws[MAXS], hs[MAXS];
    ws[0] = w; hs[0] = h;
    I0s[0] = (float*)malloc((size_t)w * h * 4); I1s[0] = (float*)malloc((size_t)w * h * 4);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            I0s[0][(size_t)y * w + x] = (float)prev[(size_t)y * stride + x];
            I1s[0][(size_t)y * w + x] = (float)next[(size_t)y * stride + x];
        }
    const double inv = 1. / P->scale_step;
    int built = 1;
    for (int s = 1; s < ns; s++) {
        int cw = (int)lrint(ws[s - 1] * P->scale_step), ch = (int)lrint(hs[s - 1] * P->scale_step);
        if (cw < 16 || ch < 16) break; /* the scale that would be too small is dropped */
        ws[s] = cw; hs[s] = ch;
        I0s[s] = (float*)malloc((size_t)cw * ch * 4); I1s[s] = (float*)malloc((size_t)cw * ch * 4);
        orc_tvl1_resize_f32(I0s[s - 1], ws[s - 1], hs[s - 1], I0s[s], cw, ch, inv, inv);
        orc_tvl1_resize_f32(I1s[s - 1], ws[s - 1], hs[s - 1], I1s[s], cw, ch, inv, inv);
        built = s + 1;
    }
    ns = built;
    for (int s = 0; s < ns; s++) {
        u1s[s] = (float*)calloc((size_t)ws[s] * hs[s], 4);
        u2s[s] = (float*)calloc((size_t)ws[s] * hs[s], 4);
    }
    int iters = 0;
    const float up = (float)(1. / P->scale_step);
    for (int s = ns - 1; s >= 0; s--) {
        iters += proc_one_scale(I0s[s], I1s[s], u1s[s], u2s[s], ws[s], hs[s], P);
        if (s == 0) break;
        orc_tvl1_resize_f32(u1s[s], ws[s], hs[s], u1s[s - 1], ws[s - 1], hs[s - 1], 1. / ((double)ws[s - 1] / ws[s]), 1. / ((double)hs[s - 1] / hs[s]));
        orc_tvl1_resize_f32(u2s[s], ws[s], hs[s], u2s[s - 1], ws[s - 1], hs[s - 1], 1. / ((double)ws[s - 1] / ws[s]), 1. / ((double)hs[s - 1] / hs[s]));
        for (size_t i = 0; i < (size_t)ws[s - 1] * hs[s - 1]; i++) { u1s[s - 1][i] *= up; u2s[s - 1][i] *= up; }
    }
    for (size_t i = 0; i < (size_t)w * h; i++) { flow[2 * i] = u1s[0][i]; flow[2 * i + 1] = u2s[0][i]; }
    for (int s = 0; s < ns; s++) { free(I0s[s]); free(I1s[s]); free(u1s[s]); free(u2s[s]); }
    return iters;
}
