/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path
 * (openfx-opencv_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.
 *
 * CPU restatement of cv::calcOpticalFlowFarneback as the reference calls it:
 *   /root/reference/VectorGenerator/VectorGenerator.cpp:403
 *     calcOpticalFlowFarneback(prev u8, next u8, flow, 0.5, levels, 3, iters, polyN, polySigma, 0)
 * The arithmetic lives in OpenCV (un-vendored dependency of the reference, found via pkg-config:
 * /root/reference/Makefile.master:8-9); we pin opencv-python-headless 4.13.0.92 (module video,
 * optflowgf.cpp) and follow the published algorithm as restated in SURVEY.md Appendix A.1.
 * Parity pin: tests/test_oracle_vs_cv2.py + tests/golden/farneback_*.npz (generated from cv2 4.13 by
 * tests/golden/make_golden.py).
 *
 * Build: plain C, no FMA contraction (the OpenCV baseline build is SSE2/SSE3, non-FMA):
 *   gcc -O2 -ffp-contract=off -fPIC -shared
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

static inline int cv_round(double v) { return (int)nearbyint(v); } /* round-half-even (default FE mode) */

/* ---- cv::getGaussianKernel(n, sigma, CV_32F)  (SURVEY A.1 "Gaussian kernel") ------------------ */
void orc_gaussian_kernel(int n, double sigma, float* out)
{
    static const float tab3[] = {0.25f, 0.5f, 0.25f};
    static const float tab1[] = {1.f};
    static const float tab5[] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
    static const float tab7[] = {0.03125f, 0.109375f, 0.21875f, 0.28125f, 0.21875f, 0.109375f, 0.03125f};
    if (sigma <= 0 && (n & 1) && n <= 7) {
        const float* t = n == 1 ? tab1 : n == 3 ? tab3 : n == 5 ? tab5 : tab7;
        memcpy(out, t, n * sizeof(float));
        return;
    }
    double s = sigma > 0 ? sigma : ((n - 1) * 0.5 - 1) * 0.3 + 0.8;
    double scale2 = -0.5 / (s * s);
    double* k = (double*)malloc(n * sizeof(double));
    double sum = 0;
    for (int i = 0; i < n; i++) {
        double x = i - (n - 1) * 0.5;
        k[i] = exp(scale2 * x * x);
        sum += k[i];
    }
    sum = 1. / sum;
    for (int i = 0; i < n; i++) out[i] = (float)(k[i] * sum);
    free(k);
}

static inline int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

/* ---- cv::GaussianBlur on f32, separable, BORDER_REFLECT_101, row pass then column pass -------- */
void orc_gaussian_blur_f32(const float* src, float* dst, int w, int h, int ksz, double sigma)
{
    float* k = (float*)malloc(ksz * sizeof(float));
    orc_gaussian_kernel(ksz, sigma, k);
    int r = ksz / 2;
    float* tmp = (float*)malloc((size_t)w * h * sizeof(float));
    for (int y = 0; y < h; y++) {
        const float* s = src + (size_t)y * w;
        float* d = tmp + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            /* symmetric form used by OpenCV's SymmRowFilter: k0*s[x] + sum k[i]*(s[x-i]+s[x+i]) */
            float acc = k[r] * s[x];
            for (int i = 1; i <= r; i++)
                acc += k[r + i] * (s[reflect101(x - i, w)] + s[reflect101(x + i, w)]);
            d[x] = acc;
        }
    }
    for (int y = 0; y < h; y++) {
        float* d = dst + (size_t)y * w;
        const float* c = tmp + (size_t)y * w;
        for (int x = 0; x < w; x++) d[x] = k[r] * c[x];
        for (int i = 1; i <= r; i++) {
            const float* a = tmp + (size_t)reflect101(y - i, h) * w;
            const float* b = tmp + (size_t)reflect101(y + i, h) * w;
            float ki = k[r + i];
            for (int x = 0; x < w; x++) d[x] += ki * (a[x] + b[x]);
        }
    }
    free(tmp);
    free(k);
}

/* ---- cv::resize(INTER_LINEAR) on f32 with `cn` interleaved channels (SURVEY A.1 resize) ------- */
static void linear_coeffs(int nsrc, int ndst, int* ofs, float* alpha)
{
    double inv_scale = (double)ndst / nsrc;
    double scale = 1. / inv_scale;
    for (int d = 0; d < ndst; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= nsrc - 1) { s = nsrc - 1; f = 0.f; }
        ofs[d] = s;
        alpha[d] = f;
    }
}

void orc_resize_linear_f32(const float* src, int sw, int sh, float* dst, int dw, int dh, int cn)
{
    if (sw == dw && sh == dh) { memcpy(dst, src, (size_t)sw * sh * cn * sizeof(float)); return; }
    int* xo = (int*)malloc(dw * sizeof(int));
    int* yo = (int*)malloc(dh * sizeof(int));
    float* xa = (float*)malloc(dw * sizeof(float));
    float* ya = (float*)malloc(dh * sizeof(float));
    linear_coeffs(sw, dw, xo, xa);
    linear_coeffs(sh, dh, yo, ya);
    float* r0 = (float*)malloc((size_t)dw * cn * sizeof(float));
    float* r1 = (float*)malloc((size_t)dw * cn * sizeof(float));
    for (int y = 0; y < dh; y++) {
        int sy0 = yo[y], sy1 = sy0 + 1 < sh ? sy0 + 1 : sh - 1;
        const float* s0 = src + (size_t)sy0 * sw * cn;
        const float* s1 = src + (size_t)sy1 * sw * cn;
        for (int x = 0; x < dw; x++) {
            int sx0 = xo[x], sx1 = sx0 + 1 < sw ? sx0 + 1 : sw - 1;
            float a1 = xa[x], a0 = 1.f - a1;
            for (int c = 0; c < cn; c++) {
                r0[x * cn + c] = s0[sx0 * cn + c] * a0 + s0[sx1 * cn + c] * a1;
                r1[x * cn + c] = s1[sx0 * cn + c] * a0 + s1[sx1 * cn + c] * a1;
            }
        }
        float b1 = ya[y], b0 = 1.f - b1;
        float* d = dst + (size_t)y * dw * cn;
        for (int x = 0; x < dw * cn; x++) d[x] = r0[x] * b0 + r1[x] * b1;
    }
    free(xo); free(yo); free(xa); free(ya); free(r0); free(r1);
}

/* ---- FarnebackPrepareGaussian (SURVEY A.1 PolyExp) --------------------------------------------- */
void orc_polyexp_setup(int n, double sigma, float* g /*[2n+1] centred*/, float* xg, float* xxg, double ig[4])
{
    if (sigma < FLT_EPSILON) sigma = n * 0.3;
    float* gc = g + n; float* xgc = xg + n; float* xxgc = xxg + n;
    double s = 0.;
    for (int x = -n; x <= n; x++) {
        gc[x] = (float)exp(-x * x / (2 * sigma * sigma));
        s += gc[x];
    }
    s = 1. / s;
    for (int x = -n; x <= n; x++) {
        gc[x] = (float)(gc[x] * s);
        xgc[x] = (float)(x * gc[x]);
        xxgc[x] = (float)(x * x * gc[x]);
    }
    double G00 = 0, G11 = 0, G33 = 0, G55 = 0;
    for (int y = -n; y <= n; y++)
        for (int x = -n; x <= n; x++) {
            G00 += gc[y] * gc[x];
            G11 += gc[y] * gc[x] * x * x;
            G33 += gc[y] * gc[x] * x * x * x * x;
            G55 += gc[y] * gc[x] * x * x * y * y;
        }
    /* G couples {1,x^2,y^2}: [[G00,G11,G11],[G11,G33,G55],[G11,G55,G33]]; x,y: G11; xy: G55.        */
    /* closed-form inverse of the 3x3 block (OpenCV uses a 6x6 Cholesky inverse; identical to ~1e-16) */
    double a = G00, b = G11, c = G33, d = G55;
    double det = a * (c * c - d * d) - 2 * b * b * (c - d);
    double inv03 = -b * (c - d) / det;      /* invG(0,3) */
    double inv33 = (a * c - b * b) / det;   /* invG(3,3) */
    ig[0] = 1. / G11;  /* ig11 */
    ig[1] = inv03;     /* ig03 */
    ig[2] = inv33;     /* ig33 */
    ig[3] = 1. / G55;  /* ig55 */
}

/* ---- FarnebackPolyExp: I (h x w f32) -> R (h x w x 5 f32, AoS) --------------------------------- */
void orc_polyexp(const float* src, int w, int h, int n, double sigma, float* dst)
{
    float* kbuf = (float*)malloc((size_t)(n * 6 + 3) * sizeof(float));
    float *g = kbuf, *xg = g + 2 * n + 1, *xxg = xg + 2 * n + 1;
    double ig[4];
    orc_polyexp_setup(n, sigma, g, xg, xxg, ig);
    g += n; xg += n; xxg += n;
    double ig11 = ig[0], ig03 = ig[1], ig33 = ig[2], ig55 = ig[3];
    float* rowbuf = (float*)malloc((size_t)(w + n * 2) * 3 * sizeof(float));
    float* row = rowbuf + n * 3;
    for (int y = 0; y < h; y++) {
        float g0 = g[0], g1, g2;
        const float* srow0 = src + (size_t)y * w;
        const float* srow1;
        float* drow = dst + (size_t)y * w * 5;
        for (int x = 0; x < w; x++) {
            row[x * 3] = srow0[x] * g0;
            row[x * 3 + 1] = row[x * 3 + 2] = 0.f;
        }
        for (int k = 1; k <= n; k++) {
            g0 = g[k]; g1 = xg[k]; g2 = xxg[k];
            srow0 = src + (size_t)(y - k > 0 ? y - k : 0) * w;
            srow1 = src + (size_t)(y + k < h - 1 ? y + k : h - 1) * w;
            for (int x = 0; x < w; x++) {
                float p = srow0[x] + srow1[x];
                float t0 = row[x * 3] + g0 * p;
                float t1 = row[x * 3 + 1] + g1 * (srow1[x] - srow0[x]);
                float t2 = row[x * 3 + 2] + g2 * p;
                row[x * 3] = t0; row[x * 3 + 1] = t1; row[x * 3 + 2] = t2;
            }
        }
        for (int x = 0; x < n * 3; x++) {
            row[-1 - x] = row[2 - x];
            row[w * 3 + x] = row[w * 3 + x - 3];
        }
        for (int x = 0; x < w; x++) {
            g0 = g[0];
            double b1 = row[x * 3] * g0, b2 = 0, b3 = row[x * 3 + 1] * g0, b4 = 0, b5 = row[x * 3 + 2] * g0, b6 = 0;
            for (int k = 1; k <= n; k++) {
                double tg = row[(x + k) * 3] + row[(x - k) * 3];
                g0 = g[k];
                b1 += tg * g0;
                b4 += tg * xxg[k];
                b2 += (row[(x + k) * 3] - row[(x - k) * 3]) * xg[k];
                b3 += (row[(x + k) * 3 + 1] + row[(x - k) * 3 + 1]) * g0;
                b6 += (row[(x + k) * 3 + 1] - row[(x - k) * 3 + 1]) * xg[k];
                b5 += (row[(x + k) * 3 + 2] + row[(x - k) * 3 + 2]) * g0;
            }
            drow[x * 5 + 1] = (float)(b2 * ig11);
            drow[x * 5] = (float)(b3 * ig11);
            drow[x * 5 + 3] = (float)(b1 * ig03 + b4 * ig33);
            drow[x * 5 + 2] = (float)(b1 * ig03 + b5 * ig33);
            drow[x * 5 + 4] = (float)(b6 * ig55);
        }
    }
    free(rowbuf);
    free(kbuf);
}

/* ---- FarnebackUpdateMatrices rows [y0,y1) (SURVEY A.1 UpdateMatrices) -------------------------- */
void orc_update_matrices(const float* R0, const float* R1, const float* flow, float* M, int w, int h, int y0, int y1)
{
    enum { BORDER = 5 };
    static const float border[BORDER] = {0.14f, 0.14f, 0.4472f, 0.4472f, 0.4472f};
    size_t step1 = (size_t)w * 5;
    for (int y = y0; y < y1; y++) {
        const float* fl = flow + (size_t)y * w * 2;
        const float* r0p = R0 + (size_t)y * w * 5;
        float* m = M + (size_t)y * w * 5;
        for (int x = 0; x < w; x++) {
            float dx = fl[x * 2], dy = fl[x * 2 + 1];
            float fx = x + dx, fy = y + dy;
            int x1 = (int)floorf(fx), yy1 = (int)floorf(fy);
            float r2, r3, r4, r5, r6;
            fx -= x1; fy -= yy1;
            if ((unsigned)x1 < (unsigned)(w - 1) && (unsigned)yy1 < (unsigned)(h - 1)) {
                const float* ptr = R1 + (size_t)yy1 * step1 + (size_t)x1 * 5;
                float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
                r2 = a00 * ptr[0] + a01 * ptr[5] + a10 * ptr[step1] + a11 * ptr[step1 + 5];
                r3 = a00 * ptr[1] + a01 * ptr[6] + a10 * ptr[step1 + 1] + a11 * ptr[step1 + 6];
                r4 = a00 * ptr[2] + a01 * ptr[7] + a10 * ptr[step1 + 2] + a11 * ptr[step1 + 7];
                r5 = a00 * ptr[3] + a01 * ptr[8] + a10 * ptr[step1 + 3] + a11 * ptr[step1 + 8];
                r6 = a00 * ptr[4] + a01 * ptr[9] + a10 * ptr[step1 + 4] + a11 * ptr[step1 + 9];
                r4 = (r0p[x * 5 + 2] + r4) * 0.5f;
                r5 = (r0p[x * 5 + 3] + r5) * 0.5f;
                r6 = (r0p[x * 5 + 4] + r6) * 0.25f;
            } else {
                r2 = r3 = 0.f;
                r4 = r0p[x * 5 + 2];
                r5 = r0p[x * 5 + 3];
                r6 = r0p[x * 5 + 4] * 0.5f;
            }
            r2 = (r0p[x * 5] - r2) * 0.5f;
            r3 = (r0p[x * 5 + 1] - r3) * 0.5f;
            r2 += r4 * dy + r6 * dx;
            r3 += r6 * dy + r5 * dx;
            if ((unsigned)(x - BORDER) >= (unsigned)(w - BORDER * 2) || (unsigned)(y - BORDER) >= (unsigned)(h - BORDER * 2)) {
                float scale = (x < BORDER ? border[x] : 1.f) * (x >= w - BORDER ? border[w - x - 1] : 1.f) *
                              (y < BORDER ? border[y] : 1.f) * (y >= h - BORDER ? border[h - y - 1] : 1.f);
                r2 *= scale; r3 *= scale; r4 *= scale; r5 *= scale; r6 *= scale;
            }
            m[x * 5] = r4 * r4 + r6 * r6;
            m[x * 5 + 1] = (r4 + r5) * r6;
            m[x * 5 + 2] = r5 * r5 + r6 * r6;
            m[x * 5 + 3] = r4 * r2 + r6 * r3;
            m[x * 5 + 4] = r6 * r2 + r5 * r3;
        }
    }
}

/* ---- FarnebackUpdateFlow_Blur: box blur of M by running sums in double + 2x2 solve ------------- */
void orc_update_flow_blur(const float* R0, const float* R1, float* flow, float* M, int w, int h, int bs, int update)
{
    int m = bs / 2;
    int y0 = 0, y1;
    int min_update_stripe = (1 << 10) / w > bs ? (1 << 10) / w : bs;
    double scale = 1. / (bs * bs);
    double* vbuf = (double*)malloc((size_t)(w + m * 2 + 2) * 5 * sizeof(double));
    double* vsum = vbuf + (m + 1) * 5;
    const float* srow0 = M;
    for (int x = 0; x < w * 5; x++) vsum[x] = srow0[x] * (m + 2);
    for (int y = 1; y < m; y++) {
        srow0 = M + (size_t)(y < h - 1 ? y : h - 1) * w * 5;
        for (int x = 0; x < w * 5; x++) vsum[x] += srow0[x];
    }
    for (int y = 0; y < h; y++) {
        double g11, g12, g22, h1, h2;
        float* fl = flow + (size_t)y * w * 2;
        srow0 = M + (size_t)(y - m - 1 > 0 ? y - m - 1 : 0) * w * 5;
        const float* srow1 = M + (size_t)(y + m < h - 1 ? y + m : h - 1) * w * 5;
        for (int x = 0; x < w * 5; x++) vsum[x] += srow1[x] - srow0[x];
        for (int x = 0; x < (m + 1) * 5; x++) {
            vsum[-1 - x] = vsum[4 - x];
            vsum[w * 5 + x] = vsum[w * 5 + x - 5];
        }
        g11 = vsum[0] * (m + 2); g12 = vsum[1] * (m + 2); g22 = vsum[2] * (m + 2);
        h1 = vsum[3] * (m + 2); h2 = vsum[4] * (m + 2);
        for (int x = 1; x < m; x++) {
            g11 += vsum[x * 5]; g12 += vsum[x * 5 + 1]; g22 += vsum[x * 5 + 2];
            h1 += vsum[x * 5 + 3]; h2 += vsum[x * 5 + 4];
        }
        for (int x = 0; x < w; x++) {
            g11 += vsum[(x + m) * 5] - vsum[(x - m) * 5 - 5];
            g12 += vsum[(x + m) * 5 + 1] - vsum[(x - m) * 5 - 4];
            g22 += vsum[(x + m) * 5 + 2] - vsum[(x - m) * 5 - 3];
            h1 += vsum[(x + m) * 5 + 3] - vsum[(x - m) * 5 - 2];
            h2 += vsum[(x + m) * 5 + 4] - vsum[(x - m) * 5 - 1];
            double g11_ = g11 * scale, g12_ = g12 * scale, g22_ = g22 * scale, h1_ = h1 * scale, h2_ = h2 * scale;
            double idet = 1. / (g11_ * g22_ - g12_ * g12_ + 1e-3);
            fl[x * 2] = (float)((g11_ * h2_ - g12_ * h1_) * idet);
            fl[x * 2 + 1] = (float)((g22_ * h1_ - g12_ * h2_) * idet);
        }
        y1 = y == h - 1 ? h : y - bs;
        if (update && (y1 == h || y1 >= y0 + min_update_stripe)) {
            orc_update_matrices(R0, R1, flow, M, w, h, y0, y1);
            y0 = y1;
        }
    }
    free(vbuf);
}

/* effective number of pyramid levels (SURVEY A.1 head) */
int orc_farneback_levels(int w, int h, double pyr_scale, int levels)
{
    int k; double scale = 1;
    for (k = 0; k < levels; k++) {
        scale *= pyr_scale;
        if (w * scale < 32 || h * scale < 32) break;
    }
    return k;
}

/* ---- driver: prev,next u8 (stride bytes) -> flow (h x w x 2 f32, dense) ------------------------ */
int orc_farneback(const uint8_t* prev, const uint8_t* next, int stride, int w, int h, float* flow_out,
                  double pyr_scale, int levels, int winsize, int iters, int poly_n, double poly_sigma)
{
    levels = orc_farneback_levels(w, h, pyr_scale, levels);
    float* fimg = (float*)malloc((size_t)w * h * sizeof(float));
    float* blur = (float*)malloc((size_t)w * h * sizeof(float));
    float* prevFlow = NULL; int pw = 0, ph = 0;
    const uint8_t* imgs[2] = {prev, next};
    for (int k = levels; k >= 0; k--) {
        double scale = 1;
        for (int i = 0; i < k; i++) scale *= pyr_scale;
        double sigma = (1. / scale - 1) * 0.5;
        int ksz = cv_round(sigma * 5) | 1;
        if (ksz < 3) ksz = 3;
        int cw = cv_round(w * scale), ch = cv_round(h * scale);
        size_t n = (size_t)cw * ch;
        float* flow = (float*)calloc(n * 2, sizeof(float));
        if (prevFlow) {
            orc_resize_linear_f32(prevFlow, pw, ph, flow, cw, ch, 2);
            float fs = (float)(1. / pyr_scale);
            for (size_t i = 0; i < n * 2; i++) flow[i] *= fs;
        }
        float* R[2];
        float* I = (float*)malloc(n * sizeof(float));
        for (int i = 0; i < 2; i++) {
            for (int y = 0; y < h; y++)
                for (int x = 0; x < w; x++) fimg[(size_t)y * w + x] = (float)imgs[i][(size_t)y * stride + x];
            orc_gaussian_blur_f32(fimg, blur, w, h, ksz, sigma);
            orc_resize_linear_f32(blur, w, h, I, cw, ch, 1);
            R[i] = (float*)malloc(n * 5 * sizeof(float));
            orc_polyexp(I, cw, ch, poly_n, poly_sigma, R[i]);
        }
        float* M = (float*)malloc(n * 5 * sizeof(float));
        orc_update_matrices(R[0], R[1], flow, M, cw, ch, 0, ch);
        for (int i = 0; i < iters; i++) orc_update_flow_blur(R[0], R[1], flow, M, cw, ch, winsize, i < iters - 1);
        free(M); free(I); free(R[0]); free(R[1]);
        free(prevFlow);
        prevFlow = flow; pw = cw; ph = ch;
    }
    memcpy(flow_out, prevFlow, (size_t)w * h * 2 * sizeof(float));
    free(prevFlow); free(fimg); free(blur);
    return 0;
}
