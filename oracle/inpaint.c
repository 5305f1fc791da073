/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/farneback.c header for the rules).
 *
 * CPU restatement of cvInpaint(image, mask, out, radius, CV_INPAINT_TELEA | CV_INPAINT_NS) as called at
 * /root/reference/opencv2fx/inpaint/inpaint.cpp:311-318 (Telea hard-wired at :311; Navier-Stokes is
 * required by BASELINE.json config 4).  The arithmetic lives in OpenCV module photo (inpaint.cpp),
 * pinned here to opencv-python-headless 4.13.0.92 and restated per SURVEY.md Appendix A.2.
 * Parity pin: tests/test_oracle_vs_cv2.py (bit-exact vs cv2 4.13) + tests/golden/inpaint_*.npz.
 *
 * Build: gcc -O2 -ffp-contract=off (the OpenCV baseline build does not contract to FMA).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { KNOWN = 0, BAND = 1, INSIDE = 2, CHANGE = 3 };

/* Priority queue with CvPriorityQueueFloat semantics (A.2): a sorted list where push inserts after every
 * entry with T' <= T and pop takes the head, i.e. pop order = ascending (T, insertion counter).  The upstream
 * container is a linked list walked from the tail (O(n) per push); the same order is produced here by a binary
 * min-heap on the composite key (T, counter) so that the oracle stays usable at 4K. */
typedef struct { float T; uint32_t cnt; int32_t id; } HeapElem;
typedef struct { HeapElem* e; int n, cap; uint32_t counter; } Heap;

static void heap_init(Heap* H, int npix)
{
    H->cap = npix > 16 ? npix : 16;
    H->e = (HeapElem*)malloc((size_t)H->cap * sizeof(HeapElem));
    H->n = 0; H->counter = 0;
}
static void heap_free(Heap* H) { free(H->e); }
static inline int heap_less(const HeapElem* a, const HeapElem* b) { return a->T < b->T || (a->T == b->T && a->cnt < b->cnt); }
static void heap_push(Heap* H, int32_t id, float T)
{
    HeapElem v = {T, H->counter++, id};
    int k = H->n++;
    while (k > 0) {
        int p = (k - 1) >> 1;
        if (!heap_less(&v, &H->e[p])) break;
        H->e[k] = H->e[p]; k = p;
    }
    H->e[k] = v;
}
static int heap_pop(Heap* H, int32_t* id)
{
    if (H->n == 0) return 0;
    *id = H->e[0].id;
    HeapElem v = H->e[--H->n];
    int k = 0, n = H->n;
    for (;;) {
        int c = 2 * k + 1;
        if (c >= n) break;
        if (c + 1 < n && heap_less(&H->e[c + 1], &H->e[c])) c++;
        if (!heap_less(&H->e[c], &v)) break;
        H->e[k] = H->e[c]; k = c;
    }
    if (n > 0) H->e[k] = v;
    return 1;
}

#define F(i, j) f[(size_t)(i) * ec + (j)]
#define TT(i, j) t[(size_t)(i) * ec + (j)]

static float fmm_solve(int i1, int j1, int i2, int j2, const uint8_t* f, const float* t, int ec)
{
    double sol, a11 = TT(i1, j1), a22 = TT(i2, j2), m12 = a11 < a22 ? a11 : a22;
    if (F(i1, j1) != INSIDE) {
        if (F(i2, j2) != INSIDE) {
            if (fabs(a11 - a22) >= 1.0) sol = 1 + m12;
            else sol = (a11 + a22 + sqrt((double)(2 - (a11 - a22) * (a11 - a22)))) * 0.5;
        } else sol = 1 + a11;
    } else if (F(i2, j2) != INSIDE) sol = 1 + a22;
    else sol = 1 + m12;
    return (float)sol;
}
static inline float min4(float a, float b, float c, float d)
{
    a = a < b ? a : b; c = c < d ? c : d; return a < c ? a : c;
}
static inline float fmm_dist(int i, int j, const uint8_t* f, const float* t, int ec)
{
    return min4(fmm_solve(i - 1, j, i, j - 1, f, t, ec), fmm_solve(i + 1, j, i, j - 1, f, t, ec),
                fmm_solve(i - 1, j, i, j + 1, f, t, ec), fmm_solve(i + 1, j, i, j + 1, f, t, ec));
}

/* icvCalcFMM: march T over the INSIDE pixels of f starting from the entries already in the heap */
static void calc_fmm(uint8_t* f, float* t, Heap* H, int er, int ec, int negate)
{
    static const int di[4] = {-1, 0, 1, 0}, dj[4] = {0, -1, 0, 1};
    int32_t id;
    while (heap_pop(H, &id)) {
        int ii = id / ec, jj = id - ii * ec;
        F(ii, jj) = (uint8_t)(negate ? CHANGE : KNOWN);
        for (int q = 0; q < 4; q++) {
            int i = ii + di[q], j = jj + dj[q];
            if (i <= 0 || j <= 0 || i > er - 1 || j > ec - 1) continue;
            if (F(i, j) == INSIDE) {
                float dist = fmm_dist(i, j, f, t, ec);
                TT(i, j) = dist;
                F(i, j) = BAND;
                heap_push(H, i * ec + j, dist);
            }
        }
    }
    if (negate)
        for (size_t k = 0; k < (size_t)er * ec; k++)
            if (f[k] == CHANGE) { f[k] = KNOWN; t[k] = -t[k]; }
}

static inline uint8_t sat_u8_round(double v)
{
    int iv = (int)nearbyint(v);
    return (uint8_t)(iv < 0 ? 0 : iv > 255 ? 255 : iv);
}

#define OUT(r, c, ch) out[(size_t)(r) * ostep + (size_t)(c) * cn + (ch)]

static void fill_telea(int i, int j, const uint8_t* f, const float* t, uint8_t* out, int ostep, int cn, int er, int ec, int range)
{
    float gTx, gTy;
    if (F(i, j + 1) != INSIDE) {
        if (F(i, j - 1) != INSIDE) gTx = (float)(TT(i, j + 1) - TT(i, j - 1)) * 0.5f;
        else gTx = (float)(TT(i, j + 1) - TT(i, j));
    } else {
        if (F(i, j - 1) != INSIDE) gTx = (float)(TT(i, j) - TT(i, j - 1));
        else gTx = 0;
    }
    if (F(i + 1, j) != INSIDE) {
        if (F(i - 1, j) != INSIDE) gTy = (float)(TT(i + 1, j) - TT(i - 1, j)) * 0.5f;
        else gTy = (float)(TT(i + 1, j) - TT(i, j));
    } else {
        if (F(i - 1, j) != INSIDE) gTy = (float)(TT(i, j) - TT(i - 1, j));
        else gTy = 0;
    }
    float Jx[3] = {0, 0, 0}, Jy[3] = {0, 0, 0}, Ia[3] = {0, 0, 0}, s[3] = {1.0e-20f, 1.0e-20f, 1.0e-20f};
    for (int k = i - range; k <= i + range; k++) {
        int km = k - 1 + (k == 1), kp = k - 1 - (k == er - 2);
        for (int l = j - range; l <= j + range; l++) {
            int lm = l - 1 + (l == 1), lp = l - 1 - (l == ec - 2);
            if (!(k > 0 && l > 0 && k < er - 1 && l < ec - 1)) continue;
            if (F(k, l) == INSIDE || (l - j) * (l - j) + (k - i) * (k - i) > range * range) continue;
            float ry = (float)(i - k), rx = (float)(j - l);
            float vl = rx * rx + ry * ry;
            float dst = (float)(1. / (vl * sqrt((double)vl)));
            float lev = (float)(1. / (1 + fabs(TT(k, l) - TT(i, j))));   /* f32 difference, then double */
            float dir = rx * gTx + ry * gTy;
            if (fabs(dir) <= 0.01) dir = 0.000001f;
            float w = (float)fabs(dst * lev * dir);
            for (int c = 0; c < cn; c++) {
                float gIx, gIy;
                if (F(k, l + 1) != INSIDE) {
                    if (F(k, l - 1) != INSIDE) gIx = (float)(OUT(km, lp + 1, c) - OUT(km, lm - 1, c)) * 2.0f;
                    else gIx = (float)(OUT(km, lp + 1, c) - OUT(km, lm, c));
                } else {
                    if (F(k, l - 1) != INSIDE) gIx = (float)(OUT(km, lp, c) - OUT(km, lm - 1, c));
                    else gIx = 0;
                }
                if (F(k + 1, l) != INSIDE) {
                    if (F(k - 1, l) != INSIDE) gIy = (float)(OUT(kp + 1, lm, c) - OUT(km - 1, lm, c)) * 2.0f;
                    else gIy = (float)(OUT(kp + 1, lm, c) - OUT(km, lm, c));
                } else {
                    if (F(k - 1, l) != INSIDE) gIy = (float)(OUT(kp, lm, c) - OUT(km - 1, lm, c));
                    else gIy = 0;
                }
                Ia[c] += (float)w * (float)(OUT(k - 1, l - 1, c));
                Jx[c] -= (float)w * (float)(gIx * rx);
                Jy[c] -= (float)w * (float)(gIy * ry);
                s[c] += w;
            }
        }
    }
    for (int c = 0; c < cn; c++) {
        /* all-f32 incl. the sqrt (C++ float overload) — pinned against cv2 4.13 on exact .5 ties (radius 1) */
        float sat = Ia[c] / s[c] + (Jx[c] + Jy[c]) / (sqrtf(Jx[c] * Jx[c] + Jy[c] * Jy[c]) + 1.0e-20f) + 0.5f;
        OUT(i - 1, j - 1, c) = sat_u8_round(sat);
    }
}

static void fill_ns(int i, int j, const uint8_t* f, uint8_t* out, int ostep, int cn, int er, int ec, int range)
{
    float Ia[3] = {0, 0, 0}, s[3] = {1.0e-20f, 1.0e-20f, 1.0e-20f};
    for (int k = i - range; k <= i + range; k++) {
        int km = k - 1 + (k == 1), kp = k - 1 - (k == er - 2);
        for (int l = j - range; l <= j + range; l++) {
            int lm = l - 1 + (l == 1), lp = l - 1 - (l == ec - 2);
            if (!(k > 0 && l > 0 && k < er - 1 && l < ec - 1)) continue;
            if (F(k, l) == INSIDE || (l - j) * (l - j) + (k - i) * (k - i) > range * range) continue;
            float ry = (float)(k - i), rx = (float)(l - j);
            float vl = rx * rx + ry * ry;
            float dst = 1 / (vl * vl + 1);
            for (int c = 0; c < cn; c++) {
                float gIx, gIy;
                if (F(k + 1, l) != INSIDE) {
                    if (F(k - 1, l) != INSIDE)
                        gIx = (float)(abs(OUT(kp + 1, lm, c) - OUT(kp, lm, c)) + abs(OUT(kp, lm, c) - OUT(km - 1, lm, c)));
                    else gIx = (float)(abs(OUT(kp + 1, lm, c) - OUT(kp, lm, c))) * 2.0f;
                } else {
                    if (F(k - 1, l) != INSIDE) gIx = (float)(abs(OUT(kp, lm, c) - OUT(km - 1, lm, c))) * 2.0f;
                    else gIx = 0;
                }
                if (F(k, l + 1) != INSIDE) {
                    if (F(k, l - 1) != INSIDE)
                        gIy = (float)(abs(OUT(km, lp + 1, c) - OUT(km, lm, c)) + abs(OUT(km, lm, c) - OUT(km, lm - 1, c)));
                    else gIy = (float)(abs(OUT(km, lp + 1, c) - OUT(km, lm, c))) * 2.0f;
                } else {
                    if (F(k, l - 1) != INSIDE) gIy = (float)(abs(OUT(km, lm, c) - OUT(km, lm - 1, c))) * 2.0f;
                    else gIy = 0;
                }
                gIx = -gIx;
                float dir = rx * gIx + ry * gIy;
                if (fabs(dir) <= 0.01) dir = 0.000001f;
                else dir = fabsf((rx * gIx + ry * gIy) / sqrtf(vl * (gIx * gIx + gIy * gIy)));   /* all-f32 (C++ float overloads), pinned vs cv2 4.13 */
                float w = dst * dir;
                Ia[c] += (float)w * (float)(OUT(k - 1, l - 1, c));
                s[c] += w;
            }
        }
    }
    for (int c = 0; c < cn; c++) OUT(i - 1, j - 1, c) = sat_u8_round((double)Ia[c] / s[c]);
}

/* main FMM with fill: pop -> KNOWN; each INSIDE 4-neighbour gets T (once), its colour, BAND, push */
static void inpaint_fmm(uint8_t* f, float* t, uint8_t* out, int ostep, int cn, int er, int ec, int range, Heap* H, int method,
                        int32_t* seq_out)
{
    static const int di[4] = {-1, 0, 1, 0}, dj[4] = {0, -1, 0, 1};
    int32_t id, seq = 0;
    while (heap_pop(H, &id)) {
        int ii = id / ec, jj = id - ii * ec;
        F(ii, jj) = KNOWN;
        for (int q = 0; q < 4; q++) {
            int i = ii + di[q], j = jj + dj[q];
            if (i <= 0 || j <= 0 || i > er - 1 || j > ec - 1) continue;
            if (F(i, j) != INSIDE) continue;
            float dist = fmm_dist(i, j, f, t, ec);
            TT(i, j) = dist;
            if (method == 0) fill_ns(i, j, f, out, ostep, cn, er, ec, range);
            else fill_telea(i, j, f, t, out, ostep, cn, er, ec, range);
            F(i, j) = BAND;
            if (seq_out) seq_out[(size_t)(i - 1) * (ec - 2) + (j - 1)] = seq;
            seq++;
            heap_push(H, i * ec + j, dist);
        }
    }
}

/* src/dst: h x w x cn u8 (cn 1 or 3), mask h x w u8 (non-zero = inpaint). method: 0 = NS, 1 = TELEA
 * (cv::INPAINT_NS / cv::INPAINT_TELEA).  Optional debug outputs: t_out ((h+2)x(w+2) f32), seq_out (h x w
 * int32 fill order, -1 where not filled). */
int orc_inpaint(const uint8_t* src, int sstep, const uint8_t* mask, int mstep, uint8_t* dst, int ostep, int w, int h, int cn,
                double radius, int method, float* t_out, int32_t* seq_out)
{
    int range = (int)nearbyint(radius);
    range = range < 1 ? 1 : range > 100 ? 100 : range;
    int er = h + 2, ec = w + 2;
    size_t np = (size_t)er * ec;
    for (int y = 0; y < h; y++) memcpy(dst + (size_t)y * ostep, src + (size_t)y * sstep, (size_t)w * cn);
    if (seq_out) for (size_t k = 0; k < (size_t)w * h; k++) seq_out[k] = -1;
    uint8_t* f = (uint8_t*)calloc(np, 1);     /* KNOWN; INSIDE on the hole ("mask" matrix of cvInpaint) */
    uint8_t* band = (uint8_t*)calloc(np, 1);
    float* t = (float*)malloc(np * sizeof(float));
    for (size_t k = 0; k < np; k++) t[k] = 1.0e6f;
    long nhole = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            if (mask[(size_t)y * mstep + x]) { F(y + 1, x + 1) = INSIDE; nhole++; }
    if (nhole == 0) { free(f); free(band); free(t); if (t_out) memset(t_out, 0, np * sizeof(float)); return 0; }
    /* band = dilate(hole, 3x3 cross) - hole, border ring cleared */
    for (int i = 1; i < er - 1; i++)
        for (int j = 1; j < ec - 1; j++)
            if (F(i, j) != INSIDE && (F(i - 1, j) == INSIDE || F(i + 1, j) == INSIDE || F(i, j - 1) == INSIDE || F(i, j + 1) == INSIDE))
                band[(size_t)i * ec + j] = 1;
    Heap H; heap_init(&H, (int)np);
    for (size_t k = 0; k < np; k++) if (band[k]) { heap_push(&H, (int32_t)k, 0.f); t[k] = 0.f; }
    if (method == 1) {
        /* out-region: (dilate(hole, (2r+1)^2 rect) - hole - band), border ring cleared; T marched outwards, negated */
        uint8_t* fo = (uint8_t*)calloc(np, 1);
        /* separable rect dilation */
        uint8_t* tmp = (uint8_t*)calloc(np, 1);
        for (int i = 0; i < er; i++)
            for (int j = 0; j < ec; j++) {
                int v = 0;
                for (int l = j - range; l <= j + range && !v; l++) if (l >= 0 && l < ec && F(i, l) == INSIDE) v = 1;
                tmp[(size_t)i * ec + j] = (uint8_t)v;
            }
        for (int i = 1; i < er - 1; i++)
            for (int j = 1; j < ec - 1; j++) {
                int v = 0;
                for (int k = i - range; k <= i + range && !v; k++) if (k >= 0 && k < er && tmp[(size_t)k * ec + j]) v = 1;
                if (v && F(i, j) != INSIDE && !band[(size_t)i * ec + j]) fo[(size_t)i * ec + j] = INSIDE;
            }
        free(tmp);
        Heap O; heap_init(&O, (int)np);
        for (size_t k = 0; k < np; k++) if (band[k]) heap_push(&O, (int32_t)k, 0.f);
        calc_fmm(fo, t, &O, er, ec, 1);
        heap_free(&O);
        free(fo);
    }
    inpaint_fmm(f, t, dst, ostep, cn, er, ec, range, &H, method, seq_out);
    if (t_out) memcpy(t_out, t, np * sizeof(float));
    heap_free(&H);
    free(f); free(band); free(t);
    return 0;
}
