/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/farneback.c header for the rules).
 *
 * CPU restatement of cv::watershed(rgb8, int32 markers), the body BASELINE.json config 3 names for the
 * segment plugin (the reference's own call is cvPyrSegmentation at
 * /root/reference/opencv2fx/segment/segment.cpp:296-302, which no longer exists in OpenCV >= 3; see
 * SURVEY.md section 0 fact 2).  Algorithm: OpenCV imgproc segmentation.cpp (Meyer flooding, 256 FIFO
 * buckets), restated per SURVEY.md Appendix A.3.  Parity pin: tests/test_oracle_vs_cv2.py +
 * tests/golden/watershed_*.npz (cv2 4.13 outputs).
 */
#include <stdint.h>
#include <stdlib.h>

#define WS_IN_QUEUE (-2)
#define WS_WSHED (-1)

static inline int cdiff(const uint8_t* a, const uint8_t* b)
{
    int d0 = abs((int)a[0] - (int)b[0]), d1 = abs((int)a[1] - (int)b[1]), d2 = abs((int)a[2] - (int)b[2]);
    int d = d0 > d1 ? d0 : d1;
    return d > d2 ? d : d2;
}

/* img: h x w x 3 u8 (istep bytes/row); markers: h x w int32 (mstep ints/row), in/out.
 * Returns the number of queue pops (statistics for the byte/latency model). */
long orc_watershed(const uint8_t* img, int istep, int32_t* mask, int mstep, int w, int h)
{
    /* intrusive FIFO: every pixel is enqueued at most once */
    int32_t* nxt = (int32_t*)malloc((size_t)w * h * sizeof(int32_t));
    int32_t head[256], tail[256];
    for (int i = 0; i < 256; i++) head[i] = tail[i] = -1;
    long pops = 0;
#define PUSH(q, p) do { int32_t _p = (p); nxt[_p] = -1; if (head[q] < 0) head[q] = _p; else nxt[tail[q]] = _p; tail[q] = _p; } while (0)

    for (int j = 0; j < w; j++) mask[j] = mask[j + (size_t)mstep * (h - 1)] = WS_WSHED;
    for (int i = 1; i < h - 1; i++) {
        int32_t* mrow = mask + (size_t)i * mstep;
        const uint8_t* irow = img + (size_t)i * istep;
        mrow[0] = mrow[w - 1] = WS_WSHED;
        for (int j = 1; j < w - 1; j++) {
            int32_t* m = mrow + j;
            if (m[0] < 0) m[0] = 0;
            if (m[0] == 0 && (m[-1] > 0 || m[1] > 0 || m[-mstep] > 0 || m[mstep] > 0)) {
                const uint8_t* ptr = irow + j * 3;
                int idx = 256, t;
                if (m[-1] > 0) idx = cdiff(ptr, ptr - 3);
                if (m[1] > 0) { t = cdiff(ptr, ptr + 3); if (t < idx) idx = t; }
                if (m[-mstep] > 0) { t = cdiff(ptr, ptr - istep); if (t < idx) idx = t; }
                if (m[mstep] > 0) { t = cdiff(ptr, ptr + istep); if (t < idx) idx = t; }
                PUSH(idx, i * w + j);
                m[0] = WS_IN_QUEUE;
            }
        }
    }
    int active;
    for (active = 0; active < 256; active++) if (head[active] >= 0) break;
    if (active == 256) { free(nxt); return 0; }
    for (;;) {
        if (head[active] < 0) {
            int i;
            for (i = active + 1; i < 256; i++) if (head[i] >= 0) break;
            if (i == 256) break;
            active = i;
        }
        int32_t p = head[active];
        head[active] = nxt[p];
        pops++;
        int pi = p / w, pj = p - pi * w;
        int32_t* m = mask + (size_t)pi * mstep + pj;
        const uint8_t* ptr = img + (size_t)pi * istep + pj * 3;
        int lab = 0, t;
        t = m[-1]; if (t > 0) lab = t;
        t = m[1]; if (t > 0) { if (lab == 0) lab = t; else if (t != lab) lab = WS_WSHED; }
        t = m[-mstep]; if (t > 0) { if (lab == 0) lab = t; else if (t != lab) lab = WS_WSHED; }
        t = m[mstep]; if (t > 0) { if (lab == 0) lab = t; else if (t != lab) lab = WS_WSHED; }
        m[0] = lab;
        if (lab == WS_WSHED) continue;
        if (m[-1] == 0) { t = cdiff(ptr, ptr - 3); PUSH(t, p - 1); if (t < active) active = t; m[-1] = WS_IN_QUEUE; }
        if (m[1] == 0) { t = cdiff(ptr, ptr + 3); PUSH(t, p + 1); if (t < active) active = t; m[1] = WS_IN_QUEUE; }
        if (m[-mstep] == 0) { t = cdiff(ptr, ptr - istep); PUSH(t, p - w); if (t < active) active = t; m[-mstep] = WS_IN_QUEUE; }
        if (m[mstep] == 0) { t = cdiff(ptr, ptr + istep); PUSH(t, p + w); if (t < active) active = t; m[mstep] = WS_IN_QUEUE; }
    }
#undef PUSH
    free(nxt);
    return pops;
}
