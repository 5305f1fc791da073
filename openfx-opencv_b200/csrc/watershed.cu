// Marker-based watershed for sm_100a: cv::watershed(rgb8, int32 markers) semantics (Meyer flooding with 256 FIFO
// buckets), the body BASELINE.json config 3 names for the segment plugin (the reference calls
// cvPyrSegmentation at /root/reference/opencv2fx/segment/segment.cpp:296-302; see SURVEY.md section 0 fact 2 and
// Appendix A.3 for the verified algorithm).  Integer-only; the label map is bit-exact with the CPU path.
//
// The flood is a strictly ordered best-first process (queue level fixed by the first parent, FIFO ties, label
// decided at pop time), so ONE thread per frame walks the queues; everything around it is data-parallel:
//   ws_prepare     border ring := -1, negatives := 0, RGB8 -> packed RGBX words (one 32-bit load per pixel)
//   ws_candidates  per pixel: is it an unlabelled neighbour of a seed, and its initial bucket
//   ws_link        per frame: link the candidates into the 256 intrusive FIFOs in row-major order (ordered
//                  block compaction), 256 threads = 256 buckets
//   ws_flood       per frame: the flood; per pop ONE round of ten independent loads (next link, 4 labels,
//                  5 packed pixels), so the cost per pop is one L2/HBM latency, and throughput over a sequence
//                  comes from many frames in flight per SM (ofxcv_watershed_u8c3_batch).
// Per-frame state: nxt int32 (intrusive FIFO link, every pixel is queued at most once) + pix u32, pitch = the
// marker pitch.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace {

constexpr int WS_IN_QUEUE = -2;
constexpr int WS_WSHED = -1;

__device__ __forceinline__ int pixdiff(uint32_t a, uint32_t b)
{
    // max over the three colour bytes of |a-b|; the 4th byte of both words is 0
    uint32_t d = __vabsdiffu4(a, b);
    int d0 = d & 0xff, d1 = (d >> 8) & 0xff, d2 = (d >> 16) & 0xff;
    return max(d0, max(d1, d2));
}

__global__ void __launch_bounds__(256) ws_prepare(const uint8_t* __restrict__ rgb, ptrdiff_t rgb_stride, size_t rgb_frame,
                                                  int32_t* __restrict__ m, ptrdiff_t ms, size_t m_frame,
                                                  uint32_t* __restrict__ pix, size_t st_frame, int w, int h)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    int f = blockIdx.z;
    if (x >= w) return;
    const uint8_t* p = rgb + f * rgb_frame + (size_t)y * rgb_stride + (size_t)x * 3;
    pix[f * st_frame + (size_t)y * ms + x] = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
    int32_t* mp = m + f * m_frame + (size_t)y * ms + x;
    if (x == 0 || y == 0 || x == w - 1 || y == h - 1) *mp = WS_WSHED;
    else if (*mp < 0) *mp = 0;
}

// nxt[p] := -1 (not a candidate) or -2-idx (candidate for bucket idx)
__global__ void __launch_bounds__(256) ws_candidates(const int32_t* __restrict__ m, ptrdiff_t ms, size_t m_frame,
                                                     const uint32_t* __restrict__ pix, int32_t* __restrict__ nxt,
                                                     size_t st_frame, int w, int h)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    int f = blockIdx.z;
    if (x >= w) return;
    size_t o = (size_t)y * ms + x;
    int32_t code = -1;
    if (x > 0 && y > 0 && x < w - 1 && y < h - 1) {
        const int32_t* mp = m + f * m_frame + o;
        if (mp[0] == 0) {
            const uint32_t* pp = pix + f * st_frame + o;
            int idx = 256;
            uint32_t c = pp[0];
            if (mp[-1] > 0) idx = pixdiff(c, pp[-1]);
            if (mp[1] > 0) idx = min(idx, pixdiff(c, pp[1]));
            if (mp[-ms] > 0) idx = min(idx, pixdiff(c, pp[-ms]));
            if (mp[ms] > 0) idx = min(idx, pixdiff(c, pp[ms]));
            if (idx < 256) code = -2 - idx;
        }
    }
    nxt[f * st_frame + o] = code;
}

// one CTA (1024 threads) per frame: ordered compaction of the candidates, tile by tile in row-major order, then
// thread b appends the tile's candidates of bucket b to FIFO b.  Marks them IN_QUEUE.
__global__ void __launch_bounds__(1024) ws_link(int32_t* __restrict__ m, ptrdiff_t ms, size_t m_frame,
                                                int32_t* __restrict__ nxt, size_t st_frame, int32_t* __restrict__ heads,
                                                int w, int h)
{
    __shared__ int32_t s_pix[1024];
    __shared__ int16_t s_idx[1024];
    __shared__ int s_warp[32];
    __shared__ int s_total;
    const int f = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int32_t* mf = m + f * m_frame;
    int32_t* nf = nxt + f * st_frame;
    int32_t head = -1, tail = -1;  // FIFO of bucket `tid` (tid < 256)
    const long total = (long)w * h;
    for (long base = 0; base < total; base += 1024) {
        long i = base + tid;
        int code = -1;
        int32_t o = 0;
        if (i < total) {
            int y = (int)(i / w), x = (int)(i - (long)y * w);
            o = (int32_t)((size_t)y * ms + x);
            code = nf[o];
        }
        bool is = code <= -2;
        unsigned bal = __ballot_sync(0xffffffffu, is);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        if (wid == 0) {
            int v = s_warp[lane];
            int inc = v;
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            s_warp[lane] = inc - v;
            if (lane == 31) s_total = inc;
        }
        __syncthreads();
        if (is) {
            int pos = s_warp[wid] + __popc(bal & ((1u << lane) - 1));
            s_pix[pos] = o;
            s_idx[pos] = (int16_t)(-2 - code);
            mf[o] = WS_IN_QUEUE;
        }
        __syncthreads();
        int cnt = s_total;
        if (tid < 256 && cnt > 0) {
            for (int k = 0; k < cnt; k++) {
                if (s_idx[k] == tid) {
                    int32_t q = s_pix[k];
                    nf[q] = -1;
                    if (head < 0) head = q;
                    else nf[tail] = q;
                    tail = q;
                }
            }
        }
        __syncthreads();
    }
    if (tid < 256) {
        heads[(size_t)f * 512 + tid] = head;
        heads[(size_t)f * 512 + 256 + tid] = tail;
    }
}

// one thread per frame
__global__ void __launch_bounds__(32) ws_flood(int32_t* __restrict__ m, ptrdiff_t ms_, size_t m_frame,
                                               const uint32_t* __restrict__ pix, int32_t* __restrict__ nxt, size_t st_frame,
                                               const int32_t* __restrict__ heads, unsigned long long* __restrict__ pops_out)
{
    __shared__ int32_t head[256], tail[256];
    const int f = blockIdx.x;
    for (int i = threadIdx.x; i < 256; i += 32) {
        head[i] = heads[(size_t)f * 512 + i];
        tail[i] = heads[(size_t)f * 512 + 256 + i];
    }
    __syncwarp();
    if (threadIdx.x != 0) return;
    int32_t* mf = m + f * m_frame;
    const uint32_t* pf = pix + f * st_frame;
    int32_t* nf = nxt + f * st_frame;
    const int ms = (int)ms_;
    unsigned long long pops = 0;
    int active = 0;
    while (active < 256 && head[active] < 0) active++;
    if (active < 256) {
        for (;;) {
            if (head[active] < 0) {
                int i = active + 1;
                while (i < 256 && head[i] < 0) i++;
                if (i == 256) break;
                active = i;
            }
            const int32_t p = head[active];
            // one round of independent loads
            const int32_t np = nf[p];
            const int32_t ml = mf[p - 1], mr = mf[p + 1], mu = mf[p - ms], md = mf[p + ms];
            const uint32_t c = pf[p], cl = pf[p - 1], cr = pf[p + 1], cu = pf[p - ms], cd = pf[p + ms];
            head[active] = np;
            pops++;
            int lab = 0;
            if (ml > 0) lab = ml;
            if (mr > 0) { if (lab == 0) lab = mr; else if (mr != lab) lab = WS_WSHED; }
            if (mu > 0) { if (lab == 0) lab = mu; else if (mu != lab) lab = WS_WSHED; }
            if (md > 0) { if (lab == 0) lab = md; else if (md != lab) lab = WS_WSHED; }
            mf[p] = lab;
            if (lab == WS_WSHED) continue;
#define WS_PUSH(cond, q, cq)                                 \
    if (cond) {                                              \
        int t = pixdiff(c, cq);                              \
        nf[q] = -1;                                          \
        if (head[t] < 0) head[t] = (q);                      \
        else nf[tail[t]] = (q);                              \
        tail[t] = (q);                                       \
        if (t < active) active = t;                          \
        mf[q] = WS_IN_QUEUE;                                 \
    }
            WS_PUSH(ml == 0, p - 1, cl)
            WS_PUSH(mr == 0, p + 1, cr)
            WS_PUSH(mu == 0, p - ms, cu)
            WS_PUSH(md == 0, p + ms, cd)
#undef WS_PUSH
        }
    }
    pops_out[f] = pops;
}

// one thread per frame, software-pipelined: the FIFO link of the pixel being popped names the pixel that will be
// popped next unless this pop pushes to a lower level or empties the queue, so its ten loads are issued BEFORE the
// label / push work of the current pop and land behind it (one memory latency and one ALU chain per pop overlap
// instead of adding up).  The prefetched labels are patched in registers with what the current pop changes (its own
// label, the IN_QUEUE marks of the pixels it pushes, the link it appends behind the prefetched pixel), so the flood is
// exactly the sequential one.
__global__ void __launch_bounds__(32) ws_flood2(int32_t* __restrict__ m, ptrdiff_t ms_, size_t m_frame,
                                                const uint32_t* __restrict__ pix, int32_t* __restrict__ nxt, size_t st_frame,
                                                const int32_t* __restrict__ heads, unsigned long long* __restrict__ pops_out)
{
    __shared__ int32_t head[256], tail[256];
    const int f = blockIdx.x;
    for (int i = threadIdx.x; i < 256; i += 32) {
        head[i] = heads[(size_t)f * 512 + i];
        tail[i] = heads[(size_t)f * 512 + 256 + i];
    }
    __syncwarp();
    if (threadIdx.x != 0) return;
    int32_t* mf = m + f * m_frame;
    const uint32_t* pf = pix + f * st_frame;
    int32_t* nf = nxt + f * st_frame;
    const int ms = (int)ms_;
    unsigned long long pops = 0;
    int active = 0;
    while (active < 256 && head[active] < 0) active++;
    if (active < 256) {
        int32_t p = head[active];
        // the current pop's loads
        int32_t np = nf[p];
        int32_t ml = mf[p - 1], mr = mf[p + 1], mu = mf[p - ms], md = mf[p + ms];
        uint32_t c = pf[p], cl = pf[p - 1], cr = pf[p + 1], cu = pf[p - ms], cd = pf[p + ms];
        for (;;) {
            // prefetch the predicted next pop (the next entry of the active queue)
            const int32_t pn = np;
            int32_t n2 = -1, nl = 0, nr = 0, nu = 0, nd = 0;
            uint32_t e = 0, el = 0, er_ = 0, eu = 0, ed = 0;
            if (pn >= 0) {
                n2 = nf[pn];
                nl = mf[pn - 1]; nr = mf[pn + 1]; nu = mf[pn - ms]; nd = mf[pn + ms];
                e = pf[pn]; el = pf[pn - 1]; er_ = pf[pn + 1]; eu = pf[pn - ms]; ed = pf[pn + ms];
            }
            // ---- the pop of p -------------------------------------------------------------------------------
            head[active] = np;
            pops++;
            int lab = 0;
            if (ml > 0) lab = ml;
            if (mr > 0) { if (lab == 0) lab = mr; else if (mr != lab) lab = WS_WSHED; }
            if (mu > 0) { if (lab == 0) lab = mu; else if (mu != lab) lab = WS_WSHED; }
            if (md > 0) { if (lab == 0) lab = md; else if (md != lab) lab = WS_WSHED; }
            mf[p] = lab;
            // the prefetched registers need patching only when the predicted pixel lies within two rows of p
            const bool near = (unsigned)(pn - p + 2 * ms + 2) <= (unsigned)(4 * ms + 4);
            // p's label as seen by the prefetched pixel
            if (near) {
                if (pn - 1 == p) nl = lab;
                if (pn + 1 == p) nr = lab;
                if (pn - ms == p) nu = lab;
                if (pn + ms == p) nd = lab;
            }
            const int level_at_pop = active;
            if (lab != WS_WSHED) {
#define WS_PUSH2(cond, q, cq)                                        \
    if (cond) {                                                      \
        const int t = pixdiff(c, cq);                                \
        nf[q] = -1;                                                  \
        if (head[t] < 0) head[t] = (q);                              \
        else {                                                       \
            if (tail[t] == pn) n2 = (q); /* appended behind the prefetched pixel */ \
            nf[tail[t]] = (q);                                       \
        }                                                            \
        tail[t] = (q);                                               \
        if (t < active) active = t;                                  \
        mf[q] = WS_IN_QUEUE;                                         \
        if (near) {                                                  \
            if (pn - 1 == (q)) nl = WS_IN_QUEUE;                     \
            if (pn + 1 == (q)) nr = WS_IN_QUEUE;                     \
            if (pn - ms == (q)) nu = WS_IN_QUEUE;                    \
            if (pn + ms == (q)) nd = WS_IN_QUEUE;                    \
        }                                                            \
    }
                WS_PUSH2(ml == 0, p - 1, cl)
                WS_PUSH2(mr == 0, p + 1, cr)
                WS_PUSH2(mu == 0, p - ms, cu)
                WS_PUSH2(md == 0, p + ms, cd)
#undef WS_PUSH2
            }
            // ---- who is next? -------------------------------------------------------------------------------
            if (active == level_at_pop && pn >= 0) {
                // the prediction holds: continue with the prefetched (and patched) registers
                p = pn;
                np = n2;
                ml = nl; mr = nr; mu = nu; md = nd;
                c = e; cl = el; cr = er_; cu = eu; cd = ed;
                continue;
            }
            if (head[active] < 0) {
                int i = active + 1;
                while (i < 256 && head[i] < 0) i++;
                if (i == 256) break;
                active = i;
            }
            p = head[active];
            np = nf[p];
            ml = mf[p - 1]; mr = mf[p + 1]; mu = mf[p - ms]; md = mf[p + ms];
            c = pf[p]; cl = pf[p - 1]; cr = pf[p + 1]; cu = pf[p - ms]; cd = pf[p + ms];
        }
    }
    pops_out[f] = pops;
}

// Measured and rejected on top of ws_flood2 (3840x2160, 256 seeds: 3.7 s, 187 instructions and ~830 cycles per pop, of
// which ncu attributes 35 % to memory and the rest to issue + dependent-ALU latency of the one warp):
//  * warming L1 for queued pixels -- `prefetch.global.L1` does not allocate in L1 on sm_100a, a 4-byte cp.async.ca into
//    a shared-memory sink does (tools/microbench/l1_prefetch.cu: 470 vs 39 cycles per dependent load), but the 20+
//    extra instructions per push cost more than the misses they remove: 6.0 s;
//  * one warp per frame with lanes 1-4 owning the four neighbours (label rule = two ballots, same-level pushes linked
//    in lane order, level scan by ballot): bit-identical, but every cross-lane step (shuffle, ballot, __syncwarp,
//    shared-memory hand-off) has a longer dependent latency than the scalar code it replaces: 5.2 s.

}  // namespace

extern "C" {

size_t ofxcv_watershed_workspace_bytes(int W, int H, int nframes)
{
    return ((size_t)W * H * 8 + 512 * 4 + 8) * (size_t)(nframes > 0 ? nframes : 1) + 1024;
}

int ofxcv_watershed_u8c3_batch(ofxcv_ctx* ctx, ofxcv_stream stream_, const uint8_t* rgb, ptrdiff_t rgb_stride,
                               size_t rgb_frame_stride, int32_t* markers, ptrdiff_t markers_stride,
                               size_t markers_frame_stride, int W, int H, int nframes)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    if (!rgb || !markers || W < 3 || H < 3 || nframes < 1 || nframes > 65535 || rgb_stride < (ptrdiff_t)W * 3 ||
        markers_stride < (ptrdiff_t)W * 4 || (markers_stride & 3) || (markers_frame_stride & 3))
        return OFXCV_ERR_BAD_ARG;
    if ((size_t)markers_stride / 4 * H >= ((size_t)1 << 31)) return OFXCV_ERR_UNSUPPORTED;
    ofxcv_device_guard guard(ctx->device);
    cudaStream_t s = stream_ ? (cudaStream_t)stream_ : ctx->stream;
    const ptrdiff_t ms = markers_stride / 4;
    const size_t st_frame = (size_t)ms * H;  // elements per frame in nxt / pix
    const size_t m_frame = markers_frame_stride / 4;
    int32_t* nxt = (int32_t*)ofxcv_ws(ctx, WS_WS_NEXT, st_frame * 4 * nframes);
    uint32_t* pix = (uint32_t*)ofxcv_ws(ctx, WS_MISC0, st_frame * 4 * nframes);
    int32_t* heads = (int32_t*)ofxcv_ws(ctx, WS_MISC1, (size_t)nframes * 512 * 4);
    unsigned long long* pops = (unsigned long long*)ofxcv_ws(ctx, WS_MISC2, (size_t)nframes * 8);
    if (!nxt || !pix || !heads || !pops) return OFXCV_ERR_MEMORY;
    dim3 grid(ofxcv_div_up(W, 256), H, nframes);
    ws_prepare<<<grid, 256, 0, s>>>(rgb, rgb_stride, rgb_frame_stride, markers, ms, m_frame, pix, st_frame, W, H);
    OFXCV_LAUNCH_CHECK(ctx);
    // Few frames: the exact intra-frame parallel flood (watershed_par.cu), one frame after the other on the whole GPU.
    // Many frames: one thread per frame, all frames in flight (the parallel flood's ~5 frames/s is reached by the
    // one-thread kernel from ~16 concurrent frames up).  OFXCV_WS_MODE=seq|par overrides.
    const char* mode = getenv("OFXCV_WS_MODE");
    const bool par = mode && !strcmp(mode, "par") ? true : mode && !strcmp(mode, "seq") ? false : nframes < 16;
    std::vector<int> todo;  // frames left to the one-thread flood
    int64_t par_pops = 0;
    if (par) {
        ofxcv_time_begin(ctx, 2, s);
        for (int f = 0; f < nframes; f++) {
            int64_t pops_f = 0;
            const int st = ofxcv_wsp_flood(ctx, s, markers + f * m_frame, ms, pix + f * st_frame, W, H, &pops_f);
            if (st < 0) return st;
            if (st == 1) todo.push_back(f);
            else par_pops += pops_f;
        }
        ofxcv_time_end(ctx, 2, s);
    } else {
        for (int f = 0; f < nframes; f++) todo.push_back(f);
    }
    ctx->watershed_par_pops = par_pops;
    ctx->watershed_seq_frames.assign(todo.begin(), todo.end());
    ctx->watershed_stats[0] = par_pops;
    ctx->watershed_stats[1] = nframes;
    if (todo.empty()) return OFXCV_OK;
    ctx->watershed_stats[2] = ctx->watershed_stats[3] = 0;
    if ((int)todo.size() == nframes) {
        ws_candidates<<<grid, 256, 0, s>>>(markers, ms, m_frame, pix, nxt, st_frame, W, H);
        OFXCV_LAUNCH_CHECK(ctx);
        ws_link<<<nframes, 1024, 0, s>>>(markers, ms, m_frame, nxt, st_frame, heads, W, H);
        OFXCV_LAUNCH_CHECK(ctx);
        ofxcv_time_begin(ctx, 2, s);
        if (getenv("OFXCV_WS_FLOOD_V1")) ws_flood<<<nframes, 32, 0, s>>>(markers, ms, m_frame, pix, nxt, st_frame, heads, pops);
        else ws_flood2<<<nframes, 32, 0, s>>>(markers, ms, m_frame, pix, nxt, st_frame, heads, pops);
        ofxcv_time_end(ctx, 2, s);
        OFXCV_LAUNCH_CHECK(ctx);
    } else {
        for (int f : todo) {  // the degenerate frames of a small batch, one at a time
            dim3 g1(ofxcv_div_up(W, 256), H, 1);
            ws_candidates<<<g1, 256, 0, s>>>(markers + f * m_frame, ms, 0, pix + f * st_frame, nxt + f * st_frame, 0, W, H);
            OFXCV_LAUNCH_CHECK(ctx);
            ws_link<<<1, 1024, 0, s>>>(markers + f * m_frame, ms, 0, nxt + f * st_frame, 0, heads + (size_t)f * 512, W, H);
            OFXCV_LAUNCH_CHECK(ctx);
            ws_flood2<<<1, 32, 0, s>>>(markers + f * m_frame, ms, 0, pix + f * st_frame, nxt + f * st_frame, 0, heads + (size_t)f * 512, pops + f);
            OFXCV_LAUNCH_CHECK(ctx);
        }
    }
    ctx->watershed_stats[0] = -1;  // resolved lazily by ofxcv_watershed_last_stats
    ctx->watershed_stats[1] = nframes;
    return OFXCV_OK;
}

int ofxcv_watershed_u8c3(ofxcv_ctx* ctx, ofxcv_stream stream, const uint8_t* rgb, ptrdiff_t rgb_stride, int32_t* markers,
                         ptrdiff_t markers_stride, int W, int H)
{
    return ofxcv_watershed_u8c3_batch(ctx, stream, rgb, rgb_stride, 0, markers, markers_stride, 0, W, H, 1);
}

int ofxcv_watershed_last_stats(const ofxcv_ctx* ctx_, int64_t stats[4])
{
    ofxcv_ctx* ctx = const_cast<ofxcv_ctx*>(ctx_);
    if (!ctx || !stats) return OFXCV_ERR_BAD_ARG;
    if (ctx->watershed_stats[0] < 0 && ctx->ws[WS_MISC2].p) {
        ofxcv_device_guard guard(ctx->device);
        int n = (int)ctx->watershed_stats[1];
        std::vector<unsigned long long> h(n);
        OFXCV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        OFXCV_CUDA(ctx, cudaMemcpy(h.data(), ctx->ws[WS_MISC2].p, (size_t)n * 8, cudaMemcpyDeviceToHost));
        int64_t tot = ctx->watershed_par_pops;
        for (int f : ctx->watershed_seq_frames) tot += (int64_t)h[f];
        ctx->watershed_stats[0] = tot;
    }
    for (int i = 0; i < 4; i++) stats[i] = ctx->watershed_stats[i];
    return OFXCV_OK;
}

int ofxcv_watershed_u8c3_host(ofxcv_ctx* ctx, const uint8_t* rgb, ptrdiff_t rgb_stride, int32_t* markers,
                              ptrdiff_t markers_stride, int W, int H)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    if (!rgb || !markers || W < 3 || H < 3 || rgb_stride < (ptrdiff_t)W * 3 || markers_stride < (ptrdiff_t)W * 4)
        return OFXCV_ERR_BAD_ARG;
    ofxcv_device_guard guard(ctx->device);
    const size_t nrgb = (size_t)W * H * 3, nm = (size_t)W * H * 4;
    uint8_t* hr = (uint8_t*)ofxcv_pin(ctx, 0, nrgb);
    int32_t* hm = (int32_t*)ofxcv_pin(ctx, 1, nm);
    uint8_t* dr = (uint8_t*)ofxcv_ws(ctx, WS_STAGE_IN0, nrgb);
    int32_t* dm = (int32_t*)ofxcv_ws(ctx, WS_STAGE_OUT, nm);
    if (!hr || !hm || !dr || !dm) return OFXCV_ERR_MEMORY;
    for (int y = 0; y < H; y++) {
        memcpy(hr + (size_t)y * W * 3, rgb + (size_t)y * rgb_stride, (size_t)W * 3);
        memcpy(hm + (size_t)y * W, (const char*)markers + (size_t)y * markers_stride, (size_t)W * 4);
    }
    cudaStream_t s = ctx->stream;
    OFXCV_CUDA(ctx, cudaMemcpyAsync(dr, hr, nrgb, cudaMemcpyHostToDevice, s));
    OFXCV_CUDA(ctx, cudaMemcpyAsync(dm, hm, nm, cudaMemcpyHostToDevice, s));
    int st = ofxcv_watershed_u8c3(ctx, s, dr, (ptrdiff_t)W * 3, dm, (ptrdiff_t)W * 4, W, H);
    if (st < 0) return st;
    OFXCV_CUDA(ctx, cudaMemcpyAsync(hm, dm, nm, cudaMemcpyDeviceToHost, s));
    OFXCV_CUDA(ctx, cudaStreamSynchronize(s));
    for (int y = 0; y < H; y++) memcpy((char*)markers + (size_t)y * markers_stride, hm + (size_t)y * W, (size_t)W * 4);
    return OFXCV_OK;
}

}  // extern "C"
