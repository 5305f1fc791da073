// Context, memory and timing plumbing of the C ABI (include/ofxcv_abi.h).  No image arithmetic here.
#include <stdlib.h>

#include <atomic>
#include <thread>
#if defined(__SSE2__) || defined(__x86_64__)
#include <emmintrin.h>
#endif

#include "common.cuh"

static std::atomic<uint64_t> g_h2d_bytes{0}, g_d2h_bytes{0};

namespace {
int xfer_workers()
{
    static const int forced = getenv("OFXCV_XFER_THREADS") ? atoi(getenv("OFXCV_XFER_THREADS")) : 0;
    if (forced > 0) return forced;
    const unsigned hc = std::thread::hardware_concurrency();
    return hc >= 32 ? 8 : hc >= 8 ? 4 : hc >= 4 ? 2 : 1;
}
// Row copies of the staging pipeline.  Rows of an image are tens of KB each -- below the size at which glibc's memcpy
// switches to non-temporal stores -- so a plain memcpy write-allocates every destination line (read for ownership): three
// memory transfers per byte instead of two, on a path that is bound by the host's memory bandwidth (266 MB of row copies per
// 4K VectorGenerator render).  Streaming stores (SSE2, baseline x86-64) skip the read.  OFXCV_XFER_NT=0 restores memcpy.
bool xfer_nt()
{
    static const bool v = !(getenv("OFXCV_XFER_NT") && atoi(getenv("OFXCV_XFER_NT")) == 0);
    return v;
}
inline void copy_row(char* dst, const char* src, size_t n)
{
#if defined(__SSE2__) || defined(__x86_64__)
    if (n >= 4096 && xfer_nt()) {
        const size_t head = (16 - ((uintptr_t)dst & 15)) & 15;
        if (head) {
            memcpy(dst, src, head);
            dst += head; src += head; n -= head;
        }
        size_t i = 0;
        for (; i + 64 <= n; i += 64) {
            const __m128i a = _mm_loadu_si128((const __m128i*)(src + i)), b = _mm_loadu_si128((const __m128i*)(src + i + 16));
            const __m128i c = _mm_loadu_si128((const __m128i*)(src + i + 32)), d = _mm_loadu_si128((const __m128i*)(src + i + 48));
            _mm_stream_si128((__m128i*)(dst + i), a);
            _mm_stream_si128((__m128i*)(dst + i + 16), b);
            _mm_stream_si128((__m128i*)(dst + i + 32), c);
            _mm_stream_si128((__m128i*)(dst + i + 48), d);
        }
        if (i < n) memcpy(dst + i, src + i, n - i);
        return;
    }
#endif
    memcpy(dst, src, n);
}
inline void copy_fence()
{
#if defined(__SSE2__) || defined(__x86_64__)
    _mm_sfence();  // streaming stores become visible (to the DMA engine, to the caller's threads) before the chunk is handed on
#endif
}

// runs work() on up to n threads; a thread that cannot be created is replaced by the calling thread doing the work
template <class F>
void run_workers(int n, F&& work, std::vector<std::thread>& th)
{
    for (int t = 0; t < n; t++) {
        try {
            th.emplace_back(work);
        } catch (...) {
            break;  // EAGAIN etc.: whoever is running (or the caller, in join_workers) picks the chunks up
        }
    }
}
}  // namespace

int ofxcv_fail(ofxcv_ctx* ctx, cudaError_t e, const char* what)
{
    if (ctx) {
        ctx->last_error = std::string(what) + ": " + cudaGetErrorString(e);
    }
    cudaGetLastError();  // clear the sticky-less error state
    if (e == cudaErrorMemoryAllocation) return OFXCV_ERR_MEMORY;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return OFXCV_ERR_NO_DEVICE;
    return OFXCV_ERR_CUDA;
}

void* ofxcv_ws(ofxcv_ctx* ctx, int slot, size_t bytes)
{
    ofxcv_buf& b = ctx->ws[slot];
    if (b.cap >= bytes && b.p) return b.p;
    if (b.p) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
    }
    size_t cap = (bytes + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc(&b.p, cap);
    if (e != cudaSuccess) {
        ofxcv_fail(ctx, e, "cudaMalloc(workspace)");
        b.p = nullptr;
        return nullptr;
    }
    b.cap = cap;
    return b.p;
}

void* ofxcv_pin(ofxcv_ctx* ctx, int slot, size_t bytes)
{
    ofxcv_buf& b = ctx->pin[slot];
    if (b.cap >= bytes && b.p) return b.p;
    if (b.p) {
        cudaStreamSynchronize(ctx->stream);
        cudaFreeHost(b.p);
        b.p = nullptr;
        b.cap = 0;
    }
    cudaError_t e = cudaMallocHost(&b.p, bytes);
    if (e != cudaSuccess) {
        ofxcv_fail(ctx, e, "cudaMallocHost(staging)");
        b.p = nullptr;
        return nullptr;
    }
    b.cap = bytes;
    return b.p;
}

void ofxcv_time_begin(ofxcv_ctx* ctx, int family, cudaStream_t s)
{
    if (!ctx->timing) return;
    ofxcv_timed_launch t;
    if (!ctx->event_pool.empty()) {
        t = ctx->event_pool.back();
        ctx->event_pool.pop_back();
    } else {
        cudaEventCreate(&t.a);
        cudaEventCreate(&t.b);
    }
    cudaEventRecord(t.a, s);
    ctx->timed[family].push_back(t);
}

void ofxcv_time_end(ofxcv_ctx* ctx, int family, cudaStream_t s)
{
    if (!ctx->timing) return;
    cudaEventRecord(ctx->timed[family].back().b, s);
}

extern "C" {

int ofxcv_abi_version(void) { return 1; }

const char* ofxcv_status_string(int status)
{
    switch (status) {
        case OFXCV_OK: return "ok";
        case OFXCV_ABORTED: return "aborted by the host";
        case OFXCV_ERR_BAD_ARG: return "bad argument";
        case OFXCV_ERR_NO_DEVICE: return "no CUDA device";
        case OFXCV_ERR_MEMORY: return "out of device or pinned memory";
        case OFXCV_ERR_CUDA: return "CUDA error";
        case OFXCV_ERR_UNSUPPORTED: return "unsupported parameter";
    }
    return "unknown status";
}

int ofxcv_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

ofxcv_ctx* ofxcv_create(int device)
{
    int n = ofxcv_device_count();
    if (n <= 0) return nullptr;
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) return nullptr;
    }
    if (device >= n) return nullptr;
    ofxcv_ctx* ctx = new ofxcv_ctx();
    ctx->device = device;
    ofxcv_device_guard g(device);
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        delete ctx;
        return nullptr;
    }
    cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device);
    return ctx;
}

void ofxcv_destroy(ofxcv_ctx* ctx)
{
    if (!ctx) return;
    for (auto& c : ctx->sub) {
        if (c) ofxcv_destroy(c);
        c = nullptr;
    }
    ofxcv_device_guard g(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& b : ctx->ws)
        if (b.p) cudaFree(b.p);
    for (auto& b : ctx->pin)
        if (b.p) cudaFreeHost(b.p);
    for (auto& y : ctx->fb_pyr) {
        if (y.buf) cudaFree(y.buf);
        if (y.built) cudaEventDestroy(y.built);
        for (auto& e : y.used)
            if (e) cudaEventDestroy(e);
    }
    for (auto& st : ctx->stream_lane)
        if (st) cudaStreamDestroy(st);
    for (auto& e : ctx->lane_done)
        if (e) cudaEventDestroy(e);
    if (ctx->lane_start) cudaEventDestroy(ctx->lane_start);
    if (ctx->stream_up) cudaStreamDestroy(ctx->stream_up);
    if (ctx->stream_down) cudaStreamDestroy(ctx->stream_down);
    for (auto& e : ctx->tv_ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : ctx->seq_ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : ctx->xfer_ev) cudaEventDestroy(e);
    if (ctx->xfer_up_done) cudaEventDestroy(ctx->xfer_up_done);
    if (ctx->order_ev) cudaEventDestroy(ctx->order_ev);
    for (int f = 0; f < 3; f++)
        for (auto& t : ctx->timed[f]) {
            cudaEventDestroy(t.a);
            cudaEventDestroy(t.b);
        }
    for (auto& r : ctx->prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    for (auto& t : ctx->event_pool) {
        cudaEventDestroy(t.a);
        cudaEventDestroy(t.b);
    }
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int ofxcv_device(const ofxcv_ctx* ctx) { return ctx ? ctx->device : -1; }
ofxcv_stream ofxcv_ctx_stream(ofxcv_ctx* ctx) { return ctx ? (ofxcv_stream)ctx->stream : nullptr; }

int ofxcv_synchronize(ofxcv_ctx* ctx)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    ofxcv_device_guard g(ctx->device);
    OFXCV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return OFXCV_OK;
}

const char* ofxcv_last_error(const ofxcv_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "no context"; }
uint64_t ofxcv_launch_count(const ofxcv_ctx* ctx) { return ctx ? ctx->launches : 0; }

void ofxcv_kernel_time_enable(ofxcv_ctx* ctx, int enable)
{
    if (!ctx) return;
    ctx->timing = enable != 0;
}

uint64_t ofxcv_kernel_time_ms(ofxcv_ctx* ctx, int family, double* total_ms)
{
    if (!ctx || family < 0 || family > 2) return 0;
    ofxcv_device_guard g(ctx->device);
    // fold finished events into the accumulators, then report and reset
    for (auto& t : ctx->timed[family]) {
        cudaEventSynchronize(t.b);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) {
            ctx->timed_ms[family] += ms;
            ctx->timed_n[family]++;
        }
        ctx->event_pool.push_back(t);
    }
    ctx->timed[family].clear();
    uint64_t n = ctx->timed_n[family];
    if (total_ms) *total_ms = ctx->timed_ms[family];
    ctx->timed_ms[family] = 0;
    ctx->timed_n[family] = 0;
    return n;
}

void ofxcv_prof_enable(ofxcv_ctx* ctx, int enable)
{
    if (ctx) ctx->prof_on = enable != 0;
}

size_t ofxcv_prof_report(ofxcv_ctx* ctx, char* buf, size_t cap)
{
    if (!ctx) return 0;
    ofxcv_device_guard g(ctx->device);
    struct agg {
        const char* name;
        int tag;
        double ms;
        uint64_t n;
    };
    std::vector<agg> rows;
    for (auto& r : ctx->prof) {
        cudaEventSynchronize(r.b);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) ms = 0.f;
        bool found = false;
        for (auto& a : rows)
            if (a.tag == r.tag && !strcmp(a.name, r.name)) {
                a.ms += ms;
                a.n++;
                found = true;
                break;
            }
        if (!found) rows.push_back({r.name, r.tag, ms, 1});
        ofxcv_timed_launch t;
        t.a = r.a;
        t.b = r.b;
        ctx->event_pool.push_back(t);
    }
    ctx->prof.clear();
    std::string out;
    char line[256];
    for (auto& a : rows) {
        snprintf(line, sizeof line, "%s %d %llu %.6f\n", a.name, a.tag, (unsigned long long)a.n, a.ms);
        out += line;
    }
    if (buf && cap) {
        size_t n = out.size() < cap - 1 ? out.size() : cap - 1;
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return out.size();
}

void* ofxcv_device_alloc(ofxcv_ctx* ctx, size_t bytes)
{
    if (!ctx) return nullptr;
    ofxcv_device_guard g(ctx->device);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        ofxcv_fail(ctx, e, "cudaMalloc");
        return nullptr;
    }
    return p;
}

void* ofxcv_scratch_device(ofxcv_ctx* ctx, int slot, size_t bytes)
{
    if (!ctx || slot < 0 || slot >= 8) return nullptr;
    ofxcv_device_guard g(ctx->device);
    return ofxcv_ws(ctx, WS_COUNT + slot, bytes ? bytes : 1);
}

void* ofxcv_scratch_pinned(ofxcv_ctx* ctx, int slot, size_t bytes)
{
    if (!ctx || slot < 0 || slot >= 8) return nullptr;
    ofxcv_device_guard g(ctx->device);
    return ofxcv_pin(ctx, 4 + slot, bytes ? bytes : 1);
}

void ofxcv_device_free(ofxcv_ctx* ctx, void* dptr)
{
    if (!ctx || !dptr) return;
    ofxcv_device_guard g(ctx->device);
    cudaFree(dptr);
}

void* ofxcv_pinned_alloc(ofxcv_ctx* ctx, size_t bytes)
{
    if (!ctx) return nullptr;
    ofxcv_device_guard g(ctx->device);
    void* p = nullptr;
    cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        ofxcv_fail(ctx, e, "cudaMallocHost");
        return nullptr;
    }
    return p;
}

void ofxcv_pinned_free(ofxcv_ctx* ctx, void* hptr)
{
    if (!ctx || !hptr) return;
    ofxcv_device_guard g(ctx->device);
    cudaFreeHost(hptr);
}

static cudaStream_t pick(ofxcv_ctx* ctx, ofxcv_stream s) { return s ? (cudaStream_t)s : ctx->stream; }

int ofxcv_upload(ofxcv_ctx* ctx, ofxcv_stream s, void* dst_dev, const void* src_host, size_t bytes)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    if (!dst_dev || !src_host) return OFXCV_ERR_BAD_ARG;
    ofxcv_device_guard g(ctx->device);
    OFXCV_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, pick(ctx, s)));
    g_h2d_bytes += bytes;
    return OFXCV_OK;
}

int ofxcv_download(ofxcv_ctx* ctx, ofxcv_stream s, void* dst_host, const void* src_dev, size_t bytes)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    if (!dst_host || !src_dev) return OFXCV_ERR_BAD_ARG;
    ofxcv_device_guard g(ctx->device);
    OFXCV_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, pick(ctx, s)));
    g_d2h_bytes += bytes;
    return OFXCV_OK;
}

int ofxcv_device_copy(ofxcv_ctx* ctx, ofxcv_stream s, void* dst_dev, const void* src_dev, size_t bytes)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    if (!dst_dev || !src_dev) return OFXCV_ERR_BAD_ARG;
    ofxcv_device_guard g(ctx->device);
    OFXCV_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, pick(ctx, s)));
    return OFXCV_OK;
}

int ofxcv_memset(ofxcv_ctx* ctx, ofxcv_stream s, void* dst_dev, int value, size_t bytes)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    if (!dst_dev) return OFXCV_ERR_BAD_ARG;
    ofxcv_device_guard g(ctx->device);
    OFXCV_CUDA(ctx, cudaMemsetAsync(dst_dev, value, bytes, pick(ctx, s)));
    return OFXCV_OK;
}

ofxcv_stream ofxcv_aux_stream(ofxcv_ctx* ctx)
{
    if (!ctx) return nullptr;
    ofxcv_device_guard g(ctx->device);
    if (!ctx->stream_up && cudaStreamCreateWithFlags(&ctx->stream_up, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return (ofxcv_stream)ctx->stream_up;
}

int ofxcv_stream_wait(ofxcv_ctx* ctx, ofxcv_stream waiter, ofxcv_stream signaller)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    ofxcv_device_guard g(ctx->device);
    if (!ctx->order_ev) OFXCV_CUDA(ctx, cudaEventCreateWithFlags(&ctx->order_ev, cudaEventDisableTiming));
    OFXCV_CUDA(ctx, cudaEventRecord(ctx->order_ev, pick(ctx, signaller)));
    OFXCV_CUDA(ctx, cudaStreamWaitEvent(pick(ctx, waiter), ctx->order_ev, 0));
    return OFXCV_OK;
}

int ofxcv_stream_synchronize(ofxcv_ctx* ctx, ofxcv_stream s)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    ofxcv_device_guard g(ctx->device);
    OFXCV_CUDA(ctx, cudaStreamSynchronize(pick(ctx, s)));
    return OFXCV_OK;
}

int ofxcv_current_device(void)
{
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return d;
}

int ofxcv_pointer_device(const void* p)
{
    cudaPointerAttributes a;
    if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? a.device : -1;
}

void ofxcv_set_abort_callback(ofxcv_ctx* ctx, int (*cb)(void*), void* user)
{
    if (!ctx) return;
    ctx->abort_cb = cb;
    ctx->abort_user = user;
}

void ofxcv_transfer_stats(uint64_t* h2d_bytes, uint64_t* d2h_bytes)
{
    if (h2d_bytes) *h2d_bytes = g_h2d_bytes.load();
    if (d2h_bytes) *d2h_bytes = g_d2h_bytes.load();
}

// rows of a host image <-> tight device rows.  Large images are cut into row chunks: a few workers copy chunks between the
// caller's (pageable) rows and the context's pinned staging while the DMA of the chunks that are ready already runs.

int ofxcv_upload_rows(ofxcv_ctx* ctx, ofxcv_stream s_, void* dst_dev, const void* src_host, ptrdiff_t src_stride, size_t row_bytes, int rows)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    if (!dst_dev || !src_host || rows <= 0 || row_bytes == 0) return OFXCV_ERR_BAD_ARG;
    ofxcv_device_guard g(ctx->device);
    cudaStream_t s = pick(ctx, s_);
    const size_t total = row_bytes * (size_t)rows;
    if (src_stride == (ptrdiff_t)row_bytes && ofxcv_is_pinned(src_host)) {  // page-locked and dense: straight over PCIe
        OFXCV_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, total, cudaMemcpyHostToDevice, s));
        g_h2d_bytes += total;
        return OFXCV_OK;
    }
    char* pinned = (char*)ofxcv_pin(ctx, 2, total);
    if (!pinned) return OFXCV_ERR_MEMORY;
    if (!ctx->xfer_up_done) OFXCV_CUDA(ctx, cudaEventCreateWithFlags(&ctx->xfer_up_done, cudaEventDisableTiming));
    else OFXCV_CUDA(ctx, cudaEventSynchronize(ctx->xfer_up_done));  // the previous DMA out of the staging buffer has finished
    const int nch = rows >= 256 && total >= ((size_t)8 << 20) ? 16 : 1;
    const char* src = (const char*)src_host;
    auto rows_of = [&](int k, int& y0, int& y1) {
        y0 = (int)((long long)rows * k / nch);
        y1 = (int)((long long)rows * (k + 1) / nch);
    };
    std::vector<std::atomic<int>> done(nch);
    for (auto& d : done) d.store(0, std::memory_order_relaxed);
    std::atomic<int> next{0};
    auto work = [&]() {
        for (;;) {
            const int k = next.fetch_add(1);
            if (k >= nch) return;
            int y0, y1;
            rows_of(k, y0, y1);
            for (int y = y0; y < y1; y++) copy_row(pinned + (size_t)y * row_bytes, src + (ptrdiff_t)y * src_stride, row_bytes);
            copy_fence();
            done[k].store(1, std::memory_order_release);
        }
    };
    std::vector<std::thread> th;
    if (nch > 1) run_workers(xfer_workers(), work, th);
    if (th.empty()) work();
    cudaError_t err = cudaSuccess;
    for (int k = 0; k < nch; k++) {
        while (!done[k].load(std::memory_order_acquire)) {
            if (th.empty()) break;
            std::this_thread::yield();
        }
        int y0, y1;
        rows_of(k, y0, y1);
        if (err == cudaSuccess)
            err = cudaMemcpyAsync((char*)dst_dev + (size_t)y0 * row_bytes, pinned + (size_t)y0 * row_bytes, (size_t)(y1 - y0) * row_bytes,
                                  cudaMemcpyHostToDevice, s);
    }
    for (auto& t : th) t.join();
    if (err != cudaSuccess) return ofxcv_fail(ctx, err, "cudaMemcpyAsync(upload rows)");
    OFXCV_CUDA(ctx, cudaEventRecord(ctx->xfer_up_done, s));
    g_h2d_bytes += total;
    return OFXCV_OK;
}

int ofxcv_download_rows(ofxcv_ctx* ctx, ofxcv_stream s_, void* dst_host, ptrdiff_t dst_stride, const void* src_dev, size_t row_bytes, int rows)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    if (!dst_host || !src_dev || rows <= 0 || row_bytes == 0) return OFXCV_ERR_BAD_ARG;
    ofxcv_device_guard g(ctx->device);
    cudaStream_t s = pick(ctx, s_);
    const size_t total = row_bytes * (size_t)rows;
    if (dst_stride == (ptrdiff_t)row_bytes && ofxcv_is_pinned(dst_host)) {
        OFXCV_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, total, cudaMemcpyDeviceToHost, s));
        OFXCV_CUDA(ctx, cudaStreamSynchronize(s));
        g_d2h_bytes += total;
        return OFXCV_OK;
    }
    char* pinned = (char*)ofxcv_pin(ctx, 3, total);
    if (!pinned) return OFXCV_ERR_MEMORY;
    const int nch = rows >= 256 && total >= ((size_t)8 << 20) ? 16 : 1;
    while ((int)ctx->xfer_ev.size() < nch) {
        cudaEvent_t e;
        OFXCV_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->xfer_ev.push_back(e);
    }
    auto rows_of = [&](int k, int& y0, int& y1) {
        y0 = (int)((long long)rows * k / nch);
        y1 = (int)((long long)rows * (k + 1) / nch);
    };
    for (int k = 0; k < nch; k++) {  // the whole download is enqueued first; the workers follow the chunks as they land
        int y0, y1;
        rows_of(k, y0, y1);
        OFXCV_CUDA(ctx, cudaMemcpyAsync(pinned + (size_t)y0 * row_bytes, (const char*)src_dev + (size_t)y0 * row_bytes,
                                        (size_t)(y1 - y0) * row_bytes, cudaMemcpyDeviceToHost, s));
        OFXCV_CUDA(ctx, cudaEventRecord(ctx->xfer_ev[k], s));
    }
    char* dst = (char*)dst_host;
    std::atomic<int> next{0};
    std::atomic<int> failed{0};
    const int device = ctx->device;
    auto work = [&]() {
        cudaSetDevice(device);
        for (;;) {
            const int k = next.fetch_add(1);
            if (k >= nch) break;
            if (cudaEventSynchronize(ctx->xfer_ev[k]) != cudaSuccess) {
                failed.store(1);
                break;
            }
            int y0, y1;
            rows_of(k, y0, y1);
            for (int y = y0; y < y1; y++) copy_row(dst + (ptrdiff_t)y * dst_stride, pinned + (size_t)y * row_bytes, row_bytes);
        }
        copy_fence();
    };
    std::vector<std::thread> th;
    if (nch > 1) run_workers(xfer_workers() - 1, work, th);
    work();  // the calling thread takes its share (and everything, should no thread have started)
    for (auto& t : th) t.join();
    if (failed.load()) return ofxcv_fail(ctx, cudaGetLastError(), "cudaEventSynchronize(download rows)");
    g_d2h_bytes += total;
    return OFXCV_OK;
}

}  // extern "C"
