// Context, memory and timing plumbing of the C ABI (include/ofxcv_abi.h).  No image arithmetic here.
#include "common.cuh"

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
    return OFXCV_OK;
}

int ofxcv_download(ofxcv_ctx* ctx, ofxcv_stream s, void* dst_host, const void* src_dev, size_t bytes)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    if (!dst_host || !src_dev) return OFXCV_ERR_BAD_ARG;
    ofxcv_device_guard g(ctx->device);
    OFXCV_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, pick(ctx, s)));
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

}  // extern "C"
