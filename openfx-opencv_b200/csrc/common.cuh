// Internal to libofxcv_b200.so: the context object and small helpers shared by the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/ofxcv_abi.h"

constexpr int OFXCV_FB_MAX_LANES = 4;  // Farneback pairs in flight in the clip entry points

struct ofxcv_buf {
    void* p = nullptr;
    size_t cap = 0;
};

// one cached Farneback frame pyramid: the polynomial expansion R (5 coefficients per pixel) of a frame at every
// scale, keyed by the caller's frame key so that frame t+1 of pair t is reused as frame t of pair t+1
struct ofxcv_fb_pyr {
    void* buf = nullptr;
    size_t cap = 0;
    uint64_t key = 0;    // 0 = anonymous (never matched)
    uint64_t sig = 0;    // size + the parameters the pyramid depends on
    uint64_t tick = 0;   // LRU
    size_t off_q[16] = {0}, off_s[16] = {0};
    cudaEvent_t built = nullptr;  // recorded on the build stream after the pyramid is complete
    cudaEvent_t used[OFXCV_FB_MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};  // recorded on a solve lane's stream after a solve that read it
    bool used_pending[OFXCV_FB_MAX_LANES] = {false, false, false, false};
};

struct ofxcv_timed_launch {
    cudaEvent_t a, b;
};

struct ofxcv_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;  // owned
    int num_sms = 148;
    std::string last_error;
    uint64_t launches = 0;
    // named device workspaces, grown on demand, reused between calls
    ofxcv_buf ws[72];
    // pinned host staging for the *_host entry points
    ofxcv_buf pin[15];  // 0-3 internal staging, 4-11 ofxcv_scratch_pinned, 12 Dual TV-L1 stop flags, 13 watershed round counters, 14 content keys
    // per-family kernel timing (bench.py roofline numerator): events recorded on the launching stream
    bool timing = false;
    std::vector<ofxcv_timed_launch> timed[3];
    std::vector<ofxcv_timed_launch> event_pool;
    double timed_ms[3] = {0, 0, 0};
    uint64_t timed_n[3] = {0, 0, 0};
    // per-launch profile (tools/prof_*.py): one event pair per labelled launch, aggregated by ofxcv_prof_report
    bool prof_on = false;
    struct prof_rec {
        const char* name;
        int tag;
        cudaEvent_t a, b;
    };
    std::vector<prof_rec> prof;
    ofxcv_fb_pyr fb_pyr[8];  // >= lanes + 2: the pyramids of the pairs in flight + the one being built
    uint64_t fb_tick = 0;
    uint64_t fb_pyr_built = 0, fb_pyr_hits = 0;
    int fb_lanes = 0;  // pairs in flight in ofxcv_farneback_sequence_u8: 0 = by frame size (ofxcv_farneback_set_lanes)
    // copy streams + events of the *_sequence_host entry points (created on first use)
    cudaStream_t stream_up = nullptr, stream_down = nullptr, stream_lane[OFXCV_FB_MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t lane_done[OFXCV_FB_MAX_LANES] = {nullptr, nullptr, nullptr, nullptr}, lane_start = nullptr;
    cudaEvent_t seq_ev[12] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    ofxcv_ctx* sub[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // workers of ofxcv_inpaint_sequence_u8 (owned)
    int ip_fill_blocks_per_sm = 8;  // persistent CTAs of the inpaint fill kernel per SM (ofxcv_inpaint_set_fill_blocks)
    cudaEvent_t tv_ev[16] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t tv_ctrl_off = 0;  // where the last ofxcv_tvl1_u8 put its control block inside WS_TV_ARENA
    std::vector<cudaEvent_t> xfer_ev;     // chunk events of ofxcv_download_rows
    cudaEvent_t xfer_up_done = nullptr;   // the last DMA out of the upload staging buffer
    cudaEvent_t order_ev = nullptr;       // ofxcv_stream_wait
    int (*abort_cb)(void*) = nullptr;     // polled between pyramid scales / pairs / frames (ofxcv_set_abort_callback)
    void* abort_user = nullptr;
    int64_t inpaint_stats[4] = {0, 0, 0, 0};
    int64_t watershed_stats[4] = {0, 0, 0, 0};  // [0] pops (-1 = still on the device), [1] frames, [2] rounds, [3] passes of the parallel flood
    std::vector<int> watershed_seq_frames;  // frames of the last call that ran on the one-thread flood
    int64_t watershed_par_pops = 0;
};

int ofxcv_fail(ofxcv_ctx* ctx, cudaError_t e, const char* what);
void* ofxcv_ws(ofxcv_ctx* ctx, int slot, size_t bytes);    // device workspace slot (nullptr on OOM)
void* ofxcv_pin(ofxcv_ctx* ctx, int slot, size_t bytes);   // pinned host slot
void ofxcv_time_begin(ofxcv_ctx* ctx, int family, cudaStream_t s);
void ofxcv_time_end(ofxcv_ctx* ctx, int family, cudaStream_t s);

// RAII label around one or more launches on stream s; free when profiling is off
struct ofxcv_prof_scope {
    ofxcv_ctx* ctx;
    cudaStream_t s;
    bool on;
    size_t idx = 0;  // scopes may nest
    ofxcv_prof_scope(ofxcv_ctx* c, cudaStream_t st, const char* name, int tag) : ctx(c), s(st), on(c->prof_on)
    {
        if (!on) return;
        ofxcv_ctx::prof_rec r;
        r.name = name;
        r.tag = tag;
        if (!ctx->event_pool.empty()) {
            r.a = ctx->event_pool.back().a;
            r.b = ctx->event_pool.back().b;
            ctx->event_pool.pop_back();
        } else {
            cudaEventCreate(&r.a);
            cudaEventCreate(&r.b);
        }
        cudaEventRecord(r.a, s);
        idx = ctx->prof.size();
        ctx->prof.push_back(r);
    }
    ~ofxcv_prof_scope()
    {
        if (on) cudaEventRecord(ctx->prof[idx].b, s);
    }
};

#define OFXCV_CUDA(ctx, call)                                         \
    do {                                                              \
        cudaError_t _e = (call);                                      \
        if (_e != cudaSuccess) return ofxcv_fail((ctx), _e, #call);   \
    } while (0)

#define OFXCV_LAUNCH_CHECK(ctx)                                              \
    do {                                                                     \
        (ctx)->launches++;                                                   \
        cudaError_t _e = cudaPeekAtLastError();                              \
        if (_e != cudaSuccess) return ofxcv_fail((ctx), _e, "kernel launch"); \
    } while (0)

struct ofxcv_device_guard {
    int prev = -1;
    explicit ofxcv_device_guard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~ofxcv_device_guard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

static inline bool ofxcv_is_pinned(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

static inline int ofxcv_div_up(int a, int b) { return (a + b - 1) / b; }
static inline bool ofxcv_aborted(const ofxcv_ctx* ctx) { return ctx->abort_cb && ctx->abort_cb(ctx->abort_user) != 0; }

// watershed_par.cu: exact intra-frame parallel flood of one prepared frame (see the header of that file); returns 1 when
// the flood is degenerate and the one-thread kernel of watershed.cu should run instead (label map restored)
int ofxcv_wsp_flood(ofxcv_ctx* ctx, cudaStream_t s, int32_t* m, ptrdiff_t pitch, const uint32_t* pix, int W, int H, int64_t* pops_out);
size_t ofxcv_wsp_workspace_bytes(int W, int H, ptrdiff_t pitch);

// workspace slot numbering
enum {
    WS_FB_TMP = 0,   // row-blurred samples
    WS_FB_I0,        // pyramid image 0 at the current scale
    WS_FB_I1,
    WS_FB_R0Q,       // polynomial expansion of image 0: channels 0..3 (float4 per pixel)
    WS_FB_R0S,       // channel 4
    WS_FB_R1Q,
    WS_FB_R1S,
    WS_FB_MAQ,       // matrix field ping
    WS_FB_MAS,
    WS_FB_MBQ,       // matrix field pong
    WS_FB_MBS,
    WS_FB_FLOWA,     // flow of the previous / current scale
    WS_FB_FLOWB,
    WS_STAGE_IN0,    // *_host staging on the device
    WS_STAGE_IN1,
    WS_STAGE_OUT,
    WS_INP_A,
    WS_INP_B,
    WS_INP_C,
    WS_INP_D,
    WS_WS_NEXT,
    WS_MISC0,
    WS_MISC1,
    WS_MISC2,
    WS_LUT,      // sRGB hipart table of the staging conversion
    WS_INP_E,
    WS_INP_F,
    WS_INP_G,
    WS_INP_H,
    WS_FB_SINT,  // per-band interior difference sums of the running column sum
    WS_FB_TOT,   // per-band totals / exclusive prefixes (ping-pong)
    WS_FB_CNT,   // per-strip finish tickets of the band kernel
    WS_FB1_MAQ,  // second solve lane of the Farneback sequence entry points (two pairs in flight)
    WS_FB1_MAS,
    WS_FB1_MBQ,
    WS_FB1_MBS,
    WS_FB1_FLOWA,
    WS_FB1_FLOWB,
    WS_FB1_TOT,
    WS_FB2_MAQ,  // lanes 2 and 3: same seven slots in the same order
    WS_FB2_MAS,
    WS_FB2_MBQ,
    WS_FB2_MBS,
    WS_FB2_FLOWA,
    WS_FB2_FLOWB,
    WS_FB2_TOT,
    WS_FB3_MAQ,
    WS_FB3_MAS,
    WS_FB3_MBQ,
    WS_FB3_MBS,
    WS_FB3_FLOWA,
    WS_FB3_FLOWB,
    WS_FB3_TOT,
    WS_TV_ARENA,  // Dual TV-L1: pyramids + J/A/P/U planes + control block, one allocation
    WS_WSP_ARENA, // parallel watershed: claims, records, level queues, sort buffers, one allocation
    WS_KEY,       // content-key accumulators
    WS_COUNT
};
static_assert(WS_COUNT + 8 <= 72, "workspace slots (the last 8 are ofxcv_scratch_device)");
