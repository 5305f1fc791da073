/*
 * ofxcv_abi.h — the drop-in C ABI of the B200-native filter bodies behind openfx-opencv's render actions.
 *
 * Every entry point replaces ONE OpenCV call the reference plugins make from their render action
 * (the reference file:line is cited on each).  Plain C types only: no OFX, OpenCV, CUDA-runtime or torch
 * types appear in a signature (`ofxcv_stream` is a `cudaStream_t` passed as void*; NULL = the context's
 * own stream).  Two flavours per op:
 *   ofxcv_<op>        all image pointers are DEVICE pointers, work is enqueued on `stream`, no sync;
 *   ofxcv_<op>_host   all image pointers are HOST pointers; the call stages them through pinned memory into
 *                     HBM once, runs the same kernels, copies the result back and synchronises (this is the
 *                     call the OFX glue makes when the host did not enable CUDA render).
 * Return value: 0 (OFXCV_OK) or a negative ofxcv_status.  There is NO CPU fallback anywhere in this library:
 * without a usable CUDA device `ofxcv_create` returns NULL and every op returns OFXCV_ERR_NO_DEVICE.
 *
 * Threading: a context is used by one thread at a time (it owns its workspace, streams and pinned staging);
 * create one per render thread (the OFX glue keeps a pool per device).  Different contexts are fully independent,
 * so the library is re-entrant as kOfxImageEffectRenderFullySafe needs (VectorGenerator.cpp:108).  Two families of
 * calls start threads of their own inside the library and join them before returning: ofxcv_upload_rows /
 * ofxcv_download_rows (row-copy workers) and ofxcv_inpaint_sequence_u8[_host] (one worker per frame in flight, each on
 * a sub-context of `ctx`); a host with its own render threads needs no extra care, it only must not share one
 * context between them.
 * Blocking: the device flavours only enqueue work, with two exceptions that loop on the host and return when done:
 * ofxcv_watershed_u8c3[_batch] with fewer than 16 frames (round loop of the parallel flood) and the inpaint calls
 * (the marching reads its batch counters back).
 */
#ifndef OFXCV_ABI_H
#define OFXCV_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define OFXCV_API __attribute__((visibility("default")))
#else
#define OFXCV_API
#endif

typedef struct ofxcv_ctx ofxcv_ctx;
typedef void* ofxcv_stream; /* cudaStream_t */

typedef enum ofxcv_status {
    OFXCV_ABORTED = 1,        /* the abort callback fired: the output is incomplete (maps to kOfxStatOK after a sync, like the
                                 reference's `if (abort()) return;`) */
    OFXCV_OK = 0,
    OFXCV_ERR_BAD_ARG = -1,   /* NULL pointer, non-positive size, unsupported channel count ... */
    OFXCV_ERR_NO_DEVICE = -2, /* no CUDA device / context creation failed */
    OFXCV_ERR_MEMORY = -3,    /* device or pinned allocation failed (maps to kOfxStatErrMemory) */
    OFXCV_ERR_CUDA = -4,      /* any other CUDA runtime error (maps to kOfxStatFailed) */
    OFXCV_ERR_UNSUPPORTED = -5
} ofxcv_status;

/* ---- context ---------------------------------------------------------------------------------------- */
OFXCV_API int ofxcv_abi_version(void); /* 1 */
OFXCV_API const char* ofxcv_status_string(int status);
OFXCV_API int ofxcv_device_count(void);
/* device < 0: use the calling thread's current CUDA device. */
OFXCV_API ofxcv_ctx* ofxcv_create(int device);
OFXCV_API void ofxcv_destroy(ofxcv_ctx* ctx);
OFXCV_API int ofxcv_device(const ofxcv_ctx* ctx);
OFXCV_API ofxcv_stream ofxcv_ctx_stream(ofxcv_ctx* ctx);
OFXCV_API int ofxcv_synchronize(ofxcv_ctx* ctx);
OFXCV_API const char* ofxcv_last_error(const ofxcv_ctx* ctx); /* text of the last CUDA error seen by ctx */
/* number of kernel launches issued through this context so far (bench.py's gpu_launches claim) */
OFXCV_API uint64_t ofxcv_launch_count(const ofxcv_ctx* ctx);
/* device-time (ms, CUDA events on the launching stream) of the dominant kernel family since the last reset:
 * family 0 = Farneback band kernel, full-resolution ITER launches; 1 = inpaint fill; 2 = watershed flood.
 * Returns launches counted. */
OFXCV_API uint64_t ofxcv_kernel_time_ms(ofxcv_ctx* ctx, int family, double* total_ms);
OFXCV_API void ofxcv_kernel_time_enable(ofxcv_ctx* ctx, int enable);
/* developer profile: while enabled every labelled launch is bracketed by CUDA events on its stream;
 * ofxcv_prof_report writes "name tag launches total_ms" lines (tag = pyramid scale for Farneback) into buf,
 * returns the full text length and resets the records. */
OFXCV_API void ofxcv_prof_enable(ofxcv_ctx* ctx, int enable);
OFXCV_API size_t ofxcv_prof_report(ofxcv_ctx* ctx, char* buf, size_t cap);

/* plain device-memory helpers so that a non-CUDA host language (ctypes, cgo, JNI) can stage frames */
OFXCV_API void* ofxcv_device_alloc(ofxcv_ctx* ctx, size_t bytes);
OFXCV_API void ofxcv_device_free(ofxcv_ctx* ctx, void* dptr);
OFXCV_API void* ofxcv_pinned_alloc(ofxcv_ctx* ctx, size_t bytes);
OFXCV_API void ofxcv_pinned_free(ofxcv_ctx* ctx, void* hptr);
/* grow-only scratch owned by the context (freed by ofxcv_destroy), slot 0..7: what the OFX glue stages clips in, so that
 * a render pays neither cudaMalloc nor cudaMallocHost (page-locking a 4K float RGBA frame takes tens of ms). */
OFXCV_API void* ofxcv_scratch_device(ofxcv_ctx* ctx, int slot, size_t bytes);
OFXCV_API void* ofxcv_scratch_pinned(ofxcv_ctx* ctx, int slot, size_t bytes);
OFXCV_API int ofxcv_upload(ofxcv_ctx* ctx, ofxcv_stream s, void* dst_dev, const void* src_host, size_t bytes);
OFXCV_API int ofxcv_download(ofxcv_ctx* ctx, ofxcv_stream s, void* dst_host, const void* src_dev, size_t bytes);
OFXCV_API int ofxcv_device_copy(ofxcv_ctx* ctx, ofxcv_stream s, void* dst_dev, const void* src_dev, size_t bytes);
OFXCV_API int ofxcv_memset(ofxcv_ctx* ctx, ofxcv_stream s, void* dst_dev, int value, size_t bytes);

/* the calling thread's current CUDA device (-1 on error) and the device that owns a device pointer (-1: host memory or
 * unknown) -- what the OFX glue needs to pick a context of the right GPU when a host hands it device images
 * (kOfxImageEffectPropCudaEnabled, ofxImageEffect.h:1013-1049). */
OFXCV_API int ofxcv_current_device(void);
OFXCV_API int ofxcv_pointer_device(const void* p);
/* Rows of a (pageable) host image -> tight device rows (pitch = row_bytes) and back, the staging of an OFX host-memory clip
 * (src_stride / dst_stride may be negative: ofxImageEffect.h:909-921).  Large images move in row chunks: worker threads
 * copy between the caller's rows and the context's pinned staging while the DMA of the finished chunks runs.
 * ofxcv_upload_rows returns once the caller's rows have been read (the DMA may still run on `stream`);
 * ofxcv_download_rows returns once the caller's rows are written. */
OFXCV_API int ofxcv_upload_rows(ofxcv_ctx* ctx, ofxcv_stream s, void* dst_dev, const void* src_host, ptrdiff_t src_stride,
                                size_t row_bytes, int rows);
OFXCV_API int ofxcv_download_rows(ofxcv_ctx* ctx, ofxcv_stream s, void* dst_host, ptrdiff_t dst_stride, const void* src_dev,
                                  size_t row_bytes, int rows);
/* A second stream of the context for staging work that should overlap the compute on the main stream (the OFX glue uploads
 * and converts frame t+1 on it while the backward flow of frames t, t-1 runs), and the two ordering primitives that
 * go with it: `waiter` waits for everything enqueued on `signaller` so far; synchronise one stream.  NULL = main stream. */
OFXCV_API ofxcv_stream ofxcv_aux_stream(ofxcv_ctx* ctx);
OFXCV_API int ofxcv_stream_wait(ofxcv_ctx* ctx, ofxcv_stream waiter, ofxcv_stream signaller);
OFXCV_API int ofxcv_stream_synchronize(ofxcv_ctx* ctx, ofxcv_stream s);
/* bytes moved host->device / device->host by the staging helpers of this library since it was loaded (all contexts) */
OFXCV_API void ofxcv_transfer_stats(uint64_t* h2d_bytes, uint64_t* d2h_bytes);
/* abort polling (the host's OfxImageEffectSuiteV1::abort, /root/reference/openfx/include/ofxImageEffect.h): when set, the
 * flow entry points call cb(user) between pyramid scales and between the pairs / frames of a clip and stop enqueueing
 * work once it returns non-zero; they then return OFXCV_ABORTED (> 0).  NULL clears it. */
OFXCV_API void ofxcv_set_abort_callback(ofxcv_ctx* ctx, int (*cb)(void*), void* user);

/* ---- dense optical flow ----------------------------------------------------------------------------- */
/* Replaces cv::calcOpticalFlowFarneback(prev, next, flow, pyr_scale, levels, winsize, iters, poly_n,
 * poly_sigma, flags) as called at /root/reference/VectorGenerator/VectorGenerator.cpp:403
 * (argument values :391-399; parameter defaults :804,:814,:824,:834).                                     */
typedef struct ofxcv_fb_params {
    double pyr_scale;  /* 0.5  (VectorGenerator.cpp:391) */
    int levels;        /* 3    (:804)  -- runs levels+1 scales, like OpenCV */
    int winsize;       /* 3    (:395)  -- only 3 is implemented (the value the plugin hard-wires) */
    int iterations;    /* 15   (:814) */
    int poly_n;        /* 5    (:824)  -- 1..16 */
    double poly_sigma; /* 1.1  (:834) */
    int flags;         /* 0    (:403)  -- OPTFLOW_USE_INITIAL_FLOW / FARNEBACK_GAUSSIAN unsupported */
} ofxcv_fb_params;

OFXCV_API void ofxcv_fb_default_params(ofxcv_fb_params* p);
/* number of scales that will run (effective levels + 1) and the algorithmic HBM bytes of one pair
 * (SURVEY.md section 8d byte model: sum_k n_k*(66+88*I) + 2*W*H per scale). */
OFXCV_API int ofxcv_farneback_scales(int W, int H, const ofxcv_fb_params* p);
OFXCV_API double ofxcv_farneback_algorithmic_bytes(int W, int H, const ofxcv_fb_params* p);
/* algorithmic bytes of ONE full-resolution iteration launch of the band kernel (88 B per pixel: SURVEY.md 8d) --
 * the numerator of bench.py's roofline for the dominant kernel; ofxcv_kernel_time_ms(ctx, 0, ..) times exactly
 * those launches (iterations-1 per pair). */
OFXCV_API double ofxcv_farneback_iter_bytes(int W, int H, const ofxcv_fb_params* p);
OFXCV_API size_t ofxcv_farneback_workspace_bytes(int W, int H, const ofxcv_fb_params* p);

/* prev/next: 8-bit gray, `stride` bytes per row; flow: interleaved (dx,dy) float32, `flow_stride` BYTES/row. */
OFXCV_API int ofxcv_farneback_u8(ofxcv_ctx* ctx, ofxcv_stream stream, const uint8_t* prev, const uint8_t* next,
                                 ptrdiff_t stride, int W, int H, float* flow, ptrdiff_t flow_stride,
                                 const ofxcv_fb_params* params);
OFXCV_API int ofxcv_farneback_u8_host(ofxcv_ctx* ctx, const uint8_t* prev, const uint8_t* next, ptrdiff_t stride,
                                      int W, int H, float* flow, ptrdiff_t flow_stride,
                                      const ofxcv_fb_params* params);

/* The same call with frame keys: a non-zero key names the CONTENT of a frame (the caller guarantees that equal keys
 * mean equal pixels, size and parameters); the context keeps the polynomial-expansion pyramids of the last few
 * keyed frames, so that frame t+1 of pair t is not blurred / expanded again as frame t of pair t+1 (or for the
 * backward flow of the same render: VectorGenerator.cpp:559-638 runs t->t+1 and t->t-1).  Calls that share keys
 * may be issued on different streams (a hit waits for the event of the build).  The context keeps eight pyramids (least
 * recently used replaced).  key 0 = anonymous (what ofxcv_farneback_u8 passes).  Returns OFXCV_ABORTED (1) when the
 * abort callback fired between two pyramid scales. */
OFXCV_API int ofxcv_farneback_u8_keyed(ofxcv_ctx* ctx, ofxcv_stream stream, const uint8_t* prev, const uint8_t* next,
                                       ptrdiff_t stride, int W, int H, float* flow, ptrdiff_t flow_stride,
                                       const ofxcv_fb_params* params, uint64_t key_prev, uint64_t key_next);
/* a content key for the call above: position-sensitive hash of a device-resident 8-bit plane (two independent 64-bit
 * lanes folded together with the geometry; never 0).
 * Synchronises `stream`.  The VectorGenerator glue keys every staged gray frame with it, so that consecutive
 * renders of a clip (and the forward / backward flow of one render) share frame pyramids. */
OFXCV_API int ofxcv_content_key_u8(ofxcv_ctx* ctx, ofxcv_stream stream, const uint8_t* img, ptrdiff_t stride, int W,
                                   int H, uint64_t* key);
/* pairs in flight inside the clip entry points: pair t is solved on lane t % lanes, each lane with its own stream and
 * workspace, so that the latency-bound coarse scales of one pair overlap the bandwidth-bound fine scales of another.
 * 0 (default) = by frame size: 2 lanes from ~4 Mpx up (two 4K solves fill the GPU), 4 below; 1..4 fixes it; 1 =
 * strictly one kernel at a time, what bench.py uses to time the dominant kernel on its own. */
OFXCV_API void ofxcv_farneback_set_lanes(ofxcv_ctx* ctx, int lanes);
OFXCV_API void ofxcv_farneback_cache_clear(ofxcv_ctx* ctx);
/* pyramids built / cache hits since the context was created */
OFXCV_API int ofxcv_farneback_cache_stats(const ofxcv_ctx* ctx, uint64_t* built, uint64_t* hits);
/* A clip: nframes gray frames (frame t at frames + t*frame_stride) -> nframes-1 forward flow fields t -> t+1
 * (flow t at flows + t*flow_frame_stride bytes).  Every frame's pyramid is built exactly once per call.  The host
 * flavour takes arrays of host pointers (page-locked for real overlap) and pipelines upload / compute / download
 * on three streams; this is the call of a host that renders a sequence (BASELINE.json configs 4 and 5). */
OFXCV_API int ofxcv_farneback_sequence_u8(ofxcv_ctx* ctx, ofxcv_stream stream, const uint8_t* frames, ptrdiff_t stride,
                                          size_t frame_stride, int W, int H, int nframes, float* flows,
                                          ptrdiff_t flow_stride, size_t flow_frame_stride,
                                          const ofxcv_fb_params* params);
OFXCV_API int ofxcv_farneback_sequence_u8_host(ofxcv_ctx* ctx, const uint8_t* const* frames, ptrdiff_t stride, int W,
                                               int H, int nframes, float* const* flows, ptrdiff_t flow_stride,
                                               const ofxcv_fb_params* params);

/* ---- Dual TV-L1 optical flow (the plugin's second method) --------------------------------------------------- */
/* Replaces createOptFlow_DualTVL1() + set{Tau,Lambda,Theta,ScalesNumber,WarpingsNumber,Epsilon,InnerIterations} +
 * calc(prev, next, flow) as called at /root/reference/VectorGenerator/VectorGenerator.cpp:436-492 (parameter
 * defaults :874-929, `iterations` :814).  The last three fields are the values OpenCV 3/4 fixes and the plugin never
 * sets.  PARITY UNPINNED for the method as a whole (no OpenCV build with DualTVL1 is available to pin it): results
 * are bit-identical to the CPU test oracle (tvl1.c), whose primitives (bicubic remap, 5x5 median, bilinear resize) are pinned to cv2. */
typedef struct ofxcv_tvl1_params {
    double tau;           /* 0.25 (VectorGenerator.cpp:877) */
    double lambda;        /* 0.15 (:887) */
    double theta;         /* 0.3  (:897) */
    double epsilon;       /* 0.01 (:926) -- a warping stops once the squared flow update <= epsilon^2 * W * H */
    int nscales;          /* 5    (:907) -- scales smaller than 16 px are dropped */
    int warps;            /* 5    (:917) */
    int iterations;       /* 15   (:814) -- inner iterations (setInnerIterations / setIterations, :481-485) */
    int outer_iterations; /* 10   OpenCV default; one 5x5 median of the flow per outer iteration */
    double scale_step;    /* 0.8  OpenCV default */
    int median_filtering; /* 5    OpenCV default; <= 1 switches the median off, other sizes are rejected */
} ofxcv_tvl1_params;

OFXCV_API void ofxcv_tvl1_default_params(ofxcv_tvl1_params* p);
OFXCV_API int ofxcv_tvl1_scales(int W, int H, const ofxcv_tvl1_params* p);
OFXCV_API size_t ofxcv_tvl1_workspace_bytes(int W, int H, const ofxcv_tvl1_params* p);
/* algorithmic bytes of ONE full-resolution inner iteration (one launch): 64 B per pixel
 * (in: warped-gradient plane 16 + flow 8 + dual variable 16; out: flow 8 + dual variable 16). */
OFXCV_API double ofxcv_tvl1_iter_bytes(int W, int H);
/* prev/next: 8-bit gray, `stride` bytes per row; flow: interleaved (dx,dy) float32, `flow_stride` BYTES/row, rows
 * 8-byte aligned.  The convergence test stays on the device; the host only paces its launches by it (it waits for the
 * flag of outer iteration n-2 before enqueuing n), so the call returns once the last outer iteration is enqueued --
 * the tail of the work is still asynchronous on `stream`, like ofxcv_farneback_u8. */
OFXCV_API int ofxcv_tvl1_u8(ofxcv_ctx* ctx, ofxcv_stream stream, const uint8_t* prev, const uint8_t* next, ptrdiff_t stride,
                            int W, int H, float* flow, ptrdiff_t flow_stride, const ofxcv_tvl1_params* params);
/* inner iterations the last ofxcv_tvl1_u8 of this context actually ran (synchronises the device); -1 if none */
OFXCV_API int64_t ofxcv_tvl1_iterations_run(ofxcv_ctx* ctx);

/* ---- inpainting ------------------------------------------------------------------------------------- */
/* Replaces cvInpaint(image0, mask, image1, radius, CV_INPAINT_TELEA) at
 * /root/reference/opencv2fx/inpaint/inpaint.cpp:311-318 (Telea hard-wired at :311; Navier-Stokes is the
 * method BASELINE.json config 4 names).  Method values are OpenCV's: cv::INPAINT_NS=0, cv::INPAINT_TELEA=1. */
#define OFXCV_INPAINT_NS 0
#define OFXCV_INPAINT_TELEA 1
/* img/out: `channels` (1 or 3) interleaved u8; mask: u8, non-zero = pixel to inpaint. strides in bytes. */
OFXCV_API int ofxcv_inpaint_u8(ofxcv_ctx* ctx, ofxcv_stream stream, const uint8_t* img, ptrdiff_t img_stride,
                               int channels, const uint8_t* mask, ptrdiff_t mask_stride, uint8_t* out,
                               ptrdiff_t out_stride, int W, int H, double radius, int method);
OFXCV_API int ofxcv_inpaint_u8_host(ofxcv_ctx* ctx, const uint8_t* img, ptrdiff_t img_stride, int channels,
                                    const uint8_t* mask, ptrdiff_t mask_stride, uint8_t* out, ptrdiff_t out_stride,
                                    int W, int H, double radius, int method);
OFXCV_API size_t ofxcv_inpaint_workspace_bytes(int W, int H, int channels);
/* The fill stage is a dataflow over a dependency chain thousands of pixels deep: one frame leaves most of the GPU idle.
 * A sequence renderer runs several frames at once (one context + host thread each) and gives every context a share
 * of the SMs: persistent fill CTAs per SM for this context, 1..8 (default 8 = the whole GPU for one frame). */
OFXCV_API void ofxcv_inpaint_set_fill_blocks(ofxcv_ctx* ctx, int blocks_per_sm);
/* A clip of independent frames (BASELINE.json config 4 is a 300-frame sequence), `frames_in_flight` of them at a time
 * (1..8, <= 0 = 8): the library runs one worker (own stream, workspaces and host thread) per frame in flight and splits
 * the persistent fill CTAs between them -- what a sequence renderer built on ofxcv_inpaint_set_fill_blocks would do
 * by hand.  imgs / masks / outs: arrays of `nframes` frame pointers with common strides (device pointers; host pointers
 * for _host, whose uploads and downloads overlap the other frames' compute).  Blocking: waits for `stream`, returns
 * when every frame is done.  Results are those of ofxcv_inpaint_u8 frame by frame. */
OFXCV_API int ofxcv_inpaint_sequence_u8(ofxcv_ctx* ctx, ofxcv_stream stream, const uint8_t* const* imgs,
                                        ptrdiff_t img_stride, int channels, const uint8_t* const* masks,
                                        ptrdiff_t mask_stride, uint8_t* const* outs, ptrdiff_t out_stride, int W, int H,
                                        int nframes, double radius, int method, int frames_in_flight);
OFXCV_API int ofxcv_inpaint_sequence_u8_host(ofxcv_ctx* ctx, const uint8_t* const* imgs, ptrdiff_t img_stride,
                                             int channels, const uint8_t* const* masks, ptrdiff_t mask_stride,
                                             uint8_t* const* outs, ptrdiff_t out_stride, int W, int H, int nframes,
                                             double radius, int method, int frames_in_flight);
/* statistics of the last inpaint call on ctx: [0]=hole pixels, [1]=marching batches (0.7-wide T windows popped
 * in parallel), [2]=T relaxation rounds over all batches, [3]=kernel launches of the call */
OFXCV_API int ofxcv_inpaint_last_stats(const ofxcv_ctx* ctx, int64_t stats[4]);
/* test hook: after an inpaint call of size WxH on ctx, copy out the marched T map ((H+2)x(W+2) f32, padded like
 * OpenCV's) and the fill order (HxW int32, -1 where nothing was filled).  Either pointer may be NULL. */
OFXCV_API int ofxcv_inpaint_debug_maps(ofxcv_ctx* ctx, int W, int H, float* t_host, int32_t* order_host);

/* ---- segmentation ----------------------------------------------------------------------------------- */
/* Replaces the segment plugin's body (cvPyrSegmentation at /root/reference/opencv2fx/segment/segment.cpp:296-302)
 * by cv::watershed(rgb8, int32 markers) as BASELINE.json config 3 requires.
 * rgb: 3-channel interleaved u8; markers: int32 in/out (>0 seeds; on return labels >0, -1 ridges + border).
 * The label map is the one of OpenCV's sequential Meyer flood, bit for bit, by two exact schedules: fewer than 16 frames
 * are flooded one after the other by the intra-frame PARALLEL flood (rounds of speculative blocks ordered by rank,
 * csrc/watershed_par.cu: a 4K frame in ~0.3 s, blocking call), 16 or more by one thread per frame with all frames in
 * flight (asynchronous).  OFXCV_WS_MODE=par|seq forces one of them. */
OFXCV_API int ofxcv_watershed_u8c3(ofxcv_ctx* ctx, ofxcv_stream stream, const uint8_t* rgb, ptrdiff_t rgb_stride,
                                   int32_t* markers, ptrdiff_t markers_stride, int W, int H);
OFXCV_API int ofxcv_watershed_u8c3_host(ofxcv_ctx* ctx, const uint8_t* rgb, ptrdiff_t rgb_stride, int32_t* markers,
                                        ptrdiff_t markers_stride, int W, int H);
/* `nframes` independent frames flooded concurrently (frame f at base + f*frame_stride bytes): the exact flood
 * is latency-bound per frame, so sequence throughput comes from frames in flight. */
OFXCV_API int ofxcv_watershed_u8c3_batch(ofxcv_ctx* ctx, ofxcv_stream stream, const uint8_t* rgb,
                                         ptrdiff_t rgb_stride, size_t rgb_frame_stride, int32_t* markers,
                                         ptrdiff_t markers_stride, size_t markers_frame_stride, int W, int H,
                                         int nframes);
OFXCV_API size_t ofxcv_watershed_workspace_bytes(int W, int H, int nframes);
/* stats of the last call: [0]=queue pops (all frames), [1]=frames, [2]=rounds and [3]=passes of the parallel flood
 * (last frame; 0 when the one-thread flood ran) */
OFXCV_API int ofxcv_watershed_last_stats(const ofxcv_ctx* ctx, int64_t stats[4]);

/* ---- staging conversions (the steps either side of the OpenCV call inside the render actions) ------- */
/* float RGBA/RGB/Alpha (linear) -> Rec.709 luma -> sRGB 8-bit gray, the semantics of
 * GenericOpenCVPlugin::fetchCVImage8UGrayscale (/root/reference/OpenCV/GenericOpenCVPlugin.cpp:223-265) with
 * Lut::to_byte_grayscale_nodither (/root/reference/SupportExt/ofxsLut.h:447-486) over the WHOLE row
 * (SURVEY.md Appendix B1).  ncomp = 4, 3 or 1; src_stride/dst_stride in bytes (src_stride may be negative). */
OFXCV_API int ofxcv_rgba32f_to_srgb_gray8(ofxcv_ctx* ctx, ofxcv_stream stream, const float* src,
                                          ptrdiff_t src_stride, int ncomp, uint8_t* dst, ptrdiff_t dst_stride,
                                          int W, int H);
/* float (linear) -> 8-bit sRGB, packed: GenericOpenCVPlugin::fetchCVImage8U
 * (/root/reference/OpenCV/GenericOpenCVPlugin.cpp:177-221) = Lut::to_byte_packed_nodither
 * (/root/reference/SupportExt/ofxsLut.h:389-444) over WHOLE rows.  src/dst component counts
 * 1 (alpha), 3 or 4; colour through the sRGB table, alpha through floatToInt<256>; a 3-component source gives alpha 0.
 * Strides in bytes (src_stride may be negative). */
OFXCV_API int ofxcv_rgba32f_to_srgb8_packed(ofxcv_ctx* ctx, ofxcv_stream stream, const float* src, ptrdiff_t src_stride,
                                            int src_ncomp, uint8_t* dst, ptrdiff_t dst_stride, int dst_ncomp, int W,
                                            int H);
/* 8-bit sRGB -> float (linear), packed: GenericOpenCVPlugin::cvImageToOfxImage (GenericOpenCVPlugin.cpp:267-325) =
 * Lut::from_byte_packed (ofxsLut.h:536-581).  Same component count on both sides (1, 3 or 4). */
OFXCV_API int ofxcv_srgb8_packed_to_rgba32f(ofxcv_ctx* ctx, ofxcv_stream stream, const uint8_t* src,
                                            ptrdiff_t src_stride, float* dst, ptrdiff_t dst_stride, int ncomp, int W,
                                            int H);
/* flow -> selected RGBA channels with the renderScale division: VectorGenerator.cpp:494-519.
 * chan_sel[c] for c = R,G,B,A: -1 = leave untouched, 0 = flow.x / scale_x, 1 = flow.y / scale_y. */
OFXCV_API int ofxcv_flow_to_rgba32f(ofxcv_ctx* ctx, ofxcv_stream stream, const float* flow, ptrdiff_t flow_stride,
                                    float* dst, ptrdiff_t dst_stride, int W, int H, const int chan_sel[4],
                                    double scale_x, double scale_y);
/* RGBA8 -> RGB8 + hole mask: cvCvtColor x3 + cvThreshold(BINARY_INV, 0) + cvDilate(3x3, iterations) of
 * /root/reference/opencv2fx/inpaint/inpaint.cpp:303-309.  mask = 255 where gray(rgb)==0, dilated. */
OFXCV_API int ofxcv_rgba8_to_rgb8_mask(ofxcv_ctx* ctx, ofxcv_stream stream, const uint8_t* rgba,
                                       ptrdiff_t rgba_stride, uint8_t* rgb, ptrdiff_t rgb_stride, uint8_t* mask,
                                       ptrdiff_t mask_stride, int W, int H, int dilate_iterations);
/* RGB8 -> RGBA8 with alpha 255: the write-back loop inpaint.cpp:320-358 / segment.cpp:307-323. */
OFXCV_API int ofxcv_rgb8_to_rgba8(ofxcv_ctx* ctx, ofxcv_stream stream, const uint8_t* rgb, ptrdiff_t rgb_stride,
                                  uint8_t* rgba, ptrdiff_t rgba_stride, int W, int H);
/* the same write-back with the plugin's optional noise on hole pixels (inpaint.cpp:320-347, `inpaintnoise`):
 * noise_div = (int)(1/inpaintnoise), 0 = no noise.  Counter-based RNG (documented, outside the parity contract). */
OFXCV_API int ofxcv_rgb8_to_rgba8_noise(ofxcv_ctx* ctx, ofxcv_stream stream, const uint8_t* rgb, ptrdiff_t rgb_stride,
                                        const uint8_t* mask, ptrdiff_t mask_stride, uint8_t* rgba, ptrdiff_t rgba_stride,
                                        int W, int H, int noise_div, unsigned seed);
/* deterministic seed grid for the segment plugin (`seeds` param): n = gx*gy squares of
 * (2*half+1)^2 pixels labelled 1..n on a regular grid, 0 elsewhere. */
OFXCV_API int ofxcv_seed_grid(ofxcv_ctx* ctx, ofxcv_stream stream, int32_t* markers, ptrdiff_t markers_stride,
                              int W, int H, int gx, int gy, int half);
/* labels -> RGBA8 visualisation: mean colour of each segment (label>0), ridges black; the segment plugin's
 * output image (cvPyrSegmentation also paints each segment with its mean colour). nlabels = max label. */
OFXCV_API int ofxcv_labels_to_rgba8(ofxcv_ctx* ctx, ofxcv_stream stream, const uint8_t* rgb, ptrdiff_t rgb_stride,
                                    const int32_t* labels, ptrdiff_t labels_stride, uint8_t* rgba,
                                    ptrdiff_t rgba_stride, int W, int H, int nlabels);

#ifdef __cplusplus
}
#endif
#endif /* OFXCV_ABI_H */
