// Exact intra-frame parallel marker watershed for sm_100a: the flood of ONE frame, bit-identical to the sequential
// Meyer flood of cv::watershed (BASELINE.json config 3; the segment plugin renders one frame per action,
// /root/reference/opencv2fx/segment/segment.cpp:197-338), spread over the whole GPU.
//
// Why the ordered flood can be parallelised exactly.  cv::watershed always pops the head of the lowest non-empty FIFO
// level (256 levels, first-parent level assignment, label decided at pop time), i.e. it runs
//     Drain(b) = for c = 0..b:  while Q[c] not empty { pop head of Q[c]; Drain(c-1) }
// so the pop sequence is a sequence of BLOCKS: one entry e popped from Q[c] while every lower level is empty, followed
// by the complete sub-flood its pushes start at levels < c.  Blocks are ordered by (phase c, FIFO rank in Q[c]); the
// entries a phase-c block appends to Q[c] form the next GENERATION of that phase, entries at levels > c wait for their
// phase in push order.  On natural frames a 4K flood is a few hundred (phase, generation) ROUNDS of up to ~10^6 blocks;
// 95 % of the blocks are the single pop, the sub-floods are small connected regions.
//
// A round runs all its blocks concurrently, one thread each, as ordered transactions on ONE 64-bit word per pixel:
//     committed pixel:  0xFFFFFFFF : label-map value            (> 0 label, 0 unvisited, -1 ridge, -2 queued)
//     claimed pixel:    rank << 1 | was-queued-entry : state    (state = label / -1 once popped, -2 queued for a later
//                                                                round, <= -3 queued inside the block's sub-flood; that
//                                                                code carries the link to the next pixel of its level)
//   * claims are made with atomicMin / atomicCAS, so the lower rank (the earlier block of the sequential order) always
//     wins; a block reads a neighbour's claim only if the claim's rank is <= its own, else it sees the committed state
//     the claim replaced (it must not see the future);
//   * each pop records its OUTCOME (label, pixels pushed) and the neighbour states it took from its own block's earlier
//     claims.  After the run a validation pass re-derives the outcome of every pop from the final claims of the lower
//     ranks: a block with a pop whose outcome would now differ, or that lost a pixel of its sub-flood to a lower rank, is
//     dirty -> its claims are retracted and it runs again (a queued-only pixel lost to a lower rank is simply struck
//     from the record: nothing else of the block depended on it).  Rank 0 is right after the first pass, and by
//     induction on the rank the iteration converges to exactly the sequential result (2-3 passes in practice);
//   * commit: claims become committed words, the pushes at levels >= c are sorted by (level, rank, pop number inside the
//     block, direction) = their sequential push order and appended to the level queues (plain arrays, every pixel is
//     queued once).
// The result does not depend on thread scheduling: the fixed point is unique.  A frame whose flood degenerates into
// very many tiny rounds (adversarial serpentine images) is handed back to the one-thread flood (watershed.cu); the label
// map is only written at the very end, so nothing has to be undone for that.
#include <stdlib.h>

#include <algorithm>
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

constexpr int WS_IN_QUEUE = -2;
constexpr int WS_WSHED = -1;
constexpr int WSP_PRIVATE = -3;         // state <= WSP_PRIVATE: queued inside the owner's sub-flood, next pixel = WSP_PRIVATE - state
constexpr int WSP_NOLINK = 0x7ffffff0;  // "no next pixel yet"
constexpr unsigned WSP_COMMITTED = 0xffffffffu;
constexpr unsigned long long WSP_EMPTY_WORD = ~0ull;  // tag COMMITTED: "no claim there"
constexpr int WSP_CHUNK0 = 4, WSP_CHUNK = 16;  // sub-flood records are allocated in chunks: the first one small, then WSP_CHUNK

typedef unsigned long long WspPx;  // the claim word of a pixel (the packed colours stay in their own plane: read-only, L1-cacheable)

struct __align__(16) WspRec {
    int pixel;  // -1 = unused slot / retracted record
    int rank;
    int popseq;
    int label;
    int view[4];  // neighbour states at pop time; only the directions of the self mask are used by the validation
    unsigned char lvl[4];
    unsigned pushmask;  // bits 0-3: directions pushed; bits 4-7: neighbour state came from a claim of this block (self mask)
    int pad[2];
};
static_assert(sizeof(WspRec) == 48, "record layout");

struct WspCtl {  // device control block, copied to the host once per pass
    unsigned long long nrec;  // record pool top
    unsigned ndirty;
    unsigned nout;   // live pushes at levels >= c (counted by the validation pass)
    unsigned nlive;  // live records = pops of the round once no block is dirty
    unsigned overflow;
    unsigned ncand;
    unsigned pad[3];
};

struct WspArgs {
    WspPx* px;
    const uint32_t* pix;  // packed RGBX
    WspRec* rec;
    const int* ent;
    int* dirty;  // set by whoever invalidates a block (a steal during the run, the validation)
    int* runf;   // the blocks this pass re-runs: dirty latched at the start of the pass
    WspCtl* ctl;
    unsigned long long rec_cap;
    int ms;  // pitch in pixels
    int N;   // entries (blocks) of this round
    int c;   // phase
};

__device__ __forceinline__ int wsp_diff(uint32_t a, uint32_t b)
{
    const uint32_t d = __vabsdiffu4(a, b);
    return max((int)(d & 0xff), max((int)((d >> 8) & 0xff), (int)((d >> 16) & 0xff)));
}

__device__ __forceinline__ unsigned long long wsp_word(unsigned tag, int state) { return ((unsigned long long)tag << 32) | (unsigned)state; }
__device__ __forceinline__ unsigned wsp_tag(unsigned long long w) { return (unsigned)(w >> 32); }
__device__ __forceinline__ int wsp_state(unsigned long long w) { return (int)(unsigned)w; }
// the committed state a claim replaced: a queued entry of this round, or an unvisited pixel
__device__ __forceinline__ int wsp_under(unsigned tag) { return (tag & 1u) ? WS_IN_QUEUE : 0; }
// what a block of rank q sees of a pixel word, all claims up to rank `upto` included
__device__ __forceinline__ int wsp_visible(unsigned long long w, unsigned upto_excl)
{
    const unsigned tag = wsp_tag(w);
    if (tag == WSP_COMMITTED || (tag >> 1) < upto_excl) return wsp_state(w);
    return wsp_under(tag);
}

__device__ __forceinline__ void wsp_flag_dirty(const WspArgs& a, unsigned r)
{
    if (atomicExch(a.dirty + r, 1) == 0) atomicAdd(&a.ctl->ndirty, 1u);
}

__device__ __forceinline__ unsigned long long wsp_ld_own(const WspPx* p)
{
    return __ldcg(p);
}

// ---- label map + packed colours -> pixel words ----------------------------------------------------------------------
__global__ void __launch_bounds__(256) wsp_pack(const int32_t* __restrict__ m, WspPx* __restrict__ px, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) px[i] = ((unsigned long long)WSP_COMMITTED << 32) | (unsigned)m[i];
}

__global__ void __launch_bounds__(256) wsp_unpack(const WspPx* __restrict__ px, int32_t* __restrict__ m, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) m[i] = (int)(unsigned)px[i];
}

// ---- candidates: unlabelled 4-neighbours of a seed, keyed (level, row-major position) = OpenCV's initial push order --
__global__ void __launch_bounds__(256) wsp_candidates(const int32_t* __restrict__ m, const uint32_t* __restrict__ pix, int ms, int w, int h,
                                                      unsigned long long* __restrict__ keys, int* __restrict__ vals, WspCtl* ctl)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x < 1 || y < 1 || x >= w - 1 || y >= h - 1) return;
    const int o = y * ms + x;
    if (m[o] != 0) return;
    int idx = 256;
    const uint32_t c = pix[o];
    if (m[o - 1] > 0) idx = wsp_diff(c, pix[o - 1]);
    if (m[o + 1] > 0) idx = min(idx, wsp_diff(c, pix[o + 1]));
    if (m[o - ms] > 0) idx = min(idx, wsp_diff(c, pix[o - ms]));
    if (m[o + ms] > 0) idx = min(idx, wsp_diff(c, pix[o + ms]));
    if (idx == 256) return;
    const unsigned k = atomicAdd(&ctl->ncand, 1u);
    keys[k] = ((unsigned long long)idx << 56) | (unsigned)o;
    vals[k] = o;
}

__global__ void __launch_bounds__(256) wsp_mark_queued(WspPx* __restrict__ px, const int* __restrict__ vals, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) px[vals[i]] = ((unsigned long long)WSP_COMMITTED << 32) | (unsigned)WS_IN_QUEUE;
}

// start[l] = first sorted key whose level (top byte) is >= l, l = 0..256
__global__ void __launch_bounds__(288) wsp_bounds(const unsigned long long* __restrict__ keys, int n, int* __restrict__ start)
{
    const int l = threadIdx.x;
    if (l > 256) return;
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((int)(keys[mid] >> 56) < l) lo = mid + 1;
        else hi = mid;
    }
    start[l] = lo;
}

struct WspSeg {
    int src, cnt, dst;
};
__global__ void __launch_bounds__(256) wsp_gather(const int* __restrict__ Q, const WspSeg* __restrict__ segs, int* __restrict__ dst)
{
    const WspSeg s = segs[blockIdx.y];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < s.cnt; i += gridDim.x * blockDim.x) dst[s.dst + i] = Q[s.src + i];
}

// start of a round: every block runs, the record pool restarts above the N entry records
__global__ void __launch_bounds__(256) wsp_round_init(int* __restrict__ dirty, int* __restrict__ runf, int n, WspCtl* ctl)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        dirty[i] = 0;
        runf[i] = 1;
    }
    if (i == 0) {
        ctl->nrec = (unsigned long long)n;
        ctl->ndirty = ctl->nout = ctl->nlive = ctl->overflow = 0;
    }
}

// start of a later pass: the blocks flagged so far are the ones whose records are retracted and that run again; flags
// raised from now on (steals during the run, the validation) belong to the next pass
__global__ void __launch_bounds__(256) wsp_latch(int* __restrict__ dirty, int* __restrict__ runf, int n, WspCtl* ctl)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        runf[i] = dirty[i];
        dirty[i] = 0;
    }
    if (i == 0) ctl->ndirty = ctl->nout = ctl->nlive = 0;
}

// ---- one pass over the blocks of a round that (re-)run: one thread = one block (entry pop + its sub-flood below c) ------
template <int NL>
__global__ void __launch_bounds__(128) wsp_run(WspArgs a)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.N) return;
    if (a.runf[q] == 0) return;
    volatile int* dflag = a.dirty + q;  // raised by a lower rank that takes a pixel of our sub-flood
    const unsigned uq = (unsigned)q;
    const unsigned mytag = uq << 1;  // tag of the pixels we queue; the entry itself carries mytag | 1
    const int e = a.ent[q];
    const int ms = a.ms;
    const int c = a.c;
    WspPx* const px = a.px;
    // private queue of the sub-flood: one FIFO per level < c; heads / tails are pixels, the link to the next pixel of a level
    // is carried by the claim word of the queued pixel itself
    int head[NL], tail[NL];
    unsigned mask[NL / 32];
#pragma unroll
    for (int i = 0; i < NL / 32; i++) mask[i] = 0;
    unsigned long long chunk = 0;  // next free sub-flood record; chunk_left of them remain reserved
    int chunk_left = 0;
    int popseq = 0;
    int x = e;
    int lx = -1;  // level x was dequeued from, when its successor still has to become the head (-1: nothing pending)
    const unsigned long long link_exp = wsp_word(mytag, WSP_PRIVATE - WSP_NOLINK);  // an unlinked tail of ours
    // The sub-flood is software-pipelined: the claims of a pop (atomics) are issued, the queue is updated as if all of them
    // succeeded, and the answers are only looked at after the NEXT pop's loads have been issued (`settle`); a claim that did
    // not go our way voids the run (it is rare: somebody else took the pixel between our load and our claim).
    struct Pending {
        unsigned long long ri;
        unsigned long long cas_exp, cas_old;  // claim of the popped pixel's label
        unsigned long long old[4];            // answers of the push claims
        unsigned long long link_old[4];       // answers of the link appends
        int x, seq, lab;
        int vis[4];
        unsigned lv, att, self;  // levels, directions attempted, self mask
        bool live;
    } p;
    p.live = false;
    bool void_run = false;
    auto settle = [&]() {
        if (!p.live) return;
        p.live = false;
        unsigned pm = 0;
        bool bad = p.cas_old != p.cas_exp;  // a lower rank took the popped pixel between our load and the claim of its label
#pragma unroll
        for (int d = 0; d < 4; d++) {
            if (!(p.att & (1u << d))) continue;
            const unsigned otag = wsp_tag(p.old[d]);
            if (otag != WSP_COMMITTED) {
                if ((otag >> 1) < uq) {  // a lower rank queued it in the meantime: not ours
                    if ((int)((p.lv >> (8 * d)) & 0xff) < c) bad = true;  // ... but it went into our queue
                    continue;
                }
                // stolen from a later block: if the pixel was part of its sub-flood that run is void (tell it now, it may
                // still be flooding); a pixel it had only queued is struck from its record by the validation
                if (wsp_state(p.old[d]) != WS_IN_QUEUE) wsp_flag_dirty(a, otag >> 1);
            }
            pm |= 1u << d;
            if (p.link_old[d] != link_exp) bad = true;  // the tail we linked it behind had been taken by a lower rank
        }
        WspRec* R = a.rec + p.ri;
        // streaming stores: the record pool must not push the claim words out of L2
        __stcs(reinterpret_cast<int4*>(&R->pixel), make_int4(p.x, q, p.seq, p.lab));
        __stcs(reinterpret_cast<int4*>(R->view), make_int4(p.vis[0], p.vis[1], p.vis[2], p.vis[3]));
        __stcs(reinterpret_cast<uint2*>(R->lvl), make_uint2(p.lv, pm | p.self));
        if (bad) void_run = true;
    };
    unsigned long long ri = (unsigned long long)q;  // record of the pop being processed (the entry's record is slot q)
    for (;;) {
        // ---- the pop of x: one round of independent loads ------------------------------------------------------
        if (x != e && chunk_left == 0) {
            const int n = popseq == 1 ? WSP_CHUNK0 : WSP_CHUNK;
            chunk = atomicAdd(&a.ctl->nrec, (unsigned long long)n);
            chunk_left = n;
        }
        const int off[4] = {-1, 1, -ms, ms};
        const unsigned long long wx = __ldcg(px + x);
        unsigned long long wn[4];
        uint32_t cn[4];
#pragma unroll
        for (int d = 0; d < 4; d++) wn[d] = __ldcg(px + x + off[d]);
        const uint32_t cx = __ldg(a.pix + x);
#pragma unroll
        for (int d = 0; d < 4; d++) cn[d] = __ldg(a.pix + x + off[d]);
        const int stolen = x != e ? *dflag : 0;
        settle();  // the previous pop, while these loads are in flight
        if (x != e) {
            // x must still be ours, queued in our sub-flood; its word carries the link to its successor
            const bool ours = wsp_tag(wx) == mytag && wsp_state(wx) <= WSP_PRIVATE;
            if (chunk + chunk_left > a.rec_cap) {
                a.ctl->overflow = 1;
                chunk_left = 0;
                void_run = true;
            }
            if (stolen || !ours) void_run = true;
        }
        if (void_run) break;
        if (x != e) {
            ri = chunk++;
            chunk_left--;
            if (lx >= 0) head[lx] = WSP_PRIVATE - wsp_state(wx);
        }
        int vis[4];
        unsigned self = 0;
#pragma unroll
        for (int d = 0; d < 4; d++) {
            vis[d] = wsp_visible(wn[d], uq + 1);
            if (wsp_tag(wn[d]) != WSP_COMMITTED && (wsp_tag(wn[d]) >> 1) == uq) self |= 16u << d;
        }
        int lab = 0;
#pragma unroll
        for (int d = 0; d < 4; d++) {
            const int t = vis[d];
            if (t > 0) {
                if (lab == 0) lab = t;
                else if (t != lab) lab = WS_WSHED;
            }
        }
        p.live = true;
        p.ri = ri;
        p.x = x;
        p.seq = popseq++;
        p.lab = lab;
        p.self = self;
        p.att = 0;
        p.lv = 0;
#pragma unroll
        for (int d = 0; d < 4; d++) p.vis[d] = vis[d];
        p.cas_exp = p.cas_old = wx;
        if (x == e) px[x] = wsp_word(mytag | 1u, lab);  // nobody else ever claims a queued entry
        else p.cas_old = atomicCAS(&px[x], wx, wsp_word(mytag, lab));
        if (lab != WS_WSHED) {
#pragma unroll
            for (int d = 0; d < 4; d++) {
                p.link_old[d] = link_exp;
                if (vis[d] != 0) continue;
                const int y = x + off[d];
                const int l = wsp_diff(cx, cn[d]);
                p.old[d] = atomicMin(&px[y], wsp_word(mytag, l < c ? WSP_PRIVATE - WSP_NOLINK : WS_IN_QUEUE));
                p.att |= 1u << d;
                p.lv |= (unsigned)l << (8 * d);
                if (l < c) {  // into our queue, as if the claim had succeeded (settle() checks)
                    if (mask[l >> 5] & (1u << (l & 31))) {
                        p.link_old[d] = atomicCAS(&px[tail[l]], link_exp, wsp_word(mytag, WSP_PRIVATE - y));
                        tail[l] = y;
                    } else {
                        mask[l >> 5] |= 1u << (l & 31);
                        head[l] = tail[l] = y;
                    }
                }
            }
        }
        // ---- next pop: head of the lowest non-empty private level ------------------------------------------------
        int l = -1;
#pragma unroll
        for (int i = NL / 32 - 1; i >= 0; i--)
            if (mask[i]) l = i * 32 + __ffs(mask[i]) - 1;
        if (l < 0) break;
        x = head[l];
        if (x == tail[l]) {
            mask[l >> 5] &= ~(1u << (l & 31));
            lx = -1;
        } else {
            lx = l;  // head[l] becomes the link stored in x's word, which is loaded with the neighbourhood of x
        }
    }
    settle();
    if (void_run) wsp_flag_dirty(a, uq);
    for (; chunk_left > 0; chunk_left--) a.rec[chunk++].pixel = -1;
}

// the record kernels run over [0, pool top) with the top read on the device (no host round trip between run and validation)
#define WSP_FOR_RECORDS(i)                                                                                        \
    const unsigned long long wsp_top = min(a.ctl->nrec, a.rec_cap);                                               \
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < wsp_top;           \
         i += (unsigned long long)gridDim.x * blockDim.x)

// ---- validation: one thread per record ---------------------------------------------------------------------------
__device__ __forceinline__ void wsp_validate_one(const WspArgs& a, unsigned long long i)
{
    WspRec* R = a.rec + i;
    const int x = R->pixel;
    if (x < 0) return;
    const unsigned q = (unsigned)R->rank;
    const int ms = a.ms;
    const int off[4] = {-1, 1, -ms, ms};
    const int4 view = *reinterpret_cast<const int4*>(R->view);
    const int vw[4] = {view.x, view.y, view.z, view.w};
    unsigned pm = R->pushmask;
    const unsigned lv = *reinterpret_cast<const unsigned*>(R->lvl);
    // the neighbour states this pop would see now: its own block's claims as recorded, lower ranks' claims as they ended
    // up, everything else as committed
    int st[4];
    unsigned long long cl[4];
#pragma unroll
    for (int d = 0; d < 4; d++) {
        cl[d] = wsp_ld_own(a.px + x + off[d]);
        st[d] = (pm & (16u << d)) ? vw[d] : wsp_visible(cl[d], q);
    }
    int lab = 0;
#pragma unroll
    for (int d = 0; d < 4; d++) {
        const int t = st[d];
        if (t > 0) {
            if (lab == 0) lab = t;
            else if (t != lab) lab = WS_WSHED;
        }
    }
    bool bad = lab != R->label;
    const unsigned xt = wsp_tag(wsp_ld_own(a.px + x));
    if (xt == WSP_COMMITTED || (xt >> 1) != q) bad = true;  // the popped pixel itself was taken by a lower rank
    unsigned nout = 0;
    bool repaired = false;
#pragma unroll
    for (int d = 0; d < 4; d++) {
        const bool should = lab != WS_WSHED && st[d] == 0;
        const bool has = (pm >> d) & 1u;
        const int l = (int)((lv >> (8 * d)) & 0xff);
        if (has) {
            const unsigned t = wsp_tag(cl[d]);
            if (t != WSP_COMMITTED && (t >> 1) == q) {
                if (!should) bad = true;
                else if (l >= a.c) nout++;
            } else if (t != WSP_COMMITTED && (t >> 1) < q && l >= a.c) {
                pm &= ~(1u << d);  // a lower rank queued it first: in the sequential order this pop finds it queued
                repaired = true;
            } else {
                bad = true;
            }
        } else if (should) {
            bad = true;  // e.g. the lower claim that kept us from queueing it has been retracted
        }
    }
    if (repaired) R->pushmask = pm;
    if (bad) wsp_flag_dirty(a, q);
    if (nout) atomicAdd(&a.ctl->nout, nout);
    atomicAdd(&a.ctl->nlive, 1u);
}

__global__ void __launch_bounds__(256) wsp_validate(WspArgs a)
{
    WSP_FOR_RECORDS(i) wsp_validate_one(a, i);
}

// ---- retraction of the records of the blocks that run again ---------------------------------------------------------
__device__ __forceinline__ void wsp_retract_one(const WspArgs& a, unsigned long long i)
{
    WspRec* R = a.rec + i;
    const int x = R->pixel;
    if (x < 0) return;
    const unsigned q = (unsigned)R->rank;
    if (!a.runf[q]) return;
    const int ms = a.ms;
    const int off[4] = {-1, 1, -ms, ms};
    const unsigned pm = R->pushmask;
    auto release = [&](WspPx* p) {  // only what is (still) ours goes back to the committed state it replaced
        const unsigned long long v = wsp_ld_own(p);
        const unsigned t = wsp_tag(v);
        if (t != WSP_COMMITTED && (t >> 1) == q) atomicCAS(p, v, wsp_word(WSP_COMMITTED, wsp_under(t)));
    };
    release(a.px + x);
#pragma unroll
    for (int d = 0; d < 4; d++)
        if (pm & (1u << d)) release(a.px + x + off[d]);
    R->pixel = -1;
}

__global__ void __launch_bounds__(256) wsp_retract(WspArgs a)
{
    WSP_FOR_RECORDS(i) wsp_retract_one(a, i);
}

// ---- commit ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void wsp_commit_one(const WspArgs& a, unsigned long long i, unsigned long long* __restrict__ keys,
                                               int* __restrict__ vals, unsigned* __restrict__ nkeys)
{
    const WspRec* R = a.rec + i;
    const int x = R->pixel;
    if (x < 0) return;
    const unsigned q = (unsigned)R->rank;
    const int ms = a.ms;
    const int off[4] = {-1, 1, -ms, ms};
    a.px[x] = wsp_word(WSP_COMMITTED, R->label);
    const unsigned pm = R->pushmask;
    const unsigned lv = *reinterpret_cast<const unsigned*>(R->lvl);
    const unsigned seq = (unsigned)R->popseq;
#pragma unroll
    for (int d = 0; d < 4; d++) {
        if (!(pm & (1u << d))) continue;
        const int y = x + off[d];
        const int l = (int)((lv >> (8 * d)) & 0xff);
        if (l >= a.c) {  // stays queued after this round (a pixel queued below c was popped by a record of its own)
            a.px[y] = wsp_word(WSP_COMMITTED, WS_IN_QUEUE);
            const unsigned k = atomicAdd(nkeys, 1u);
            keys[k] = ((unsigned long long)l << 56) | ((unsigned long long)q << 29) | ((unsigned long long)seq << 2) | (unsigned)d;
            vals[k] = y;
        }
    }
}

__global__ void __launch_bounds__(256) wsp_commit(WspArgs a, unsigned long long* __restrict__ keys, int* __restrict__ vals,
                                                  unsigned* __restrict__ nkeys)
{
    WSP_FOR_RECORDS(i) wsp_commit_one(a, i, keys, vals, nkeys);
}

size_t wsp_align(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

// key layout of a queued push: level 8 | rank 27 | pop number inside the block 27 | direction 2
constexpr int WSP_RANK_BITS = 27;

size_t ofxcv_wsp_workspace_bytes(int W, int H, ptrdiff_t pitch)
{
    const size_t st = (size_t)pitch * H;
    size_t sort_tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int*)nullptr, (int*)nullptr,
                                    (int)st, 0, 64, (cudaStream_t)0);
    (void)W;
    return wsp_align(st * sizeof(WspPx)) + wsp_align((st + 65536) * sizeof(WspRec)) + 4 * wsp_align(st * 4 + 4096) + 2 * wsp_align(st * 8) +
           wsp_align(st * 4) + wsp_align(sort_tmp) + 65536;
}

// Flood of one prepared frame (border = -1, negatives = 0, pix = packed RGBX; what ws_prepare leaves).  Blocking: the
// round loop reads a few counters back per pass.  Returns OFXCV_OK, an error, or 1 = "degenerate flood: run the
// one-thread flood instead" (the label map is untouched in that case).
int ofxcv_wsp_flood(ofxcv_ctx* ctx, cudaStream_t s, int32_t* m, ptrdiff_t pitch, const uint32_t* pix, int W, int H, int64_t* pops_out)
{
    const size_t st = (size_t)pitch * H;
    if (st >= ((size_t)1 << WSP_RANK_BITS)) return 1;  // ranks / pop numbers would not fit the sort key
    size_t sort_tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int*)nullptr,
                                    (int*)nullptr, (int)st, 0, 64, s);
    const size_t rec_cap = st + 65536;
    char* arena = (char*)ofxcv_ws(ctx, WS_WSP_ARENA, ofxcv_wsp_workspace_bytes(W, H, pitch));
    WspCtl* hctl = (WspCtl*)ofxcv_pin(ctx, 13, 4096);
    if (!arena || !hctl) return OFXCV_ERR_MEMORY;
    int* hstart = (int*)(hctl + 1);  // 257 level boundaries
    size_t o = 0;
    auto carve = [&](size_t bytes) {
        char* p = arena + o;
        o += wsp_align(bytes);
        return p;
    };
    WspPx* px = (WspPx*)carve(st * sizeof(WspPx));
    WspRec* rec = (WspRec*)carve(rec_cap * sizeof(WspRec));
    int* Q = (int*)carve(st * 4 + 4096);
    int* ent0 = (int*)carve(st * 4 + 4096);
    int* dirty = (int*)carve(st * 4 + 4096);
    int* runf = (int*)carve(st * 4 + 4096);
    unsigned long long* keys = (unsigned long long*)carve(st * 8);
    unsigned long long* keys2 = (unsigned long long*)carve(st * 8);
    int* vals = (int*)carve(st * 4);
    void* sort_tmp = carve(sort_tmp_bytes);
    WspCtl* ctl = (WspCtl*)carve(4096);
    int* dstart = (int*)(ctl + 1);
    WspSeg* dsegs = (WspSeg*)carve(32768);
    constexpr int MAX_GATHER = 32768 / (int)sizeof(WspSeg);

    OFXCV_CUDA(ctx, cudaMemsetAsync(ctl, 0, sizeof(WspCtl), s));
    wsp_pack<<<(unsigned)((st + 255) / 256), 256, 0, s>>>(m, px, st);
    OFXCV_LAUNCH_CHECK(ctx);
    auto readback = [&](size_t bytes) -> int {
        OFXCV_CUDA(ctx, cudaMemcpyAsync(hctl, ctl, bytes, cudaMemcpyDeviceToHost, s));
        OFXCV_CUDA(ctx, cudaStreamSynchronize(s));
        return OFXCV_OK;
    };
    // per-level lists of queue segments (offset into Q, count), in push order
    std::vector<std::pair<int, int>> segs[256];
    size_t qtop = 0;
    auto append_sorted = [&](int n, int cmin, int* next_off, int* next_cnt) -> int {
        // keys/vals[0..n) -> sorted by key, values land at Q + qtop; registers one segment per level present
        OFXCV_CUDA(ctx, cub::DeviceRadixSort::SortPairs(sort_tmp, sort_tmp_bytes, keys, keys2, vals, Q + qtop, n, 0, 64, s));
        ctx->launches += 2;
        wsp_bounds<<<1, 288, 0, s>>>(keys2, n, dstart);
        OFXCV_LAUNCH_CHECK(ctx);
        OFXCV_CUDA(ctx, cudaMemcpyAsync(hstart, dstart, 257 * sizeof(int), cudaMemcpyDeviceToHost, s));
        OFXCV_CUDA(ctx, cudaStreamSynchronize(s));
        for (int l = 0; l < 256; l++) {
            const int cnt = hstart[l + 1] - hstart[l];
            if (cnt <= 0) continue;
            if (l == cmin && next_off) {
                *next_off = (int)qtop + hstart[l];
                *next_cnt = cnt;
            } else {
                segs[l].push_back({(int)qtop + hstart[l], cnt});
            }
        }
        qtop += (size_t)n;
        return OFXCV_OK;
    };

    // ---- initial candidates ---------------------------------------------------------------------------------------
    wsp_candidates<<<dim3(ofxcv_div_up(W, 256), H), 256, 0, s>>>(m, pix, (int)pitch, W, H, keys, vals, ctl);
    OFXCV_LAUNCH_CHECK(ctx);
    int st_ = readback(sizeof(WspCtl));
    if (st_ < 0) return st_;
    const int ncand = (int)hctl->ncand;
    if (ncand > 0) {
        wsp_mark_queued<<<ofxcv_div_up(ncand, 256), 256, 0, s>>>(px, vals, ncand);
        OFXCV_LAUNCH_CHECK(ctx);
        if ((st_ = append_sorted(ncand, -1, nullptr, nullptr)) < 0) return st_;
    }

    WspArgs a;
    a.px = px;
    a.pix = pix;
    a.rec = rec;
    a.dirty = dirty;
    a.runf = runf;
    a.ctl = ctl;
    a.rec_cap = rec_cap;
    a.ms = (int)pitch;
    long rounds = 0, passes = 0;
    int64_t pops = 0;
    const bool prof = getenv("OFXCV_WS_PROF") != nullptr;
    std::vector<cudaEvent_t> pev;
    std::vector<int> pinfo;  // per pass: N, c
    auto pmark = [&]() {
        if (!prof) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        pev.push_back(e);
    };
    const int rgrid = ctx->num_sms * 8;  // grid of the record kernels (grid-stride over the pool)
    for (int c = 0; c < 256; c++) {
        if (segs[c].empty()) continue;
        // generation 0 of phase c: every segment queued for this level so far, in push order
        const int* ent = nullptr;
        int N = 0;
        if (segs[c].size() == 1) {
            ent = Q + segs[c][0].first;
            N = segs[c][0].second;
        } else {
            std::vector<WspSeg> hs;
            for (auto& sg : segs[c]) {
                hs.push_back({sg.first, sg.second, N});
                N += sg.second;
            }
            for (size_t b = 0; b < hs.size(); b += MAX_GATHER) {
                const int nb = (int)std::min(hs.size() - b, (size_t)MAX_GATHER);
                OFXCV_CUDA(ctx, cudaMemcpyAsync(dsegs, hs.data() + b, nb * sizeof(WspSeg), cudaMemcpyHostToDevice, s));
                wsp_gather<<<dim3(32, nb), 256, 0, s>>>(Q, dsegs, ent0);
                OFXCV_LAUNCH_CHECK(ctx);
                OFXCV_CUDA(ctx, cudaStreamSynchronize(s));  // hs is pageable host memory, dsegs is reused
            }
            ent = ent0;
        }
        segs[c].clear();
        while (N > 0) {
            // a flood that degenerates into very many tiny rounds is faster on the one-thread kernel
            if (++rounds >= 2048 && (rounds & 1023) == 0 && pops < rounds * 256) return 1;
            a.ent = ent;
            a.N = N;
            a.c = c;
            wsp_round_init<<<ofxcv_div_up(N, 256), 256, 0, s>>>(dirty, runf, N, ctl);
            OFXCV_LAUNCH_CHECK(ctx);
            for (bool first = true;; first = false) {
                passes++;
                if (!first) {
                    wsp_latch<<<ofxcv_div_up(N, 256), 256, 0, s>>>(dirty, runf, N, ctl);
                    OFXCV_LAUNCH_CHECK(ctx);
                    wsp_retract<<<rgrid, 256, 0, s>>>(a);
                    OFXCV_LAUNCH_CHECK(ctx);
                }
                pmark();
                if (c <= 32) wsp_run<32><<<ofxcv_div_up(N, 128), 128, 0, s>>>(a);
                else wsp_run<256><<<ofxcv_div_up(N, 128), 128, 0, s>>>(a);
                OFXCV_LAUNCH_CHECK(ctx);
                pmark();
                wsp_validate<<<rgrid, 256, 0, s>>>(a);
                OFXCV_LAUNCH_CHECK(ctx);
                pmark();
                if (prof) {
                    pinfo.push_back(N);
                    pinfo.push_back(c);
                }
                if ((st_ = readback(sizeof(WspCtl))) < 0) return st_;
                if (hctl->overflow) return 1;
                if (hctl->ndirty == 0) break;
            }
            const int nout = (int)hctl->nout;
            pops += (int64_t)hctl->nlive;
            int next_off = 0, next_cnt = 0;
            OFXCV_CUDA(ctx, cudaMemsetAsync(&ctl->nout, 0, sizeof(unsigned), s));
            wsp_commit<<<rgrid, 256, 0, s>>>(a, keys, vals, &ctl->nout);
            OFXCV_LAUNCH_CHECK(ctx);
            if (nout > 0) {
                if (qtop + (size_t)nout > st + 1024) return 1;
                if ((st_ = append_sorted(nout, c, &next_off, &next_cnt)) < 0) return st_;
            }
            ent = Q + next_off;
            N = next_cnt;
        }
    }
    wsp_unpack<<<(unsigned)((st + 255) / 256), 256, 0, s>>>(px, m, st);
    OFXCV_LAUNCH_CHECK(ctx);
    if (prof) {
        cudaStreamSynchronize(s);
        double t_run = 0, t_val = 0;
        std::vector<std::pair<float, size_t>> top;
        for (size_t i = 0; i + 2 < pev.size() && i / 3 < pinfo.size() / 2; i += 3) {
            float a_ = 0, b_ = 0;
            cudaEventElapsedTime(&a_, pev[i], pev[i + 1]);
            cudaEventElapsedTime(&b_, pev[i + 1], pev[i + 2]);
            t_run += a_;
            t_val += b_;
            top.push_back({a_, i / 3});
        }
        std::sort(top.begin(), top.end(), [](auto& x, auto& y) { return x.first > y.first; });
        fprintf(stderr, "[wsp] rounds %ld passes %ld: run kernels %.1f ms, validate %.1f ms; longest runs:", rounds, passes, t_run, t_val);
        for (size_t i = 0; i < top.size() && i < 12; i++) fprintf(stderr, " %.2fms(#%zu,N=%d,c=%d)", top[i].first, top[i].second, pinfo[2 * top[i].second], pinfo[2 * top[i].second + 1]);
        fprintf(stderr, "\n");
        for (auto e : pev) cudaEventDestroy(e);
    }
    if (pops_out) *pops_out = pops;
    ctx->watershed_stats[2] = rounds;
    ctx->watershed_stats[3] = passes;
    return OFXCV_OK;
}
