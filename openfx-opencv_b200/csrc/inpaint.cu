// FMM inpainting (Telea and Navier-Stokes) for sm_100a — the body behind the inpaint plugin's render action
// (/root/reference/opencv2fx/inpaint/inpaint.cpp:311-318: cvInpaint(image0, mask, image1, radius, TELEA)).
// Algorithm per SURVEY.md Appendix A.2 (OpenCV photo/inpaint.cpp as pinned by cv2 4.13); the result is bit-exact
// with the CPU path by construction.  THIS FILE IS COMPILED WITH -fmad=false.
//
// The CPU algorithm is one sequential priority process.  It is split here into two exactly-equivalent stages:
//
//  Stage A (mask only): the fast-marching ORDER and the T values.  Entries leave the heap in (T, push-counter)
//    order and a pixel reached by a pop at T=tau gets T' >= tau + 0.7071, so every heap entry with
//    T in [tau_min, tau_min+0.7) can be popped as ONE parallel batch: the pixels reached by the batch get their
//    pusher (smallest (T,counter) selected neighbour), a fill key (pusher key, neighbour slot), and -- after a
//    radix sort of the keys -- the same push counters the CPU would hand out.  T of a reached pixel depends only
//    on 4-neighbours reached earlier; inside a batch that is a short dependency chain resolved by relaxation
//    rounds.  Telea runs this twice (outside band, negated; then the hole), Navier-Stokes once.
//
//  Stage B (colours): every hole pixel, in fill order, is a weighted sum over its radius-disc of pixels that were
//    known or filled before it.  One WARP per pixel: tickets are handed out in fill order, a warp spins until the
//    earlier-filled hole pixels inside its (2r+3)^2 box are done (dataflow, only ever waits on lower tickets),
//    the lanes evaluate the taps' weights in parallel, and the f32 accumulators are then summed in the CPU's
//    row-major tap order (one lane per accumulator) so that every rounding matches.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <thread>
#include <vector>

#include <cub/cub.cuh>

#include "common.cuh"

namespace {

enum : uint8_t { ST_OUT = 0, ST_INSIDE = 1, ST_HEAP = 2, ST_POPPED = 3, ST_SELECTED = 4, ST_NEW = 5 };
constexpr float T_FAR = 1.0e6f;
constexpr uint32_t CNT_NONE = 0xffffffffu;

struct IpGeom {
    int W, H, er, ec;  // image size, padded map size (H+2, W+2)
    int np;            // er*ec
};

// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ip_init(const uint8_t* __restrict__ mask, ptrdiff_t mstride, uint8_t* __restrict__ hole,
                                               float* __restrict__ t, uint32_t* __restrict__ cnt, uint16_t* __restrict__ rnd,
                                               uint8_t* __restrict__ done, unsigned* __restrict__ counters, IpGeom g)
{
    int id = blockIdx.x * blockDim.x + threadIdx.x;
    uint8_t h = 0;
    if (id < g.np) {
        int i = id / g.ec, j = id - i * g.ec;
        if (i >= 1 && i <= g.H && j >= 1 && j <= g.W) h = mask[(size_t)(i - 1) * mstride + (j - 1)] != 0;
        hole[id] = h;
        t[id] = T_FAR;
        cnt[id] = CNT_NONE;
        rnd[id] = 0;
        done[id] = 0;
    }
    // number of hole pixels: one atomic per block (same-address atomics run at ~2 per ns)
    const int c = __syncthreads_count(h != 0);
    if (threadIdx.x == 0 && c) atomicAdd(&counters[0], (unsigned)c);
}

// band = known interior pixels with a hole 4-neighbour; heap entries with T = 0 pushed in row-major order
// (counter = row-major id).  region: 0 -> FMM over the hole (st INSIDE on holes), 1 -> FMM over `outreg`.
__global__ void __launch_bounds__(256) ip_setup_pass(const uint8_t* __restrict__ hole, const uint8_t* __restrict__ outreg,
                                                     uint8_t* __restrict__ st, float* __restrict__ t, uint32_t* __restrict__ cnt,
                                                     uint16_t* __restrict__ rnd, int outer, IpGeom g)
{
    int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= g.np) return;
    int i = id / g.ec, j = id - i * g.ec;
    uint8_t s = ST_OUT;
    bool interior = i >= 1 && i <= g.H && j >= 1 && j <= g.W;
    if (interior) {
        bool h = hole[id];
        bool band = !h && (hole[id - g.ec] || hole[id + g.ec] || hole[id - 1] || hole[id + 1]);
        if (band) {
            s = ST_HEAP;
            t[id] = 0.f;
            cnt[id] = (uint32_t)id;
        } else if (outer ? (outreg[id] != 0) : h) {
            s = ST_INSIDE;
        }
    }
    st[id] = s;
    rnd[id] = 0;
}

// (2r+1)^2 rect dilation of the hole, separable; out-region = dilated & !hole & !band, border ring cleared
__global__ void __launch_bounds__(256) ip_dilate_rows(const uint8_t* __restrict__ hole, uint8_t* __restrict__ tmp, int range, IpGeom g)
{
    int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= g.np) return;
    int i = id / g.ec, j = id - i * g.ec;
    uint8_t v = 0;
    for (int l = max(j - range, 0); l <= min(j + range, g.ec - 1); l++) v |= hole[i * g.ec + l];
    tmp[id] = v;
}
__global__ void __launch_bounds__(256) ip_dilate_cols(const uint8_t* __restrict__ hole, const uint8_t* __restrict__ tmp,
                                                      uint8_t* __restrict__ outreg, int range, IpGeom g)
{
    int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= g.np) return;
    int i = id / g.ec, j = id - i * g.ec;
    uint8_t r = 0;
    if (i >= 1 && i <= g.H && j >= 1 && j <= g.W && !hole[id]) {
        uint8_t v = 0;
        for (int k = max(i - range, 0); k <= min(i + range, g.er - 1); k++) v |= tmp[k * g.ec + j];
        bool band = hole[id - g.ec] || hole[id + g.ec] || hole[id - 1] || hole[id + 1];
        r = v && !band;
    }
    outreg[id] = r;
}

// ---- one batch of the marching ---------------------------------------------------------------------------
// pass A: retire last batch (SELECTED -> POPPED, NEW -> HEAP) and find tau = min T over the heap
__global__ void __launch_bounds__(256) ip_pass_a(uint8_t* __restrict__ st, const float* __restrict__ t, unsigned* __restrict__ tau_bits, IpGeom g)
{
    int id = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned local = 0xffffffffu;
    if (id < g.np) {
        uint8_t s = st[id];
        if (s == ST_SELECTED) st[id] = ST_POPPED;
        else if (s == ST_NEW) { st[id] = ST_HEAP; s = ST_HEAP; }
        if (s == ST_HEAP) local = __float_as_uint(t[id]);  // T >= 0: the bit pattern orders like the value
    }
    __shared__ unsigned wmin[8];
    local = __reduce_min_sync(0xffffffffu, local);
    if ((threadIdx.x & 31) == 0) wmin[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {  // one atomic per block
        local = __reduce_min_sync(0xffffffffu, threadIdx.x < 8 ? wmin[threadIdx.x] : 0xffffffffu);
        if (threadIdx.x == 0 && local != 0xffffffffu) atomicMin(tau_bits, local);
    }
}

// pass B: select the window [tau, tau+0.7)
__global__ void __launch_bounds__(256) ip_pass_b(uint8_t* __restrict__ st, const float* __restrict__ t, const unsigned* __restrict__ tau_bits, IpGeom g)
{
    int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= g.np) return;
    if (st[id] == ST_HEAP) {
        float lim = __uint_as_float(*tau_bits) + 0.7f;
        if (t[id] < lim) st[id] = ST_SELECTED;
    }
}

// pass C: every INSIDE pixel next to a selected entry is reached in this batch; pusher = smallest (T, counter),
// slot = position of the pixel in the pusher's neighbour order (up, left, down, right of the PUSHER)
__global__ void __launch_bounds__(256) ip_pass_c(uint8_t* __restrict__ st, const float* __restrict__ t, const uint32_t* __restrict__ cnt,
                                                 unsigned long long* __restrict__ keys, uint32_t* __restrict__ ids,
                                                 unsigned* __restrict__ n_new, IpGeom g)
{
    int id = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long best = ~0ull;
    if (id < g.np && st[id] == ST_INSIDE) {
        // neighbour p of q        q = p + dir[slot]
        //  p above  (id-ec)       slot 2 (down)
        //  p left   (id-1)        slot 3 (right)
        //  p below  (id+ec)       slot 0 (up)
        //  p right  (id+1)        slot 1 (left)
        const int nb[4] = {id - g.ec, id - 1, id + g.ec, id + 1};
        const unsigned slot[4] = {2u, 3u, 0u, 1u};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            int p = nb[q];
            if (st[p] == ST_SELECTED) {
                unsigned long long k = ((unsigned long long)__float_as_uint(t[p]) << 32) | ((unsigned long long)cnt[p] << 2) | slot[q];
                best = k < best ? k : best;
            }
        }
    }
    // append (the sort that follows fixes the order): one global atomic per block
    __shared__ unsigned wcount[8], base;
    const bool hit = best != ~0ull;
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) wcount[w] = __popc(bal);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { unsigned c = wcount[k]; wcount[k] = tot; tot += c; }
        base = tot ? atomicAdd(n_new, tot) : 0u;
    }
    __syncthreads();
    if (hit) {
        unsigned pos = base + wcount[w] + __popc(bal & ((1u << lane) - 1u));
        keys[pos] = best;
        ids[pos] = (uint32_t)id;
    }
}

// ---- the same three passes over 16 / 4 pixels per thread ----------------------------------------------------------
// The state map is one byte per pixel and most of it is OUT / POPPED / INSIDE-far-from-the-front in every batch: a
// thread reads 16 (passes A, B) or 4 (pass C) states with one load, tests them with byte-wise SIMD compares and only
// looks at T / neighbours for the few bytes that matter.  The map is padded to a multiple of 16 with OUT bytes.
__device__ __forceinline__ bool ip_any_byte(uint32_t w, uint32_t v) { return __vcmpeq4(w, v * 0x01010101u) != 0; }

__global__ void __launch_bounds__(256) ip_pass_a16(uint4* __restrict__ st16, const float* __restrict__ t, unsigned* __restrict__ tau_bits, int n16)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned local = 0xffffffffu;
    if (i < n16) {
        const uint4 w4 = st16[i];
        uint32_t v[4] = {w4.x, w4.y, w4.z, w4.w};
        bool changed = false;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t w = v[q];
            if (!(ip_any_byte(w, ST_HEAP) || ip_any_byte(w, ST_SELECTED) || ip_any_byte(w, ST_NEW))) continue;
            uint32_t o = w;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                uint32_t sb = (w >> (8 * b)) & 0xffu;
                if (sb == ST_SELECTED) sb = ST_POPPED;
                else if (sb == ST_NEW) sb = ST_HEAP;
                o = (o & ~(0xffu << (8 * b))) | (sb << (8 * b));
                if (sb == ST_HEAP) local = min(local, __float_as_uint(t[(size_t)i * 16 + q * 4 + b]));
            }
            changed |= o != w;
            v[q] = o;
        }
        if (changed) st16[i] = make_uint4(v[0], v[1], v[2], v[3]);
    }
    __shared__ unsigned wmin[8];
    local = __reduce_min_sync(0xffffffffu, local);
    if ((threadIdx.x & 31) == 0) wmin[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
        local = __reduce_min_sync(0xffffffffu, threadIdx.x < 8 ? wmin[threadIdx.x] : 0xffffffffu);
        if (threadIdx.x == 0 && local != 0xffffffffu) atomicMin(tau_bits, local);
    }
}

__global__ void __launch_bounds__(256) ip_pass_b16(uint4* __restrict__ st16, const float* __restrict__ t, const unsigned* __restrict__ tau_bits, int n16)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n16) return;
    const uint4 w4 = st16[i];
    uint32_t v[4] = {w4.x, w4.y, w4.z, w4.w};
    bool changed = false;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint32_t w = v[q];
        if (!ip_any_byte(w, ST_HEAP)) continue;
        const float lim = __uint_as_float(*tau_bits) + 0.7f;
        uint32_t o = w;
#pragma unroll
        for (int b = 0; b < 4; b++)
            if (((w >> (8 * b)) & 0xffu) == ST_HEAP && t[(size_t)i * 16 + q * 4 + b] < lim)
                o = (o & ~(0xffu << (8 * b))) | ((uint32_t)ST_SELECTED << (8 * b));
        changed |= o != w;
        v[q] = o;
    }
    if (changed) st16[i] = make_uint4(v[0], v[1], v[2], v[3]);
}

__global__ void __launch_bounds__(256) ip_pass_c4(const uint8_t* __restrict__ st, const float* __restrict__ t, const uint32_t* __restrict__ cnt,
                                                  unsigned long long* __restrict__ keys, uint32_t* __restrict__ ids,
                                                  unsigned* __restrict__ n_new, int n4, IpGeom g)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long best[4] = {~0ull, ~0ull, ~0ull, ~0ull};
    int nhit = 0;
    if (i < n4) {
        const uint32_t w = reinterpret_cast<const uint32_t*>(st)[i];
        if (ip_any_byte(w, ST_INSIDE)) {
#pragma unroll
            for (int b = 0; b < 4; b++) {
                if (((w >> (8 * b)) & 0xffu) != ST_INSIDE) continue;
                const int id = i * 4 + b;  // INSIDE pixels are interior: all four neighbours exist
                const int nb[4] = {id - g.ec, id - 1, id + g.ec, id + 1};
                const unsigned slot[4] = {2u, 3u, 0u, 1u};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int p = nb[q];
                    if (st[p] == ST_SELECTED) {
                        const unsigned long long k = ((unsigned long long)__float_as_uint(t[p]) << 32) | ((unsigned long long)cnt[p] << 2) | slot[q];
                        best[b] = k < best[b] ? k : best[b];
                    }
                }
                nhit += best[b] != ~0ull;
            }
        }
    }
    // append: exclusive scan of the per-thread counts inside the warp, warp totals in shared memory, one global atomic
    __shared__ unsigned wcount[8], base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned incl = nhit;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) wcount[wid] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { const unsigned c = wcount[k]; wcount[k] = tot; tot += c; }
        base = tot ? atomicAdd(n_new, tot) : 0u;
    }
    __syncthreads();
    unsigned pos = base + wcount[wid] + incl - nhit;
#pragma unroll
    for (int b = 0; b < 4; b++)
        if (best[b] != ~0ull) {
            keys[pos] = best[b];
            ids[pos] = (uint32_t)(i * 4 + b);
            pos++;
        }
}

// after the sort: hand out push counters in fill order; the hole pass also records the fill order itself
__global__ void __launch_bounds__(256) ip_assign(const uint32_t* __restrict__ ids_sorted, unsigned n, uint32_t base,
                                                 uint8_t* __restrict__ st, uint32_t* __restrict__ cnt, uint32_t* __restrict__ order,
                                                 uint32_t order_base)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t id = ids_sorted[i];
    cnt[id] = base + i;
    st[id] = ST_NEW;
    if (order) order[order_base + i] = id;
}

__device__ __forceinline__ float fmm_solve(bool k1, bool k2, float t1, float t2)
{
    double sol, a11 = t1, a22 = t2, m12 = a11 < a22 ? a11 : a22;
    if (k1) {
        if (k2) {
            if (fabs(a11 - a22) >= 1.0) sol = 1 + m12;
            else sol = (a11 + a22 + sqrt((double)(2 - (a11 - a22) * (a11 - a22)))) * 0.5;
        } else sol = 1 + a11;
    } else if (k2) sol = 1 + a22;
    else sol = 1 + m12;
    return (float)sol;
}

// relaxation round `round` (>=1): a pixel reached in this batch gets its T once every 4-neighbour reached EARLIER
// in the same batch has its T (rnd != 0 and < round: values written in this very round are ignored -> race-free)
__global__ void __launch_bounds__(256) ip_round(const uint32_t* __restrict__ ids_sorted, unsigned n, const uint8_t* __restrict__ st,
                                                float* __restrict__ t, const uint32_t* __restrict__ cnt, uint16_t* __restrict__ rnd,
                                                unsigned round, unsigned* __restrict__ pending, IpGeom g)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t id = ids_sorted[i];
    if (rnd[id] != 0) return;
    const uint32_t mycnt = cnt[id];
    const int nb[4] = {(int)id - g.ec, (int)id + g.ec, (int)id - 1, (int)id + 1};  // up, down, left, right
    bool known[4];
    float tv[4];
    bool ready = true;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        int p = nb[q];
        uint8_t s = st[p];
        bool k;
        if (s == ST_INSIDE) k = false;
        else if (s == ST_NEW) {
            k = cnt[p] < mycnt;
            if (k) {
                unsigned r = rnd[p];
                if (r == 0 || r >= round) ready = false;
            }
        } else k = true;
        known[q] = k;
        tv[q] = k ? t[p] : T_FAR;  // an unreached pixel still carries T = 1e6 on the CPU at this moment
    }
    if (!ready) {
        const unsigned nb_ = __activemask();  // lanes that are not ready in this warp: one atomic for all of them
        if ((threadIdx.x & 31) == (unsigned)(__ffs(nb_) - 1)) atomicAdd(pending, (unsigned)__popc(nb_));
        return;
    }
    // min4(solve(i-1,j,i,j-1), solve(i+1,j,i,j-1), solve(i-1,j,i,j+1), solve(i+1,j,i,j+1))
    float a = fmm_solve(known[0], known[2], tv[0], tv[2]);
    float b = fmm_solve(known[1], known[2], tv[1], tv[2]);
    float c = fmm_solve(known[0], known[3], tv[0], tv[3]);
    float d = fmm_solve(known[1], known[3], tv[1], tv[3]);
    a = a < b ? a : b;
    c = c < d ? c : d;
    t[id] = a < c ? a : c;
    rnd[id] = (uint16_t)round;
}

// The same relaxation in ONE launch per batch.  A reached pixel depends only on 4-neighbours reached EARLIER in the batch,
// i.e. on entries with a smaller index in the sorted list, so the chains can be followed inside the kernel: blocks take
// their index from a ticket (a block that holds a ticket is running and every lower ticket is running or finished, so
// waiting on lower indices cannot deadlock), a warp polls the stamps of the neighbours its lanes still miss and every
// lane computes as soon as its own are there (no lane ever blocks another lane of its warp).  T is published with a
// fence before the stamp; readers take the stamp with an acquire load.  Each T is computed exactly once from final
// neighbour values, as in the round-by-round version.
__global__ void __launch_bounds__(256) ip_round_chain(const uint32_t* __restrict__ ids_sorted, unsigned n, const uint8_t* __restrict__ st,
                                                      float* t, const uint32_t* __restrict__ cnt, uint16_t* rnd, unsigned* __restrict__ ticket,
                                                      IpGeom g)
{
    __shared__ unsigned bid;
    if (threadIdx.x == 0) bid = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned i = bid * blockDim.x + threadIdx.x;
    bool active = i < n;
    uint32_t id = 0;
    int nb[4] = {0, 0, 0, 0};
    bool known[4] = {false, false, false, false}, wait[4] = {false, false, false, false};
    float tv[4] = {T_FAR, T_FAR, T_FAR, T_FAR};
    if (active) {
        id = ids_sorted[i];
        const uint32_t mycnt = cnt[id];
        nb[0] = (int)id - g.ec; nb[1] = (int)id + g.ec; nb[2] = (int)id - 1; nb[3] = (int)id + 1;  // up, down, left, right
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint8_t s = st[nb[q]];
            if (s == ST_INSIDE) known[q] = false;
            else if (s == ST_NEW) {
                known[q] = cnt[nb[q]] < mycnt;
                wait[q] = known[q];      // its T is being computed in this very launch
            } else known[q] = true;
            if (known[q] && !wait[q]) tv[q] = t[nb[q]];
        }
    }
    bool done = !active;
    while (!__all_sync(0xffffffffu, done)) {
        if (!done) {
            bool ready = true;
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (wait[q]) {
                    if (*((volatile uint16_t*)rnd + nb[q]) != 0) {  // relaxed poll, then ONE acquire load (see ip_fill_staged)
                        unsigned stamp;
                        asm volatile("ld.acquire.gpu.global.u16 %0, [%1];" : "=r"(stamp) : "l"(rnd + nb[q]) : "memory");
                        tv[q] = *((volatile float*)t + nb[q]);
                        wait[q] = false;
                    } else ready = false;
                }
            if (ready) {
                // min4(solve(i-1,j,i,j-1), solve(i+1,j,i,j-1), solve(i-1,j,i,j+1), solve(i+1,j,i,j+1))
                float a = fmm_solve(known[0], known[2], tv[0], tv[2]);
                float b = fmm_solve(known[1], known[2], tv[1], tv[2]);
                float c = fmm_solve(known[0], known[3], tv[0], tv[3]);
                float d = fmm_solve(known[1], known[3], tv[1], tv[3]);
                a = a < b ? a : b;
                c = c < d ? c : d;
                *((volatile float*)t + id) = a < c ? a : c;
                __threadfence();
                *((volatile uint16_t*)rnd + id) = 1;
                done = true;
            }
        }
    }
}

__global__ void __launch_bounds__(256) ip_negate(const uint8_t* __restrict__ st, const uint8_t* __restrict__ outreg, float* __restrict__ t, IpGeom g)
{
    int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= g.np) return;
    if (outreg[id] && st[id] != ST_INSIDE && st[id] != ST_OUT) t[id] = -t[id];
}

__global__ void __launch_bounds__(256) ip_copy(const uint8_t* __restrict__ src, ptrdiff_t sstride, uint8_t* __restrict__ dst, ptrdiff_t dstride,
                                               int rowbytes, int H)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= rowbytes) return;
    dst[(size_t)y * dstride + x] = src[(size_t)y * sstride + x];
}

// ---- Stage B: colour fill, one warp per hole pixel, dataflow in fill order -------------------------------
constexpr int IP_WARPS = 8;
constexpr int IP_MAXACC = 12;  // Telea, 3 channels: (Ia, Jx, Jy, s) x 3

__device__ __forceinline__ int ld_u8_cg(const uint8_t* p) { return (int)__ldcg(p); }

template <int METHOD, int CN>
__global__ void __launch_bounds__(IP_WARPS * 32)
ip_fill(const uint32_t* __restrict__ order, unsigned nfill, uint32_t cnt_base, const uint8_t* __restrict__ hole,
        const uint32_t* __restrict__ cnt, const float* __restrict__ t, uint8_t* out, ptrdiff_t ostride, uint8_t* done,
        unsigned* __restrict__ ticket, int range, IpGeom g)
{
    constexpr int NACC = METHOD == OFXCV_INPAINT_TELEA ? 4 * CN : 2 * CN;
    __shared__ float s_term[IP_WARPS][32][IP_MAXACC + 1];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int ec = g.ec, er = g.er;
    const int side = 2 * range + 1, ntaps = side * side;
    const int bside = 2 * range + 3, nbox = bside * bside;
    volatile uint8_t* vdone = done;

    for (;;) {
        unsigned tk = 0;
        if (lane == 0) tk = atomicAdd(ticket, 1u);
        tk = __shfl_sync(0xffffffffu, tk, 0);
        if (tk >= nfill) break;
        const int id = (int)order[tk];
        const uint32_t mycnt = cnt_base + tk;
        const int i = id / ec, j = id - i * ec;

        // 1. wait for every earlier-filled hole pixel of the (2r+3)^2 box
        for (int b = lane; b < nbox; b += 32) {
            int k = i - range - 1 + b / bside, l = j - range - 1 + b % bside;
            if (k < 1 || l < 1 || k > g.H || l > g.W) continue;
            int n = k * ec + l;
            if (hole[n] && cnt[n] < mycnt) {
                while (vdone[n] == 0) { }
            }
        }
        __syncwarp();
        __threadfence();

        // a pixel is "not INSIDE" at this pixel's fill time iff it is not a hole or was filled earlier
        auto known = [&](int n) -> bool { return !hole[n] || cnt[n] < mycnt; };

        float gTx = 0.f, gTy = 0.f, ti = 0.f;
        if (METHOD == OFXCV_INPAINT_TELEA) {
            ti = t[id];
            if (known(id + 1)) {
                if (known(id - 1)) gTx = (float)(t[id + 1] - t[id - 1]) * 0.5f;
                else gTx = (float)(t[id + 1] - ti);
            } else {
                if (known(id - 1)) gTx = (float)(ti - t[id - 1]);
                else gTx = 0;
            }
            if (known(id + ec)) {
                if (known(id - ec)) gTy = (float)(t[id + ec] - t[id - ec]) * 0.5f;
                else gTy = (float)(t[id + ec] - ti);
            } else {
                if (known(id - ec)) gTy = (float)(ti - t[id - ec]);
                else gTy = 0;
            }
        }

        float acc = 0.f;  // lane a < NACC owns accumulator a
        if (METHOD == OFXCV_INPAINT_TELEA) { if ((lane & 3) == 3) acc = 1.0e-20f; }
        else { if ((lane & 1) == 1) acc = 1.0e-20f; }

        for (int base = 0; base < ntaps; base += 32) {
            const int tp = base + lane;
            bool valid = false;
            float term[IP_MAXACC];
#pragma unroll
            for (int a = 0; a < IP_MAXACC; a++) term[a] = 0.f;
            if (tp < ntaps) {
                const int k = i - range + tp / side, l = j - range + tp % side;
                if (k > 0 && l > 0 && k < er - 1 && l < ec - 1) {
                    const int n = k * ec + l;
                    if (known(n) && (l - j) * (l - j) + (k - i) * (k - i) <= range * range) {
                        valid = true;
                        const int km = k - 1 + (k == 1), kp = k - 1 - (k == er - 2);
                        const int lm = l - 1 + (l == 1), lp = l - 1 - (l == ec - 2);
                        const bool fr = known(n + 1), fl = known(n - 1), fd = known(n + ec), fu = known(n - ec);
#define OUTP(r, c, ch) ld_u8_cg(out + (size_t)(r) * ostride + (size_t)(c) * CN + (ch))
                        if (METHOD == OFXCV_INPAINT_TELEA) {
                            float ry = (float)(i - k), rx = (float)(j - l);
                            float vl = rx * rx + ry * ry;
                            float dst = (float)(1. / (vl * sqrt((double)vl)));
                            float lev = (float)(1. / (1 + (double)fabsf(t[n] - ti)));  // f32 difference, f64 sum (C fabs)
                            float dir = rx * gTx + ry * gTy;
                            if (fabs(dir) <= 0.01) dir = 0.000001f;
                            float w = (float)fabs(dst * lev * dir);
#pragma unroll
                            for (int c = 0; c < CN; c++) {
                                float gIx, gIy;
                                if (fr) {
                                    if (fl) gIx = (float)(OUTP(km, lp + 1, c) - OUTP(km, lm - 1, c)) * 2.0f;
                                    else gIx = (float)(OUTP(km, lp + 1, c) - OUTP(km, lm, c));
                                } else {
                                    if (fl) gIx = (float)(OUTP(km, lp, c) - OUTP(km, lm - 1, c));
                                    else gIx = 0;
                                }
                                if (fd) {
                                    if (fu) gIy = (float)(OUTP(kp + 1, lm, c) - OUTP(km - 1, lm, c)) * 2.0f;
                                    else gIy = (float)(OUTP(kp + 1, lm, c) - OUTP(km, lm, c));
                                } else {
                                    if (fu) gIy = (float)(OUTP(kp, lm, c) - OUTP(km - 1, lm, c));
                                    else gIy = 0;
                                }
                                term[c * 4 + 0] = w * (float)OUTP(k - 1, l - 1, c);
                                term[c * 4 + 1] = -(w * (gIx * rx));
                                term[c * 4 + 2] = -(w * (gIy * ry));
                                term[c * 4 + 3] = w;
                            }
                        } else {
                            float ry = (float)(k - i), rx = (float)(l - j);
                            float vl = rx * rx + ry * ry;
                            float dst = 1 / (vl * vl + 1);
#pragma unroll
                            for (int c = 0; c < CN; c++) {
                                float gIx, gIy;
                                if (fd) {
                                    if (fu) gIx = (float)(abs(OUTP(kp + 1, lm, c) - OUTP(kp, lm, c)) + abs(OUTP(kp, lm, c) - OUTP(km - 1, lm, c)));
                                    else gIx = (float)(abs(OUTP(kp + 1, lm, c) - OUTP(kp, lm, c))) * 2.0f;
                                } else {
                                    if (fu) gIx = (float)(abs(OUTP(kp, lm, c) - OUTP(km - 1, lm, c))) * 2.0f;
                                    else gIx = 0;
                                }
                                if (fr) {
                                    if (fl) gIy = (float)(abs(OUTP(km, lp + 1, c) - OUTP(km, lm, c)) + abs(OUTP(km, lm, c) - OUTP(km, lm - 1, c)));
                                    else gIy = (float)(abs(OUTP(km, lp + 1, c) - OUTP(km, lm, c))) * 2.0f;
                                } else {
                                    if (fl) gIy = (float)(abs(OUTP(km, lm, c) - OUTP(km, lm - 1, c))) * 2.0f;
                                    else gIy = 0;
                                }
                                gIx = -gIx;
                                float dir = rx * gIx + ry * gIy;
                                if (fabs(dir) <= 0.01) dir = 0.000001f;
                                else dir = fabsf((rx * gIx + ry * gIy) / sqrtf(vl * (gIx * gIx + gIy * gIy)));
                                float w = dst * dir;
                                term[c * 2 + 0] = w * (float)OUTP(k - 1, l - 1, c);
                                term[c * 2 + 1] = w;
                            }
                        }
#undef OUTP
                    }
                }
            }
            const unsigned vmask = __ballot_sync(0xffffffffu, valid);
            if (valid) {
#pragma unroll
                for (int a = 0; a < NACC; a++) s_term[wid][lane][a] = term[a];
            }
            __syncwarp();
            if (lane < NACC) {
                unsigned m = vmask;
                while (m) {
                    int tl = __ffs(m) - 1;
                    m &= m - 1;
                    acc = acc + s_term[wid][tl][lane];
                }
            }
            __syncwarp();
        }

        // 3. finish: lane c gathers its channel's accumulators
        uint8_t result = 0;
        if (METHOD == OFXCV_INPAINT_TELEA) {
            int c = lane < CN ? lane : 0;
            float Ia = __shfl_sync(0xffffffffu, acc, c * 4 + 0);
            float Jx = __shfl_sync(0xffffffffu, acc, c * 4 + 1);
            float Jy = __shfl_sync(0xffffffffu, acc, c * 4 + 2);
            float s = __shfl_sync(0xffffffffu, acc, c * 4 + 3);
            float sat = Ia / s + (Jx + Jy) / (sqrtf(Jx * Jx + Jy * Jy) + 1.0e-20f) + 0.5f;
            int iv = __double2int_rn((double)sat);
            result = (uint8_t)(iv < 0 ? 0 : iv > 255 ? 255 : iv);
        } else {
            int c = lane < CN ? lane : 0;
            float Ia = __shfl_sync(0xffffffffu, acc, c * 2 + 0);
            float s = __shfl_sync(0xffffffffu, acc, c * 2 + 1);
            int iv = __double2int_rn((double)Ia / s);
            result = (uint8_t)(iv < 0 ? 0 : iv > 255 ? 255 : iv);
        }
        if (lane < CN) {
            volatile uint8_t* o = out + (size_t)(i - 1) * ostride + (size_t)(j - 1) * CN + lane;
            *o = result;
        }
        __syncwarp();
        __threadfence();
        if (lane == 0) vdone[id] = 1;
    }
}

// fidx[n] = position of hole pixel n in the fill order, -1 for every other map position: "known at the fill time of
// ticket tk" is then fidx[n] < tk and "earlier-filled hole" is 0 <= fidx[n] < tk -- one load instead of hole + cnt
__global__ void __launch_bounds__(256) ip_fill_index(const uint8_t* __restrict__ hole, const uint32_t* __restrict__ cnt,
                                                     uint32_t cnt_base, int32_t* __restrict__ fidx, IpGeom g)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= g.np) return;
    // a hole pixel the march never reached (cnt == CNT_NONE) is never filled: it stays "not known" for everyone
    fidx[id] = hole[id] ? (cnt[id] == CNT_NONE ? 0x7fffffff : (int32_t)(cnt[id] - cnt_base)) : -1;
}

// ---- Stage B, staged: the (2r+3)^2 neighbourhood of the pixel lives in shared memory --------------------------------
// ip_fill above issues ~1700 dependent L2 loads per pixel from inside the tap arithmetic, and the fill order is a
// dependency chain thousands of pixels long (a pixel needs every earlier-filled hole pixel of its box), so the kernel
// runs at chain length x per-pixel latency.  Here a warp (a) loads the fill index of the whole box in one round --
// "known at this pixel's fill time" is fidx < tk --, (b) loads the colours (+T for Telea) of the box in one more round
// into shared memory with the known flag packed beside them, (c) evaluates the taps from shared memory only, with the
// tap / box geometry of each lane computed once per warp and constant offsets for pixels away from the image border
// (the per-pixel instruction count IS the chain latency: one warp executes it serially), and (d) sums the accumulators
// in tap order from registers.  Same arithmetic as ip_fill.  Two schedulers:
//   READYQ = false: tickets in fill order, a warp spins on the done flags of the earlier-filled holes of its box;
//   READYQ = true : dependency counters + a ready queue -- a warp only ever waits for a queue slot, so independent
//                   dependency chains advance in parallel even when the fill order lays them out one after the other
//                   (scratches, several blobs); costs ~2 us more per chain step.
// Used for radius <= IP2_MAXR (the plugin's range is [1, 10]).
constexpr int IP2_MAXR = 10;
constexpr int IP2_NPOS = 4;  // box positions per lane kept in flight / tabulated (covers r <= 4: 121 positions)
constexpr int IP2_NTAP = 3;  // taps per lane tabulated (covers r <= 4: 81 taps)

template <int METHOD, int CN, bool READYQ>
__global__ void __launch_bounds__(IP_WARPS * 32, READYQ ? 2 : 4)  // in-order tickets: the resident-warp count is the window over the chains
ip_fill_staged(const uint32_t* __restrict__ order, unsigned nfill, const int32_t* __restrict__ fidx, const float* __restrict__ t,
               uint8_t* out, ptrdiff_t ostride, uint8_t* done, int32_t* dep, int32_t* rq, unsigned* __restrict__ ticket,
               unsigned* __restrict__ rtail, uint32_t* pub, const uint32_t* __restrict__ orig, int range, IpGeom g)
{
    constexpr int NACC = METHOD == OFXCV_INPAINT_TELEA ? 4 * CN : 2 * CN;
    extern __shared__ __align__(16) unsigned char ip2_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int ec = g.ec, er = g.er;
    const int side = 2 * range + 1, ntaps = side * side;
    const int bside = 2 * range + 3, nbox = bside * bside;
    // per warp: packed colour+known word and T of every box position, then the term transpose buffer
    const int nbox_pad = (nbox + 3) & ~3;
    const size_t per_warp = (size_t)nbox_pad * 8 + 32 * (IP_MAXACC + 1) * 4;
    uint32_t* s_px = reinterpret_cast<uint32_t*>(ip2_smem + per_warp * wid);
    float* s_t = reinterpret_cast<float*>(s_px + nbox_pad);
    float(*s_term)[IP_MAXACC + 1] = reinterpret_cast<float(*)[IP_MAXACC + 1]>(s_t + nbox_pad);
    volatile uint8_t* vdone = done;

    // geometry of this lane's box positions and taps: the same for every pixel
    const bool tabulated = nbox <= 32 * IP2_NPOS && ntaps <= 32 * IP2_NTAP;
    int pk[IP2_NPOS], pl[IP2_NPOS];  // box position -> offset from the box origin (row, col); row < 0 = none
#pragma unroll
    for (int q = 0; q < IP2_NPOS; q++) {
        const int b = q * 32 + lane;
        pk[q] = b < nbox ? b / bside : -1;
        pl[q] = b < nbox ? b - (b / bside) * bside : 0;
    }
    int tdk[IP2_NTAP], tdl[IP2_NTAP];  // tap -> offset from the pixel; tdk = INT_MIN/2 = no tap / outside the disc
#pragma unroll
    for (int u = 0; u < IP2_NTAP; u++) {
        const int tp = u * 32 + lane;
        const int dk = tp / side - range, dl = tp % side - range;
        const bool in = tp < ntaps && dk * dk + dl * dl <= range * range;
        tdk[u] = in ? dk : -(1 << 20);
        tdl[u] = dl;
    }

    for (;;) {
        unsigned tk;
        int id;
        if (!READYQ) {
            // a ticket is taken only when the warp is free: a ticket parked on a busy warp (tried: tickets one pixel ahead)
            // delays the whole chain behind it
            unsigned v = 0;
            if (lane == 0) v = atomicAdd(ticket, 1u);
            tk = __shfl_sync(0xffffffffu, v, 0);
            if (tk >= nfill) break;
            id = (int)order[tk];
        } else {
            unsigned my = 0;
            if (lane == 0) my = atomicAdd(ticket, 1u);
            my = __shfl_sync(0xffffffffu, my, 0);
            if (my >= nfill) break;
            int got = 0;
            if (lane == 0) {
                volatile int32_t* slot = rq + my;
                while (*slot < 0) { }
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(got) : "l"(rq + my) : "memory");
            }
            id = __shfl_sync(0xffffffffu, got, 0);  // the shuffle orders every lane after the acquire
            tk = (unsigned)__ldg(fidx + id);
        }
        const int i = id / ec, j = id - i * ec;
        const int k0 = i - range - 1, l0 = j - range - 1;  // map coordinates of box position (0, 0)

        // (a) fill index of every box position in one round (SPIN: wait for the earlier-filled holes among them)
        auto stage_a = [&](int b, int k, int l, int& fi, int& n) {
            fi = -1;
            n = -1;
            if (k >= 0 && l >= 0 && k < er && l < ec) {
                n = k * ec + l;
                fi = __ldg(fidx + n);
            }
        };
        auto stage_a2 = [&](int b, int fi, int n) {
            if (!READYQ && fi >= 0 && fi < (int)tk) {
                // poll relaxed (an acquire load drags a CCTL.IVALL = L1 invalidate into every iteration), then ONE
                // acquire load: what the finished pixel's warp wrote before its release is visible after it
                while (vdone[n] == 0) { }
                unsigned d;
                asm volatile("ld.acquire.gpu.global.u8 %0, [%1];" : "=r"(d) : "l"(done + n) : "memory");
            }
            s_px[b] = (fi < (int)tk ? 1u : 0u) << 24;
        };
        // (b) colours (+T) of the box in one round
        auto stage_b = [&](int k, int l, uint32_t& px, float& tv, bool colour = true) {
            px = 0;
            tv = 0.f;
            if (colour && k >= 1 && l >= 1 && k <= g.H && l <= g.W) {
                // a hole filled LATER than this pixel is not known, but a tap on the image's border ring reads neighbours
                // whose flags it did not test (the CPU code's clamped indices): what the CPU sees there is the source colour,
                // and `out` may already hold what a warp running ahead has written
                const int f = __ldg(fidx + k * ec + l);
                if (f > (int)tk && f != 0x7fffffff) {
                    px = __ldg(orig + f);
                } else {
                    const uint8_t* o = out + (size_t)(k - 1) * ostride + (size_t)(l - 1) * CN;
#pragma unroll
                    for (int c = 0; c < CN; c++) px |= (uint32_t)ld_u8_cg(o + c) << (8 * c);
                }
            }
            if (METHOD == OFXCV_INPAINT_TELEA && k >= 0 && l >= 0 && k < er && l < ec) tv = __ldg(t + k * ec + l);
        };
        if (tabulated && !READYQ) {
            // In-order tickets, small boxes: an earlier-filled hole pixel publishes ONE word, its colour with a "done" bit
            // (pub[fill index]).  Flag and data travel together, so the dependent warp needs neither an acquire nor a second
            // round trip for the colour, and the finishing warp no release fence; everything that does not depend on a
            // pending pixel is loaded BEFORE the wait, so the chain step is: poll -> stage -> taps -> publish.
            int fi[IP2_NPOS], nn[IP2_NPOS];
            uint32_t px[IP2_NPOS];
            float tv[IP2_NPOS];
#pragma unroll
            for (int q = 0; q < IP2_NPOS; q++)
                if (pk[q] >= 0) stage_a(q * 32 + lane, k0 + pk[q], l0 + pl[q], fi[q], nn[q]);
#pragma unroll
            for (int q = 0; q < IP2_NPOS; q++)
                if (pk[q] >= 0) stage_b(k0 + pk[q], l0 + pl[q], px[q], tv[q], !(fi[q] >= 0 && fi[q] < (int)tk));
            // first look at all of this lane's dependencies at once (one round trip when they are already there, the
            // usual case), then spin only on the ones still missing
            uint32_t pv[IP2_NPOS];
#pragma unroll
            for (int q = 0; q < IP2_NPOS; q++) {
                pv[q] = 0x80000000u;
                if (pk[q] >= 0 && fi[q] >= 0 && fi[q] < (int)tk) pv[q] = *((volatile uint32_t*)pub + fi[q]);
            }
#pragma unroll
            for (int q = 0; q < IP2_NPOS; q++)
                if (pk[q] >= 0) {
                    if (fi[q] >= 0 && fi[q] < (int)tk) {
                        volatile uint32_t* w = pub + fi[q];
                        while ((pv[q] & 0x80000000u) == 0) pv[q] = *w;
                        px[q] = pv[q] & 0x00ffffffu;
                    }
                    s_px[q * 32 + lane] = ((fi[q] < (int)tk ? 1u : 0u) << 24) | px[q];
                    s_t[q * 32 + lane] = tv[q];
                }
        } else if (tabulated) {
            int fi[IP2_NPOS], nn[IP2_NPOS];
#pragma unroll
            for (int q = 0; q < IP2_NPOS; q++)
                if (pk[q] >= 0) stage_a(q * 32 + lane, k0 + pk[q], l0 + pl[q], fi[q], nn[q]);
#pragma unroll
            for (int q = 0; q < IP2_NPOS; q++)
                if (pk[q] >= 0) stage_a2(q * 32 + lane, fi[q], nn[q]);
            __syncwarp();  // orders every lane's loads below after the acquire loads of the lanes that waited
            uint32_t px[IP2_NPOS];
            float tv[IP2_NPOS];
#pragma unroll
            for (int q = 0; q < IP2_NPOS; q++)
                if (pk[q] >= 0) stage_b(k0 + pk[q], l0 + pl[q], px[q], tv[q]);
#pragma unroll
            for (int q = 0; q < IP2_NPOS; q++)
                if (pk[q] >= 0) {
                    s_px[q * 32 + lane] |= px[q];
                    s_t[q * 32 + lane] = tv[q];
                }
        } else {
            for (int b = lane; b < nbox; b += 32) {
                const int bk = b / bside, bl = b - bk * bside;
                int fi, n;
                stage_a(b, k0 + bk, l0 + bl, fi, n);
                stage_a2(b, fi, n);
            }
            __syncwarp();
            for (int b = lane; b < nbox; b += 32) {
                const int bk = b / bside, bl = b - bk * bside;
                uint32_t px;
                float tv;
                stage_b(k0 + bk, l0 + bl, px, tv);
                s_px[b] |= px;
                s_t[b] = tv;
            }
        }
        __syncwarp();

        // box accessors: map position (k, l) / image position (r, c) = map (r+1, c+1)
        auto bidx = [&](int k, int l) -> int { return (k - k0) * bside + (l - l0); };
        auto known = [&](int k, int l) -> bool { return (s_px[bidx(k, l)] >> 24) != 0; };
#define OUTP(r, c, ch) (int)((s_px[bidx((r) + 1, (c) + 1)] >> (8 * (ch))) & 0xffu)
#define TV(k, l) s_t[bidx((k), (l))]
// the same through a box index (pixels away from the image border: every stencil index is a constant offset)
#define KNOWN_AT(x) ((s_px[(x)] >> 24) != 0)
#define OUT_AT(x, ch) (int)((s_px[(x)] >> (8 * (ch))) & 0xffu)

        float gTx = 0.f, gTy = 0.f, ti = 0.f;
        if (METHOD == OFXCV_INPAINT_TELEA) {
            ti = TV(i, j);
            if (known(i, j + 1)) {
                if (known(i, j - 1)) gTx = (float)(TV(i, j + 1) - TV(i, j - 1)) * 0.5f;
                else gTx = (float)(TV(i, j + 1) - ti);
            } else {
                if (known(i, j - 1)) gTx = (float)(ti - TV(i, j - 1));
                else gTx = 0;
            }
            if (known(i + 1, j)) {
                if (known(i - 1, j)) gTy = (float)(TV(i + 1, j) - TV(i - 1, j)) * 0.5f;
                else gTy = (float)(TV(i + 1, j) - ti);
            } else {
                if (known(i - 1, j)) gTy = (float)(ti - TV(i - 1, j));
                else gTy = 0;
            }
        }

        float acc = 0.f;  // lane a < NACC owns accumulator a
        if (METHOD == OFXCV_INPAINT_TELEA) { if ((lane & 3) == 3) acc = 1.0e-20f; }
        else { if ((lane & 1) == 1) acc = 1.0e-20f; }

        // every tap's whole stencil inside the image without clamping?
        const bool interior = tabulated && i - range >= 2 && i + range <= g.H - 1 && j - range >= 2 && j + range <= g.W - 1;
        const int nrounds = (ntaps + 31) >> 5;
        for (int u = 0; u < nrounds; u++) {
            bool valid = false;
            float term[IP_MAXACC];
#pragma unroll
            for (int a = 0; a < IP_MAXACC; a++) term[a] = 0.f;
            if (interior) {
                // u < IP2_NTAP here; the tables are indexed with compile-time constants below
                int dk = -(1 << 20), dl = 0;
#pragma unroll
                for (int uu = 0; uu < IP2_NTAP; uu++)
                    if (uu == u) { dk = tdk[uu]; dl = tdl[uu]; }
                if (dk > -(1 << 19)) {
                    const int x = (dk + range + 1) * bside + (dl + range + 1);  // box index of the tap
                    if (KNOWN_AT(x)) {
                        valid = true;
                        const bool fr = KNOWN_AT(x + 1), fl = KNOWN_AT(x - 1), fd = KNOWN_AT(x + bside), fu = KNOWN_AT(x - bside);
                        if (METHOD == OFXCV_INPAINT_TELEA) {
                            float ry = (float)(-dk), rx = (float)(-dl);
                            float vl = rx * rx + ry * ry;
                            float dst = (float)(1. / (vl * sqrt((double)vl)));
                            float lev = (float)(1. / (1 + (double)fabsf(s_t[x] - ti)));  // f32 difference, f64 sum (C fabs)
                            float dir = rx * gTx + ry * gTy;
                            if (fabs(dir) <= 0.01) dir = 0.000001f;
                            float w = (float)fabs(dst * lev * dir);
#pragma unroll
                            for (int c = 0; c < CN; c++) {
                                float gIx, gIy;
                                if (fr) {
                                    if (fl) gIx = (float)(OUT_AT(x + 1, c) - OUT_AT(x - 1, c)) * 2.0f;
                                    else gIx = (float)(OUT_AT(x + 1, c) - OUT_AT(x, c));
                                } else {
                                    if (fl) gIx = (float)(OUT_AT(x, c) - OUT_AT(x - 1, c));
                                    else gIx = 0;
                                }
                                if (fd) {
                                    if (fu) gIy = (float)(OUT_AT(x + bside, c) - OUT_AT(x - bside, c)) * 2.0f;
                                    else gIy = (float)(OUT_AT(x + bside, c) - OUT_AT(x, c));
                                } else {
                                    if (fu) gIy = (float)(OUT_AT(x, c) - OUT_AT(x - bside, c));
                                    else gIy = 0;
                                }
                                term[c * 4 + 0] = w * (float)OUT_AT(x, c);
                                term[c * 4 + 1] = -(w * (gIx * rx));
                                term[c * 4 + 2] = -(w * (gIy * ry));
                                term[c * 4 + 3] = w;
                            }
                        } else {
                            float ry = (float)dk, rx = (float)dl;
                            float vl = rx * rx + ry * ry;
                            float dst = 1 / (vl * vl + 1);
#pragma unroll
                            for (int c = 0; c < CN; c++) {
                                float gIx, gIy;
                                if (fd) {
                                    if (fu) gIx = (float)(abs(OUT_AT(x + bside, c) - OUT_AT(x, c)) + abs(OUT_AT(x, c) - OUT_AT(x - bside, c)));
                                    else gIx = (float)(abs(OUT_AT(x + bside, c) - OUT_AT(x, c))) * 2.0f;
                                } else {
                                    if (fu) gIx = (float)(abs(OUT_AT(x, c) - OUT_AT(x - bside, c))) * 2.0f;
                                    else gIx = 0;
                                }
                                if (fr) {
                                    if (fl) gIy = (float)(abs(OUT_AT(x + 1, c) - OUT_AT(x, c)) + abs(OUT_AT(x, c) - OUT_AT(x - 1, c)));
                                    else gIy = (float)(abs(OUT_AT(x + 1, c) - OUT_AT(x, c))) * 2.0f;
                                } else {
                                    if (fl) gIy = (float)(abs(OUT_AT(x, c) - OUT_AT(x - 1, c))) * 2.0f;
                                    else gIy = 0;
                                }
                                gIx = -gIx;
                                float dir = rx * gIx + ry * gIy;
                                if (fabs(dir) <= 0.01) dir = 0.000001f;
                                else dir = fabsf((rx * gIx + ry * gIy) / sqrtf(vl * (gIx * gIx + gIy * gIy)));
                                float w = dst * dir;
                                term[c * 2 + 0] = w * (float)OUT_AT(x, c);
                                term[c * 2 + 1] = w;
                            }
                        }
                    }
                }
            } else {
                const int tp = u * 32 + lane;
                if (tp < ntaps) {
                    const int k = i - range + tp / side, l = j - range + tp % side;
                    if (k > 0 && l > 0 && k < er - 1 && l < ec - 1) {
                        if (known(k, l) && (l - j) * (l - j) + (k - i) * (k - i) <= range * range) {
                            valid = true;
                            const int km = k - 1 + (k == 1), kp = k - 1 - (k == er - 2);
                            const int lm = l - 1 + (l == 1), lp = l - 1 - (l == ec - 2);
                            const bool fr = known(k, l + 1), fl = known(k, l - 1), fd = known(k + 1, l), fu = known(k - 1, l);
                            if (METHOD == OFXCV_INPAINT_TELEA) {
                                float ry = (float)(i - k), rx = (float)(j - l);
                                float vl = rx * rx + ry * ry;
                                float dst = (float)(1. / (vl * sqrt((double)vl)));
                                float lev = (float)(1. / (1 + (double)fabsf(TV(k, l) - ti)));  // f32 difference, f64 sum (C fabs)
                                float dir = rx * gTx + ry * gTy;
                                if (fabs(dir) <= 0.01) dir = 0.000001f;
                                float w = (float)fabs(dst * lev * dir);
#pragma unroll
                                for (int c = 0; c < CN; c++) {
                                    float gIx, gIy;
                                    if (fr) {
                                        if (fl) gIx = (float)(OUTP(km, lp + 1, c) - OUTP(km, lm - 1, c)) * 2.0f;
                                        else gIx = (float)(OUTP(km, lp + 1, c) - OUTP(km, lm, c));
                                    } else {
                                        if (fl) gIx = (float)(OUTP(km, lp, c) - OUTP(km, lm - 1, c));
                                        else gIx = 0;
                                    }
                                    if (fd) {
                                        if (fu) gIy = (float)(OUTP(kp + 1, lm, c) - OUTP(km - 1, lm, c)) * 2.0f;
                                        else gIy = (float)(OUTP(kp + 1, lm, c) - OUTP(km, lm, c));
                                    } else {
                                        if (fu) gIy = (float)(OUTP(kp, lm, c) - OUTP(km - 1, lm, c));
                                        else gIy = 0;
                                    }
                                    term[c * 4 + 0] = w * (float)OUTP(k - 1, l - 1, c);
                                    term[c * 4 + 1] = -(w * (gIx * rx));
                                    term[c * 4 + 2] = -(w * (gIy * ry));
                                    term[c * 4 + 3] = w;
                                }
                            } else {
                                float ry = (float)(k - i), rx = (float)(l - j);
                                float vl = rx * rx + ry * ry;
                                float dst = 1 / (vl * vl + 1);
#pragma unroll
                                for (int c = 0; c < CN; c++) {
                                    float gIx, gIy;
                                    if (fd) {
                                        if (fu) gIx = (float)(abs(OUTP(kp + 1, lm, c) - OUTP(kp, lm, c)) + abs(OUTP(kp, lm, c) - OUTP(km - 1, lm, c)));
                                        else gIx = (float)(abs(OUTP(kp + 1, lm, c) - OUTP(kp, lm, c))) * 2.0f;
                                    } else {
                                        if (fu) gIx = (float)(abs(OUTP(kp, lm, c) - OUTP(km - 1, lm, c))) * 2.0f;
                                        else gIx = 0;
                                    }
                                    if (fr) {
                                        if (fl) gIy = (float)(abs(OUTP(km, lp + 1, c) - OUTP(km, lm, c)) + abs(OUTP(km, lm, c) - OUTP(km, lm - 1, c)));
                                        else gIy = (float)(abs(OUTP(km, lp + 1, c) - OUTP(km, lm, c))) * 2.0f;
                                    } else {
                                        if (fl) gIy = (float)(abs(OUTP(km, lm, c) - OUTP(km, lm - 1, c))) * 2.0f;
                                        else gIy = 0;
                                    }
                                    gIx = -gIx;
                                    float dir = rx * gIx + ry * gIy;
                                    if (fabs(dir) <= 0.01) dir = 0.000001f;
                                    else dir = fabsf((rx * gIx + ry * gIy) / sqrtf(vl * (gIx * gIx + gIy * gIy)));
                                    float w = dst * dir;
                                    term[c * 2 + 0] = w * (float)OUTP(k - 1, l - 1, c);
                                    term[c * 2 + 1] = w;
                                }
                            }
                        }
                    }
                }
            }
            const unsigned vmask = __ballot_sync(0xffffffffu, valid);
#pragma unroll
            for (int a = 0; a < NACC; a++) s_term[lane][a] = valid ? term[a] : 0.f;
            __syncwarp();
            if (lane < NACC) {
                // the 32 terms of this accumulator into registers (independent loads), then the ordered sum
                float v[32];
#pragma unroll
                for (int q = 0; q < 32; q++) v[q] = s_term[q][lane];
#pragma unroll
                for (int q = 0; q < 32; q++)
                    if ((vmask >> q) & 1u) acc = acc + v[q];
            }
            __syncwarp();
        }
#undef OUTP
#undef TV
#undef KNOWN_AT
#undef OUT_AT

        // finish: lane c gathers its channel's accumulators
        uint8_t result = 0;
        if (METHOD == OFXCV_INPAINT_TELEA) {
            int c = lane < CN ? lane : 0;
            float Ia = __shfl_sync(0xffffffffu, acc, c * 4 + 0);
            float Jx = __shfl_sync(0xffffffffu, acc, c * 4 + 1);
            float Jy = __shfl_sync(0xffffffffu, acc, c * 4 + 2);
            float s = __shfl_sync(0xffffffffu, acc, c * 4 + 3);
            float sat = Ia / s + (Jx + Jy) / (sqrtf(Jx * Jx + Jy * Jy) + 1.0e-20f) + 0.5f;
            int iv = __double2int_rn((double)sat);
            result = (uint8_t)(iv < 0 ? 0 : iv > 255 ? 255 : iv);
        } else {
            int c = lane < CN ? lane : 0;
            float Ia = __shfl_sync(0xffffffffu, acc, c * 2 + 0);
            float s = __shfl_sync(0xffffffffu, acc, c * 2 + 1);
            int iv = __double2int_rn((double)Ia / s);
            result = (uint8_t)(iv < 0 ? 0 : iv > 255 ? 255 : iv);
        }
        if (!READYQ && tabulated) {  // colour + done bit in one word: what the dependents poll
            uint32_t word = 0x80000000u;
#pragma unroll
            for (int c = 0; c < CN; c++) word |= (uint32_t)__shfl_sync(0xffffffffu, (unsigned)result, c) << (8 * c);
            if (lane == 0) *((volatile uint32_t*)pub + tk) = word;
        }
        if (lane < CN) {
            volatile uint8_t* o = out + (size_t)(i - 1) * ostride + (size_t)(j - 1) * CN + lane;
            *o = result;
        }
        __syncwarp();
        if (!READYQ) {
            if (!tabulated && lane == 0) {  // release: the colour bytes stored by lanes 0..CN-1 (ordered by the barrier above) before the flag
                asm volatile("st.release.gpu.global.u8 [%0], %1;" ::"l"(done + id), "r"(1u) : "memory");
            }
        } else {
            // notify: every later-filled hole pixel whose box contains this pixel loses one dependency; whoever takes the
            // last one publishes it.  acq_rel read-modify-writes: the colour bytes stored above (ordered by the barrier) are
            // visible to the warp that later acquires the queue slot.
            for (int b = lane; b < nbox; b += 32) {
                const int bk = b / bside, bl = b - bk * bside;
                const int k = k0 + bk, l = l0 + bl;
                if (k < 0 || l < 0 || k >= er || l >= ec) continue;
                const int n = k * ec + l;
                const int fq = __ldg(fidx + n);
                if (fq > (int)tk && fq != 0x7fffffff) {
                    int old;
                    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(dep + fq), "r"(-1) : "memory");
                    if (old == 1) {
                        const unsigned slot = atomicAdd(rtail, 1u);
                        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(rq + slot), "r"(n) : "memory");
                    }
                }
            }
        }
    }
}

// source colour of every hole pixel by fill index (packed like the published words), taken before the fill starts
template <int CN>
__global__ void __launch_bounds__(256) ip_orig(const uint32_t* __restrict__ order, unsigned nfill, const uint8_t* __restrict__ out,
                                               ptrdiff_t ostride, uint32_t* __restrict__ orig, IpGeom g)
{
    const unsigned tk = blockIdx.x * blockDim.x + threadIdx.x;
    if (tk >= nfill) return;
    const int id = (int)order[tk];
    const int i = id / g.ec, j = id - i * g.ec;
    const uint8_t* o = out + (size_t)(i - 1) * ostride + (size_t)(j - 1) * CN;
    uint32_t v = 0;
#pragma unroll
    for (int c = 0; c < CN; c++) v |= (uint32_t)o[c] << (8 * c);
    orig[tk] = v;
}

// ---- Stage B, incremental (the default for radius <= 4 with the in-order scheduler) ----------------------------------
// ip_fill_staged waits for ALL earlier-filled pixels of the (2r+3)^2 box and only then evaluates its 49-81 taps: the chain
// step is poll + every tap + the ordered sum (~6200 cycles after the wait at radius 3), and the chain of an iid 10 % mask at
// 4K is ~2700 steps deep.  Two observations shorten both factors:
//   1. A pending pixel p only enters the taps at p and at p's 4-neighbours (a tap reads its own colour and those of its
//      4-neighbours; on the image's border ring the CPU code's clamped indices stay inside the 3x3 around the tap).  So the
//      only pending pixels that can hold a pixel up are those around a tap that CONTRIBUTES (inside the disc, known): the box
//      corners, the positions around unknown taps ... are struck from the wait list.  That alone cuts the chain of the iid
//      mask from ~2700 to ~900 steps (tools measured it on the CPU from the oracle's fill order) and the fill from 11.7 to
//      5.9 ms.
//   2. Everything that does not touch a pending pixel can be evaluated BEFORE the wait.  A warp stages what is there at first
//      look, keeps a warp-uniform bit mask P of the (relevant) box positions still pending, evaluates -- compacted over the
//      lanes, one tap per lane -- every tap with nothing pending around it into a per-tap term table in shared memory, THEN
//      spins on the pending words, and evaluates the taps they held back in one more round (wait_all = 1, the default; with
//      wait_all = 0 the taps are released arrival by arrival, which costs a pass over the tap arithmetic per arrival for
//      the same critical path: measured equal).  The accumulators are then summed from the table in the CPU's tap order,
//      one lane per accumulator, as before.  What follows the LAST arrival is one round over a few taps + the ordered sum +
//      the publication instead of all of it: 5.9 -> 4.9 ms (NS), 4.4 ms (Telea); NS 81 -> 182 frames/s, Telea 75 -> 140 at 4K.
// Same arithmetic, same expression forms, same order of the accumulator sums as ip_fill / ip_fill_staged: bit-identical.
// Three resident CTAs per SM (80 registers): four (64 registers, spills) measured 20 % slower, two the same.
// OFXCV_IP_FILL_INC=0 selects ip_fill_staged, OFXCV_IP_FILL_WAITALL=0 the arrival-by-arrival release.
template <int METHOD, int CN, int MINB>
__global__ void __launch_bounds__(IP_WARPS * 32, MINB)
ip_fill_inc(const uint32_t* __restrict__ order, unsigned nfill, const int32_t* __restrict__ fidx, const float* __restrict__ t,
            uint8_t* out, ptrdiff_t ostride, unsigned* __restrict__ ticket, uint32_t* pub, const uint32_t* __restrict__ orig, int range,
            int wait_all, IpGeom g)
{
    constexpr int NACC = METHOD == OFXCV_INPAINT_TELEA ? 4 * CN : 2 * CN;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char ip3_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int ec = g.ec, er = g.er;
    const int side = 2 * range + 1, ntaps = side * side;
    const int bside = 2 * range + 3, nbox = bside * bside;
    const int nbox_pad = (nbox + 3) & ~3;
    const int nrounds = (ntaps + 31) >> 5;  // <= IP2_NTAP
    // CTA: the 3x3 "affected by" masks of every tap (4 words over the box positions); per warp: packed colour+known word
    // and T of every box position, the term table [tap][accumulator], the compaction list
    uint4* s_aff = reinterpret_cast<uint4*>(ip3_smem);
    const size_t per_warp = (size_t)nbox_pad * 8 + (size_t)nrounds * 32 * (IP_MAXACC + 1) * 4 + (size_t)nrounds * 32 * 4;
    unsigned char* wbase = ip3_smem + (size_t)2 * IP2_NTAP * 32 * 16 + per_warp * wid;
    uint32_t* s_px = reinterpret_cast<uint32_t*>(wbase);
    float* s_t = reinterpret_cast<float*>(s_px + nbox_pad);
    float(*s_term)[IP_MAXACC + 1] = reinterpret_cast<float(*)[IP_MAXACC + 1]>(s_t + nbox_pad);
    uint32_t* s_list = reinterpret_cast<uint32_t*>(s_term + nrounds * 32);

    // geometry of this lane's box positions and taps: the same for every pixel
    int pk[IP2_NPOS], pl[IP2_NPOS];  // box position -> offset from the box origin (row, col); row < 0 = none
#pragma unroll
    for (int q = 0; q < IP2_NPOS; q++) {
        const int b = q * 32 + lane;
        pk[q] = b < nbox ? b / bside : -1;
        pl[q] = b < nbox ? b - (b / bside) * bside : 0;
    }
    int tdk[IP2_NTAP], tdl[IP2_NTAP];  // tap -> offset from the pixel; tdk = -2^20 = no tap / outside the disc
#pragma unroll
    for (int u = 0; u < IP2_NTAP; u++) {
        const int tp = u * 32 + lane;
        const int dk = tp / side - range, dl = tp % side - range;
        const bool in = tp < ntaps && dk * dk + dl * dl <= range * range;
        tdk[u] = in ? dk : -(1 << 20);
        tdl[u] = dl;
    }
    // table 0: the tap and its 4-neighbours (what a tap away from the image border reads); table 1: the whole 3x3 (the
    // clamped indices of a tap on the border ring stay inside it)
    for (int e = threadIdx.x; e < 2 * IP2_NTAP * 32; e += blockDim.x) {
        const int tp = e % (IP2_NTAP * 32), full3 = e / (IP2_NTAP * 32);
        unsigned m[4] = {0u, 0u, 0u, 0u};
        if (tp < ntaps) {
            const int x = (tp / side + 1) * bside + (tp % side + 1);  // box index of the tap
            for (int dr = -1; dr <= 1; dr++)
                for (int dc = -1; dc <= 1; dc++) {
                    if (!full3 && dr != 0 && dc != 0) continue;
                    const int p = x + dr * bside + dc;
#pragma unroll
                    for (int w = 0; w < 4; w++)
                        if ((p >> 5) == w) m[w] |= 1u << (p & 31);
                }
        }
        s_aff[e] = make_uint4(m[0], m[1], m[2], m[3]);
    }
    __syncthreads();

    for (;;) {
        unsigned tk = 0;
        if (lane == 0) tk = atomicAdd(ticket, 1u);
        tk = __shfl_sync(FULL, tk, 0);
        if (tk >= nfill) break;
        const int id = (int)order[tk];
        const int i = id / ec, j = id - i * ec;
        const int k0 = i - range - 1, l0 = j - range - 1;  // map coordinates of box position (0, 0)

        // first look: fill index of every box position, colours of what is not a dependency, T, the published words of
        // the dependencies
        int fi[IP2_NPOS];
        bool pend[IP2_NPOS];
        uint32_t px[IP2_NPOS];
        float tv[IP2_NPOS];
#pragma unroll
        for (int q = 0; q < IP2_NPOS; q++) {
            fi[q] = -1;
            if (pk[q] >= 0) {
                const int k = k0 + pk[q], l = l0 + pl[q];
                if (k >= 0 && l >= 0 && k < er && l < ec) fi[q] = __ldg(fidx + k * ec + l);
            }
        }
#pragma unroll
        for (int q = 0; q < IP2_NPOS; q++) {
            px[q] = 0;
            tv[q] = 0.f;
            pend[q] = false;
            if (pk[q] >= 0) {
                const int k = k0 + pk[q], l = l0 + pl[q];
                const bool dep = fi[q] >= 0 && fi[q] < (int)tk;
                if (dep) {
                    const uint32_t v = *((volatile uint32_t*)pub + fi[q]);
                    pend[q] = (v >> 31) == 0;
                    px[q] = v & 0x00ffffffu;
                } else if (fi[q] > (int)tk && fi[q] != 0x7fffffff) {
                    // a hole pixel filled LATER: not known, but a tap on the image's border ring reads neighbours whose flags it
                    // did not test (the clamped indices of the CPU code).  What the CPU sees there at this pixel's time is the
                    // source colour; `out` may already hold the colour a warp running ahead has written.
                    px[q] = __ldg(orig + fi[q]);
                } else if (k >= 1 && l >= 1 && k <= g.H && l <= g.W) {
                    const uint8_t* o = out + (size_t)(k - 1) * ostride + (size_t)(l - 1) * CN;
#pragma unroll
                    for (int c = 0; c < CN; c++) px[q] |= (uint32_t)ld_u8_cg(o + c) << (8 * c);
                }
                if (METHOD == OFXCV_INPAINT_TELEA && k >= 0 && l >= 0 && k < er && l < ec) tv[q] = __ldg(t + k * ec + l);
            }
        }
#pragma unroll
        for (int q = 0; q < IP2_NPOS; q++)
            if (pk[q] >= 0) {
                s_px[q * 32 + lane] = ((fi[q] < (int)tk ? 1u : 0u) << 24) | (pend[q] ? 0u : px[q]);
                s_t[q * 32 + lane] = tv[q];
            }
        __syncwarp();

        // box accessors: map position (k, l) / image position (r, c) = map (r+1, c+1)
        auto bidx = [&](int k, int l) -> int { return (k - k0) * bside + (l - l0); };
        auto known = [&](int k, int l) -> bool { return (s_px[bidx(k, l)] >> 24) != 0; };
#define OUTP(r, c, ch) (int)((s_px[bidx((r) + 1, (c) + 1)] >> (8 * (ch))) & 0xffu)
#define TV(k, l) s_t[bidx((k), (l))]

        float gTx = 0.f, gTy = 0.f, ti = 0.f;
        if (METHOD == OFXCV_INPAINT_TELEA) {  // the known flags and T never wait for anybody
            ti = TV(i, j);
            if (known(i, j + 1)) {
                if (known(i, j - 1)) gTx = (float)(TV(i, j + 1) - TV(i, j - 1)) * 0.5f;
                else gTx = (float)(TV(i, j + 1) - ti);
            } else {
                if (known(i, j - 1)) gTx = (float)(ti - TV(i, j - 1));
                else gTx = 0;
            }
            if (known(i + 1, j)) {
                if (known(i - 1, j)) gTy = (float)(TV(i + 1, j) - TV(i - 1, j)) * 0.5f;
                else gTy = (float)(TV(i + 1, j) - ti);
            } else {
                if (known(i - 1, j)) gTy = (float)(ti - TV(i - 1, j));
                else gTy = 0;
            }
        }

        // the taps that contribute (known, inside the disc, not on the map's border ring) and their masks in tap order
        bool todo[IP2_NTAP];
        unsigned vm[IP2_NTAP];
#pragma unroll
        for (int u = 0; u < IP2_NTAP; u++) {
            todo[u] = false;
            if (u < nrounds && tdk[u] > -(1 << 19)) {
                const int k = i + tdk[u], l = j + tdl[u];
                todo[u] = k > 0 && l > 0 && k < er - 1 && l < ec - 1 && known(k, l);
            }
            vm[u] = __ballot_sync(FULL, todo[u]);
        }
        // Only a pending pixel that one of these taps reads can hold this pixel up: the box corners, the positions around
        // taps that are themselves unknown ... are struck from the wait list (this alone cuts the dependency chain of an
        // iid 10 % mask at 4K from ~2700 to ~900 steps).  P = the pending positions that matter, one word per 32 positions.
        const bool interior = i - range >= 2 && i + range <= g.H - 1 && j - range >= 2 && j + range <= g.W - 1;
        const uint4* aff = s_aff + (interior ? 0 : IP2_NTAP * 32);
        unsigned P[IP2_NPOS];
        {
            unsigned rel[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int u = 0; u < IP2_NTAP; u++)
                if (todo[u]) {
                    const uint4 a = aff[u * 32 + lane];
                    rel[0] |= a.x; rel[1] |= a.y; rel[2] |= a.z; rel[3] |= a.w;
                }
#pragma unroll
            for (int q = 0; q < IP2_NPOS; q++) {
                const unsigned r = __reduce_or_sync(FULL, rel[q]);
                pend[q] = pend[q] && ((r >> lane) & 1u);
                P[q] = __ballot_sync(FULL, pend[q]);
            }
        }

        // spin until every pending word of this lane (and of the warp) has arrived, staging the colours
        auto spin_all = [&](bool left) {
            while (__any_sync(FULL, left)) {
                left = false;
#pragma unroll
                for (int q = 0; q < IP2_NPOS; q++)
                    if (pend[q]) {
                        const uint32_t v = *((volatile uint32_t*)pub + fi[q]);
                        if (v >> 31) {
                            s_px[q * 32 + lane] = (1u << 24) | (v & 0x00ffffffu);
                            pend[q] = false;
                        } else {
                            left = true;
                        }
                    }
            }
        };
        // Telea: the weight of a tap (two f64 divisions and an f64 square root) depends on T and the geometry only, never on
        // a colour: the first round computes it for EVERY contributing tap, also for those whose colours are still held back
        // (list entry bit 31 = "weight only"), and parks it in the spare column of the term table
        bool first_round = true;
        for (;;) {
            // hand every tap with nothing pending around it to a lane
            int total = 0;
#pragma unroll
            for (int u = 0; u < IP2_NTAP; u++) {
                bool r = false;
                if (todo[u]) {
                    const uint4 a = aff[u * 32 + lane];
                    r = ((a.x & P[0]) | (a.y & P[1]) | (a.z & P[2]) | (a.w & P[3])) == 0u;
                }
                const bool sub = r || (METHOD == OFXCV_INPAINT_TELEA && first_round && todo[u]);
                const unsigned m = __ballot_sync(FULL, sub);
                if (sub) {
                    s_list[total + __popc(m & lt)] = (uint32_t)(u * 32 + lane) | ((uint32_t)(tdk[u] + 64) << 8) | ((uint32_t)(tdl[u] + 64) << 16) |
                                                     (r ? 0u : 0x80000000u);
                    if (r) todo[u] = false;
                }
                total += __popc(m);
            }
            __syncwarp();
            for (int base = 0; base < total; base += 32) {
                if (base + lane < total) {
                    const uint32_t e = s_list[base + lane];
                    const int tp = (int)(e & 0xffu), dk = (int)((e >> 8) & 0xffu) - 64, dl = (int)((e >> 16) & 0xffu) - 64;
                    const bool weight_only = (e >> 31) != 0;
                    const int k = i + dk, l = j + dl;
                    float term[IP_MAXACC];
                    const int km = k - 1 + (k == 1), kp = k - 1 - (k == er - 2);
                    const int lm = l - 1 + (l == 1), lp = l - 1 - (l == ec - 2);
                    const bool fr = known(k, l + 1), fl = known(k, l - 1), fd = known(k + 1, l), fu = known(k - 1, l);
                    if (METHOD == OFXCV_INPAINT_TELEA) {
                        float ry = (float)(i - k), rx = (float)(j - l);
                        float w;
                        if (first_round) {
                            float vl = rx * rx + ry * ry;
                            float dst = (float)(1. / (vl * sqrt((double)vl)));
                            float lev = (float)(1. / (1 + (double)fabsf(TV(k, l) - ti)));  // f32 difference, f64 sum (C fabs)
                            float dir = rx * gTx + ry * gTy;
                            if (fabs(dir) <= 0.01) dir = 0.000001f;
                            w = (float)fabs(dst * lev * dir);
                            s_term[tp][IP_MAXACC] = w;
                        } else {
                            w = s_term[tp][IP_MAXACC];
                        }
                        if (!weight_only) {
#pragma unroll
                        for (int c = 0; c < CN; c++) {
                            float gIx, gIy;
                            if (fr) {
                                if (fl) gIx = (float)(OUTP(km, lp + 1, c) - OUTP(km, lm - 1, c)) * 2.0f;
                                else gIx = (float)(OUTP(km, lp + 1, c) - OUTP(km, lm, c));
                            } else {
                                if (fl) gIx = (float)(OUTP(km, lp, c) - OUTP(km, lm - 1, c));
                                else gIx = 0;
                            }
                            if (fd) {
                                if (fu) gIy = (float)(OUTP(kp + 1, lm, c) - OUTP(km - 1, lm, c)) * 2.0f;
                                else gIy = (float)(OUTP(kp + 1, lm, c) - OUTP(km, lm, c));
                            } else {
                                if (fu) gIy = (float)(OUTP(kp, lm, c) - OUTP(km - 1, lm, c));
                                else gIy = 0;
                            }
                            term[c * 4 + 0] = w * (float)OUTP(k - 1, l - 1, c);
                            term[c * 4 + 1] = -(w * (gIx * rx));
                            term[c * 4 + 2] = -(w * (gIy * ry));
                            term[c * 4 + 3] = w;
                        }
                        }
                    } else {
                        float ry = (float)(k - i), rx = (float)(l - j);
                        float vl = rx * rx + ry * ry;
                        float dst = 1 / (vl * vl + 1);
#pragma unroll
                        for (int c = 0; c < CN; c++) {
                            float gIx, gIy;
                            if (fd) {
                                if (fu) gIx = (float)(abs(OUTP(kp + 1, lm, c) - OUTP(kp, lm, c)) + abs(OUTP(kp, lm, c) - OUTP(km - 1, lm, c)));
                                else gIx = (float)(abs(OUTP(kp + 1, lm, c) - OUTP(kp, lm, c))) * 2.0f;
                            } else {
                                if (fu) gIx = (float)(abs(OUTP(kp, lm, c) - OUTP(km - 1, lm, c))) * 2.0f;
                                else gIx = 0;
                            }
                            if (fr) {
                                if (fl) gIy = (float)(abs(OUTP(km, lp + 1, c) - OUTP(km, lm, c)) + abs(OUTP(km, lm, c) - OUTP(km, lm - 1, c)));
                                else gIy = (float)(abs(OUTP(km, lp + 1, c) - OUTP(km, lm, c))) * 2.0f;
                            } else {
                                if (fl) gIy = (float)(abs(OUTP(km, lm, c) - OUTP(km, lm - 1, c))) * 2.0f;
                                else gIy = 0;
                            }
                            gIx = -gIx;
                            float dir = rx * gIx + ry * gIy;
                            if (fabs(dir) <= 0.01) dir = 0.000001f;
                            else dir = fabsf((rx * gIx + ry * gIy) / sqrtf(vl * (gIx * gIx + gIy * gIy)));
                            float w = dst * dir;
                            term[c * 2 + 0] = w * (float)OUTP(k - 1, l - 1, c);
                            term[c * 2 + 1] = w;
                        }
                    }
                    if (!weight_only) {
#pragma unroll
                        for (int a = 0; a < NACC; a++) s_term[tp][a] = term[a];
                    }
                }
            }
            __syncwarp();
            first_round = false;
            if ((P[0] | P[1] | P[2] | P[3]) == 0u) break;  // nothing was pending: that round was the last
            // poll the pending words until some have arrived, stage them, release the taps around them
            uint32_t pv[IP2_NPOS];
            bool got;
            do {
                got = false;
#pragma unroll
                for (int q = 0; q < IP2_NPOS; q++) {
                    pv[q] = 0;
                    if (pend[q]) pv[q] = *((volatile uint32_t*)pub + fi[q]);
                    got |= (pv[q] >> 31) != 0;
                }
            } while (!__any_sync(FULL, got));
            bool left = false;
#pragma unroll
            for (int q = 0; q < IP2_NPOS; q++) {
                if (pend[q] && (pv[q] >> 31)) {
                    s_px[q * 32 + lane] = (1u << 24) | (pv[q] & 0x00ffffffu);
                    pend[q] = false;
                }
                left |= pend[q];
            }
            // one more round only: every evaluation round costs the whole warp a pass over the tap arithmetic whatever the
            // number of taps in it, and the round after the LAST arrival is as long with 29 taps as with 3
            if (wait_all) spin_all(left);
#pragma unroll
            for (int q = 0; q < IP2_NPOS; q++) P[q] = __ballot_sync(FULL, pend[q]);
            __syncwarp();
        }
#undef OUTP
#undef TV

        // the accumulators, summed in tap order: lane a < NACC owns accumulator a
        float acc = 0.f;
        if (METHOD == OFXCV_INPAINT_TELEA) { if ((lane & 3) == 3) acc = 1.0e-20f; }
        else { if ((lane & 1) == 1) acc = 1.0e-20f; }
        if (lane < NACC) {
#pragma unroll
            for (int u = 0; u < IP2_NTAP; u++)
                if (u < nrounds) {
                    float v[32];
#pragma unroll
                    for (int q = 0; q < 32; q++) v[q] = s_term[u * 32 + q][lane];
#pragma unroll
                    for (int q = 0; q < 32; q++)
                        if ((vm[u] >> q) & 1u) acc = acc + v[q];
                }
        }
        __syncwarp();

        // finish: lane c gathers its channel's accumulators
        uint8_t result = 0;
        if (METHOD == OFXCV_INPAINT_TELEA) {
            int c = lane < CN ? lane : 0;
            float Ia = __shfl_sync(FULL, acc, c * 4 + 0);
            float Jx = __shfl_sync(FULL, acc, c * 4 + 1);
            float Jy = __shfl_sync(FULL, acc, c * 4 + 2);
            float s = __shfl_sync(FULL, acc, c * 4 + 3);
            float sat = Ia / s + (Jx + Jy) / (sqrtf(Jx * Jx + Jy * Jy) + 1.0e-20f) + 0.5f;
            int iv = __double2int_rn((double)sat);
            result = (uint8_t)(iv < 0 ? 0 : iv > 255 ? 255 : iv);
        } else {
            int c = lane < CN ? lane : 0;
            float Ia = __shfl_sync(FULL, acc, c * 2 + 0);
            float s = __shfl_sync(FULL, acc, c * 2 + 1);
            int iv = __double2int_rn((double)Ia / s);
            result = (uint8_t)(iv < 0 ? 0 : iv > 255 ? 255 : iv);
        }
        // colour + done bit in one word: what the dependents poll
        uint32_t word = 0x80000000u;
#pragma unroll
        for (int c = 0; c < CN; c++) word |= (uint32_t)__shfl_sync(FULL, (unsigned)result, c) << (8 * c);
        if (lane == 0) *((volatile uint32_t*)pub + tk) = word;
        if (lane < CN) out[(size_t)(i - 1) * ostride + (size_t)(j - 1) * CN + lane] = result;
        __syncwarp();
    }
}

// dependency counts of the ready-queue fill: dep[tk] = number of earlier-filled hole pixels in the (2r+3)^2 box of the
// pixel with fill index tk; pixels without any go straight into the ready queue.
// Also marks the tickets whose LATEST dependency is the ticket right before them (seq[tk]): a run of such tickets is a
// dependency chain laid out contiguously in the fill order, which in-order tickets can only walk one step at a time (the
// window of resident warps then covers one chain instead of many).
__global__ void __launch_bounds__(256) ip_deps(const uint32_t* __restrict__ order, unsigned nfill, const int32_t* __restrict__ fidx,
                                               int32_t* __restrict__ dep, int32_t* __restrict__ rq, unsigned* __restrict__ rtail,
                                               uint8_t* __restrict__ seq, int range, IpGeom g)
{
    const unsigned tk = blockIdx.x * blockDim.x + threadIdx.x;
    if (tk >= nfill) return;
    const int id = (int)order[tk];
    const int i = id / g.ec, j = id - i * g.ec;
    int count = 0, latest = -1;
    for (int k = max(i - range - 1, 0); k <= min(i + range + 1, g.er - 1); k++)
        for (int l = max(j - range - 1, 0); l <= min(j + range + 1, g.ec - 1); l++) {
            const int f = fidx[k * g.ec + l];
            if (f >= 0 && f < (int)tk) {
                count++;
                latest = max(latest, f);
            }
        }
    dep[tk] = count;
    if (count == 0) rq[atomicAdd(rtail, 1u)] = id;
    seq[tk] = latest >= 0 && latest == (int)tk - 1;
}

// number of tickets inside sequential runs and number of runs (a run starts where seq turns on)
__global__ void __launch_bounds__(256) ip_runs(const uint8_t* __restrict__ seq, unsigned nfill, unsigned* __restrict__ nseq,
                                               unsigned* __restrict__ nruns)
{
    const unsigned tk = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = tk < nfill && seq[tk];
    const bool start = in && (tk == 0 || !seq[tk - 1]);
    const unsigned mi = __ballot_sync(0xffffffffu, in), ms = __ballot_sync(0xffffffffu, start);
    if ((threadIdx.x & 31) == 0) {
        if (mi) atomicAdd(nseq, (unsigned)__popc(mi));
        if (ms) atomicAdd(nruns, (unsigned)__popc(ms));
    }
}

struct IpCounters {
    unsigned nholes, tau_bits, n_new, pending, ticket, pad[3];
};

}  // namespace

extern "C" {

size_t ofxcv_inpaint_workspace_bytes(int W, int H, int channels)
{
    (void)channels;
    size_t np = (size_t)(W + 2) * (H + 2);
    // hole, outreg, tmp, st, done (1 B), rnd (2 B), t, cnt, order, ids x2 (4 B), keys x2 (8 B), + sort temp
    return np * (5 + 2 + 4 * 5 + 16) + np * 8 + 4096;
}

int ofxcv_inpaint_u8(ofxcv_ctx* ctx, ofxcv_stream stream_, const uint8_t* img, ptrdiff_t img_stride, int channels,
                     const uint8_t* mask, ptrdiff_t mask_stride, uint8_t* out, ptrdiff_t out_stride, int W, int H, double radius,
                     int method)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    if (!img || !mask || !out || W < 2 || H < 2 || (channels != 1 && channels != 3) || img_stride < (ptrdiff_t)W * channels ||
        out_stride < (ptrdiff_t)W * channels || mask_stride < W || (method != OFXCV_INPAINT_NS && method != OFXCV_INPAINT_TELEA))
        return OFXCV_ERR_BAD_ARG;
    if ((size_t)(W + 2) * (H + 2) >= ((size_t)1 << 28)) return OFXCV_ERR_UNSUPPORTED;
    ofxcv_device_guard guard(ctx->device);
    cudaStream_t s = stream_ ? (cudaStream_t)stream_ : ctx->stream;
    int range = (int)nearbyint(radius);
    range = range < 1 ? 1 : range > 100 ? 100 : range;

    IpGeom g;
    g.W = W; g.H = H; g.er = H + 2; g.ec = W + 2; g.np = g.er * g.ec;
    const size_t np = (size_t)g.np;
    const size_t npa = (np + 15) & ~(size_t)15;  // the byte maps start on 16-byte boundaries (16-pixel passes over `st`)
    uint8_t* bytes = (uint8_t*)ofxcv_ws(ctx, WS_INP_A, npa * 5);
    uint16_t* rnd = (uint16_t*)ofxcv_ws(ctx, WS_INP_B, np * 2);
    float* t = (float*)ofxcv_ws(ctx, WS_INP_C, np * 4);
    uint32_t* u32s = (uint32_t*)ofxcv_ws(ctx, WS_INP_D, np * 4 * 4);
    unsigned long long* keys = (unsigned long long*)ofxcv_ws(ctx, WS_INP_E, np * 8 * 2);
    IpCounters* ctr = (IpCounters*)ofxcv_ws(ctx, WS_INP_F, sizeof(IpCounters));
    IpCounters* hctr = (IpCounters*)ofxcv_pin(ctx, 2, sizeof(IpCounters));
    if (!bytes || !rnd || !t || !u32s || !keys || !ctr || !hctr) return OFXCV_ERR_MEMORY;
    uint8_t *hole = bytes, *outreg = bytes + npa, *tmp = bytes + 2 * npa, *st = bytes + 3 * npa, *done = bytes + 4 * npa;
    uint32_t *cnt = u32s, *order = u32s + np, *ids = u32s + 2 * np, *ids_sorted = u32s + 3 * np;
    unsigned long long* keys_sorted = keys + np;
    size_t sort_tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp_bytes, keys, keys_sorted, ids, ids_sorted, g.np, 0, 64, s);
    void* sort_tmp = ofxcv_ws(ctx, WS_INP_G, sort_tmp_bytes);
    if (!sort_tmp) return OFXCV_ERR_MEMORY;

    const int nblk = ofxcv_div_up(g.np, 256);
    uint64_t launches0 = ctx->launches;
    int64_t batches = 0, rounds_total = 0;
    ofxcv_prof_scope ps_all(ctx, s, "ip_total", 0);
    OFXCV_CUDA(ctx, cudaMemsetAsync(ctr, 0, sizeof(IpCounters), s));
    ip_init<<<nblk, 256, 0, s>>>(mask, mask_stride, hole, t, cnt, rnd, done, &ctr->nholes, g);
    OFXCV_LAUNCH_CHECK(ctx);
    if (out != img) {
        if (img_stride >= (ptrdiff_t)W * channels && out_stride >= (ptrdiff_t)W * channels) {  // the driver's pitched D2D copy
            OFXCV_CUDA(ctx, cudaMemcpy2DAsync(out, (size_t)out_stride, img, (size_t)img_stride, (size_t)W * channels, H,
                                              cudaMemcpyDeviceToDevice, s));
        } else {  // bottom-up (negative stride) images
            ip_copy<<<dim3(ofxcv_div_up(W * channels, 256), H), 256, 0, s>>>(img, img_stride, out, out_stride, W * channels, H);
            OFXCV_LAUNCH_CHECK(ctx);
        }
    }
    OFXCV_CUDA(ctx, cudaMemcpyAsync(hctr, ctr, sizeof(IpCounters), cudaMemcpyDeviceToHost, s));
    OFXCV_CUDA(ctx, cudaStreamSynchronize(s));
    const unsigned nholes = hctr->nholes;
    ctx->inpaint_stats[0] = nholes;
    ctx->inpaint_stats[1] = ctx->inpaint_stats[2] = ctx->inpaint_stats[3] = 0;
    if (nholes == 0) return OFXCV_OK;

    if (method == OFXCV_INPAINT_TELEA) {
        ip_dilate_rows<<<nblk, 256, 0, s>>>(hole, tmp, range, g);
        OFXCV_LAUNCH_CHECK(ctx);
        ip_dilate_cols<<<nblk, 256, 0, s>>>(hole, tmp, outreg, range, g);
        OFXCV_LAUNCH_CHECK(ctx);
    }
    uint32_t nfilled = 0;
    static const bool wide_passes = !getenv("OFXCV_IP_PASSES_V1");
    static const bool chain_rounds = !getenv("OFXCV_IP_ROUNDS_V1");
    if (npa > np) OFXCV_CUDA(ctx, cudaMemsetAsync(st + np, ST_OUT, npa - np, s));  // padding of the state map: never in the heap
    for (int pass = (method == OFXCV_INPAINT_TELEA ? 1 : 0); pass >= 0; pass--) {
        const int outer = pass;  // Telea: pass 1 = outside band (negated afterwards), pass 0 = the hole itself
        ip_setup_pass<<<nblk, 256, 0, s>>>(hole, outreg, st, t, cnt, rnd, outer, g);
        OFXCV_LAUNCH_CHECK(ctx);
        uint32_t next_cnt = (uint32_t)g.np;  // reached pixels get counters above every band counter
        for (;;) {
            ofxcv_prof_scope ps(ctx, s, "ip_march_batch", pass);
            OFXCV_CUDA(ctx, cudaMemsetAsync(&ctr->tau_bits, 0xff, sizeof(unsigned), s));
            OFXCV_CUDA(ctx, cudaMemsetAsync(&ctr->n_new, 0, sizeof(unsigned), s));
            if (wide_passes) {
                const int n16 = (int)(npa / 16), n4 = (int)(npa / 4);
                ip_pass_a16<<<ofxcv_div_up(n16, 256), 256, 0, s>>>((uint4*)st, t, &ctr->tau_bits, n16);
                OFXCV_LAUNCH_CHECK(ctx);
                ip_pass_b16<<<ofxcv_div_up(n16, 256), 256, 0, s>>>((uint4*)st, t, &ctr->tau_bits, n16);
                OFXCV_LAUNCH_CHECK(ctx);
                ip_pass_c4<<<ofxcv_div_up(n4, 256), 256, 0, s>>>(st, t, cnt, keys, ids, &ctr->n_new, n4, g);
                OFXCV_LAUNCH_CHECK(ctx);
            } else {
                ip_pass_a<<<nblk, 256, 0, s>>>(st, t, &ctr->tau_bits, g);
                OFXCV_LAUNCH_CHECK(ctx);
                ip_pass_b<<<nblk, 256, 0, s>>>(st, t, &ctr->tau_bits, g);
                OFXCV_LAUNCH_CHECK(ctx);
                ip_pass_c<<<nblk, 256, 0, s>>>(st, t, cnt, keys, ids, &ctr->n_new, g);
                OFXCV_LAUNCH_CHECK(ctx);
            }
            OFXCV_CUDA(ctx, cudaMemcpyAsync(hctr, ctr, sizeof(IpCounters), cudaMemcpyDeviceToHost, s));
            OFXCV_CUDA(ctx, cudaStreamSynchronize(s));
            if (hctr->tau_bits == 0xffffffffu) break;  // heap empty
            batches++;
            const unsigned n_new = hctr->n_new;
            if (n_new == 0) continue;
            OFXCV_CUDA(ctx, cub::DeviceRadixSort::SortPairs(sort_tmp, sort_tmp_bytes, keys, keys_sorted, ids, ids_sorted, (int)n_new, 0, 64, s));
            ctx->launches += 4;
            const int nb2 = ofxcv_div_up((int)n_new, 256);
            ip_assign<<<nb2, 256, 0, s>>>(ids_sorted, n_new, next_cnt, st, cnt, outer ? nullptr : order, nfilled);
            OFXCV_LAUNCH_CHECK(ctx);
            next_cnt += n_new;
            if (!outer) nfilled += n_new;
            // T of the reached pixels: dependency chains followed inside one launch (round by round with OFXCV_IP_ROUNDS_V1)
            unsigned round = 1;
            if (chain_rounds) {
                OFXCV_CUDA(ctx, cudaMemsetAsync(&ctr->ticket, 0, sizeof(unsigned), s));
                ip_round_chain<<<nb2, 256, 0, s>>>(ids_sorted, n_new, st, t, cnt, rnd, &ctr->ticket, g);
                OFXCV_LAUNCH_CHECK(ctx);
                round = 2;
            } else
                for (;;) {
                    const int burst = round == 1 ? 4 : 16;
                    for (int b = 0; b < burst; b++, round++) {
                        if (round >= 65535) return OFXCV_ERR_UNSUPPORTED;
                        OFXCV_CUDA(ctx, cudaMemsetAsync(&ctr->pending, 0, sizeof(unsigned), s));
                        ip_round<<<nb2, 256, 0, s>>>(ids_sorted, n_new, st, t, cnt, rnd, round, &ctr->pending, g);
                        OFXCV_LAUNCH_CHECK(ctx);
                    }
                    OFXCV_CUDA(ctx, cudaMemcpyAsync(hctr, ctr, sizeof(IpCounters), cudaMemcpyDeviceToHost, s));
                    OFXCV_CUDA(ctx, cudaStreamSynchronize(s));
                    if (hctr->pending == 0) break;
                }
            rounds_total += round - 1;
        }
        if (outer) {
            ip_negate<<<nblk, 256, 0, s>>>(st, outreg, t, g);
            OFXCV_LAUNCH_CHECK(ctx);
        }
    }

    // Stage B
    if (nfilled > 0) {
        OFXCV_CUDA(ctx, cudaMemsetAsync(&ctr->ticket, 0, sizeof(unsigned), s));
        int blocks = ctx->num_sms * ctx->ip_fill_blocks_per_sm;
        int need = ofxcv_div_up((int)nfilled, IP_WARPS);
        if (blocks > need) blocks = need;
        ofxcv_prof_scope ps(ctx, s, "ip_fill", 0);
        ofxcv_time_begin(ctx, 1, s);
        const bool v2 = range <= IP2_MAXR && !getenv("OFXCV_IP_FILL_V1");
        int32_t* fidx = nullptr;
        if (v2) {
            fidx = (int32_t*)ofxcv_ws(ctx, WS_INP_H, np * 4);
            if (!fidx) return OFXCV_ERR_MEMORY;
            ip_fill_index<<<nblk, 256, 0, s>>>(hole, cnt, (uint32_t)g.np, fidx, g);
            OFXCV_LAUNCH_CHECK(ctx);
        }
        // scheduler: dependency counters + ready queue when the fill order lays the dependency chains out one after the
        // other (most pixels depend on one of the 64 tickets right before them), in-order tickets + flag spinning otherwise
        // (cheaper per chain step; what an iid mask wants).  OFXCV_IP_READYQ=0/1 forces one.
        int32_t* dep = (int32_t*)keys;  // the sort buffers are free once the march is over
        int32_t* rq = dep + np;
        uint32_t* pub = (uint32_t*)(rq + np);  // colour + done word per fill index (in-order scheduler)
        uint32_t* orig = pub + np;             // source colour per fill index (ip_fill_inc)
        OFXCV_CUDA(ctx, cudaMemsetAsync(pub, 0, (size_t)nfilled * 4, s));
        bool ready_queue = false;
        if (v2) {
            OFXCV_CUDA(ctx, cudaMemsetAsync(rq, 0xff, (size_t)nfilled * 4, s));
            OFXCV_CUDA(ctx, cudaMemsetAsync(&ctr->pending, 0, sizeof(unsigned), s));
            OFXCV_CUDA(ctx, cudaMemsetAsync(&ctr->n_new, 0, sizeof(unsigned), s));
            OFXCV_CUDA(ctx, cudaMemsetAsync(&ctr->pad[0], 0, sizeof(unsigned), s));
            ip_deps<<<ofxcv_div_up((int)nfilled, 256), 256, 0, s>>>(order, nfilled, fidx, dep, rq, &ctr->pending, tmp, range, g);
            OFXCV_LAUNCH_CHECK(ctx);
            ip_runs<<<ofxcv_div_up((int)nfilled, 256), 256, 0, s>>>(tmp, nfilled, &ctr->n_new, &ctr->pad[0]);
            OFXCV_LAUNCH_CHECK(ctx);
            OFXCV_CUDA(ctx, cudaMemcpyAsync(hctr, ctr, sizeof(IpCounters), cudaMemcpyDeviceToHost, s));
            OFXCV_CUDA(ctx, cudaStreamSynchronize(s));
            // several long chains laid out one after the other in the fill order (horizontal scratches, bars): the ready
            // queue advances them side by side at ~7.5 us per step (measured at 4K: 45 lines of 3000 px 24 ms against 94 ms
            // with in-order tickets).  Everything else -- iid masks, blobs, diagonal scratches, ONE line: chains interleaved
            // by the fill order, or a single chain -- runs 2-5x faster on the in-order incremental fill (~2 us per step).
            const unsigned nseq = hctr->n_new, nruns = hctr->pad[0];
            const char* rqenv = getenv("OFXCV_IP_READYQ");
            ready_queue = rqenv ? *rqenv == '1' : (nruns >= 4 && (double)nseq >= 256.0 * (double)nruns);
        }
        const int nbox_pad = ((2 * range + 3) * (2 * range + 3) + 3) & ~3;
        const size_t fill_smem = v2 ? ((size_t)nbox_pad * 8 + 32 * (IP_MAXACC + 1) * 4) * IP_WARPS : 0;
        // incremental fill (ip_fill_inc): in-order scheduler, tabulated geometry (radius <= 4)
        const int ntaps = (2 * range + 1) * (2 * range + 1), nrounds = (ntaps + 31) / 32;
        static const int inc_env = getenv("OFXCV_IP_FILL_INC") ? atoi(getenv("OFXCV_IP_FILL_INC")) : 1;
        static const int inc_wait_all = getenv("OFXCV_IP_FILL_WAITALL") ? atoi(getenv("OFXCV_IP_FILL_WAITALL")) : 1;
        const bool inc = v2 && !ready_queue && inc_env != 0 && nbox_pad <= 32 * IP2_NPOS && ntaps <= 32 * IP2_NTAP;
        const size_t inc_smem = (size_t)2 * IP2_NTAP * 32 * 16 +
                                ((size_t)nbox_pad * 8 + (size_t)nrounds * 32 * (IP_MAXACC + 1) * 4 + (size_t)nrounds * 32 * 4) * IP_WARPS;
#define IP_FILL_INC(M, C, B)                                                                                                       \
    do {                                                                                                                           \
        if (inc_smem > 48 * 1024)                                                                                                  \
            OFXCV_CUDA(ctx, cudaFuncSetAttribute(ip_fill_inc<M, C, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)inc_smem)); \
        ip_fill_inc<M, C, B><<<blocks, IP_WARPS * 32, inc_smem, s>>>(order, nfilled, fidx, t, out, out_stride, &ctr->ticket, pub,  \
                                                                     orig, range, inc_wait_all, g);                                \
    } while (0)
#define IP_FILL(M, C)                                                                                                              \
    do {                                                                                                                           \
        if (v2) ip_orig<C><<<ofxcv_div_up((int)nfilled, 256), 256, 0, s>>>(order, nfilled, out, out_stride, orig, g);              \
        if (inc) {                                                                                                                 \
            IP_FILL_INC(M, C, 3);                                                                                                  \
        } else if (v2) {                                                                                                              \
            if (fill_smem > 48 * 1024) {                                                                                           \
                OFXCV_CUDA(ctx, cudaFuncSetAttribute(ip_fill_staged<M, C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                                     (int)fill_smem));                                                             \
                OFXCV_CUDA(ctx, cudaFuncSetAttribute(ip_fill_staged<M, C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                                     (int)fill_smem));                                                             \
            }                                                                                                                      \
            if (ready_queue)                                                                                                       \
                ip_fill_staged<M, C, true><<<blocks, IP_WARPS * 32, fill_smem, s>>>(order, nfilled, fidx, t, out, out_stride,     \
                                                                                    done, dep, rq, &ctr->ticket, &ctr->pending,   \
                                                                                    pub, orig, range, g);                          \
            else                                                                                                                   \
                ip_fill_staged<M, C, false><<<blocks, IP_WARPS * 32, fill_smem, s>>>(order, nfilled, fidx, t, out, out_stride,    \
                                                                                     done, dep, rq, &ctr->ticket, &ctr->pending,  \
                                                                                     pub, orig, range, g);                         \
        } else {                                                                                                                   \
            ip_fill<M, C><<<blocks, IP_WARPS * 32, 0, s>>>(order, nfilled, (uint32_t)g.np, hole, cnt, t, out, out_stride, done,    \
                                                           &ctr->ticket, range, g);                                                \
        }                                                                                                                          \
    } while (0)
        if (method == OFXCV_INPAINT_TELEA) {
            if (channels == 3) IP_FILL(OFXCV_INPAINT_TELEA, 3);
            else IP_FILL(OFXCV_INPAINT_TELEA, 1);
        } else {
            if (channels == 3) IP_FILL(OFXCV_INPAINT_NS, 3);
            else IP_FILL(OFXCV_INPAINT_NS, 1);
        }
#undef IP_FILL
#undef IP_FILL_INC
        ofxcv_time_end(ctx, 1, s);
        OFXCV_LAUNCH_CHECK(ctx);
    }
    ctx->inpaint_stats[1] = batches;
    ctx->inpaint_stats[2] = rounds_total;
    ctx->inpaint_stats[3] = (int64_t)(ctx->launches - launches0);
    return OFXCV_OK;
}

void ofxcv_inpaint_set_fill_blocks(ofxcv_ctx* ctx, int blocks_per_sm)
{
    if (ctx) ctx->ip_fill_blocks_per_sm = blocks_per_sm < 1 ? 1 : blocks_per_sm > 8 ? 8 : blocks_per_sm;
}

int ofxcv_inpaint_debug_maps(ofxcv_ctx* ctx, int W, int H, float* t_host, int32_t* order_host)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    const size_t np = (size_t)(W + 2) * (H + 2);
    if (!ctx->ws[WS_INP_C].p || ctx->ws[WS_INP_C].cap < np * 4 || !ctx->ws[WS_INP_D].p) return OFXCV_ERR_BAD_ARG;
    ofxcv_device_guard guard(ctx->device);
    OFXCV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (t_host) OFXCV_CUDA(ctx, cudaMemcpy(t_host, ctx->ws[WS_INP_C].p, np * 4, cudaMemcpyDeviceToHost));
    if (order_host) {
        // cnt map -> fill order per image pixel (-1 where not a filled hole pixel)
        std::vector<uint32_t> cnt(np);
        std::vector<uint8_t> hole(np);
        OFXCV_CUDA(ctx, cudaMemcpy(cnt.data(), ctx->ws[WS_INP_D].p, np * 4, cudaMemcpyDeviceToHost));
        OFXCV_CUDA(ctx, cudaMemcpy(hole.data(), ctx->ws[WS_INP_A].p, np, cudaMemcpyDeviceToHost));
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                size_t id = (size_t)(y + 1) * (W + 2) + (x + 1);
                order_host[(size_t)y * W + x] = (hole[id] && cnt[id] != 0xffffffffu) ? (int32_t)(cnt[id] - (uint32_t)np) : -1;
            }
    }
    return OFXCV_OK;
}

int ofxcv_inpaint_last_stats(const ofxcv_ctx* ctx, int64_t stats[4])
{
    if (!ctx || !stats) return OFXCV_ERR_BAD_ARG;
    for (int i = 0; i < 4; i++) stats[i] = ctx->inpaint_stats[i];
    return OFXCV_OK;
}

int ofxcv_inpaint_u8_host(ofxcv_ctx* ctx, const uint8_t* img, ptrdiff_t img_stride, int channels, const uint8_t* mask,
                          ptrdiff_t mask_stride, uint8_t* out, ptrdiff_t out_stride, int W, int H, double radius, int method)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    if (!img || !mask || !out || W < 2 || H < 2 || (channels != 1 && channels != 3)) return OFXCV_ERR_BAD_ARG;
    ofxcv_device_guard guard(ctx->device);
    const size_t rb = (size_t)W * channels, nimg = rb * H, nmask = (size_t)W * H;
    uint8_t* hi = (uint8_t*)ofxcv_pin(ctx, 0, nimg + nmask);
    uint8_t* ho = (uint8_t*)ofxcv_pin(ctx, 1, nimg);
    uint8_t* di = (uint8_t*)ofxcv_ws(ctx, WS_STAGE_IN0, nimg);
    uint8_t* dm = (uint8_t*)ofxcv_ws(ctx, WS_STAGE_IN1, nmask);
    uint8_t* dout = (uint8_t*)ofxcv_ws(ctx, WS_STAGE_OUT, nimg);
    if (!hi || !ho || !di || !dm || !dout) return OFXCV_ERR_MEMORY;
    for (int y = 0; y < H; y++) {
        memcpy(hi + (size_t)y * rb, img + (size_t)y * img_stride, rb);
        memcpy(hi + nimg + (size_t)y * W, mask + (size_t)y * mask_stride, W);
    }
    cudaStream_t s = ctx->stream;
    OFXCV_CUDA(ctx, cudaMemcpyAsync(di, hi, nimg, cudaMemcpyHostToDevice, s));
    OFXCV_CUDA(ctx, cudaMemcpyAsync(dm, hi + nimg, nmask, cudaMemcpyHostToDevice, s));
    int st = ofxcv_inpaint_u8(ctx, s, di, (ptrdiff_t)rb, channels, dm, W, dout, (ptrdiff_t)rb, W, H, radius, method);
    if (st < 0) return st;
    OFXCV_CUDA(ctx, cudaMemcpyAsync(ho, dout, nimg, cudaMemcpyDeviceToHost, s));
    OFXCV_CUDA(ctx, cudaStreamSynchronize(s));
    for (int y = 0; y < H; y++) memcpy(out + (size_t)y * out_stride, ho + (size_t)y * rb, rb);
    return OFXCV_OK;
}

// A clip of independent frames, several in flight: worker k (own sub-context = own stream and workspaces, own host
// thread because the marching reads its batch counters back) takes frames k, k+K, ...; the persistent fill CTAs of
// the K frames share the SMs.  Blocking: waits for `stream`, returns when every frame is done.
static int ip_sequence(ofxcv_ctx* ctx, ofxcv_stream stream, bool host, const uint8_t* const* imgs, ptrdiff_t img_stride, int channels,
                       const uint8_t* const* masks, ptrdiff_t mask_stride, uint8_t* const* outs, ptrdiff_t out_stride, int W, int H,
                       int nframes, double radius, int method, int frames_in_flight)
{
    if (!ctx) return OFXCV_ERR_NO_DEVICE;
    if (!imgs || !masks || !outs || nframes < 0) return OFXCV_ERR_BAD_ARG;
    if (nframes == 0) return OFXCV_OK;
    for (int f = 0; f < nframes; f++)
        if (!imgs[f] || !masks[f] || !outs[f]) return OFXCV_ERR_BAD_ARG;
    ofxcv_device_guard guard(ctx->device);
    OFXCV_CUDA(ctx, cudaStreamSynchronize(stream ? (cudaStream_t)stream : ctx->stream));
    int K = frames_in_flight <= 0 ? 8 : frames_in_flight;
    K = std::min(std::min(K, 8), nframes);
    for (int k = 0; k < K; k++) {
        if (!ctx->sub[k]) ctx->sub[k] = ofxcv_create(ctx->device);
        if (!ctx->sub[k]) return OFXCV_ERR_MEMORY;
        ofxcv_inpaint_set_fill_blocks(ctx->sub[k], std::max(1, 8 / K));
    }
    std::vector<int> status(K, OFXCV_OK);
    std::vector<int64_t> holes(K, 0), launches(K, 0);
    auto work = [&](int k) {
        ofxcv_ctx* c = ctx->sub[k];
        cudaSetDevice(ctx->device);
        for (int f = k; f < nframes && status[k] == OFXCV_OK; f += K) {
            status[k] = host ? ofxcv_inpaint_u8_host(c, imgs[f], img_stride, channels, masks[f], mask_stride, outs[f], out_stride, W, H, radius, method)
                             : ofxcv_inpaint_u8(c, nullptr, imgs[f], img_stride, channels, masks[f], mask_stride, outs[f], out_stride, W, H, radius,
                                                method);
            holes[k] += c->inpaint_stats[0];
            launches[k] += c->inpaint_stats[3];
        }
        if (cudaStreamSynchronize(c->stream) != cudaSuccess && status[k] == OFXCV_OK) status[k] = OFXCV_ERR_CUDA;
    };
    std::vector<std::thread> th;
    for (int k = 1; k < K; k++) th.emplace_back(work, k);
    work(0);
    for (auto& t : th) t.join();
    ctx->inpaint_stats[0] = ctx->inpaint_stats[1] = ctx->inpaint_stats[2] = ctx->inpaint_stats[3] = 0;
    for (int k = 0; k < K; k++) {
        ctx->inpaint_stats[0] += holes[k];
        ctx->inpaint_stats[3] += launches[k];
        ctx->launches += (uint64_t)launches[k];
        if (status[k] != OFXCV_OK) {
            ctx->last_error = ctx->sub[k]->last_error;
            return status[k];
        }
    }
    return OFXCV_OK;
}

int ofxcv_inpaint_sequence_u8(ofxcv_ctx* ctx, ofxcv_stream stream, const uint8_t* const* imgs, ptrdiff_t img_stride, int channels,
                              const uint8_t* const* masks, ptrdiff_t mask_stride, uint8_t* const* outs, ptrdiff_t out_stride, int W, int H,
                              int nframes, double radius, int method, int frames_in_flight)
{
    return ip_sequence(ctx, stream, false, imgs, img_stride, channels, masks, mask_stride, outs, out_stride, W, H, nframes, radius, method,
                       frames_in_flight);
}

int ofxcv_inpaint_sequence_u8_host(ofxcv_ctx* ctx, const uint8_t* const* imgs, ptrdiff_t img_stride, int channels,
                                   const uint8_t* const* masks, ptrdiff_t mask_stride, uint8_t* const* outs, ptrdiff_t out_stride, int W,
                                   int H, int nframes, double radius, int method, int frames_in_flight)
{
    return ip_sequence(ctx, nullptr, true, imgs, img_stride, channels, masks, mask_stride, outs, out_stride, W, H, nframes, radius, method,
                       frames_in_flight);
}

}  // extern "C"
