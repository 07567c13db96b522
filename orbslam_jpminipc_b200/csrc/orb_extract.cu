// orb_extract.cu — sm_100a kernels for ORBextractor::operator() (reference
// src/ORBextractor.cc:718-779) over a batch of frames.  One launch per stage covers every
// frame (and, where the stage allows, every pyramid level) of the batch:
//
//   K0 k_level0        copyMakeBorder(image, REFLECT_101)                      (:814)
//   K1 k_resize        resize(prev level, INTER_LINEAR) + copyMakeBorder        (:800,:806)
//   K2 k_fast_nms      FAST-9/16 score + per-cell 3x3 NMS -> score map          (:607,:613)
//   K3 k_cell_compact  raster-ordered per-cell candidate lists + th=7 fallback  (:609-614)
//   K4 k_select        quota redistribution + retainBest per cell and per level (:622-701)
//   K5 k_blur          GaussianBlur 7x7 sigma 2 on the level ROI                (:760)
//   K6 k_describe      IC_Angle + rotated BRIEF-256 + keypoint emission         (:124-194,:769-777)
//
// Integer stages are bit-exact by construction; the two floating-point stages (blur, angle /
// rotation) spell out every rounding with __f*_rn / fmaf so the compiler cannot contract them.
#include "orb_internal.h"
#include "orb_trig.h"
#include <mutex>
#include <type_traits>
#include <climits>
#include "introselect.h"

namespace {

__device__ __forceinline__ int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

// ---- TMA / mbarrier helpers (tile loads run on the copy engine and overlap the previous tile's math) ----
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
// 3-D tiled TMA load (x, y, frame) -> shared memory, completion on the mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int x, int y, int z, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_addr(dst)), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(smem_addr(bar)) : "memory");
}

// PDL (launch_k below): let the dependent grid start scheduling, then wait until the prerequisite grid has completed and its memory is
// visible.  First statement of every kernel of the pass: all pipeline data is touched behind it; a no-op without the launch attribute.
#ifdef ORB_TIMELINE      // diagnostic build (tools/timeline.py): when did each kernel's first CTA arrive, and when did its dependency resolve
__device__ unsigned long long g_tl[4 * 16];     // per kernel id: min arrival, min start (behind the wait), max arrival, max start
__device__ __forceinline__ unsigned long long tl_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif
__device__ __forceinline__ void tl_stamp(int id)      // diagnostic build: latest time any CTA passed this point (slot 3 of the id)
{
#ifdef ORB_TIMELINE
    if (threadIdx.x == 0 && threadIdx.y == 0) atomicMax(&g_tl[4 * id + 3], tl_now());
#endif
}
__device__ __forceinline__ void pdl_sync(int id = 0)
{
#ifdef ORB_TIMELINE
    const bool rec = threadIdx.x == 0 && threadIdx.y == 0;
    if (rec) { const unsigned long long a = tl_now(); atomicMin(&g_tl[4 * id], a); atomicMax(&g_tl[4 * id + 2], a); }
#endif
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef ORB_TIMELINE
    if (rec) { const unsigned long long b = tl_now(); atomicMin(&g_tl[4 * id + 1], b); atomicMax(&g_tl[4 * id + 3], b); }
#endif
}

// The tile queues of the later kernels (32 counters) and k_pyramid's per-(frame, level) completion counters start every pass at zero:
// the first kernel of the pass clears them (first block of every frame) instead of a memset node in front of it — one dependent node
// fewer per pass, and a pass that consists of kernels only.
__device__ __forceinline__ void zero_counters(int* __restrict__ counters)
{
    if (blockIdx.x == 0 && blockIdx.y == 0) {
        const int tid = threadIdx.y * blockDim.x + threadIdx.x;
        if (blockIdx.z == 0 && tid < 32) counters[tid] = 0;
        if (tid < ORB_MAX_LEVELS) counters[32 + blockIdx.z * ORB_MAX_LEVELS + tid] = 0;
    }
}

// ------------------------------------------------------------------ K0
// Level 0: copy the caller's image into the ROI of the padded plane (the 16-px reflect-101 frame
// of every level is written afterwards by k_border).  4 pixels per thread.
__global__ void __launch_bounds__(256)
k_level0(const uint8_t* __restrict__ src, int w, int h, int sstride, size_t spitch, int aligned4,
         uint8_t* __restrict__ planes, size_t fbytes, int pstride, int* __restrict__ counters)
{
    pdl_sync(1);
    zero_counters(counters);
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int f = blockIdx.z;
    if (x4 >= w || y >= h) return;
    const uint8_t* S = src + (size_t)f * spitch + (size_t)y * sstride + x4;
    uint32_t v;
    if (aligned4 && x4 + 3 < w) v = __ldg(reinterpret_cast<const uint32_t*>(S));
    else {
        v = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) if (x4 + k < w) v |= (uint32_t)__ldg(S + k) << (8 * k);
    }
    // bytes past the right ROI edge (w not a multiple of 4) fall into the border and are rewritten by k_border
    *reinterpret_cast<uint32_t*>(planes + (size_t)f * fbytes + (size_t)(y + ORB_EDGE) * pstride + ORB_EDGE + x4) = v;
}

// the same copy, 16 bytes per thread, for sources whose base, row stride, frame pitch and width are multiples of 16 (640, 752, 1280,
// 1920 ... wide frames): a quarter of the loads in flight per byte moved
__global__ void __launch_bounds__(256)
k_level0_v16(const uint8_t* __restrict__ src, int w, int h, int sstride, size_t spitch, uint8_t* __restrict__ planes, size_t fbytes, int pstride,
             int* __restrict__ counters)
{
    pdl_sync(1);
    zero_counters(counters);
    const int x16 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int f = blockIdx.z;
    if (x16 >= w || y >= h) return;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + (size_t)f * spitch + (size_t)y * sstride + x16));
    *reinterpret_cast<uint4*>(planes + (size_t)f * fbytes + (size_t)(y + ORB_EDGE) * pstride + ORB_EDGE + x16) = v;
}


// ------------------------------------------------------------------ K1
// cv::resize(prev ROI -> this ROI, INTER_LINEAR), 8-bit fixed-point recipe (DESIGN.md "K1").
// Persistent kernel over 128x64 output tiles: TMA stages the source footprint of the next tile while
// the current one is computed.  A thread owns 4 adjacent output columns (source offsets and weights
// in registers) and walks down 8 output rows; the horizontal pass of a source row is kept in
// registers and reused when the next output row needs the same source row (4 rows in 5 at 1.2).
// tile = tile_w columns x (1024 / tile_w) * rs_rows rows: 128 x 64 normally, 64 x 64 with 4 rows per thread when the
// scale factor is so large that the 128-wide source footprint would exceed the 256-element TMA box limit

__global__ void __launch_bounds__(ORB_RESIZE_THREADS)
k_resize(const __grid_constant__ CUtensorMap tm, uint8_t* __restrict__ planes, size_t fbytes, LevelGeom D,
         const int2* __restrict__ xtab, const int2* __restrict__ ytab, int box_w, int box_h, int buf_bytes,
         int nimg, int* __restrict__ work_counter, int RT_W, int RS_ROWS)
{
    pdl_sync(2);
    extern __shared__ __align__(128) uint8_t rs_sm[];       // two source tiles of buf_bytes each
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ int s_next[2];
    __shared__ int s_org[2][5];                             // x0, y0, frame, source origin x / y of the item in each buffer
    const int tid = threadIdx.x;
    const int ncg = RT_W >> 2;                              // column groups (4 columns each) per tile
    const int RT_H = (ORB_RESIZE_THREADS / ncg) * RS_ROWS;
    const int tiles_x = (D.w + RT_W - 1) / RT_W, tiles_y = (D.h + RT_H - 1) / RT_H;
    const int ntiles = tiles_x * tiles_y, total = ntiles * nimg;
    const int2* xt = xtab + D.xtab_off;
    const int2* yt = ytab + D.ytab_off;
    // thread 0 only: decode an item, remember its origin for everybody and start its copy
    auto issue = [&](int item, int buf) {
        const int ti = item % ntiles, fr = item / ntiles;
        const int by = ti / tiles_x, bx = ti - by * tiles_x;
        const int x0 = bx * RT_W, y0 = by * RT_H;
        const int sxo = (__ldg(&xt[x0]).x & 0xffff) & ~15;   // 16-byte aligned TMA origin (ROI starts at padded x = 16)
        const int syo = __ldg(&yt[y0]).x & 0xffff;
        s_org[buf][0] = x0; s_org[buf][1] = y0; s_org[buf][2] = fr; s_org[buf][3] = sxo; s_org[buf][4] = syo;
        mbar_expect_tx(&bar[buf], (uint32_t)(box_w * box_h));
        tma_load_3d(rs_sm + (size_t)buf * buf_bytes, &tm, sxo + ORB_EDGE, syo + ORB_EDGE, fr, &bar[buf]);
    };
    if (tid == 0) {
        mbar_init(&bar[0], 1); mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int item = blockIdx.x;
    if (tid == 0 && item < total) issue(item, 0);
    __syncthreads();
    const int rg = tid / ncg, cgx = (tid - rg * ncg) * 4;
    for (int it = 0; item < total; it++) {
        const int buf = it & 1;
        const int x0 = s_org[buf][0], y0 = s_org[buf][1], f = s_org[buf][2], sxo = s_org[buf][3], syo = s_org[buf][4];
        if (tid == 0) {
            const int nxt = atomicAdd(work_counter, 1) + (int)gridDim.x;
            s_next[buf] = nxt;
            if (nxt < total) issue(nxt, buf ^ 1);
        }
        const int gx = x0 + cgx, ys = y0 + rg * RS_ROWS;
        const bool active = gx < D.w && ys < D.h && rg < ORB_RESIZE_THREADS / ncg;     // tile widths that do not divide 1024 leave a few threads over
        // Horizontal taps of the thread's four columns.  Their source bytes lie within 8 bytes of the first one for every scale
        // factor <= 2, so a source row costs three aligned words, two funnel shifts and per column one PRMT (both taps into
        // bytes 0,1) + one IDP.2A with the packed 16-bit weights; otherwise bytes are fetched one by one.
        int c0[4], c1[4];
        uint32_t aw[4], sel[4];
        bool fastp = true;
        int wofs = 0, sh = 0;
        if (active) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int2 e = __ldg(&xt[min(gx + k, D.w - 1)]);
                c0[k] = (e.x & 0xffff) - sxo; c1[k] = (e.x >> 16) - sxo;
                aw[k] = (uint32_t)e.y;                                  // a0 | a1 << 16, both in [0, 2048]
            }
            const int cb = c0[0];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int d0 = c0[k] - cb, d1 = c1[k] - cb;
                fastp = fastp && d0 >= 0 && d1 >= 0 && d0 <= 7 && d1 <= 7;
                sel[k] = (uint32_t)(d0 & 7) | ((uint32_t)(d1 & 7) << 4);
            }
            wofs = cb & ~3; sh = (cb & 3) * 8;
        }
        mbar_wait(&bar[buf], (uint32_t)((it >> 1) & 1));
        if (active) {
            const uint8_t* src = rs_sm + (size_t)buf * buf_bytes;
            int id0 = -1, id1 = -1, G0[4], G1[4];                       // G = horizontal sum >> 4 of source rows id0 / id1
            auto hrow = [&](int srow, int (&G)[4]) {
                const uint8_t* r = src + srow * box_w;
                if (fastp) {
                    const uint32_t* rw = reinterpret_cast<const uint32_t*>(r + wofs);
                    const uint32_t w0 = rw[0], w1 = rw[1], w2 = rw[2];
                    const uint32_t W0 = __funnelshift_r(w0, w1, sh), W1 = __funnelshift_r(w1, w2, sh);
#pragma unroll
                    for (int k = 0; k < 4; k++) G[k] = (int)(__dp2a_lo(aw[k], __byte_perm(W0, W1, sel[k]), 0u) >> 4);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; k++) G[k] = (int)((r[c0[k]] * (aw[k] & 0xffffu) + r[c1[k]] * (aw[k] >> 16)) >> 4);
                }
            };
            uint8_t* drow = planes + (size_t)f * fbytes + D.plane_off + (size_t)(ys + ORB_EDGE) * D.stride + ORB_EDGE + gx;
            const int yend = min(ys + RS_ROWS, D.h);
            for (int y = ys; y < yend; y++, drow += D.stride) {
                const int2 e = __ldg(&yt[y]);
                const int s0 = (e.x & 0xffff) - syo, s1 = (e.x >> 16) - syo;
                const int b0 = e.y & 0xffff, b1 = (int)((uint32_t)e.y >> 16);
                if (s0 != id0) {
                    if (s0 == id1) {
#pragma unroll
                        for (int k = 0; k < 4; k++) G0[k] = G1[k];
                    } else hrow(s0, G0);
                    id0 = s0;
                }
                if (s1 != id1) {
                    if (s1 == id0) {
#pragma unroll
                        for (int k = 0; k < 4; k++) G1[k] = G0[k];
                    } else hrow(s1, G1);
                    id1 = s1;
                }
                // (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2 ; weights sum to <= 2048, so the result is already in 0..255
                uint32_t o[4];
#pragma unroll
                for (int k = 0; k < 4; k++) o[k] = (uint32_t)((((b0 * G0[k]) >> 16) + ((b1 * G1[k]) >> 16) + 2) >> 2);
                // bytes past the right ROI edge fall into the border and are rewritten by k_border
                *reinterpret_cast<uint32_t*>(drow) = (o[0] | (o[1] << 8)) | ((o[2] << 16) | (o[3] << 24));
            }
        }
        __syncthreads();          // the other buffer is refilled by the next iteration's prefetch
        item = s_next[buf];
    }
}

// ------------------------------------------------------------------ K1, unrolled form (round 2)
// Same tiles, same TMA staging and the same arithmetic as k_resize, with the per-row overhead of that kernel removed: by its SASS
// k_resize spends ~95 instructions per output row of 4 pixels against ~45 of arithmetic — a dependent global load of the row's table
// entry, its decoding, 64-bit pointer stepping, four register moves whenever source row s1 becomes s0 (4 rows in 5 at 1.2), both
// forms of the horizontal pass — and ~150 per item on decoding the column taps.  Here
//   * ROWS is a compile-time constant and the row loop is fully unrolled: the ROWS table entries are fetched up front (independent
//     loads), the row offsets are immediates;
//   * the two horizontal-pass register sets swap roles every row (set B of row r is set A of row r + 1 when s0' == s1), so nothing
//     is moved; a set is recomputed only when it does not already hold the source row it should (warp-uniform tests);
//   * the column taps of a 4-column group come ready-made from a per-level table built with the plan (first source byte, the four
//     PRMT selectors, the four weight pairs: two 128-bit loads);
//   * only the packed horizontal pass exists (plan: every group's taps lie within 8 bytes, true for every scale factor <= 2);
//     other levels keep k_resize.
template <int ROWS>
__global__ void __launch_bounds__(ORB_RESIZE_THREADS)
k_resize_u(const __grid_constant__ CUtensorMap tm, uint8_t* __restrict__ planes, size_t fbytes, LevelGeom D,
           const uint4* __restrict__ xgrp, const int2* __restrict__ ytab, int box_w, int box_h, int buf_bytes,
           int nimg, int* __restrict__ work_counter, int RT_W)
{
    pdl_sync(2);
    extern __shared__ __align__(128) uint8_t rs_sm[];       // two source tiles of buf_bytes each
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ int s_next[2];
    __shared__ int s_org[2][5];                             // x0, y0, frame, source origin x / y of the item in each buffer
    const int tid = threadIdx.x;
    const int ncg = RT_W >> 2;
    const int RT_H = (ORB_RESIZE_THREADS / ncg) * ROWS;
    const int tiles_x = (D.w + RT_W - 1) / RT_W, tiles_y = (D.h + RT_H - 1) / RT_H;
    const int ntiles = tiles_x * tiles_y, total = ntiles * nimg;
    const int2* yt = ytab + D.ytab_off;
    auto issue = [&](int item, int buf) {
        const int ti = item % ntiles, fr = item / ntiles;
        const int by = ti / tiles_x, bx = ti - by * tiles_x;
        const int x0 = bx * RT_W, y0 = by * RT_H;
        const int sxo = (int)__ldg(&xgrp[(x0 >> 2) * 2]).x & ~15;      // 16-byte aligned TMA origin (ROI starts at padded x = 16)
        const int syo = __ldg(&yt[y0]).x & 0xffff;
        s_org[buf][0] = x0; s_org[buf][1] = y0; s_org[buf][2] = fr; s_org[buf][3] = sxo; s_org[buf][4] = syo;
        mbar_expect_tx(&bar[buf], (uint32_t)(box_w * box_h));
        tma_load_3d(rs_sm + (size_t)buf * buf_bytes, &tm, sxo + ORB_EDGE, syo + ORB_EDGE, fr, &bar[buf]);
    };
    if (tid == 0) {
        mbar_init(&bar[0], 1); mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int item = blockIdx.x;
    if (tid == 0 && item < total) issue(item, 0);
    __syncthreads();
    const int rg = tid / ncg, cgx = (tid - rg * ncg) * 4;
    const uint32_t sm_base = (uint32_t)__cvta_generic_to_shared(rs_sm);
    for (int it = 0; item < total; it++) {
        const int buf = it & 1;
        const int x0 = s_org[buf][0], y0 = s_org[buf][1], f = s_org[buf][2], sxo = s_org[buf][3], syo = s_org[buf][4];
        if (tid == 0) {
            const int nxt = atomicAdd(work_counter, 1) + (int)gridDim.x;
            s_next[buf] = nxt;
            if (nxt < total) issue(nxt, buf ^ 1);
        }
        const int gx = x0 + cgx, ys = y0 + rg * ROWS;
        const bool active = gx < D.w && ys < D.h && rg < ORB_RESIZE_THREADS / ncg;
        uint4 ga = make_uint4(0, 0, 0, 0), gb = ga;
        int2 ye[ROWS];
        if (active) {
            ga = __ldg(&xgrp[(gx >> 2) * 2]); gb = __ldg(&xgrp[(gx >> 2) * 2 + 1]);
#pragma unroll
            for (int r = 0; r < ROWS; r++) ye[r] = __ldg(&yt[min(ys + r, D.h - 1)]);
        }
        mbar_wait(&bar[buf], (uint32_t)((it >> 1) & 1));
        if (active) {
            const int cb = (int)ga.x - sxo;
            const uint32_t sh = (uint32_t)(cb & 3) * 8u;
            const uint32_t row0 = sm_base + (uint32_t)(buf * buf_bytes + (cb & ~3) - syo * box_w);   // + srow_abs * box_w
            const uint32_t sel0 = ga.y, sel1 = ga.y >> 16, sel2 = ga.z, sel3 = ga.z >> 16;
            uint32_t GA[4], GB[4];
            int idA = -1, idB = -1;
            auto hrow = [&](int srow, uint32_t (&G)[4]) {
                const uint32_t a = row0 + (uint32_t)(srow * box_w);
                uint32_t w0, w1, w2;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(a));
                asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(a));
                asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(w2) : "r"(a));
                const uint32_t W0 = __funnelshift_r(w0, w1, sh), W1 = __funnelshift_r(w1, w2, sh);
                G[0] = __dp2a_lo(gb.x, __byte_perm(W0, W1, sel0), 0u) >> 4;
                G[1] = __dp2a_lo(gb.y, __byte_perm(W0, W1, sel1), 0u) >> 4;
                G[2] = __dp2a_lo(gb.z, __byte_perm(W0, W1, sel2), 0u) >> 4;
                G[3] = __dp2a_lo(gb.w, __byte_perm(W0, W1, sel3), 0u) >> 4;
            };
            uint8_t* drow = planes + (size_t)f * fbytes + D.plane_off + (size_t)(ys + ORB_EDGE) * D.stride + ORB_EDGE + gx;
            const int nrow = min(ROWS, D.h - ys);
            auto out_row = [&](int r, const uint32_t (&G0)[4], const uint32_t (&G1)[4]) {
                const uint32_t b0 = (uint32_t)ye[r].y & 0xffffu, b1 = (uint32_t)ye[r].y >> 16;
                // (((b0*G0)>>16) + ((b1*G1)>>16) + 2) >> 2, the + 2 riding on the second product; weights sum to <= 2048: 0..255
                uint32_t o[4];
#pragma unroll
                for (int k = 0; k < 4; k++) o[k] = (((b0 * G0[k]) >> 16) + ((b1 * G1[k] + 0x20000u) >> 16)) >> 2;
                if (r < nrow)          // bytes past the right ROI edge fall into the border and are rewritten by k_border
                    *reinterpret_cast<uint32_t*>(drow + (size_t)r * D.stride) = (o[0] | (o[1] << 8)) | ((o[2] << 16) | (o[3] << 24));
            };
#pragma unroll
            for (int r = 0; r < ROWS; r++) {
                const int s0 = ye[r].x & 0xffff, s1 = (int)((uint32_t)ye[r].x >> 16);
                if ((r & 1) == 0) {            // even rows: A = s0, B = s1
                    if (idA != s0) { hrow(s0, GA); idA = s0; }
                    if (idB != s1) { hrow(s1, GB); idB = s1; }
                    out_row(r, GA, GB);
                } else {                       // odd rows: B = s0 (the previous row's s1 when the source advanced by one), A = s1
                    if (idB != s0) { hrow(s0, GB); idB = s0; }
                    if (idA != s1) { hrow(s1, GA); idA = s1; }
                    out_row(r, GB, GA);
                }
            }
        }
        __syncthreads();          // the other buffer is refilled by the next iteration's prefetch
        item = s_next[buf];
    }
}

// ------------------------------------------------------------------ K1, fused
// The whole resize cascade (levels 1 .. nlevels-1 of every frame) in ONE persistent launch instead of one launch per level: seven
// serial launches cost seven ramp-ups and seven tails (levels 4..7 are only a few hundred tiles per 64 frames).  Work items are
// numbered LEVEL-MAJOR over the batch (all tiles of level 1 of all frames, then level 2, ...) and handed out in that order by an
// atomic counter; an item of level l >= 2 needs level l-1 of ITS frame complete, which is tracked by one counter of finished tiles
// per (frame, level).  Every dependency of an item has a smaller item number, i.e. has been claimed by a running CTA before, and a
// CTA publishes the completion of its current item BEFORE it blocks on the dependency of its next one, so the waits cannot form a
// cycle (for a single frame they simply serialise the levels).  With a whole batch in flight the dependencies of an item were
// finished thousands of items earlier and the wait is one acquire load.
// Visibility: producers store with the generic proxy, __syncthreads, then thread 0 fences (gpu scope, cumulative) and bumps the
// counter; the consumer's thread 0 acquires the counter and issues fence.proxy.async before the TMA (async proxy) read.
struct PyrLevel { int tile_w, rows, box_w, box_h, tiles_x, ntiles, item_base, pad; };
struct PyrParams { PyrLevel L[ORB_MAX_LEVELS]; int nlevels, total, buf_bytes, nimg; };

__global__ void __launch_bounds__(ORB_RESIZE_THREADS)
k_pyramid(const __grid_constant__ TmapSet tm, uint8_t* __restrict__ planes, size_t fbytes, const Plan* __restrict__ plan,
          const __grid_constant__ PyrParams P, const int2* __restrict__ xtab, const int2* __restrict__ ytab,
          int* __restrict__ work_counter, int* __restrict__ done)
{
    pdl_sync(2);
    extern __shared__ __align__(128) uint8_t rs_sm[];       // two source tiles of buf_bytes each
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ int s_next[2];
    __shared__ int s_org[2][6];                             // x0, y0, frame, source origin x / y, level of the item in each buffer
    const int tid = threadIdx.x;
    const int buf_bytes = P.buf_bytes;
    // thread 0 only: which level / frame / tile an item is, where its source footprint starts
    auto decode = [&](int item, int buf) {
        int l = 1;
        while (l + 1 < P.nlevels && item >= P.L[l + 1].item_base) l++;
        const PyrLevel& Q = P.L[l];
        const int idx = item - Q.item_base;
        const int fr = idx / Q.ntiles, ti = idx - fr * Q.ntiles;
        const int by = ti / Q.tiles_x, bx = ti - by * Q.tiles_x;
        const int RT_H = (ORB_RESIZE_THREADS / (Q.tile_w >> 2)) * Q.rows;
        const int x0 = bx * Q.tile_w, y0 = by * RT_H;
        const LevelGeom& D = plan->L[l];
        const int sxo = (__ldg(&xtab[D.xtab_off + x0]).x & 0xffff) & ~15;   // 16-byte aligned TMA origin (ROI starts at padded x = 16)
        const int syo = __ldg(&ytab[D.ytab_off + y0]).x & 0xffff;
        s_org[buf][0] = x0; s_org[buf][1] = y0; s_org[buf][2] = fr; s_org[buf][3] = sxo; s_org[buf][4] = syo; s_org[buf][5] = l;
    };
    auto dep_ready = [&](int buf) -> bool {
        const int l = s_org[buf][5];
#ifdef ORB_PYR_NOSYNC
        return true;
#endif
        if (l < 2) return true;                             // level 1 reads level 0, written by the previous kernel
        int v;
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(done + s_org[buf][2] * ORB_MAX_LEVELS + (l - 1)) : "memory");
        return v >= P.L[l - 1].ntiles;
    };
    auto issue = [&](int buf) {
        const int l = s_org[buf][5];
        const PyrLevel& Q = P.L[l];
#ifndef ORB_PYR_NOSYNC
        if (l >= 2) asm volatile("fence.proxy.async.global;" ::: "memory");    // other CTAs' generic-proxy stores (acquired above) before this async-proxy read
#endif
        mbar_expect_tx(&bar[buf], (uint32_t)(Q.box_w * Q.box_h));
        tma_load_3d(rs_sm + (size_t)buf * buf_bytes, &tm.m[l], s_org[buf][3] + ORB_EDGE, s_org[buf][4] + ORB_EDGE, s_org[buf][2], &bar[buf]);
    };
    if (tid == 0) {
        mbar_init(&bar[0], 1); mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // the FIRST item is claimed from the queue as well (not blockIdx.x): item numbers then follow the order in which CTAs actually
    // start, so every dependency belongs to a CTA that is already running — no assumption about dispatch order or residency
    if (tid == 0) {
        const int first = atomicAdd(work_counter, 1);
        s_next[1] = first;
        if (first < P.total) {
            decode(first, 0);
            while (!dep_ready(0)) __nanosleep(64);
            issue(0);
        }
    }
    __syncthreads();
    int item = s_next[1];
    for (int it = 0; item < P.total; it++) {
        const int buf = it & 1;
        const int x0 = s_org[buf][0], y0 = s_org[buf][1], f = s_org[buf][2], sxo = s_org[buf][3], syo = s_org[buf][4], lvl = s_org[buf][5];
        bool pending = false;
        int nxt = P.total;
        if (tid == 0) {
            nxt = atomicAdd(work_counter, 1);
            s_next[buf] = nxt;
            if (nxt < P.total) {
                decode(nxt, buf ^ 1);
                if (dep_ready(buf ^ 1)) issue(buf ^ 1); else pending = true;     // never block here: the dependency may be THIS item
            }
        }
        const PyrLevel& Q = P.L[lvl];
        const LevelGeom& D = plan->L[lvl];
        const int RT_W = Q.tile_w, RS_ROWS = Q.rows, box_w = Q.box_w;
        const int ncg = RT_W >> 2;                              // column groups (4 columns each) per tile
        const int2* xt = xtab + D.xtab_off;
        const int2* yt = ytab + D.ytab_off;
        const int Dw = D.w, Dh = D.h, Dstride = D.stride;
        const int rg = tid / ncg, cgx = (tid - rg * ncg) * 4;
        const int gx = x0 + cgx, ys = y0 + rg * RS_ROWS;
        const bool active = gx < Dw && ys < Dh && rg < ORB_RESIZE_THREADS / ncg;     // tile widths that do not divide 1024 leave a few threads over
        // horizontal taps of the thread's four columns (see k_resize)
        int c0[4], c1[4];
        uint32_t aw[4], sel[4];
        bool fastp = true;
        int wofs = 0, sh = 0;
        if (active) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int2 e = __ldg(&xt[min(gx + k, Dw - 1)]);
                c0[k] = (e.x & 0xffff) - sxo; c1[k] = (e.x >> 16) - sxo;
                aw[k] = (uint32_t)e.y;                                  // a0 | a1 << 16, both in [0, 2048]
            }
            const int cb = c0[0];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int d0 = c0[k] - cb, d1 = c1[k] - cb;
                fastp = fastp && d0 >= 0 && d1 >= 0 && d0 <= 7 && d1 <= 7;
                sel[k] = (uint32_t)(d0 & 7) | ((uint32_t)(d1 & 7) << 4);
            }
            wofs = cb & ~3; sh = (cb & 3) * 8;
        }
        mbar_wait(&bar[buf], (uint32_t)((it >> 1) & 1));
        if (active) {
            const uint8_t* src = rs_sm + (size_t)buf * buf_bytes;
            int id0 = -1, id1 = -1, G0[4], G1[4];                       // G = horizontal sum >> 4 of source rows id0 / id1
            auto hrow = [&](int srow, int (&G)[4]) {
                const uint8_t* r = src + srow * box_w;
                if (fastp) {
                    const uint32_t* rw = reinterpret_cast<const uint32_t*>(r + wofs);
                    const uint32_t w0 = rw[0], w1 = rw[1], w2 = rw[2];
                    const uint32_t W0 = __funnelshift_r(w0, w1, sh), W1 = __funnelshift_r(w1, w2, sh);
#pragma unroll
                    for (int k = 0; k < 4; k++) G[k] = (int)(__dp2a_lo(aw[k], __byte_perm(W0, W1, sel[k]), 0u) >> 4);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; k++) G[k] = (int)((r[c0[k]] * (aw[k] & 0xffffu) + r[c1[k]] * (aw[k] >> 16)) >> 4);
                }
            };
            uint8_t* drow = planes + (size_t)f * fbytes + D.plane_off + (size_t)(ys + ORB_EDGE) * Dstride + ORB_EDGE + gx;
            const int yend = min(ys + RS_ROWS, Dh);
            for (int y = ys; y < yend; y++, drow += Dstride) {
                const int2 e = __ldg(&yt[y]);
                const int s0 = (e.x & 0xffff) - syo, s1 = (e.x >> 16) - syo;
                const int b0 = e.y & 0xffff, b1 = (int)((uint32_t)e.y >> 16);
                if (s0 != id0) {
                    if (s0 == id1) {
#pragma unroll
                        for (int k = 0; k < 4; k++) G0[k] = G1[k];
                    } else hrow(s0, G0);
                    id0 = s0;
                }
                if (s1 != id1) {
                    if (s1 == id0) {
#pragma unroll
                        for (int k = 0; k < 4; k++) G1[k] = G0[k];
                    } else hrow(s1, G1);
                    id1 = s1;
                }
                uint32_t o[4];
#pragma unroll
                for (int k = 0; k < 4; k++) o[k] = (uint32_t)((((b0 * G0[k]) >> 16) + ((b1 * G1[k]) >> 16) + 2) >> 2);
                *reinterpret_cast<uint32_t*>(drow) = (o[0] | (o[1] << 8)) | ((o[2] << 16) | (o[3] << 24));
            }
        }
        __syncthreads();          // every store of this item is issued; the other buffer may be refilled
        if (tid == 0) {
            // release at gpu scope, cumulative over the CTA's stores (ordered before it by the barrier); nobody reads the last level here
#ifndef ORB_PYR_NOSYNC
            if (lvl + 1 < P.nlevels)
#else
            if (false)
#endif
                asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(done + f * ORB_MAX_LEVELS + lvl) : "memory");
            if (pending) {                                              // own item published: now it is safe to block on the next one's level
                while (!dep_ready(buf ^ 1)) __nanosleep(64);
                issue(buf ^ 1);
            }
        }
        item = s_next[buf];
    }
}

// copyMakeBorder(..., 16, BORDER_REFLECT_101) for every level of every frame in one launch
// (reference src/ORBextractor.cc:806,814).  Only blur and the descriptor sampler read the frame;
// resize, FAST and IC_Angle stay inside the ROI.  One thread per 32-bit word of the frame region
// (a warp-per-row variant measured slower: 0.51 vs 0.45 ms for resize+border per 256 frames).
__global__ void __launch_bounds__(256)
k_border(uint8_t* __restrict__ planes, uint8_t* __restrict__ blurred, size_t fbytes, const Plan* __restrict__ plan)
{
    pdl_sync(3);
    // blockIdx.y = frame * nlevels + level; blockIdx.x walks the level's frame words (levels with fewer words exit).  Two words per
    // thread, both fetched before either is stored: the kernel is bound by load latency, not by instructions.
    const int l = blockIdx.y % plan->nlevels, f = blockIdx.y / plan->nlevels;
    const LevelGeom& L = plan->L[l];
    const int first = blockIdx.x * (2 * blockDim.x) + threadIdx.x;
    if (first >= L.border_items) return;
    const int w = L.w, h = L.h;
    const int wpr = L.stride >> 2;                      // words per padded row
    const int RING = L.ring;                            // ORB_RING, or the whole 16 px frame for a level whose cells reach past size - 16
    const int band = RING * wpr;                        // words in the top (or bottom) band: RING rows next to the ROI
    const int rw0 = (ORB_EDGE + w) >> 2;                // first word that contains right-frame pixels
    const int LW = RING >> 2;                           // words of the left ring
    const int side = 2 * LW + 1;                        // ring words per middle row: left ring + the words covering [w, w + RING)
    const size_t poff = (size_t)f * fbytes + L.plane_off;
    auto fetch = [&](int item, size_t& o) -> uint32_t {
    int py, wx;
    if (item < 2 * band) {                              // float reciprocal division is exact here (item < 2^23)
        const int bi = item >= band;
        if (bi) item -= band;
        py = __float2int_rz(__fdividef((float)item + 0.5f, (float)wpr));
        wx = item - py * wpr;
        py += bi ? ORB_EDGE + h : ORB_EDGE - RING;
    } else {
        item -= 2 * band;
        py = item / side;
        wx = item - py * side;
        py += ORB_EDGE;
        wx = wx < LW ? (ORB_EDGE - RING) / 4 + wx : rw0 + (wx - LW);
        if (wx >= wpr) wx = wpr - 1;                    // a ROI that ends on the last word of the pitch: rewrite that word
    }
    int sy = py - ORB_EDGE;
    if (sy < 0) sy = -sy; else if (sy >= h) sy = 2 * h - 2 - sy;
    if ((unsigned)sy >= (unsigned)h) sy = reflect101(sy, h);            // only for ROIs lower than the frame
    const uint8_t* srow = planes + poff + (size_t)(sy + ORB_EDGE) * L.stride + ORB_EDGE;
    uint32_t v = 0;
    const int xw = wx * 4 - ORB_EDGE;                   // ROI column of the word's first byte; ROI rows start 16-byte aligned
    if (xw >= 0 && xw + 3 < w) {
        v = *reinterpret_cast<const uint32_t*>(srow + xw);              // band word above / below the ROI: an aligned copy
    } else if (xw < 0 && w > ORB_EDGE + 3) {
        // left frame: bytes x = xw..xw+3 mirror columns -xw, -xw-1, -xw-2, -xw-3 = byte 0 of the aligned word at -xw and
        // bytes 3, 2, 1 of the word before it
        const uint32_t hi = *reinterpret_cast<const uint32_t*>(srow - xw), lo = *reinterpret_cast<const uint32_t*>(srow - xw - 4);
        v = __byte_perm(lo, hi, 0x1234);
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int x = xw + k;                                 // ROI column of this byte
            if (x < w + ORB_EDGE) {
                if (x < 0) x = -x; else if (x >= w) x = 2 * w - 2 - x;
                if ((unsigned)x >= (unsigned)w) x = reflect101(x, w);       // only for ROIs narrower than the frame
                v |= (uint32_t)srow[x] << (8 * k);
            }
        }
    }
    o = poff + (size_t)py * L.stride + wx * 4;
    return v;
    };
    size_t o0 = 0, o1 = 0;
    const int second = first + blockDim.x;
    const bool two = second < L.border_items;
    const uint32_t v0 = fetch(first, o0);
    const uint32_t v1 = two ? fetch(second, o1) : 0u;
    // the in-place blur of the reference leaves the frame un-blurred (:760): give the blurred buffer the same frame so
    // that the descriptor sampler reads one buffer only (k_blur later rewrites exactly the ROI bytes)
    *reinterpret_cast<uint32_t*>(planes + o0) = v0;
    *reinterpret_cast<uint32_t*>(blurred + o0) = v0;
    if (two) {
        *reinterpret_cast<uint32_t*>(planes + o1) = v1;
        *reinterpret_cast<uint32_t*>(blurred + o1) = v1;
    }
}

// ------------------------------------------------------------------ K2
// FAST-9/16 + per-cell NMS on a 64x32 tile.  Arithmetic is packed two pixels per register in
// unsigned 16-bit lanes so the 9-of-16 arc tests run on VIMNMX3.U16x2 (3-input min/max):
//   bright strength = max_k( min of ring[k..k+8] ) - v,   dark strength = v - min_k( max of ring[k..k+8] )
// corner at threshold t  <=>  max(bright, dark) > t ;  OpenCV's response = that maximum - 1.
constexpr int FT_W = ORB_TILE_W, FT_H = ORB_TILE_H;
constexpr int FIW = FT_W / 4 + 8, FI_H = FT_H + 8;   // image tile: cols x0-16..x0+FT_W+15, rows y0-4..y0+FT_H+3; TMA needs a 16-byte aligned x origin
constexpr int FSW = FT_W / 4 + 2, FS_H = FT_H + 2;   // score tile: cols x0-4..x0+FT_W+3, rows y0-1..y0+FT_H
#ifndef ORB_FAST_THREADS
#define ORB_FAST_THREADS 128      // measured on B200: 320x3 1.75 ms, 256x4 1.56, 128x8 1.46, 64x16 1.43 per 256 frames
#define ORB_FAST_CTAS 8
#endif
constexpr int FAST_THREADS = ORB_FAST_THREADS, FAST_CTAS = ORB_FAST_CTAS, FAST_THREADS_SMALL = 512;

__device__ __forceinline__ uint32_t lo16x2(uint32_t w) { return __byte_perm(w, 0, 0x4140); }   // bytes 0,1 -> u16 lanes
__device__ __forceinline__ uint32_t hi16x2(uint32_t w) { return __byte_perm(w, 0, 0x4342); }   // bytes 2,3 -> u16 lanes

// ring[k] packed for two pixels -> (max over arcs of arc-min, min over arcs of arc-max), packed
#ifndef ORB_FAST_ARC_V1
// 34 three-input operations per polarity instead of 40.  The four arcs starting at k..k+3 share the six samples C = r[k+3..k+8];
// with max(min(S,a), min(S,b)) = min(S, max(a,b)):
//   max(arc k,   arc k+1) = min(C, r[k+1], r[k+2],  max(r[k],   r[k+9]))
//   max(arc k+2, arc k+3) = min(C, r[k+9], r[k+10], max(r[k+2], r[k+11]))
// so one group of four arcs costs 8 operations (k = 0, 4, 8, 12) and two more combine the groups.
__device__ __forceinline__ void arc_minmax(const uint32_t (&r)[16], uint32_t& Mn, uint32_t& Mx)
{
    uint32_t g[4], h[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int k = 4 * q;
#define R_(i) r[(k + (i)) & 15]
        {
            const uint32_t eA = __vimin3_u16x2(R_(1), R_(2), __vmaxu2(R_(0), R_(9)));
            const uint32_t eB = __vimin3_u16x2(R_(9), R_(10), __vmaxu2(R_(2), R_(11)));
            const uint32_t c1 = __vimin3_u16x2(R_(3), R_(4), R_(5)), c2 = __vimin3_u16x2(R_(6), R_(7), R_(8));
            g[q] = __vimin3_u16x2(c1, c2, __vmaxu2(eA, eB));
        }
        {
            const uint32_t eA = __vimax3_u16x2(R_(1), R_(2), __vminu2(R_(0), R_(9)));
            const uint32_t eB = __vimax3_u16x2(R_(9), R_(10), __vminu2(R_(2), R_(11)));
            const uint32_t c1 = __vimax3_u16x2(R_(3), R_(4), R_(5)), c2 = __vimax3_u16x2(R_(6), R_(7), R_(8));
            h[q] = __vimax3_u16x2(c1, c2, __vminu2(eA, eB));
        }
#undef R_
    }
    Mn = __vmaxu2(__vimax3_u16x2(g[0], g[1], g[2]), g[3]);
    Mx = __vminu2(__vimin3_u16x2(h[0], h[1], h[2]), h[3]);
}
#else
__device__ __forceinline__ void arc_minmax(const uint32_t (&r)[16], uint32_t& Mn, uint32_t& Mx)
{
    uint32_t a[16], b[16];
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = __vimin3_u16x2(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
#pragma unroll
    for (int k = 0; k < 16; k++) b[k] = __vimin3_u16x2(a[k], a[(k + 3) & 15], a[(k + 6) & 15]);      // min of 9 contiguous
    Mn = __vimax3_u16x2(__vimax3_u16x2(b[0], b[1], b[2]), __vimax3_u16x2(b[3], b[4], b[5]), __vimax3_u16x2(b[6], b[7], b[8]));
    Mn = __vimax3_u16x2(Mn, __vimax3_u16x2(b[9], b[10], b[11]), __vimax3_u16x2(b[12], b[13], b[14]));
    Mn = __vmaxu2(Mn, b[15]);
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = __vimax3_u16x2(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
#pragma unroll
    for (int k = 0; k < 16; k++) b[k] = __vimax3_u16x2(a[k], a[(k + 3) & 15], a[(k + 6) & 15]);      // max of 9 contiguous
    Mx = __vimin3_u16x2(__vimin3_u16x2(b[0], b[1], b[2]), __vimin3_u16x2(b[3], b[4], b[5]), __vimin3_u16x2(b[6], b[7], b[8]));
    Mx = __vimin3_u16x2(Mx, __vimin3_u16x2(b[9], b[10], b[11]), __vimin3_u16x2(b[12], b[13], b[14]));
    Mx = __vminu2(Mx, b[15]);
}
#endif

// packed epilogue for two pixels.  strength T = max(Mn - v, v - Mx); a corner has T > th and OpenCV's response is T - 1.  The score
// tile holds the EXCESS q = max(T - th, 0) instead (same order among corners, 0 for the rest, so the strict NMS decides identically);
// the survivors get th - 1 added back when they are written.  With nvp = -(v + th) and vm1 = v - th + 1 per lane (once per pixel
// pair) one evaluation is a NOT and two VIADDMNMX.RELU:  q = max(Mn + nvp, ~Mx + vm1, 0)   (~Mx = -Mx - 1 in a 16-bit lane).
__device__ __forceinline__ uint32_t excess2(uint32_t Mn, uint32_t Mx, uint32_t nvp, uint32_t vm1)
{
    const uint32_t qb = __viaddmax_s16x2_relu(Mn, nvp, 0u);
    return __viaddmax_s16x2_relu(~Mx, vm1, qb);
}

// ---- round 2: part of the arc network on the FMA pipe ----
// ncu puts k_fast_nms at 87 % of the ALU pipe with the FMA pipe idle (4.6 %).  Both pipes issue a warp instruction every second
// cycle per scheduler, so a 2-input min/max that leaves the ALU pipe frees it for two cycles even if it costs two FMA-pipe
// instructions.  With a lane = 0x6400 | pixel (the fp16 number 1024 + pixel; positive halves order like their bit patterns, so
// VIMNMX3.U16x2 keeps working on the same registers) and d = relu(a - b) = fma.rn.relu.f16x2(b, -1, a):
//     max(a, b) = b + d        min(a, b) = a - d          (all values are integers below 2048: exact in fp16)
// and the min tree and the max tree want max / min of the SAME pairs (r[k], r[k+9]) and (r[k+2], r[k+11]), so those share d:
// 3 FMA-pipe instructions replace 2 ALU-pipe ones.  Per call 26 of the 68 operations move (ORB_FAST_HALF_LEVEL: 1 = the shared
// pairs only, 2 = also max(eA, eB) / min(eA, eB) and the final combine).
#ifndef ORB_FAST_HALF_LEVEL
#define ORB_FAST_HALF_LEVEL 2
#endif
__device__ __forceinline__ uint32_t h2_relu_sub(uint32_t a, uint32_t b)       // relu(a - b) per fp16 lane
{
    uint32_t d;
    asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(b), "r"(0xBC00BC00u), "r"(a));
    return d;
}
#ifdef ORB_FAST_HALF_FMAONLY
__device__ __forceinline__ uint32_t h2_add(uint32_t a, uint32_t b) { uint32_t r; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(0x3C003C00u), "r"(b)); return r; }
__device__ __forceinline__ uint32_t h2_sub(uint32_t a, uint32_t b) { uint32_t r; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(0xBC00BC00u), "r"(a)); return r; }
#else
__device__ __forceinline__ uint32_t h2_add(uint32_t a, uint32_t b) { uint32_t r; asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t h2_sub(uint32_t a, uint32_t b) { uint32_t r; asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
#endif
__device__ __forceinline__ uint32_t h2_max(uint32_t a, uint32_t b) { return h2_add(b, h2_relu_sub(a, b)); }
__device__ __forceinline__ uint32_t h2_min(uint32_t a, uint32_t b) { return h2_sub(a, h2_relu_sub(a, b)); }

__device__ __forceinline__ void arc_minmax_h(const uint32_t (&r)[16], uint32_t& Mn, uint32_t& Mx)
{
    uint32_t g[4], h[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int k = 4 * q;
#define R_(i) r[(k + (i)) & 15]
#if ORB_FAST_HALF_LEVEL >= 1
        const uint32_t d09 = h2_relu_sub(R_(0), R_(9)), d211 = h2_relu_sub(R_(2), R_(11));
        const uint32_t mx09 = h2_add(R_(9), d09), mn09 = h2_sub(R_(0), d09);
        const uint32_t mx211 = h2_add(R_(11), d211), mn211 = h2_sub(R_(2), d211);
#else
        const uint32_t mx09 = __vmaxu2(R_(0), R_(9)), mn09 = __vminu2(R_(0), R_(9)), mx211 = __vmaxu2(R_(2), R_(11)), mn211 = __vminu2(R_(2), R_(11));
#endif
        {
            const uint32_t eA = __vimin3_u16x2(R_(1), R_(2), mx09);
            const uint32_t eB = __vimin3_u16x2(R_(9), R_(10), mx211);
            const uint32_t c1 = __vimin3_u16x2(R_(3), R_(4), R_(5)), c2 = __vimin3_u16x2(R_(6), R_(7), R_(8));
            g[q] = __vimin3_u16x2(c1, c2, ORB_FAST_HALF_LEVEL >= 2 ? h2_max(eA, eB) : __vmaxu2(eA, eB));
        }
        {
            const uint32_t eA = __vimax3_u16x2(R_(1), R_(2), mn09);
            const uint32_t eB = __vimax3_u16x2(R_(9), R_(10), mn211);
            const uint32_t c1 = __vimax3_u16x2(R_(3), R_(4), R_(5)), c2 = __vimax3_u16x2(R_(6), R_(7), R_(8));
            h[q] = __vimax3_u16x2(c1, c2, ORB_FAST_HALF_LEVEL >= 2 ? h2_min(eA, eB) : __vminu2(eA, eB));
        }
#undef R_
    }
    const uint32_t a = __vimax3_u16x2(g[0], g[1], g[2]), b = __vimin3_u16x2(h[0], h[1], h[2]);
    Mn = ORB_FAST_HALF_LEVEL >= 2 ? h2_max(a, g[3]) : __vmaxu2(a, g[3]);
    Mx = ORB_FAST_HALF_LEVEL >= 2 ? h2_min(b, h[3]) : __vminu2(b, h[3]);
}

// Persistent kernel: each CTA walks (tile, frame) work items; the image tile of item i+1 is fetched by
// TMA into the other shared-memory buffer while item i is being scored.
// ETILE (round 2, ORB_FAST_ETILE): the raw tile is first re-encoded into a tile of half lanes (one 16-bit lane 0x6400 | pixel per pixel,
// columns x0-4 .. x0+67), so that a ring sample pair is an aligned 32-bit word of that tile when its first pixel is even and ONE
// PRMT of two neighbouring words when it is odd: 18 PRMT per 4-pixel task instead of 42 + 10 LOP3, and the 0x64 comes out of
// memory instead of out of an instruction.  The raw tile is then dead (the halo pass, which still reads it, runs before the score
// pass), so ONE raw buffer suffices: the next item's TMA copy is issued after the re-encoding and lands during the score pass.
constexpr int EW = FT_W / 2 + 4;        // 32-bit words per row of the half-lane tile (72 pixels)
// THREADS: 128 for batches; FAST_THREADS_SMALL for calls of a few frames, where every SM holds at most ONE tile and a tile's time is
// the issue latency of the four warps that walk it (16.7 of the 115 us of one 640x480 frame)
template <bool ETILE, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == FAST_THREADS ? (ETILE ? 5 : 6) : 1)      // residency is set by shared memory (42.6 / 36.1 KB per CTA), so the register cap may follow it
k_fast_nms(const __grid_constant__ TmapSet tm, uint8_t* __restrict__ nms, uint8_t* __restrict__ bitmap, size_t fbytes,
           const Plan* __restrict__ plan, const Tile* __restrict__ tiles, int ntiles, int total, int* __restrict__ work_counter,
           const uint8_t* __restrict__ coltab, const int16_t* __restrict__ rowtab)
{
    pdl_sync(4);
    __shared__ __align__(128) uint32_t img2[ETILE ? 1 : 2][(FI_H * FIW + 31) & ~31];      // each buffer 128-byte aligned (TMA destination) for any tile height
    __shared__ __align__(16) uint32_t et[ETILE ? FI_H * EW : 4];
    __shared__ int s_next[2], s_ti[2], s_fr[2];             // next work item; tile index and frame of the item in each buffer
    __shared__ __align__(16) uint32_t sc[FS_H * FSW];
    __shared__ short rowcell[FS_H];
    __shared__ __align__(16) uint32_t mask3[3 * FSW];       // byte masks per score-tile column: in region / left, right neighbour in the same cell
    const uint8_t* m_in = reinterpret_cast<const uint8_t*>(mask3);
    const uint8_t* m_l = m_in + FSW * 4;
    const uint8_t* m_r = m_in + FSW * 8;
    __shared__ __align__(8) uint64_t bar[2];
    const int tid = threadIdx.x;
    // th = 0 (fastTh 0): a corner of strength 1 has response 0, never survives the strict NMS and never suppresses anything, exactly
    // like a non-corner, so the kernel may treat it as one
    const int th = max(plan->th_lo, 1);
    const uint32_t c1 = (uint32_t)((1 - th) & 0xffff) * 0x00010001u;       // (1 - th) in both lanes
    // half lanes carry 0x6400 + value: the epilogue's lane constants absorb it (all sums stay inside a signed 16-bit lane)
    const uint32_t c1n = (uint32_t)((1 - th - 0x6400) & 0xffff) * 0x00010001u, c1p = (uint32_t)((1 - th + 0x6400) & 0xffff) * 0x00010001u;
    const uint32_t th_m1 = (uint32_t)(th - 1);
    constexpr uint32_t TILE_BYTES = FI_H * FIW * 4;

    auto issue = [&](int item, int buf) {      // one thread: arm the barrier, start the copy
        const int ti = item % ntiles, fr = item / ntiles;
        const Tile t = tiles[ti];
        s_ti[buf] = ti; s_fr[buf] = fr;                     // decoded once here instead of by every thread
        const int rb = ETILE ? 0 : buf;
        mbar_expect_tx(&bar[rb], TILE_BYTES);
        tma_load_3d(&img2[rb][0], &tm.m[t.level], t.x0 - 16 + ORB_EDGE, t.y0 - 4 + ORB_EDGE, fr, &bar[rb]);
    };
    if (tid == 0) {
        mbar_init(&bar[0], 1); mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // dynamic work distribution: items differ a lot in cost (flat areas are rejected early)
    int item = blockIdx.x;
    if (tid == 0 && item < total) issue(item, 0);
    __syncthreads();

    for (int it = 0; item < total; it++) {
        const int buf = it & 1;
        const int f = s_fr[buf];
        const Tile t = tiles[s_ti[buf]];
        const LevelGeom& L = plan->L[t.level];
        // claim + prefetch the next item; its buffer was last read before the post-scoring barrier of the previous iteration
        if (tid == 0) {
            const int nxt = atomicAdd(work_counter, 1) + (int)gridDim.x;
            s_next[buf] = nxt;
            if (!ETILE && nxt < total) issue(nxt, buf ^ 1);
        }

        // detection cells of the score tile's columns (as byte masks) and rows: copied from the per-level tables the host built
        // (orb_plan.cu); computing them here cost two integer divisions per entry and one more barrier per tile
        {
            const uint32_t* ct = reinterpret_cast<const uint32_t*>(coltab + L.ct_off) + (t.x0 >> 2);     // column x0 - 4 is entry x0
            const int16_t* rt = rowtab + L.rt_off + t.y0;                                                // row y0 - 1 is entry y0
            const int ctw = L.ct_len >> 2;
            for (int i = tid; i < 3 * FSW + FS_H; i += THREADS) {
                if (i < 3 * FSW) {
                    const int a = i / FSW;
                    mask3[i] = __ldg(ct + a * ctw + (i - a * FSW));
                } else rowcell[i - 3 * FSW] = __ldg(rt + (i - 3 * FSW));
            }
        }
        if (tid < 32) mbar_wait(&bar[ETILE ? 0 : buf], (uint32_t)((ETILE ? it : it >> 1) & 1));      // image tile has landed; one warp polls, the rest sleep in the barrier
        __syncthreads();
        const uint32_t* img = img2[ETILE ? 0 : buf];
        // Tiles on the right / bottom edge of the detection region are only partly filled (8 % of the tile area at 752x480):
        // tasks are numbered over the filled part only, so that a narrow tile takes fewer rounds instead of idle lanes.
        // vw x vh = detection pixels of the tile; the NMS walks nq 16-pixel groups per row and reads one score word more on
        // either side, one row more above and below.  Score words outside [nr) x [nw) keep stale values and are never read.
        const int vw = min(FT_W, L.xend - t.x0), vh = min(FT_H, L.yend - t.y0);       // xend / yend: size - 16, or further where an inner cell of a degenerate grid reaches past it
        const int nq = (vw + 15) >> 4, nwi = 4 * nq, nr = min(FS_H, vh + 2);
        const uint32_t inv_nq = ((1u << 20) + nq - 1) / nq;   // floor(task / nq) = task * inv >> 20, exact for task < 4000, nq <= 66

        // ---- corner strength: one task = 4 horizontally adjacent pixels ----
        // (score words 1 .. 4 nq; the two halo words of a row are one pixel each and have their own pass below)
        // Full tiles (the common case) keep compile-time task counts, partly filled ones take the run-time form (both instantiated).
        auto score_pass = [&](auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        const int ntask = FULL ? FS_H * (FT_W / 4) : nr * nwi;
        for (int task = tid; task < ntask; task += THREADS) {
            const int r = FULL ? task / (FT_W / 4) : (int)(((uint32_t)(task >> 2) * inv_nq) >> 20), g = task - r * (FULL ? FT_W / 4 : nwi) + 1;
            const uint32_t cm = FULL ? 0xffffffffu : reinterpret_cast<const uint32_t*>(m_in)[g];   // every column of a full tile is a detection column
            uint32_t outw = 0;
            if (ETILE) {
              if (cm != 0 && rowcell[r] >= 0) {
                // words 2g-2 .. 2g+3 of rows r..r+6 of the half-lane tile = pixels x-4 .. x+7 (x = first pixel of the task); the rows
                // three and two away only need words 2g-1 .. 2g+2
                const uint32_t* ep = et + r * EW + 2 * g;
                uint32_t W[7][6];
#pragma unroll
                for (int q = 0; q < 7; q++) {
                    const uint2 m = *reinterpret_cast<const uint2*>(ep + q * EW);
                    W[q][2] = m.x; W[q][3] = m.y;
                    if (q >= 2 && q <= 4) {
                        const uint2 lo = *reinterpret_cast<const uint2*>(ep + q * EW - 2), hi = *reinterpret_cast<const uint2*>(ep + q * EW + 2);
                        W[q][0] = lo.x; W[q][1] = lo.y; W[q][4] = hi.x; W[q][5] = hi.y;
                    } else { W[q][0] = 0; W[q][1] = ep[q * EW - 1]; W[q][4] = ep[q * EW + 2]; W[q][5] = 0; }
                }
                // pixel pair starting s pixels from x in row q: a word of the tile for even s, the upper half of one and the lower half of the next for odd s
#define S_(q, s) ((((s) + 4) & 1) == 0 ? W[q][((s) + 4) >> 1] : __byte_perm(W[q][((s) + 3) >> 1], W[q][((s) + 5) >> 1], 0x5432))
#define RING_E(o, R)                                                                                                     \
                R[0]  = S_(6, 0 + o);  R[1]  = S_(6, 1 + o);  R[2]  = S_(5, 2 + o);  R[3]  = S_(4, 3 + o);                  \
                R[4]  = S_(3, 3 + o);  R[5]  = S_(2, 3 + o);  R[6]  = S_(1, 2 + o);  R[7]  = S_(0, 1 + o);                  \
                R[8]  = S_(0, 0 + o);  R[9]  = S_(0, -1 + o); R[10] = S_(1, -2 + o); R[11] = S_(2, -3 + o);                 \
                R[12] = S_(3, -3 + o); R[13] = S_(4, -3 + o); R[14] = S_(5, -2 + o); R[15] = S_(6, -1 + o);
                const uint32_t vlo = W[3][2] & 0x00ff00ffu, vhi = W[3][3] & 0x00ff00ffu;
                uint32_t ring[16], Mn, Mx;
                RING_E(0, ring)
                const uint32_t h0 = S_(6, 2), h4 = S_(3, 5), h8 = S_(0, 2), h12 = S_(3, -1);
                const uint32_t nvp_lo = __vadd2(~vlo, c1n), vm1_lo = __vadd2(vlo, c1p);    // -(v + th) - 0x6400, v - th + 1 + 0x6400
                const uint32_t nvp_hi = __vadd2(~vhi, c1n), vm1_hi = __vadd2(vhi, c1p);
                bool any;
                {
                    const uint32_t bl = __vimax3_u16x2(ring[0], ring[4], __vmaxu2(ring[8], ring[12]));
                    const uint32_t dl = __vimin3_u16x2(ring[0], ring[4], __vminu2(ring[8], ring[12]));
                    const uint32_t bh = __vimax3_u16x2(h0, h4, __vmaxu2(h8, h12));
                    const uint32_t dh = __vimin3_u16x2(h0, h4, __vminu2(h8, h12));
                    any = (excess2(bl, dl, nvp_lo, vm1_lo) | excess2(bh, dh, nvp_hi, vm1_hi)) != 0;
                }
                if (any) {
                    arc_minmax_h(ring, Mn, Mx);
                    const uint32_t slo = excess2(Mn, Mx, nvp_lo, vm1_lo);
                    RING_E(2, ring)
                    arc_minmax_h(ring, Mn, Mx);
                    const uint32_t shi = excess2(Mn, Mx, nvp_hi, vm1_hi);
                    outw = __byte_perm(slo, shi, 0x6420) & cm;              // low byte of each of the four lanes
                }
#undef RING_E
#undef S_
              }
            } else
            if (cm != 0 && rowcell[r] >= 0) {
                const uint32_t* ip = img + r * FIW + g + 2;          // rows r..r+6 (y-3..y+3), words of cols x-4..x+7
                uint32_t w0[7], w1[7], w2[7];
#pragma unroll
                for (int q = 0; q < 7; q++) { w0[q] = ip[q * FIW]; w1[q] = ip[q * FIW + 1]; w2[q] = ip[q * FIW + 2]; }
                // Ring sample k for pixels (x, x+1) resp. (x+2, x+3), packed as two u16 lanes holding value*257
                // (byte duplicated): one PRMT straight from the two source words, order preserving.
#ifdef ORB_FAST_INTLANES
#define RPAIR(A, B, o, hi) __byte_perm(A, B, (hi) ? ((((o) + 3) << 12) | (((o) + 3) << 8) | (((o) + 2) << 4) | ((o) + 2)) \
                                                   : ((((o) + 1) << 12) | (((o) + 1) << 8) | ((o) << 4) | (o)))
#else
                // half lanes: 0x64 above the pixel byte.  Both bytes in one source word: one PRMT against the constant word;
                // the pair (byte 3 of A, byte 0 of B) needs both words as PRMT sources and gets its 0x64 from a LOP3.
                constexpr uint32_t K64 = 0x64646464u;
#define RP_(i0) ((i0) + 1 <= 3 ? __byte_perm(A_, K64, 0x4040 | (((i0) + 1) << 8) | (i0))                               \
                 : (i0) >= 4   ? __byte_perm(B_, K64, 0x4040 | (((i0) - 3) << 8) | ((i0) - 4))                         \
                               : ((__byte_perm(A_, B_, 0x0403) & 0x00ff00ffu) | 0x64006400u))
#define RPAIR(A, B, o, hi) ([&] { const uint32_t A_ = (A), B_ = (B); return RP_((o) + ((hi) ? 2 : 0)); }())
#endif
#define RING_ALL(hi, R)                                                                      \
                R[0]  = RPAIR(w1[6], w2[6], 0, hi);  R[1]  = RPAIR(w1[6], w2[6], 1, hi);     \
                R[2]  = RPAIR(w1[5], w2[5], 2, hi);  R[3]  = RPAIR(w1[4], w2[4], 3, hi);     \
                R[4]  = RPAIR(w1[3], w2[3], 3, hi);  R[5]  = RPAIR(w1[2], w2[2], 3, hi);     \
                R[6]  = RPAIR(w1[1], w2[1], 2, hi);  R[7]  = RPAIR(w1[0], w2[0], 1, hi);     \
                R[8]  = RPAIR(w1[0], w2[0], 0, hi);  R[9]  = RPAIR(w0[0], w1[0], 3, hi);     \
                R[10] = RPAIR(w0[1], w1[1], 2, hi);  R[11] = RPAIR(w0[2], w1[2], 1, hi);     \
                R[12] = RPAIR(w0[3], w1[3], 1, hi);  R[13] = RPAIR(w0[4], w1[4], 1, hi);     \
                R[14] = RPAIR(w0[5], w1[5], 2, hi);  R[15] = RPAIR(w0[6], w1[6], 3, hi);
                const uint32_t cw = w1[3];
                const uint32_t vlo = lo16x2(cw), vhi = hi16x2(cw);
                uint32_t ring[16], Mn, Mx;
                RING_ALL(0, ring)
                // quick reject (exact, never rejects a corner): every 9-arc contains a compass point (0,4,8,12), so a
                // corner needs one of them brighter than v+th or darker than v-th.  Flat areas leave here.
                const uint32_t h0 = RPAIR(w1[6], w2[6], 0, 1), h4 = RPAIR(w1[3], w2[3], 3, 1);
                const uint32_t h8 = RPAIR(w1[0], w2[0], 0, 1), h12 = RPAIR(w0[3], w1[3], 1, 1);
#ifdef ORB_FAST_INTLANES
                const uint32_t nvp_lo = __vadd2(~vlo, c1), vm1_lo = __vadd2(vlo, c1);      // -(v + th), v - th + 1
                const uint32_t nvp_hi = __vadd2(~vhi, c1), vm1_hi = __vadd2(vhi, c1);
#define LANEVAL(x) __byte_perm(x, 0, 0x4240)
#define ARC_MINMAX arc_minmax
#else
                const uint32_t nvp_lo = __vadd2(~vlo, c1n), vm1_lo = __vadd2(vlo, c1p);    // -(v + th) - 0x6400, v - th + 1 + 0x6400
                const uint32_t nvp_hi = __vadd2(~vhi, c1n), vm1_hi = __vadd2(vhi, c1p);
#define LANEVAL(x) (x)
#define ARC_MINMAX arc_minmax_h
#endif
                bool any;
                {
                    const uint32_t bl = __vimax3_u16x2(ring[0], ring[4], __vmaxu2(ring[8], ring[12]));
                    const uint32_t dl = __vimin3_u16x2(ring[0], ring[4], __vminu2(ring[8], ring[12]));
                    const uint32_t bh = __vimax3_u16x2(h0, h4, __vmaxu2(h8, h12));
                    const uint32_t dh = __vimin3_u16x2(h0, h4, __vminu2(h8, h12));
                    any = (excess2(LANEVAL(bl), LANEVAL(dl), nvp_lo, vm1_lo) | excess2(LANEVAL(bh), LANEVAL(dh), nvp_hi, vm1_hi)) != 0;
                }
                if (any) {
                    ARC_MINMAX(ring, Mn, Mx);
                    const uint32_t slo = excess2(LANEVAL(Mn), LANEVAL(Mx), nvp_lo, vm1_lo);
                    RING_ALL(1, ring)
                    ARC_MINMAX(ring, Mn, Mx);
                    const uint32_t shi = excess2(LANEVAL(Mn), LANEVAL(Mx), nvp_hi, vm1_hi);
                    outw = __byte_perm(slo, shi, 0x6420) & cm;              // low byte of each of the four lanes
                }
#undef LANEVAL
#undef ARC_MINMAX
#undef RING_ALL
#undef RPAIR
#ifndef ORB_FAST_INTLANES
#undef RP_
#endif
            }
            sc[r * FSW + g] = outw;
        }
        };
        // ---- halo columns: the NMS of the tile's first / last pixel column needs the strength of ONE pixel to the left (x0 - 1,
        //      byte 3 of score word 0) and to the right (x0 + 16 nq, byte 0 of score word 4 nq + 1).  Scoring those two words like
        //      the others would spend 2 of 18 tasks per row on 2 useful pixels of 8; here one task scores the pair (left, right)
        //      of a row in the two 16-bit lanes: 1 task per row instead of 2, and with 2 pixels instead of 4. ----
        auto halo_pass = [&](auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        const int nrh = FULL ? FS_H : nr, gr = (FULL ? FT_W / 4 : nwi) + 1;
        for (int r = tid; r < nrh; r += THREADS) {
            const uint32_t mL = m_in[3], mR = m_in[4 * gr];
            uint32_t sL = 0, sR = 0;
            if (rowcell[r] >= 0 && (mL | mR) != 0) {
                const uint32_t* ip = img + r * FIW + 3;              // image words 3,4 hold columns x0-4 .. x0+3; words gr+2, gr+3 hold x0+16nq-4 .. +3
                uint32_t a[7], b[7], c[7], d[7];
#pragma unroll
                for (int q = 0; q < 7; q++) { a[q] = ip[q * FIW]; b[q] = ip[q * FIW + 1]; c[q] = ip[q * FIW + gr - 1]; d[q] = ip[q * FIW + gr]; }
                // sample (dx, row q) of both pixels, each duplicated into its 16-bit lane: left pixel = byte 3 of a[], right = byte 0 of d[]
#define HP(q, dx) __byte_perm((dx) <= 0 ? a[q] : b[q], (dx) < 0 ? c[q] : d[q],                                               \
                              (((dx) <= 0 ? 3 + (dx) : (dx) - 1) * 0x11) | ((4 + ((dx) < 0 ? 4 + (dx) : (dx))) * 0x1100))
                uint32_t ring[16], Mn, Mx;
                ring[0]  = HP(6, 0);  ring[1]  = HP(6, 1);  ring[2]  = HP(5, 2);  ring[3]  = HP(4, 3);
                ring[4]  = HP(3, 3);  ring[5]  = HP(2, 3);  ring[6]  = HP(1, 2);  ring[7]  = HP(0, 1);
                ring[8]  = HP(0, 0);  ring[9]  = HP(0, -1); ring[10] = HP(1, -2); ring[11] = HP(2, -3);
                ring[12] = HP(3, -3); ring[13] = HP(4, -3); ring[14] = HP(5, -2); ring[15] = HP(6, -1);
                const uint32_t v = HP(3, 0) & 0x00ff00ffu;
#undef HP
                const uint32_t nvp = __vadd2(~v, c1), vm1 = __vadd2(v, c1);
                const uint32_t bl = __vimax3_u16x2(ring[0], ring[4], __vmaxu2(ring[8], ring[12]));
                const uint32_t dl = __vimin3_u16x2(ring[0], ring[4], __vminu2(ring[8], ring[12]));
                if (excess2(__byte_perm(bl, 0, 0x4240), __byte_perm(dl, 0, 0x4240), nvp, vm1) != 0) {
                    arc_minmax(ring, Mn, Mx);
                    const uint32_t sp2 = excess2(__byte_perm(Mn, 0, 0x4240), __byte_perm(Mx, 0, 0x4240), nvp, vm1);
                    sL = (sp2 & 0xffu) & mL; sR = (sp2 >> 16) & mR;
                }
            }
            sc[r * FSW] = sL << 24; sc[r * FSW + gr] = sR;
        }
        };
        const bool full = vw == FT_W && nr == FS_H;
        if (ETILE) {
            // re-encode the raw tile: raw word 3 + j of a row (columns x0-4+4j ..) -> words 2j, 2j+1 of the half-lane tile
            // thread = (raw word column, row residue): no index arithmetic per word, compile-time row offsets (the flat form with a
            // division per word was 7 % of the kernel's instructions)
            {
                constexpr int EC = EW / 2, ER = THREADS / EC;          // 18 word columns, 7 rows per pass
                if (tid < EC * ER) {
                    const int er0 = tid / EC, ej = tid - er0 * EC;
                    const uint32_t* src = img + er0 * FIW + 3 + ej;
                    uint32_t* dst = et + er0 * EW + 2 * ej;
#pragma unroll
                    for (int k = 0; k < (FI_H + ER - 1) / ER; k++) {
                        if ((k + 1) * ER <= FI_H || er0 + k * ER < FI_H) {
                            const uint32_t rw = src[k * ER * FIW];
                            *reinterpret_cast<uint2*>(dst + k * ER * EW) = make_uint2(__byte_perm(rw, 0x64646464u, 0x4140), __byte_perm(rw, 0x64646464u, 0x4342));
                        }
                    }
                }
            }
            if (full) halo_pass(std::true_type{}); else halo_pass(std::false_type{});
            __syncthreads();       // the raw tile is dead: fetch the next item's into the same buffer
            if (tid == 0 && s_next[buf] < total) issue(s_next[buf], buf ^ 1);
        }
        if (full) score_pass(std::true_type{}); else score_pass(std::false_type{});
        if (!ETILE) { if (full) halo_pass(std::true_type{}); else halo_pass(std::false_type{}); }
        __syncthreads();

        // ---- NMS restricted to the pixel's own cell: one task = 16 output pixels (four words): one 128-bit store
        //      of responses to the score map and one 16-bit store to the survivor bitmap ----
        uint8_t* out = nms + (size_t)f * fbytes + L.plane_off;
        uint8_t* bm = bitmap + (size_t)f * plan->bm_total + L.bm_off;
        auto nms_pass = [&](auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        const int ntask = FULL ? FT_H * (FT_W / 16) : vh * nq;
        for (int task = tid; task < ntask; task += THREADS) {
            const int ro = FULL ? task / (FT_W / 16) : (int)(((uint32_t)task * inv_nq) >> 20), q4 = task - ro * (FULL ? FT_W / 16 : nq);
            const int r = ro + 1;
            const int rc = rowcell[r];
            const bool up = rowcell[r - 1] == rc, dn = rowcell[r + 1] == rc;
            uint32_t vv[4];
            uint32_t bits = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int g = q4 * 4 + j + 1;
                const uint32_t* sp = sc + r * FSW + g;
                const uint32_t c = sp[0];
                uint32_t v = 0;
                if (c) {
                    const uint32_t ml = reinterpret_cast<const uint32_t*>(m_l)[g], mr = reinterpret_cast<const uint32_t*>(m_r)[g];
#ifndef ORB_NMS_LOBYTE
                    // A 16-bit unsigned max looks at a lane's low byte only on ties of the high byte, so a lane may carry its pixel in
                    // the HIGH byte and anything below it.  The score word itself is then the
                    // lane pair of pixels (1, 3), `word << 8` that of pixels (0, 2), and every neighbour is a shift or funnel shift of
                    // the nine words: no PRMT unpacking (16 per word in the ORB_NMS_LOBYTE form below).  For the strict comparison the lanes are
                    // halved first (score << 7 against max << 7 | 0x7f) so that adding 0x7fff cannot carry into the next lane.
                    // Host model: tests/test_kernel_arith_models.py, test_nms_high_byte_lanes_model.  Measured on B200 (752x480, 256 frames):
                    // 0.957 -> 0.933 ms against the low-byte form, whole GPU suite bit-exact with either.
                    const uint32_t pc = sp[-1], nc = sp[1];
                    uint32_t ul = sp[-FSW - 1], uc = sp[-FSW], ur = sp[-FSW + 1], dl = sp[FSW - 1], dc = sp[FSW], dr = sp[FSW + 1];
                    if (!up) { ul = 0u; uc = 0u; ur = 0u; }
                    if (!dn) { dl = 0u; dc = 0u; dr = 0u; }
                    const uint32_t c8 = c << 8, u8 = uc << 8, d8 = dc << 8;
                    const uint32_t Lo = __vimax3_u16x2(c8, u8, d8) & ml;                                      // pixels 1, 3: left neighbours
                    const uint32_t Ro = __vimax3_u16x2(__funnelshift_r(c, nc, 8), __funnelshift_r(uc, ur, 8), __funnelshift_r(dc, dr, 8)) & mr;
                    const uint32_t Mo = __vimax3_u16x2(Lo, Ro, __vmaxu2(uc, dc));
                    const uint32_t Le = __vimax3_u16x2(__funnelshift_r(pc, c, 16), __funnelshift_r(ul, uc, 16), __funnelshift_r(dl, dc, 16)) & (ml << 8);
                    const uint32_t Re = __vimax3_u16x2(c, uc, dc) & (mr << 8);                                // pixels 0, 2: right neighbours
                    const uint32_t Me = __vimax3_u16x2(Le, Re, __vmaxu2(u8, d8));
                    const uint32_t ao = (c >> 1) & 0x7f807f80u, bo = ((Mo >> 1) | 0x007f007fu) & 0x7fff7fffu;
                    const uint32_t ae = (c8 >> 1) & 0x7f807f80u, be = ((Me >> 1) | 0x007f007fu) & 0x7fff7fffu;
                    const uint32_t yo = ao - __vminu2(ao, bo) + 0x7fff7fffu, ye = ae - __vminu2(ae, be) + 0x7fff7fffu;
                    uint32_t m4;
                    asm("prmt.b32 %0, %1, %2, 0xfbd9;" : "=r"(m4) : "r"(ye), "r"(yo));         // byte i = 0xff iff pixel i survives
#else
                    uint32_t nb[8];
                    nb[0] = __funnelshift_r(sp[-1], c, 24) & ml;
                    nb[1] = __funnelshift_r(c, sp[1], 8) & mr;
                    const uint32_t ul = sp[-FSW - 1], uc = sp[-FSW], ur = sp[-FSW + 1];
                    nb[2] = up ? uc : 0u;
                    nb[3] = up ? (__funnelshift_r(ul, uc, 24) & ml) : 0u;
                    nb[4] = up ? (__funnelshift_r(uc, ur, 8) & mr) : 0u;
                    const uint32_t dl = sp[FSW - 1], dc = sp[FSW], dr = sp[FSW + 1];
                    nb[5] = dn ? dc : 0u;
                    nb[6] = dn ? (__funnelshift_r(dl, dc, 24) & ml) : 0u;
                    nb[7] = dn ? (__funnelshift_r(dc, dr, 8) & mr) : 0u;
                    const uint32_t mlo = __vimax3_u16x2(__vimax3_u16x2(lo16x2(nb[0]), lo16x2(nb[1]), lo16x2(nb[2])),
                                                        __vimax3_u16x2(lo16x2(nb[3]), lo16x2(nb[4]), lo16x2(nb[5])),
                                                        __vmaxu2(lo16x2(nb[6]), lo16x2(nb[7])));
                    const uint32_t mhi = __vimax3_u16x2(__vimax3_u16x2(hi16x2(nb[0]), hi16x2(nb[1]), hi16x2(nb[2])),
                                                        __vimax3_u16x2(hi16x2(nb[3]), hi16x2(nb[4]), hi16x2(nb[5])),
                                                        __vmaxu2(hi16x2(nb[6]), hi16x2(nb[7])));
                    // strictly greater than all 8 neighbours (a score of 0 never passes), for the four pixels at once: in 16-bit lanes
                    // score - min(score, neighbour max) is positive exactly for a survivor; adding 0x7fff moves that into the lane's
                    // sign bit and one PRMT in sign-replication mode turns the four sign bits into byte masks.
                    // survivors: excess -> OpenCV's response (T - 1 = excess + th - 1, at most 254, so the byte-wise add cannot carry)
                    const uint32_t clo = lo16x2(c), chi = hi16x2(c);
                    const uint32_t ylo = clo - __vminu2(clo, mlo) + 0x7fff7fffu, yhi = chi - __vminu2(chi, mhi) + 0x7fff7fffu;
                    uint32_t m4;
                    asm("prmt.b32 %0, %1, %2, 0xfdb9;" : "=r"(m4) : "r"(ylo), "r"(yhi));       // byte i = 0xff iff pixel i survives
#endif
                    v = (c + th_m1 * 0x01010101u) & m4;
                    bits |= (((m4 & 0x01010101u) * 0x01020408u) >> 24) << (4 * j);             // mask bits 0, 8, 16, 24 -> bits 0..3
                }
                vv[j] = v;
            }
            const int py = t.y0 + ro + ORB_EDGE, px = t.x0 + q4 * 16 + ORB_EDGE;
            // The response map is only ever read where the bitmap has a bit set (k_cell_compact), so a 16-pixel group without a
            // survivor (about 60 % of them) writes nothing: no zeros to DRAM, stale bytes there are never looked at.
            if (bits == 0) { /* bitmap word below is still written: it is what says "nothing here" */ }
            else if (FULL || (py < L.prows && px + 15 < L.stride))       // a full tile lies inside the detection region
                *reinterpret_cast<uint4*>(out + (size_t)py * L.stride + px) = make_uint4(vv[0], vv[1], vv[2], vv[3]);
            else if (py < L.prows) {
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (px + 4 * j + 3 < L.stride) *reinterpret_cast<uint32_t*>(out + (size_t)py * L.stride + px + 4 * j) = vv[j];
            }
            // bitmap: bit i of row y (ROI) = ROI column 16+i ; tiles start at ROI x0 = 16 + 64k
            *reinterpret_cast<uint16_t*>(bm + (size_t)(t.y0 + ro - ORB_EDGE) * L.bm_pitch + ((t.x0 - ORB_EDGE) >> 3) + q4 * 2) = (uint16_t)bits;
        }
        };
        if (full && vh == FT_H) nms_pass(std::true_type{}); else nms_pass(std::false_type{});
        __syncthreads();       // masks / score tile are rewritten by the next item
        item = s_next[buf];
    }
}

// ------------------------------------------------------------------ K3
// One warp per (frame, cell): scans the cell's detection rectangle of the NMS map in raster
// order and appends survivors with ballot/popc prefix sums, so the list order is exactly the
// order cv::FAST emits (y, then x).  Each lane reads 4 pixels (one 32-bit word) per step.  Then
// applies the reference's fallback: if fewer than 4 survive at fastTh the cell is re-detected at
// threshold 7 (src/ORBextractor.cc:609-614) — both sets are sub-sequences of the th_lo list
// ("one-pass fallback").
// record = score<<24 | y_local<<12 | x_local   (cell-image coordinates, as cv::FAST reports)
// A warp's life here is short (one or two steps over a cell) and mostly a chain of dependent memory round trips: cell count -> cell geometry ->
// level geometry -> bitmap word -> response.  The plan's part of it (ORB_COMPACT_ARGS, default) travels in the kernel's parameter block instead,
// which leaves three round trips.
#ifndef ORB_COMPACT_ARGS
#define ORB_COMPACT_ARGS 1
#endif
struct CompactGeom { int plane_off[ORB_MAX_LEVELS], stride[ORB_MAX_LEVELS], bm_off[ORB_MAX_LEVELS], bm_pitch[ORB_MAX_LEVELS]; int ncells, bm_total, cand_total, fast_th; };
__global__ void __launch_bounds__(256, 8)      // 32 registers (20 bytes spilled): latency bound, 0.128 -> 0.108 ms per 256 frames against 47 registers
k_cell_compact(const uint8_t* __restrict__ nms, const uint8_t* __restrict__ bitmap, size_t fbytes, const Plan* __restrict__ plan,
               const CellGeom* __restrict__ cells, uint32_t* __restrict__ cand, int* __restrict__ ntotal, const __grid_constant__ CompactGeom cg)
{
    pdl_sync(5);
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
#if ORB_COMPACT_ARGS
    const int ncells = cg.ncells;
    if (warp >= ncells) return;
    const int f = blockIdx.y;
    const CellGeom g = cells[warp];
    struct { int plane_off, stride, bm_off, bm_pitch; } L = { cg.plane_off[g.level], cg.stride[g.level], cg.bm_off[g.level], cg.bm_pitch[g.level] };
    const int bm_total = cg.bm_total, cand_total = cg.cand_total, thP = cg.fast_th;
#else
    const int ncells = plan->ncells;
    if (warp >= ncells) return;
    const int f = blockIdx.y;
    const CellGeom g = cells[warp];
    const LevelGeom& L = plan->L[g.level];
    const int bm_total = plan->bm_total, cand_total = plan->cand_total, thP = plan->fast_th;
#endif
    const uint8_t* map = nms + (size_t)f * fbytes + L.plane_off + (size_t)ORB_EDGE * L.stride + ORB_EDGE;
    const uint32_t* bm = reinterpret_cast<const uint32_t*>(bitmap + (size_t)f * bm_total + L.bm_off);
    uint32_t* out = cand + (size_t)f * cand_total + g.cand_off;
    int count = 0, nP = 0, n7 = 0;
    // The rectangle is walked as (row, 32-pixel bitmap word) items in raster order; a lane owns one item per step.
    const int b0 = g.x0 - ORB_EDGE, b1 = g.x1 - ORB_EDGE;            // bit range [b0, b1) of a bitmap row
    const int w0 = b0 >> 5, wpr = b1 > b0 ? ((b1 - 1) >> 5) - w0 + 1 : 0;
    const int nitems = (g.y1 - g.y0) * wpr;
    const int pitchw = L.bm_pitch >> 2;
    for (int i0 = 0; i0 < nitems; i0 += 32) {
        const int i = i0 + lane;
        uint32_t word = 0;
        int y = 0, xbase = 0;
        if (i < nitems) {
            const int rr = i / wpr, wi = i - rr * wpr;
            y = g.y0 + rr;
            word = bm[(size_t)(y - ORB_EDGE) * pitchw + w0 + wi];
            const int bit0 = (w0 + wi) << 5;                         // bit index of the word's LSB
            if (bit0 < b0) word &= 0xffffffffu << (b0 - bit0);
            if (bit0 + 32 > b1) word &= 0xffffffffu >> (bit0 + 32 - b1);
            xbase = bit0 + ORB_EDGE;                                 // ROI x of the word's LSB
        }
        const int cnt = __popc(word);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int tv = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += tv; }
        int pos = count + incl - cnt;
        const uint8_t* row = map + (size_t)y * L.stride;
        while (word) {
            const int b = __ffs(word) - 1;
            word &= word - 1;
            const int x = xbase + b;
            const uint32_t s = row[x];
            nP += (int)s >= thP;
            n7 += s >= 7;
            out[pos++] = (s << 24) | ((uint32_t)(y - g.iniy) << 12) | (uint32_t)(x - g.inix);
        }
        count += __shfl_sync(0xffffffffu, incl, 31);
    }
    nP = __reduce_add_sync(0xffffffffu, nP);
    n7 = __reduce_add_sync(0xffffffffu, n7);
    const uint32_t lt = (1u << lane) - 1;
    const int thr = nP > 3 ? thP : 7;
    const int want = nP > 3 ? nP : n7;
    if (want < count) {           // drop the weaker corners, keeping raster order (in place, warp-synchronous)
        __syncwarp();
        int w = 0;
        for (int b = 0; b < count; b += 32) {
            const int i = b + lane;
            const uint32_t r = i < count ? out[i] : 0u;
            const bool k = i < count && (int)(r >> 24) >= thr;
            const uint32_t m = __ballot_sync(0xffffffffu, k);
            __syncwarp();
            if (k) out[w + __popc(m & lt)] = r;
            w += __popc(m);
            __syncwarp();
        }
        count = w;
    }
    if (lane == 0) ntotal[(size_t)f * ncells + warp] = count;
}

// Calls of a few frames (single-frame latency): one CTA per (frame, cell) instead of one warp.  A warp's serial walk over a whole
// cell — a dependent bitmap load, then a dependent response load per survivor, per 32-word step — was 21 of the 116 us of one 640x480
// frame.  Here the eight warps take an eighth of the cell's rows each and build their part of the list in shared memory; the
// per-warp counts of survivors at fastTh and at 7 then give every warp its place in the cell's list (raster order = warp order), and
// the fallback rule (:609-614) is applied while the parts are written out.  A part that overflows its segment (more than CCW_SEG
// survivors in an eighth of a cell) sends the cell to the serial walk of the batch kernel, done by warp 0.
constexpr int CCW_SEG = 512;
__global__ void __launch_bounds__(256, 2)
k_cell_compact_wide(const uint8_t* __restrict__ nms, const uint8_t* __restrict__ bitmap, size_t fbytes, const Plan* __restrict__ plan,
                    const CellGeom* __restrict__ cells, uint32_t* __restrict__ cand, int* __restrict__ ntotal)
{
    pdl_sync(5);
    __shared__ uint32_t seg[8][CCW_SEG];
    __shared__ int s_cnt[8], s_nP[8], s_n7[8];
    const int cell = blockIdx.x, f = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const CellGeom g = cells[cell];
    const LevelGeom& L = plan->L[g.level];
    const uint8_t* map = nms + (size_t)f * fbytes + L.plane_off + (size_t)ORB_EDGE * L.stride + ORB_EDGE;
    const uint32_t* bm = reinterpret_cast<const uint32_t*>(bitmap + (size_t)f * plan->bm_total + L.bm_off);
    uint32_t* out = cand + (size_t)f * plan->cand_total + g.cand_off;
    const int thP = plan->fast_th;
    const int b0 = g.x0 - ORB_EDGE, b1 = g.x1 - ORB_EDGE;            // bit range [b0, b1) of a bitmap row
    const int w0 = b0 >> 5, wpr = b1 > b0 ? ((b1 - 1) >> 5) - w0 + 1 : 0;
    const int pitchw = L.bm_pitch >> 2;
    const uint32_t lt = (1u << lane) - 1;
    auto walk = [&](int ya, int yb, auto put) {                      // rows [ya, yb) in raster order; returns (count, nP, n7) summed over the warp
        int count = 0, nP = 0, n7 = 0;
        const int nitems = max(yb - ya, 0) * wpr;
        for (int i0 = 0; i0 < nitems; i0 += 32) {
            const int i = i0 + lane;
            uint32_t word = 0;
            int y = 0, xbase = 0;
            if (i < nitems) {
                const int rr = i / wpr, wi = i - rr * wpr;
                y = ya + rr;
                word = bm[(size_t)(y - ORB_EDGE) * pitchw + w0 + wi];
                const int bit0 = (w0 + wi) << 5;
                if (bit0 < b0) word &= 0xffffffffu << (b0 - bit0);
                if (bit0 + 32 > b1) word &= 0xffffffffu >> (bit0 + 32 - b1);
                xbase = bit0 + ORB_EDGE;
            }
            const int cnt = __popc(word);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int tv = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += tv; }
            int pos = count + incl - cnt;
            const uint8_t* row = map + (size_t)y * L.stride;
            while (word) {
                const int b = __ffs(word) - 1;
                word &= word - 1;
                const int x = xbase + b;
                const uint32_t s = row[x];
                nP += (int)s >= thP;
                n7 += s >= 7;
                put(pos++, (s << 24) | ((uint32_t)(y - g.iniy) << 12) | (uint32_t)(x - g.inix));
            }
            count += __shfl_sync(0xffffffffu, incl, 31);
        }
        nP = __reduce_add_sync(0xffffffffu, nP);
        n7 = __reduce_add_sync(0xffffffffu, n7);
        return make_int3(count, nP, n7);
    };
    const int rows = g.y1 - g.y0, rpw = (rows + 7) >> 3;
    {
        const int ya = g.y0 + warp * rpw, yb = min(ya + rpw, g.y1);
        uint32_t* sg = seg[warp];
        const int3 r = walk(ya, yb, [&](int pos, uint32_t rec) { if (pos < CCW_SEG) sg[pos] = rec; });
        if (lane == 0) { s_cnt[warp] = r.x; s_nP[warp] = r.y; s_n7[warp] = r.z; }
    }
    __syncthreads();
    int nP = 0, n7 = 0, before = 0;
    bool overflow = false;
#pragma unroll
    for (int k = 0; k < 8; k++) { nP += s_nP[k]; n7 += s_n7[k]; overflow |= s_cnt[k] > CCW_SEG; }
    const bool useP = nP > 3;
    const int thr = useP ? thP : 7;
    if (overflow) {                   // the batch kernel's walk and in-place fallback filter, by one warp
        if (warp != 0) return;
        const int3 r = walk(g.y0, g.y1, [&](int pos, uint32_t rec) { out[pos] = rec; });
        int count = r.x;
        const int want = useP ? r.y : r.z;
        if (want < count) {
            __syncwarp();
            int w = 0;
            for (int b = 0; b < count; b += 32) {
                const int i = b + lane;
                const uint32_t rec = i < count ? out[i] : 0u;
                const bool k = i < count && (int)(rec >> 24) >= thr;
                const uint32_t m = __ballot_sync(0xffffffffu, k);
                __syncwarp();
                if (k) out[w + __popc(m & lt)] = rec;
                w += __popc(m);
                __syncwarp();
            }
            count = w;
        }
        if (lane == 0) ntotal[(size_t)f * plan->ncells + cell] = count;
        return;
    }
    // every record with a response >= thr stays (all of them when none is weaker: the serial form's want == count)
#pragma unroll
    for (int k = 0; k < 8; k++) if (k < warp) before += useP ? s_nP[k] : s_n7[k];
    {
        const uint32_t* sg = seg[warp];
        const int n = s_cnt[warp];
        int w = before;
        for (int b = 0; b < n; b += 32) {
            const int i = b + lane;
            const uint32_t rec = i < n ? sg[i] : 0u;
            const bool k = i < n && (int)(rec >> 24) >= thr;
            const uint32_t m = __ballot_sync(0xffffffffu, k);
            if (k) out[w + __popc(m & lt)] = rec;
            w += __popc(m);
        }
    }
    if (threadIdx.x == 0) ntotal[(size_t)f * plan->ncells + cell] = useP ? nP : n7;
}

// ------------------------------------------------------------------ K4
// One CTA per (frame, level).  Thread 0 replays the quota redistribution loop (:622-670); every
// cell then runs retainBest (= libstdc++ introselect, first n survivors, see introselect.h) on
// its own list in parallel; the survivors are concatenated in cell order, converted to level
// coordinates, and capped to nDesired by a second introselect (:697-701).
// level record = score<<32 | y<<16 | x  (level ROI coordinates)
constexpr int SEL_STAGE = 6144;     // candidate records staged in shared memory per (frame, level)

// HARRIS_SCORE (src/ORBextractor.cc:616-620 -> HarrisResponses :79-120, blockSize 7, k = 0.04): one thread per candidate of the
// per-cell lists replaces the FAST response by the Harris measure; integer moments a, b, c over the 7x7 block of Sobel-like
// gradients, then the float expression of :117-118 with every operation rounded on its own.
__global__ void __launch_bounds__(256)
k_harris(const uint8_t* __restrict__ planes, size_t fbytes, const Plan* __restrict__ plan, const CellGeom* __restrict__ cells,
         const uint32_t* __restrict__ cand, const int* __restrict__ ntotal, unsigned long long* __restrict__ cand64)
{
    pdl_sync(6);
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= plan->ncells) return;
    const int f = blockIdx.y;
    const CellGeom g = cells[warp];
    const LevelGeom& L = plan->L[g.level];
    const int n = ntotal[(size_t)f * plan->ncells + warp];
    const uint8_t* roi = planes + (size_t)f * fbytes + L.plane_off + (size_t)ORB_EDGE * L.stride + ORB_EDGE;
    const uint32_t* in = cand + (size_t)f * plan->cand_total + g.cand_off;
    unsigned long long* out = cand64 + (size_t)f * plan->cand_total + g.cand_off;
    const int step = L.stride;
    float scale = __fmul_rn((float)((1 << 2) * 7), 255.0f);
    scale = __fdiv_rn(1.0f, scale);
    const float scale_sq_sq = __fmul_rn(__fmul_rn(__fmul_rn(scale, scale), scale), scale);
    for (int i = lane; i < n; i += 32) {
        const uint32_t r = in[i];
        const int x = (int)(r & 0xfff) + g.inix, y = (int)((r >> 12) & 0xfff) + g.iniy;      // level (ROI) coordinates
        const uint8_t* p0 = roi + (size_t)(y - 3) * step + (x - 3);
        int a = 0, b = 0, c = 0;
        for (int yy = 0; yy < 7; yy++) {
            const uint8_t* pu = p0 + (ptrdiff_t)(yy - 1) * step, *pc = pu + step, *pd = pc + step;
            int u0 = pu[-1], u1 = pu[0], c0 = pc[-1], c1 = pc[0], d0 = pd[-1], d1 = pd[0];
#pragma unroll
            for (int xx = 0; xx < 7; xx++) {
                const int u2 = pu[xx + 1], c2 = pc[xx + 1], d2 = pd[xx + 1];
                const int Ix = (c2 - c0) * 2 + (u2 - u0) + (d2 - d0);
                const int Iy = (d1 - u1) * 2 + (d0 - u0) + (d2 - u2);
                a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
                u0 = u1; u1 = u2; c0 = c1; c1 = c2; d0 = d1; d1 = d2;
            }
        }
        const float fa = (float)a, fb = (float)b, fc = (float)c;
        const float sum = __fadd_rn(fa, fb);
        const float resp = __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc)), __fmul_rn(__fmul_rn(0.04f, sum), sum)), scale_sq_sq);
        out[i] = ((unsigned long long)__float_as_uint(resp) << 32) | (r & 0xffffffu);
    }
}

// ---- warp-cooperative, exact replay of std::nth_element (introselect.h explains why the partition may be evaluated in parallel) ----
constexpr int SEL_WCAP = 512;       // candidates one warp stages in shared memory (longer cell lists are selected in place in global memory); 1024: 0.140 ms, 512: 0.104, 256: 0.110 per 256 frames
#ifndef ORB_SEL_WARPS
#define ORB_SEL_WARPS 8
#define ORB_SEL_SERIAL_BELOW 8
#endif
constexpr int SEL_WARPS = ORB_SEL_WARPS, SEL_WARPS_WIDE = 32, SEL_SERIAL_BELOW = ORB_SEL_SERIAL_BELOW;   // below that range length lane 0 finishes alone

template <typename T, typename C>
__device__ __forceinline__ int warp_partition(T* v, int first, int last, C lt, unsigned short* scratch, int lane)
{
    const T pivot = v[first];
    const unsigned below = (1u << lane) - 1u;
    int nB = 0;                                                     // b_1, b_2, ...: right-scan stoppers, from the right
    for (int base = last - 1; base > first; base -= 128) {          // four chunks per trip: the loads are independent
        bool isB[4];
        unsigned bal[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const int i = base - 32 * u - lane; isB[u] = i > first && !lt(pivot, v[i]); }
#pragma unroll
        for (int u = 0; u < 4; u++) bal[u] = __ballot_sync(0xffffffffu, isB[u]);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (isB[u]) scratch[nB + __popc(bal[u] & below)] = (unsigned short)(base - 32 * u - lane);
            nB += __popc(bal[u]);
        }
    }
    __syncwarp();
    int m = 0, cut = INT_MAX;
    for (int base = first + 1; base < last; base += 32) {
        const int i = base + lane;
        const bool isA = i < last && !lt(v[i], pivot);                // left-scan stopper
        const unsigned bal = __ballot_sync(0xffffffffu, isA);
        const int rank = m + __popc(bal & below);
        const int partner = (isA && rank < nB) ? (int)scratch[rank] : -1;
        const bool ok = isA && partner > i;                          // a_k < b_k: the scan would swap this pair
        const unsigned okm = __ballot_sync(0xffffffffu, ok), fail = bal & ~okm;
        if (ok) { const T t = v[i]; v[i] = v[partner]; v[partner] = t; }
        m += __popc(okm);
        if (fail) { cut = base + __ffs(fail) - 1; break; }
        __syncwarp();      // a partner may lie in the next chunk: order this trip's stores before the next trip's loads
    }
    __syncwarp();
    const int bm = m > 0 ? (int)scratch[m - 1] : INT_MAX;
    cut = min(cut, bm);
    return cut == INT_MAX ? last : cut;
}

template <typename T, typename C>
__device__ __forceinline__ void warp_nth_element(T* v, int n, int nth, C lt, unsigned short* scratch, int lane)
{
    if (n <= 0 || nth >= n) return;
    int first = 0, last = n, depth = orbsel::depth_limit(n);
    while (last - first >= SEL_SERIAL_BELOW && depth > 0) {
        --depth;
        const int a = first + 1, b = first + (last - first) / 2, c = last - 1;      // median of three -> first
        const T va = v[a], vb = v[b], vc = v[c];
        int pick;
        if (lt(va, vb)) pick = lt(vb, vc) ? b : (lt(va, vc) ? c : a);
        else pick = lt(va, vc) ? a : (lt(vb, vc) ? c : b);
        __syncwarp();
        if (lane == 0) orbsel::swp(v, first, pick);
        __syncwarp();
        const int cut = warp_partition(v, first, last, lt, scratch, lane);
        if (cut <= nth) first = cut; else last = cut;
    }
    if (lane == 0) orbsel::nth_element_from(v, first, last, nth, depth, lt);       // short ranges, the heap-select fallback, the final insertion sort
    __syncwarp();
}

// Selection: CTA per (frame, level), warp per cell.  HARRIS: 64-bit records (float response | position) in global memory.
// WARPS: 8 for batches (seven CTAs per SM stay resident); WARPS_WIDE for a handful of frames, where the (frame, level) CTAs do
// not fill the machine and the kernel's time is the serial walk of one CTA's warps over the level's cells (single-frame latency).
template <bool HARRIS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_select_fast(const Plan* __restrict__ plan, const CellGeom* __restrict__ cells, uint32_t* __restrict__ cand,
              unsigned long long* __restrict__ cand64, const int* __restrict__ ntotal,
              unsigned long long* __restrict__ lvl, int* __restrict__ nkept, int* __restrict__ status, uint8_t* __restrict__ spare, size_t fbytes)
{
    pdl_sync(7);
    // sel_list_cap u64 | WARPS x SEL_WCAP u32 | WARPS x SEL_WCAP u16 | per-cell tables sized for the plan's largest grid
    // (sel_cells_cap: static arrays of ORB_MAX_CELLS_LEVEL entries cost 1 % of the whole pipeline in residency next to k_blur)
    extern __shared__ unsigned long long s_list[];
    const int level = blockIdx.x, f = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const LevelGeom& L = plan->L[level];
    const int nCells = L.ncells;
    const CellGeom* cg = cells + L.cell_base;
    const int* nt = ntotal + (size_t)f * plan->ncells + L.cell_base;
    uint32_t* s_wbuf = reinterpret_cast<uint32_t*>(s_list + plan->sel_list_cap);
    unsigned short* s_scr = reinterpret_cast<unsigned short*>(s_wbuf + WARPS * SEL_WCAP);
    const int ccap = plan->sel_cells_cap;                      // multiple of 4
    int* s_total = reinterpret_cast<int*>(s_scr + WARPS * SEL_WCAP);
    int* s_retain = s_total + ccap;
    int* s_off = s_retain + ccap;                              // ccap + 4 entries
    unsigned char* noMore = reinterpret_cast<unsigned char*>(s_off + ccap + 4);
    unsigned char* s_skip = noMore + ccap;                     // cg[c].skipped, fetched by all threads: thread 0's quota loop below would pay one global-load latency per cell
    for (int c = tid; c < nCells; c += blockDim.x) { s_total[c] = nt[c]; s_skip[c] = (unsigned char)(cg[c].skipped != 0); }
    __syncthreads();
    if (warp == 0) {                                           // quota redistribution, src/ORBextractor.cc:622-670
        // by one WARP: every pass of the reference's loop treats the cells independently of each other given nNew, and what it carries
        // from pass to pass are two integer sums — so the lanes take the cells and the sums are warp reductions (a single thread pays a
        // shared-memory round trip per step: 5-9 us of the 24 us this kernel took for one frame)
        const int nfc = L.nfCell;
        int nNoMore = 0, nToDistribute = 0;
        for (int c = lane; c < nCells; c += 32) {
            int ret = 0; unsigned char nm = 0;
            if (!s_skip[c]) {                                  // a skipped cell stays open with nTotal = 0
                const int nKeys = s_total[c];
                if (nKeys > nfc) ret = nfc;
                else { ret = nKeys; nToDistribute += nfc - nKeys; nm = 1; nNoMore++; }
            }
            s_retain[c] = ret; noMore[c] = nm;
        }
        nNoMore = __reduce_add_sync(0xffffffffu, nNoMore); nToDistribute = __reduce_add_sync(0xffffffffu, nToDistribute);
        while (nToDistribute > 0 && nNoMore < nCells) {
            const int nNew = nfc + (int)ceilf(__fdiv_rn((float)nToDistribute, (float)(nCells - nNoMore)));
            int d = 0, m = 0;
            for (int c = lane; c < nCells; c += 32) {
                if (noMore[c]) continue;
                const int nt_c = s_total[c];
                if (nt_c > nNew) s_retain[c] = nNew;
                else { s_retain[c] = nt_c; d += nNew - nt_c; noMore[c] = 1; m++; }
            }
            nToDistribute = __reduce_add_sync(0xffffffffu, d); nNoMore += __reduce_add_sync(0xffffffffu, m);
        }
        __syncwarp();
        int o = 0;
        for (int c0 = 0; c0 < nCells; c0 += 32) {              // exclusive prefix of the quotas in cell order
            const int c = c0 + lane, v = c < nCells ? s_retain[c] : 0;
            int incl = v;
#pragma unroll
            for (int k = 1; k < 32; k <<= 1) { const int tv = __shfl_up_sync(0xffffffffu, incl, k); if (lane >= k) incl += tv; }
            if (c < nCells) s_off[c] = o + incl - v;
            o += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) {
            s_off[nCells] = o;
            if (o > L.lvl_cap) { atomicExch(status, ORB_ERR_CAPACITY); s_off[nCells] = -1; }
        }
    }
    __syncthreads();
    tl_stamp(10);
    int total = s_off[nCells];
    if (total < 0) { if (tid == 0) nkept[f * plan->nlevels + level] = 0; return; }
    uint32_t* gbase = cand + (size_t)f * plan->cand_total;
    unsigned long long* gbase64 = HARRIS ? cand64 + (size_t)f * plan->cand_total : nullptr;
    const orbsel::KeyGreater<uint32_t, 24> lt32;
    const orbsel::FloatKeyGreater64 ltf;
    for (int c = warp; c < nCells; c += WARPS) {           // retainBest per cell (:683-685)
        const int n = s_total[c], keep = s_retain[c];
        if (keep <= 0) continue;
        const int ix = cg[c].inix, iy = cg[c].iniy;
        // index scratch for lists that are selected in place in global memory: the frame's NMS score map, which nothing reads after
        // k_cell_compact (2 bytes per candidate slot)
        unsigned short* gs = spare ? reinterpret_cast<unsigned short*>(spare + (size_t)f * fbytes) + cg[c].cand_off : nullptr;
        if (HARRIS) {
            unsigned long long* v = gbase64 + cg[c].cand_off;
            if (n > keep) {
                if (n <= 65535 && gs) warp_nth_element(v, n, keep - 1, ltf, gs, lane);
                else { if (lane == 0) orbsel::nth_element(v, n, keep - 1, ltf); __syncwarp(); }
            }
            for (int k = lane; k < keep; k += 32) {
                const unsigned long long r = v[k];
                const unsigned long long x = (r & 0xfff) + ix, y = ((r >> 12) & 0xfff) + iy;
                s_list[s_off[c] + k] = (r & 0xffffffff00000000ull) | (y << 16) | x;
            }
        } else {
            uint32_t* v = gbase + cg[c].cand_off;
            if (n > keep) {
                if (n <= SEL_WCAP) {
                    uint32_t* w = s_wbuf + warp * SEL_WCAP;
                    for (int k = lane; k < n; k += 32) w[k] = v[k];
                    __syncwarp();
                    warp_nth_element(w, n, keep - 1, lt32, s_scr + warp * SEL_WCAP, lane);
                    v = w;
                } else if (n <= 65535 && gs) {                     // longer than the staging buffer (few features on a large image)
                    warp_nth_element(v, n, keep - 1, lt32, gs, lane);
                } else {
                    if (lane == 0) orbsel::nth_element(v, n, keep - 1, lt32);
                    __syncwarp();
                }
            }
            for (int k = lane; k < keep; k += 32) {
                const uint32_t r = v[k];
                const unsigned long long x = (r & 0xfff) + ix, y = ((r >> 12) & 0xfff) + iy;
                s_list[s_off[c] + k] = ((unsigned long long)(r >> 24) << 32) | (y << 16) | x;
            }
        }
        __syncwarp();
    }
    __syncthreads();
    tl_stamp(11);
    if (total > L.nDesired) {                                  // retainBest per level (:697-701)
        const orbsel::KeyGreater<unsigned long long, 32> lt64;
        if (warp == 0) {
            if (total <= WARPS * SEL_WCAP) {
                if (HARRIS) warp_nth_element(s_list, total, L.nDesired - 1, ltf, s_scr, lane);
                else warp_nth_element(s_list, total, L.nDesired - 1, lt64, s_scr, lane);
            } else if (lane == 0) {
                if (HARRIS) orbsel::nth_element(s_list, total, L.nDesired - 1, ltf);
                else orbsel::nth_element(s_list, total, L.nDesired - 1, lt64);
            }
        }
        total = L.nDesired;
        __syncthreads();
    }
    tl_stamp(12);
    unsigned long long* dst = lvl + (size_t)f * plan->lvl_total + L.lvl_base;
    for (int k = tid; k < total; k += blockDim.x) dst[k] = s_list[k];
    if (tid == 0) nkept[f * plan->nlevels + level] = total;
    tl_stamp(13);
}

template <bool HARRIS>
__global__ void __launch_bounds__(128)
k_select(const Plan* __restrict__ plan, const CellGeom* __restrict__ cells, uint32_t* __restrict__ cand,
         unsigned long long* __restrict__ cand64, const int* __restrict__ ntotal, unsigned long long* __restrict__ lvl,
         int* __restrict__ nkept, int* __restrict__ status)
{
    pdl_sync(7);
    extern __shared__ unsigned long long s_list[];            // lvl_cap records, then SEL_STAGE u32 records
    __shared__ int s_total[ORB_MAX_CELLS_LEVEL], s_retain[ORB_MAX_CELLS_LEVEL], s_off[ORB_MAX_CELLS_LEVEL + 1];
    __shared__ int s_coff[ORB_MAX_CELLS_LEVEL + 1];
    const int level = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const LevelGeom& L = plan->L[level];
    const int nCells = L.ncells;
    const CellGeom* cg = cells + L.cell_base;
    const int* nt = ntotal + (size_t)f * plan->ncells + L.cell_base;
    uint32_t* s_cand = reinterpret_cast<uint32_t*>(s_list + plan->sel_list_cap);
    for (int c = tid; c < nCells; c += blockDim.x) s_total[c] = nt[c];
    __syncthreads();
    if (tid == 0) {
        const int nfc = L.nfCell;
        int nNoMore = 0, nToDistribute = 0;
        unsigned char noMore[ORB_MAX_CELLS_LEVEL];
        for (int c = 0; c < nCells; c++) {
            noMore[c] = 0; s_retain[c] = 0;
            if (cg[c].skipped) continue;                       // stays open with nTotal = 0
            const int nKeys = s_total[c];
            if (nKeys > nfc) { s_retain[c] = nfc; }
            else { s_retain[c] = nKeys; nToDistribute += nfc - nKeys; noMore[c] = 1; nNoMore++; }
        }
        while (nToDistribute > 0 && nNoMore < nCells) {
            const int nNew = nfc + (int)ceilf(__fdiv_rn((float)nToDistribute, (float)(nCells - nNoMore)));
            nToDistribute = 0;
            for (int c = 0; c < nCells; c++) {
                if (noMore[c]) continue;
                if (s_total[c] > nNew) s_retain[c] = nNew;
                else { s_retain[c] = s_total[c]; nToDistribute += nNew - s_total[c]; noMore[c] = 1; nNoMore++; }
            }
        }
        int o = 0, co = 0;
        for (int c = 0; c < nCells; c++) {
            s_off[c] = o; o += s_retain[c];
            // only cells that really run the selection are staged
            s_coff[c] = co; if (s_total[c] > s_retain[c] && s_retain[c] > 0) co += s_total[c];
        }
        s_off[nCells] = o; s_coff[nCells] = co;
        if (o > L.lvl_cap) { atomicExch(status, ORB_ERR_CAPACITY); s_off[nCells] = -1; }
    }
    __syncthreads();
    int total = s_off[nCells];
    if (total < 0) { if (tid == 0) nkept[f * plan->nlevels + level] = 0; return; }
    const bool staged = !HARRIS && s_coff[nCells] <= SEL_STAGE;
    uint32_t* gbase = cand + (size_t)f * plan->cand_total;
    unsigned long long* gbase64 = HARRIS ? cand64 + (size_t)f * plan->cand_total : nullptr;
    if (staged) {          // coalesced copy of the lists that need a selection into shared memory
        const int warp = tid >> 5, lane = tid & 31;
        for (int c = warp; c < nCells; c += 4) {
            const int n = s_coff[c + 1] - s_coff[c];
            const uint32_t* src = gbase + cg[c].cand_off;
            for (int k = lane; k < n; k += 32) s_cand[s_coff[c] + k] = src[k];
        }
        __syncthreads();
    }
    for (int c = tid; c < nCells; c += blockDim.x) {
        const int n = s_total[c], keep = s_retain[c];
        const int ix = cg[c].inix, iy = cg[c].iniy;
        if (HARRIS) {             // 64-bit records (float response | position) straight in global memory
            unsigned long long* v = gbase64 + cg[c].cand_off;
            if (n > keep && keep > 0) orbsel::nth_element(v, n, keep - 1, orbsel::FloatKeyGreater64());
            for (int k = 0; k < keep; k++) {
                const unsigned long long r = v[k];
                const unsigned long long x = (r & 0xfff) + ix, y = ((r >> 12) & 0xfff) + iy;
                s_list[s_off[c] + k] = (r & 0xffffffff00000000ull) | (y << 16) | x;
            }
        } else {
            uint32_t* v = gbase + cg[c].cand_off;
            if (n > keep && keep > 0) {
                if (staged) v = s_cand + s_coff[c];
                orbsel::nth_element(v, n, keep - 1, orbsel::KeyGreater<uint32_t, 24>());
            }
            for (int k = 0; k < keep; k++) {
                const uint32_t r = v[k];
                const unsigned long long x = (r & 0xfff) + ix, y = ((r >> 12) & 0xfff) + iy;
                s_list[s_off[c] + k] = ((unsigned long long)(r >> 24) << 32) | (y << 16) | x;
            }
        }
    }
    __syncthreads();
    if (total > L.nDesired) {
        if (tid == 0) {
            if (HARRIS) orbsel::nth_element(s_list, total, L.nDesired - 1, orbsel::FloatKeyGreater64());
            else orbsel::nth_element(s_list, total, L.nDesired - 1, orbsel::KeyGreater<unsigned long long, 32>());
        }
        total = L.nDesired;
        __syncthreads();
    }
    unsigned long long* dst = lvl + (size_t)f * plan->lvl_total + L.lvl_base;
    for (int k = tid; k < total; k += blockDim.x) dst[k] = s_list[k];
    if (tid == 0) nkept[f * plan->nlevels + level] = total;
}

// ------------------------------------------------------------------ K5
// 7x7 sigma=2 Gaussian, the FP32 separable path OpenCV 4.x takes for an 8-bit non-isolated
// sub-matrix (DESIGN.md "K5"): row pass s = fma(I[x+i-3], k[i], s) from s = 0, column pass
// s = k3*R[y]; s = fma(R[y+d]+R[y-d], k[3+d], s), d = 1..3; round-half-even, saturate.
// The order of the FMAs is part of the result, so the kernel keeps it; what it avoids are the
// slow conversion instructions: u8 -> f32 is PRMT into the mantissa of 2^23 followed by an exact
// FADD, f32 -> u8 is the 1.5*2^23 magic add (round-to-nearest-even by the FP adder).
constexpr int BT_W = ORB_BLUR_TILE_W, BT_H = ORB_BLUR_TILE_H;   // 64 x 56 outputs per work item
constexpr int BIW = 24;                       // input tile words per row: padded cols x0 .. x0+95 (16-byte aligned TMA origin)
constexpr int BI_H = BT_H + 6;                // rows y0-3 .. y0+BT_H+2
constexpr int BI_BYTES = BI_H * BIW * 4;      // 5952
constexpr int BI_BUF = (BI_BYTES + 127) & ~127;

__device__ __forceinline__ float u8f(uint32_t w, int j)     // byte j of w as float, no I2F
{
    const uint32_t sel = 0x7440u | (uint32_t)j;
    return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, sel)), 8388608.0f);
}

// Persistent kernel, same pipeline shape as k_fast_nms: TMA fetches the input tile of the next work
// item while the current one runs its row and column passes.
#ifndef ORB_BLUR_THREADS
#define ORB_BLUR_THREADS 128     // measured on B200 (ms per 256 frames, CTAs/SM): 256x4 0.475, 192x5 0.497, 128x7 0.408, 128x8 0.388, 64x8 0.446
#endif
constexpr int BLUR_THREADS = ORB_BLUR_THREADS;
__global__ void __launch_bounds__(BLUR_THREADS)
k_blur(const __grid_constant__ TmapSet tm, uint8_t* __restrict__ blurred, size_t fbytes,
       const Plan* __restrict__ plan, const Tile* __restrict__ tiles, int ntiles, int total, int* __restrict__ work_counter)
{
    pdl_sync(8);
    __shared__ __align__(128) uint8_t img2[2][BI_BUF];
    __shared__ __align__(16) float rowp[BI_H * BT_W];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ int s_next[2], s_ti[2], s_fr[2];             // next work item; tile index and frame of the item in each buffer
    const float k0 = __uint_as_float(0x3d8fafb1u), k1 = __uint_as_float(0x3e06387eu),
                k2 = __uint_as_float(0x3e434a39u), k3 = __uint_as_float(0x3e5d4ae0u);
    const int tid = threadIdx.x;
    auto issue = [&](int item, int buf) {
        const int ti = item % ntiles, fr = item / ntiles;
        const Tile t = tiles[ti];
        s_ti[buf] = ti; s_fr[buf] = fr;                     // decoded once here instead of by every thread (two integer divisions)
        mbar_expect_tx(&bar[buf], BI_BYTES);
        tma_load_3d(&img2[buf][0], &tm.m[t.level], t.x0, t.y0 - 3 + ORB_EDGE, fr, &bar[buf]);   // padded x0 = ROI x0 - 16
    };
    if (tid == 0) {
        mbar_init(&bar[0], 1); mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int item = blockIdx.x;
    if (tid == 0 && item < total) issue(item, 0);
    for (int it = 0; item < total; it++) {
        const int buf = it & 1;
        if (tid == 0) {
            const int nxt = atomicAdd(work_counter, 1) + (int)gridDim.x;
            s_next[buf] = nxt;
            if (nxt < total) issue(nxt, buf ^ 1);
        }
        if (tid < 32) mbar_wait(&bar[buf], (uint32_t)((it >> 1) & 1));      // one warp polls, the rest sleep in the barrier
        __syncthreads();
        const int f = s_fr[buf];
        const Tile t = tiles[s_ti[buf]];
        const LevelGeom& L = plan->L[t.level];
        const int Lw = L.w, Lh = L.h, Lstride = L.stride;
        const uint32_t* img = reinterpret_cast<const uint32_t*>(img2[buf]);
        // Tiles on the right / bottom edge of the ROI are only partly filled (11 % of the tile area at 752x480): there, tasks are
        // numbered over the filled part only (rowp entries outside it keep stale values and feed only outputs that are not stored).
        // Full tiles keep compile-time task counts: with run-time bounds the compiler no longer overlaps the loads of consecutive
        // tasks and the kernel as a whole gets slower (0.349 -> 0.365 ms per 256 frames), so both forms are instantiated.
        const int ncg = (min(BT_W, Lw - t.x0) + 3) >> 2, nrg = (min(BT_H, Lh - t.y0) + 3) >> 2;     // 4x4 output blocks
        const int nseg = (ncg + 1) >> 1, nrow = min(BI_H, 4 * nrg + 6);
        const uint32_t inv_nseg = (65536u + nseg - 1) / nseg, inv_ncg = (65536u + ncg - 1) / ncg;   // floor(task / n) = task * inv >> 16, exact here
        uint8_t* out = blurred + (size_t)f * fbytes + L.plane_off;
        // row pass: one task = 8 adjacent outputs of one row; ROI column x0+c sits at tile byte 16+c
        auto row_pass = [&](auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
            const int ntask = FULL ? BI_H * (BT_W / 8) : nrow * nseg;
            for (int task = tid; task < ntask; task += BLUR_THREADS) {
                const int r = FULL ? task >> 3 : (int)(((uint32_t)task * inv_nseg) >> 16), seg = FULL ? task & 7 : task - r * nseg;
                const uint32_t* ip = img + r * BIW + 3 + 2 * seg;          // bytes 12+8seg .. 27+8seg
                const uint32_t ax = ip[0], ay = ip[1], bx = ip[2], by = ip[3];
                float v[14];
                v[0] = u8f(ax, 1); v[1] = u8f(ax, 2); v[2] = u8f(ax, 3);
                v[3] = u8f(ay, 0); v[4] = u8f(ay, 1); v[5] = u8f(ay, 2); v[6] = u8f(ay, 3);
                v[7] = u8f(bx, 0); v[8] = u8f(bx, 1); v[9] = u8f(bx, 2); v[10] = u8f(bx, 3);
                v[11] = u8f(by, 0); v[12] = u8f(by, 1); v[13] = u8f(by, 2);
                float o[8];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    float s = __fmul_rn(v[q], k0);
                    s = fmaf(v[q + 1], k1, s); s = fmaf(v[q + 2], k2, s); s = fmaf(v[q + 3], k3, s);
                    s = fmaf(v[q + 4], k2, s); s = fmaf(v[q + 5], k1, s); s = fmaf(v[q + 6], k0, s);
                    o[q] = s;
                }
                float4* op = reinterpret_cast<float4*>(rowp + r * BT_W + 8 * seg);
                op[0] = make_float4(o[0], o[1], o[2], o[3]);
                op[1] = make_float4(o[4], o[5], o[6], o[7]);
            }
        };
        // column pass: one task = 4 columns x 4 rows of outputs (10 row-pass rows, float4 loads)
        auto col_pass = [&](auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
            const int ntask = FULL ? (BT_W / 4) * (BT_H / 4) : ncg * nrg;
            for (int task = tid; task < ntask; task += BLUR_THREADS) {
                const int rg = FULL ? task >> 4 : (int)(((uint32_t)task * inv_ncg) >> 16), cg = FULL ? task & 15 : task - rg * ncg;
                const float4* rp = reinterpret_cast<const float4*>(rowp + (rg * 4) * BT_W + cg * 4);
                float4 R[10];
#pragma unroll
                for (int q = 0; q < 10; q++) R[q] = rp[q * (BT_W / 4)];
                const int x = t.x0 + cg * 4;
                uint8_t* o = out + (size_t)(t.y0 + rg * 4 + ORB_EDGE) * Lstride + x + ORB_EDGE;
#pragma unroll
                for (int q = 0; q < 4; q++, o += Lstride) {
                    const int y = t.y0 + rg * 4 + q;
                    uint32_t iv[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
#define RV(i) (e == 0 ? R[i].x : e == 1 ? R[i].y : e == 2 ? R[i].z : R[i].w)
                        float s = __fmul_rn(k3, RV(q + 3));
                        s = fmaf(__fadd_rn(RV(q + 4), RV(q + 2)), k2, s);
                        s = fmaf(__fadd_rn(RV(q + 5), RV(q + 1)), k1, s);
                        s = fmaf(__fadd_rn(RV(q + 6), RV(q)), k0, s);
#undef RV
                        // rint via the 1.5*2^23 magic constant: the low mantissa byte is the result.  No saturation needed: the taps
                        // sum to 1 within 1e-7, so s < 255.5 for 8-bit inputs.
                        iv[e] = __float_as_uint(__fadd_rn(s, 12582912.0f));
                    }
                    const uint32_t w = __byte_perm(__byte_perm(iv[0], iv[1], 0x0040), __byte_perm(iv[2], iv[3], 0x0040), 0x5410);
                    if (FULL) *reinterpret_cast<uint32_t*>(o) = w;       // a full tile lies inside the ROI
                    else if (y < Lh && x < Lw) {    // exactly the ROI bytes: the frame of the blurred buffer holds the un-blurred reflection
                        if (x + 3 < Lw) *reinterpret_cast<uint32_t*>(o) = w;
                        else for (int e = 0; x + e < Lw; e++) o[e] = (uint8_t)(w >> (8 * e));
                    }
                }
            }
        };
        if (Lw - t.x0 >= BT_W && Lh - t.y0 >= BT_H) {
            row_pass(std::true_type{});
            __syncthreads();
            col_pass(std::true_type{});
        } else {
            row_pass(std::false_type{});
            __syncthreads();
            col_pass(std::false_type{});
        }
        __syncthreads();          // rowp and the other image buffer are reused by the next item
        item = s_next[buf];
    }
}

// ------------------------------------------------------------------ K6
const int8_t h_pattern[1024] = {
#include "orb_pattern.inc"
};
// rBRIEF pattern transposed for the warp
__device__ float4 g_pattern_t[8 * 32];    // entry [k * 32 + lane] = both sample points of test k of descriptor byte `lane`: one 128-bit load per test
__device__ uint32_t g_rowmask[32];      // IC_Angle disc: bit (v+15) of entry `lane` set iff |lane-15| <= umax[|v|]
__constant__ int c_umax[16];
__device__ uint4 g_icmask[32 * 3];        // IC_Angle by rows: entry [r * 3 + q], q < 2: byte (u+15) of the 32 bytes = 0xff iff |u| <= umax[|r-15|]; row 31 and q = 2 are zero

// cv::fastAtan2 (degrees), every operation individually rounded to FP32 (no contraction)
__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    const float scale = (float)(180 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float eps = 2.2204460492503131e-16f;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// cvRound for |v| < 2^22 without the slow F2I: the 1.5*2^23 magic add rounds to nearest-even
__device__ __forceinline__ int rint_magic(float v)
{
    return (int)__float_as_uint(__fadd_rn(v, 12582912.0f)) - 0x4B400000;
}

// One warp per output keypoint slot.  Lanes 0..30 own patch column u = lane-15 for the moments;
// lane i then owns descriptor byte i (8 tests, 16 rotated samples).
// ORB_DESC_STAGE (default): the 512 rotated samples of a keypoint are a gather over a 37x37 window of the blurred level (the pattern's
// largest radius is 18.4, so a rotated coordinate rounds to at most 18): taken straight from global memory every sample instruction
// touches about 22 different 128-byte lines (~350 L1 wavefronts per keypoint against ~80 for copying the window once).  The warp therefore
// copies the window into shared memory first — 13 four-byte cp.async per lane, three rows of ten aligned words per instruction, issued
// ahead of IC_Angle — and samples it with LDS.U8.
#ifndef ORB_DESC_STAGE
#define ORB_DESC_STAGE 1
#endif
#ifndef ORB_DESC_WIDE
#define ORB_DESC_WIDE 0       // 1: the descriptor window as four 16-byte chunks per row (eight rows per cp.async instruction, 64-byte pitch) instead of ten words:
#endif                        // 5 instead of 13 copies, but 0.945 against 0.835 ms per 1024 frames — more bytes per row and a pitch of 16 banks doubles the gather's conflicts
#if ORB_DESC_WIDE
constexpr int DS_R = 18, DS_ROWS = 2 * DS_R + 1, DS_PITCH = 64, DS_AMASK = 15;
#else
constexpr int DS_R = 18, DS_ROWS = 2 * DS_R + 1, DS_WORDS = 10, DS_PITCH = 44, DS_AMASK = 3;      // window radius / rows, aligned words fetched per row, row pitch in shared memory (bytes)
#endif
constexpr int DS_WARP_BYTES = (DS_ROWS * DS_PITCH + 15) & ~15;
// ORB_DESC_ROWS (default): the orientation patch is copied as three 16-byte chunks per row (ten rows per cp.async instruction) and a lane
// owns a ROW: three LDS.128 (a 48-byte pitch is conflict free for them), the row's bytes moved to u = -15..16 by funnel shifts (the word
// part of the alignment is warp-uniform: a switch), the disc applied as a byte mask from a 1.5 KB table (L1-resident), and the row's
// two sums taken by IDP.4A against immediate weights: three shared-memory and two table loads and ~45 instructions per keypoint instead of 31 loads and ~85.
#ifndef ORB_DESC_ROWS
#define ORB_DESC_ROWS 1
#endif
#if ORB_DESC_ROWS
constexpr int IS_R = 15, IS_ROWS = 2 * IS_R + 1, IS_CHUNKS = 3, IS_PITCH = 16 * IS_CHUNKS;
#else
constexpr int IS_R = 15, IS_ROWS = 2 * IS_R + 1, IS_WORDS = 9, IS_PITCH = 36;       // the same for IC_Angle's 31x31 patch of the un-blurred level
#endif
constexpr int IS_WARP_BYTES = (IS_ROWS * IS_PITCH + 15) & ~15;
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ uint4 lds_u128(uint32_t saddr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void cp_async4(uint32_t saddr, const void* g)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
#ifndef ORB_DESC_MINB
#define ORB_DESC_MINB 6
#endif
__global__ void __launch_bounds__(256, ORB_DESC_MINB)      // 32 registers: the kernel is bound by gather latency, 0.262 -> 0.248 ms per 256 frames against 40 registers
k_describe(const uint8_t* __restrict__ planes, const uint8_t* __restrict__ blurred, size_t fbytes,
           const Plan* __restrict__ plan, const unsigned long long* __restrict__ lvl, const int* __restrict__ nkept,
           orb_keypoint* __restrict__ kps, uint8_t* __restrict__ desc, int cap, int32_t* __restrict__ counts)
{
    pdl_sync(9);
    const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.y;
    const int nl = plan->nlevels;
    // per-level keypoint counts: lane l holds level l, prefix by shuffles
    const int mycnt = lane < nl ? nkept[f * nl + lane] : 0;
    int incl = mycnt;
#pragma unroll
    for (int o = 1; o < ORB_MAX_LEVELS; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    const int total = __shfl_sync(0xffffffffu, incl, ORB_MAX_LEVELS - 1);
    if (slot == 0 && lane == 0) counts[f] = total;        // > cap means the caller's buffers truncated the output
    if (slot >= total || slot >= cap) return;
    const unsigned below = __ballot_sync(0xffffffffu, lane < nl && incl <= slot);
    const int level = __popc(below);
    const int idx = slot - (__shfl_sync(0xffffffffu, incl, level) - __shfl_sync(0xffffffffu, mycnt, level));
    const LevelGeom& L = plan->L[level];
    const int stride = L.stride;
    const unsigned long long rec = lvl[(size_t)f * plan->lvl_total + L.lvl_base + idx];
    const int x = (int)(rec & 0xffff), y = (int)((rec >> 16) & 0xffff), score = (int)(rec >> 32);
#if !ORB_DESC_STAGE
    const size_t poff = (size_t)f * fbytes + L.plane_off + (size_t)(ORB_EDGE + y) * stride + ORB_EDGE + x;
    const uint8_t* center = planes + poff;                  // un-blurred plane: orientation
    const uint8_t* bcenter = blurred + poff;                // blurred ROI + un-blurred frame: descriptor
#endif

#if ORB_DESC_STAGE
    __shared__ __align__(16) uint8_t s_patch[8][DS_WARP_BYTES + IS_WARP_BYTES];
    const int xa = ORB_EDGE + x - DS_R;                      // padded column of the descriptor window's first byte; the copy starts at the aligned word below it
    const int xi = ORB_EDGE + x - IS_R;                      // the same for the orientation patch
    const uint32_t s_ic = (uint32_t)__cvta_generic_to_shared(&s_patch[threadIdx.x >> 5][0]), s_win = s_ic + IS_WARP_BYTES;
    {
        const size_t lev_off = (size_t)f * fbytes + L.plane_off;       // byte offset of the level inside either buffer
#if ORB_DESC_ROWS
        {   // un-blurred 31x31 patch: ten rows of three 16-byte chunks per instruction (lanes 30, 31 idle)
            const int rg = (lane * 11) >> 5, ch = lane - 3 * rg;           // lane / 3, lane % 3
            const uint8_t* gp = planes + lev_off + (size_t)(ORB_EDGE + y - IS_R + rg) * stride + (xi & ~15) + 16 * ch;
            const uint32_t sw = s_ic + rg * IS_PITCH + 16 * ch;
#pragma unroll
            for (int i = 0; i < (IS_ROWS + 9) / 10; i++, gp += 10 * stride)
                if (lane < 30 && 10 * i + rg < IS_ROWS) cp_async16(sw + i * 10 * IS_PITCH, gp);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
#else
        {   // un-blurred 31x31 patch: three rows of nine words per instruction (lanes 27..31 idle)
            const int rg = lane >= 18 ? 2 : lane >= 9 ? 1 : 0, wd = lane - IS_WORDS * rg;
            const uint8_t* gp = planes + lev_off + (size_t)(ORB_EDGE + y - IS_R + rg) * stride + (xi & ~3) + 4 * wd;
            const uint32_t sw = s_ic + rg * IS_PITCH + 4 * wd;
#pragma unroll
            for (int i = 0; i < (IS_ROWS + 2) / 3; i++, gp += 3 * stride)
                if (lane < 3 * IS_WORDS && 3 * i + rg < IS_ROWS) cp_async4(sw + i * 3 * IS_PITCH, gp);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
#endif
#if ORB_DESC_WIDE
        {   // blurred 37x37 window: eight rows of four 16-byte chunks per instruction
            const int rg = lane >> 2, ch = lane & 3;
            const uint8_t* gp = blurred + lev_off + (size_t)(ORB_EDGE + y - DS_R + rg) * stride + (xa & ~15) + 16 * ch;
            const uint32_t sw = s_win + rg * DS_PITCH + 16 * ch;
#pragma unroll
            for (int i = 0; i < (DS_ROWS + 7) / 8; i++, gp += 8 * stride)
                if (8 * i + rg < DS_ROWS) cp_async16(sw + i * 8 * DS_PITCH, gp);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
#else
        {   // blurred 37x37 window: three rows of ten words per instruction (lanes 30, 31 idle)
            const int rg = lane >= 20 ? 2 : lane >= 10 ? 1 : 0, wd = lane - DS_WORDS * rg;
            const uint8_t* gp = blurred + lev_off + (size_t)(ORB_EDGE + y - DS_R + rg) * stride + (xa & ~3) + 4 * wd;
            const uint32_t sw = s_win + rg * DS_PITCH + 4 * wd;
#pragma unroll
            for (int i = 0; i < (DS_ROWS + 2) / 3; i++, gp += 3 * stride)
                if (lane < 3 * DS_WORDS && 3 * i + rg < DS_ROWS) cp_async4(sw + i * 3 * DS_PITCH, gp);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
#endif
    }
    // IC_Angle (:124-151): m10 = sum u*I, m01 = sum v*I over the radius-15 disc; lane = column u.  Rows +v and -v of a column are inside
    // the disc together (one predicate), and both sums ride in one register: acc = colsum * 2^19 + sum v*I  (|sum v*I| <= 255 * 240 < 2^18,
    // colsum <= 31 * 255 < 2^13), so a pixel costs one LDS.U8 with an immediate row offset and one IMAD with an immediate multiplier.
    int m10, m01;
    {
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
#if ORB_DESC_ROWS
        // lane = row v = lane - 15 (lane 31: its mask row is zero).  w[] = the row's 48 bytes; the patch starts at byte xi & 15.
        const uint32_t ra = s_ic + lane * IS_PITCH;
        const uint4 q0 = lds_u128(ra), q1 = lds_u128(ra + 16), q2 = lds_u128(ra + 32);
        // the disc as byte masks of this lane's row, straight from the 1.5 KB table in global memory (L1-resident): a copy of the table in
        // shared memory cost a CTA barrier in front of everything else — 0.828 against 0.792 ms per 1024 frames
        const uint4 k0 = __ldg(g_icmask + lane * 3), k1 = __ldg(g_icmask + lane * 3 + 1);
        const uint32_t w[12] = { q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w };
        const uint32_t mk[8] = { k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w };
        const uint32_t sh = 8u * (uint32_t)(xi & 3);
        uint32_t d[8];
        switch ((xi >> 2) & 3) {                               // warp-uniform word offset of the patch inside the row
#define ORB_IC_ALIGN(K) _Pragma("unroll") for (int i = 0; i < 8; i++) d[i] = __funnelshift_r(w[K + i], w[K + i + 1], sh) & mk[i];
        case 0: ORB_IC_ALIGN(0) break;
        case 1: ORB_IC_ALIGN(1) break;
        case 2: ORB_IC_ALIGN(2) break;
        default: ORB_IC_ALIGN(3) break;
#undef ORB_IC_ALIGN
        }
        uint32_t rowsum = 0, wsum = 0;                         // sum I and sum (u + 15) * I over the row's part of the disc
#pragma unroll
        for (int i = 0; i < 8; i++) {
            rowsum = __dp4a(d[i], 0x01010101u, rowsum);
            wsum = __dp4a(d[i], 0x03020100u + 0x04040404u * (uint32_t)i, wsum);
        }
        m10 = __reduce_add_sync(0xffffffffu, (int)wsum - 15 * (int)rowsum);
        m01 = __reduce_add_sync(0xffffffffu, (lane - 15) * (int)rowsum);
#else
        const uint32_t rows = g_rowmask[lane];              // bit (v+15): |u| <= umax[|v|]; symmetric in v; lane 31 holds 0
        const uint32_t sp = s_ic + (uint32_t)(xi & 3) + lane;
        uint32_t acc = 0;
        if ((rows >> 15) & 1) acc = lds_u8(sp + IS_R * IS_PITCH) << 19;
#pragma unroll
        for (int v = 1; v <= IS_R; v++)
            if ((rows >> (15 + v)) & 1) {
                acc += lds_u8(sp + (IS_R + v) * IS_PITCH) * ((1u << 19) + (uint32_t)v);
                acc += lds_u8(sp + (IS_R - v) * IS_PITCH) * ((1u << 19) - (uint32_t)v);
            }
        const int mv = ((int)(acc << 13)) >> 13;            // sign-extended low 19 bits
        const int colsum = (int)((acc - (uint32_t)mv) >> 19);
        m10 = __reduce_add_sync(0xffffffffu, (lane - 15) * colsum);
        m01 = __reduce_add_sync(0xffffffffu, mv);
#endif
    }
#else
    // IC_Angle (:124-151): m10 = sum u*I, m01 = sum v*I over the radius-15 disc; lane = column u
    int m10 = 0, m01 = 0;
    {
        const int u = lane - 15;
        const uint32_t rows = g_rowmask[lane];              // bit (v+15): |u| <= umax[|v|]
        const uint8_t* p = center + u - 15 * stride;
        int colsum = 0;
#pragma unroll
        for (int v = -15; v <= 15; v++, p += stride) {
            if ((rows >> (v + 15)) & 1) {
                const int val = *p;
                colsum += val; m01 += v * val;
            }
        }
        m10 = u * colsum;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, o); m01 += __shfl_xor_sync(0xffffffffu, m01, o); }
#endif
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // computeOrbDescriptor (:155-194); cos / sin = glibc's cosf / sinf restated in orb_trig.h (pinned on every float angle)
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    const float arad = __fmul_rn(angle, factorPI);
    float a, b;
#ifdef ORB_DESC_TRIG_CR       // the earlier pin: correctly rounded through double (differs from glibc's cosf / sinf by one ulp on 0.13 % of all angles)
    if (lane == 0) { double sd, cd; sincos((double)arad, &sd, &cd); a = (float)cd; b = (float)sd; }
    a = __shfl_sync(0xffffffffu, a, 0); b = __shfl_sync(0xffffffffu, b, 0);
#else
    orbtrig::sincosf_glibc(arad, b, a);                    // glibc's cosf / sinf, the reference's own calls (orb_trig.h); every lane computes it (~35 instructions, no shuffle)
#endif
    const float4* pat = g_pattern_t + lane;
    const bool fma_form = plan->desc_fma != 0;
    // offset = cvRound(y')*stride + cvRound(x'); the 1.5*2^23 magic add leaves the rounded integer in the mantissa
#if ORB_DESC_STAGE
    // the same in the staged window: byte (ry + 18) * pitch + (rx + 18 + alignment offset) of the warp's copy
    const uint32_t stage_fix = s_win + (uint32_t)(DS_R * DS_PITCH + DS_R + (xa & DS_AMASK)) - 0x4B400000u * (uint32_t)(DS_PITCH + 1);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
#else
    const uint32_t magic_fix = 0u - 0x4B400000u * (uint32_t)(stride + 1);      // modulo-2^32 arithmetic, exact for the in-range result
#endif
    // x*b + y*a and x*a - y*b (:166-167): two roundings each as written, or, when the reference is built with its own
    // -O3 -march=native on an FMA host, GCC's contraction fma(x, b, y*a) / fma(x, a, -(y*b)) (orb_set_descriptor_fma).  The flag is
    // uniform, so the choice is made once around the loop instead of per sample.
    auto sample_bits = [&](auto fma_tag) {
        constexpr bool FMA = decltype(fma_tag)::value;
        int bits = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            int t[2];
            const float4 pp = __ldg(pat + k * 32);
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const float2 p = e ? make_float2(pp.z, pp.w) : make_float2(pp.x, pp.y);
                const float ry = FMA ? __fmaf_rn(p.x, b, __fmul_rn(p.y, a)) : __fadd_rn(__fmul_rn(p.x, b), __fmul_rn(p.y, a));
                const float rx = FMA ? __fmaf_rn(p.x, a, -__fmul_rn(p.y, b)) : __fsub_rn(__fmul_rn(p.x, a), __fmul_rn(p.y, b));
                const uint32_t yb = __float_as_uint(__fadd_rn(ry, 12582912.0f));
                const uint32_t xb = __float_as_uint(__fadd_rn(rx, 12582912.0f));
#if ORB_DESC_STAGE
                t[e] = (int)lds_u8(yb * (uint32_t)DS_PITCH + xb + stage_fix);
#else
                t[e] = bcenter[(int)(yb * (uint32_t)stride + xb + magic_fix)];
#endif
            }
            bits |= (t[0] < t[1]) << k;
        }
        return bits;
    };
    const int val = fma_form ? sample_bits(std::true_type{}) : sample_bits(std::false_type{});
    desc[((size_t)f * cap + slot) * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        orb_keypoint kp;
        float fx = (float)x, fy = (float)y;
        if (level != 0) { fx = __fmul_rn(fx, L.scale); fy = __fmul_rn(fy, L.scale); }   // :769-775
        kp.x = fx; kp.y = fy; kp.size = (float)L.patch_size; kp.angle = angle;
        kp.response = plan->harris ? __int_as_float(score) : (float)score; kp.octave = level; kp.class_id = -1;
        kps[(size_t)f * cap + slot] = kp;
    }
    tl_stamp(14);
}

} // namespace

#ifdef ORB_TIMELINE
extern "C" int orb_debug_timeline(unsigned long long* out, int reset)
{
    unsigned long long h[4 * 16];
    if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpyFromSymbol(h, g_tl, sizeof h) != cudaSuccess) return -1;
    if (out) memcpy(out, h, sizeof h);
    if (reset) {
        for (int i = 0; i < 16; i++) { h[4 * i] = h[4 * i + 1] = ~0ull; h[4 * i + 2] = h[4 * i + 3] = 0ull; }
        if (cudaMemcpyToSymbol(g_tl, h, sizeof h) != cudaSuccess) return -1;
    }
    return 0;
}
#endif

int orb_upload_constants(const int* umax)
{
    ORB_CUDA(cudaMemcpyToSymbol(c_umax, umax, sizeof(int) * 16));
    float4 t[8 * 32];
    for (int lane = 0; lane < 32; lane++)
        for (int k = 0; k < 8; k++) {
            const int8_t* p = h_pattern + lane * 32 + 4 * k;
            t[k * 32 + lane] = make_float4((float)p[0], (float)p[1], (float)p[2], (float)p[3]);
        }
    ORB_CUDA(cudaMemcpyToSymbol(g_pattern_t, t, sizeof t));
    uint32_t rm[32];
    for (int lane = 0; lane < 32; lane++) {
        rm[lane] = 0;
        const int au = lane < 31 ? (lane > 15 ? lane - 15 : 15 - lane) : 99;
        for (int v = -15; v <= 15; v++) if (au <= umax[v < 0 ? -v : v]) rm[lane] |= 1u << (v + 15);
    }
    ORB_CUDA(cudaMemcpyToSymbol(g_rowmask, rm, sizeof rm));
    uint8_t im[32 * 48];
    memset(im, 0, sizeof im);
    for (int r = 0; r < 31; r++) {
        const int av = r > 15 ? r - 15 : 15 - r;
        for (int u = -15; u <= 15; u++) if ((u < 0 ? -u : u) <= umax[av]) im[r * 48 + u + 15] = 0xff;
    }
    ORB_CUDA(cudaMemcpyToSymbol(g_icmask, im, sizeof im));
    return ORB_OK;
}

// Every kernel of the pass is launched with programmatic stream serialization (PDL): a kernel signals griddepcontrol.launch_dependents
// at its top, so the NEXT kernel's CTAs are scheduled as soon as all CTAs of this one have started and SM resources free up, run their
// prologue and park in griddepcontrol.wait until this grid has completed and its memory is visible.  Stream-order semantics are
// unchanged (every access to pipeline data sits behind the wait); what disappears is the launch latency between dependent kernels:
// the 14 launches of ONE frame cost about 4.7 us each, of which about half is launch.  ORB_PDL=0 switches it off (A/B timing).
template <typename... KArgs, typename... Args>
static inline void launch_k(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(std::forward<Args>(args))...);      // errors surface in cudaGetLastError at the end of the pass
}

#ifdef ORB_DEBUG
#define ORB_SKIP(bit) (c->debug_skip & (bit))
#else
#define ORB_SKIP(bit) 0
#endif
int orb_launch_extract(orb_ctx* c, WorkSet& W, const uint8_t* d_imgs, int nimg, int w, int h, int stride, size_t frame_pitch,
                       orb_keypoint* d_kps, uint8_t* d_desc, int cap, int32_t* d_counts, cudaStream_t s)
{
    const Plan& P = c->plan;
    const size_t fb = (size_t)P.frame_bytes;
    int launches = 0;
    const bool pdl = c->pdl_call && !c->profile;      // chosen per call (orb_api.cu launch_extract): batches lose 1.6 % to parked CTAs (160.3 -> 157.8 K frames/s at 1024 frames)
    const dim3 blk(64, 4);
    auto mark = [&]() {           // stage boundary event (profiling mode only)
        if (!c->profile) return;
        cudaEvent_t e;
        if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); }
        else if (cudaEventCreate(&e) != cudaSuccess) return;
        cudaEventRecord(e, s);
        c->prof_events.push_back(e);
    };
    // every tile queue of this pass (FAST 1, blur 2, resize 4+l) and the per-(frame, level) completion counters of k_pyramid are cleared by k_level0
    mark();
    {
        const LevelGeom& L = P.L[0];
        dim3 grid((w + 255) / 256, (h + 3) / 4, nimg);
        const int aligned4 = (((uintptr_t)d_imgs | (uintptr_t)stride | (uintptr_t)frame_pitch) & 3) == 0;
        const bool aligned16 = ((((uintptr_t)d_imgs | (uintptr_t)stride | (uintptr_t)frame_pitch | (uintptr_t)w) & 15) == 0) && (L.stride & 15) == 0 &&
                               (L.plane_off & 15) == 0 && (fb & 15) == 0 && ((uintptr_t)W.d_planes & 15) == 0;
        if (aligned16)
            launch_k(pdl, k_level0_v16, dim3((w / 16 + 15) / 16, (h + 15) / 16, nimg), dim3(16, 16), 0, s, d_imgs, w, h, stride, frame_pitch, W.d_planes, fb, L.stride, W.d_counters);
        else
            launch_k(pdl, k_level0, grid, blk, 0, s, d_imgs, w, h, stride, frame_pitch, aligned4, W.d_planes, fb, L.stride, W.d_counters);
        launches++;
    }
    mark();
    // Which small-call forms a call takes.  ORB_SMALL_CALL = 0 switches all of them off; otherwise every form has the limit at which it
    // stopped paying on B200 (tools/latency_small_batch.py --forms4 / --each [--hd] [--noise]: blocking calls, pinned buffers, us):
    //   short resize tiles        up to 8 frames   (VGA 1 frame 93 against 102, 4: 140 / 149, 8: 195 / 195; 1080p 1 frame 211 / 223, dense noise 598 / 641)
    //   512-thread FAST CTAs      up to 2 frames, up to 4 while the call's tiles number at most 4 per SM
    //                             (VGA 1: 93 / 115, 2: 109 / 131, 4: 140 / 149, 6: 177 / 168, 16: 317 / 293; 1080p 1: 211 / 237, 2: 306 / 320, 4: 478 / 437)
    //   CTA-per-cell compaction   up to 32 frames  (VGA 16: 318 / 332, 32: 577 / 587; 1080p 1: 237 / 355; 1024 frames 0.68 / 0.32 ms)
    //   border on the side stream up to small_call_frames (VGA 8: 202 / 215; 1080p 1: 237 / 265)
    const bool small_on = c->small_call_frames > 0;
    const bool sm_resize = small_on && nimg <= 8;
    const bool sm_fast = small_on && (c->fast_wide == 2 || nimg <= 2 || (nimg <= 4 && P.ntiles_fast * nimg <= 4 * c->num_sms));   // ORB_FAST_WIDE=2: always (A/B timing)
    const bool sm_compact = small_on && nimg <= 32;
    const int rv = sm_resize ? 1 : 0;      // tiling of the resize cascade
    if (c->pyr_fused && P.nlevels > 1) {
        PyrParams Q;
        memset(&Q, 0, sizeof Q);
        Q.nlevels = P.nlevels; Q.nimg = nimg;
        int total = 0, bufb = 1024;
        for (int l = 1; l < P.nlevels; l++) {
            const LevelGeom& D = P.L[l];
            const int tw = c->rs_tile_w[rv][l], rr = c->rs_rows[rv][l], th = (4 * ORB_RESIZE_THREADS / tw) * rr;
            Q.L[l].tile_w = tw; Q.L[l].rows = rr; Q.L[l].box_w = c->rs_box_w[rv][l]; Q.L[l].box_h = c->rs_box_h[rv][l];
            Q.L[l].tiles_x = (D.w + tw - 1) / tw;
            Q.L[l].ntiles = Q.L[l].tiles_x * ((D.h + th - 1) / th);
            Q.L[l].item_base = total;
            total += Q.L[l].ntiles * nimg;
            bufb = std::max(bufb, (c->rs_box_w[rv][l] * c->rs_box_h[rv][l] + 127) & ~127);
        }
        Q.total = total; Q.buf_bytes = bufb;
        const int grid = std::min(total, c->num_sms * ORB_RESIZE_CTAS);
        launch_k(pdl, k_pyramid, grid, ORB_RESIZE_THREADS, 2 * bufb + 16, s, W.tm_resize[rv], W.d_planes, fb, c->d_plan, Q, c->d_xtab, c->d_ytab,
                                                                  W.d_counters + 4, W.d_counters + 32);
        launches++;
    } else
    for (int l = 1; l < P.nlevels; l++) {
        const LevelGeom& D = P.L[l];
        const int tw = c->rs_tile_w[rv][l], rr = c->rs_rows[rv][l], th = (4 * ORB_RESIZE_THREADS / tw) * rr;
        const int tiles = ((D.w + tw - 1) / tw) * ((D.h + th - 1) / th) * nimg;
        const int bufb = (c->rs_box_w[rv][l] * c->rs_box_h[rv][l] + 127) & ~127;
        const int grid = std::min(tiles, c->num_sms * ORB_RESIZE_CTAS);
        const bool unrolled = c->rs_unrolled && c->rs_packed[l];
        if (unrolled && rr == 2)
            launch_k(pdl, k_resize_u<2>, grid, ORB_RESIZE_THREADS, 2 * bufb + 16, s, W.tm_resize[rv].m[l], W.d_planes, fb, D, reinterpret_cast<const uint4*>(c->d_xtab + c->rs_xg_off[l]),
                                                                        c->d_ytab, c->rs_box_w[rv][l], c->rs_box_h[rv][l], bufb, nimg, W.d_counters + 4 + l, tw);
        else if (unrolled && rr == 4)
            launch_k(pdl, k_resize_u<4>, grid, ORB_RESIZE_THREADS, 2 * bufb + 16, s, W.tm_resize[rv].m[l], W.d_planes, fb, D, reinterpret_cast<const uint4*>(c->d_xtab + c->rs_xg_off[l]),
                                                                        c->d_ytab, c->rs_box_w[rv][l], c->rs_box_h[rv][l], bufb, nimg, W.d_counters + 4 + l, tw);
        else if (unrolled && rr == 8)
            launch_k(pdl, k_resize_u<8>, grid, ORB_RESIZE_THREADS, 2 * bufb + 16, s, W.tm_resize[rv].m[l], W.d_planes, fb, D, reinterpret_cast<const uint4*>(c->d_xtab + c->rs_xg_off[l]),
                                                                        c->d_ytab, c->rs_box_w[rv][l], c->rs_box_h[rv][l], bufb, nimg, W.d_counters + 4 + l, tw);
        else
        launch_k(pdl, k_resize, grid, ORB_RESIZE_THREADS, 2 * bufb + 16, s, W.tm_resize[rv].m[l], W.d_planes, fb, D, c->d_xtab, c->d_ytab, c->rs_box_w[rv][l], c->rs_box_h[rv][l],
                                               bufb, nimg, W.d_counters + 4 + l, tw, rr);
        launches++;
    }
    int border_max = 1;                 // a small level with the whole 16 px frame may hold more ring words than level 0 with its 4 px ring
    for (int l = 0; l < P.nlevels; l++) border_max = std::max(border_max, P.L[l].border_items);
    // k_blur only needs the finished pyramid: outside profiling mode it runs on a second stream,
    // concurrently with FAST -> compaction -> selection (the selection kernel is latency bound and
    // leaves most of the machine idle); k_describe joins both.
    const bool fork = !c->profile;
    auto launch_blur = [&](cudaStream_t bs) {
#ifdef ORB_DEBUG
        if (c->debug_skip & 1) return;
#endif
        const int total = P.ntiles_blur * nimg;
        const int grid = std::min(total, c->num_sms * c->blur_ctas);
        launch_k(false, k_blur, grid, BLUR_THREADS, 0, bs, W.tm_blur, W.d_blur, fb, c->d_plan, c->d_tiles_blur, P.ntiles_blur, total, W.d_counters + 2);
    };
    // Small calls: the reflect-101 ring is read by k_blur (3 px) and k_describe only — FAST, IC_Angle and the compaction stay inside
    // the ROI as long as every keypoint lies 16 px from the edge — so k_border leaves the critical path and runs, followed by k_blur,
    // on the second stream beside FAST -> compaction -> selection (timeline of one frame, tools/timeline.py: k_border held FAST back
    // by 4.7 of the pass's 60 us).  Not for HARRIS responses or levels whose cell grid reaches into the 16 px frame (LevelGeom::ring).
    bool side_border = fork && c->side_border && nimg <= c->small_call_frames && !P.harris;
    for (int l = 0; l < P.nlevels; l++) side_border = side_border && P.L[l].ring == ORB_RING;
    if (side_border) {
        ORB_CUDA(cudaEventRecord(W.ev_fork, s));
        ORB_CUDA(cudaStreamWaitEvent(W.aux_stream, W.ev_fork, 0));
        launch_k(false, k_border, dim3((border_max + 511) / 512, nimg * P.nlevels), 256, 0, W.aux_stream, W.d_planes, W.d_blur, fb, c->d_plan);
        launch_blur(W.aux_stream);
        ORB_CUDA(cudaEventRecord(W.ev_join, W.aux_stream));
    } else
    launch_k(pdl, k_border, dim3((border_max + 511) / 512, nimg * P.nlevels), 256, 0, s, W.d_planes, W.d_blur, fb, c->d_plan);
    launches++;
    if (fork && !side_border && c->fork_early == 1) {          // blur next to FAST: FAST saturates the ALU pipe and leaves the FMA pipe idle
        ORB_CUDA(cudaEventRecord(W.ev_fork, s));
        ORB_CUDA(cudaStreamWaitEvent(W.aux_stream, W.ev_fork, 0));
    }
    mark();
    {
        const int total = P.ntiles_fast * nimg;
        const int grid = std::min(total, c->num_sms * (fork && !side_border && c->fork_early == 1 ? c->fast_ctas : FAST_CTAS));
        if (c->fast_etile && c->fast_wide && sm_fast)
            launch_k(pdl, k_fast_nms<true, FAST_THREADS_SMALL>, std::min(total, c->num_sms), FAST_THREADS_SMALL, 0, s, W.tm_fast, W.d_work, W.d_bitmap, fb, c->d_plan, c->d_tiles_fast, P.ntiles_fast, total,
                     W.d_counters + 1, c->d_fast_coltab, c->d_fast_rowtab);
        else if (c->fast_etile)
            launch_k(pdl, k_fast_nms<true, FAST_THREADS>, grid, FAST_THREADS, 0, s, W.tm_fast, W.d_work, W.d_bitmap, fb, c->d_plan, c->d_tiles_fast, P.ntiles_fast, total, W.d_counters + 1,
                     c->d_fast_coltab, c->d_fast_rowtab);
        else
            launch_k(pdl, k_fast_nms<false, FAST_THREADS>, grid, FAST_THREADS, 0, s, W.tm_fast, W.d_work, W.d_bitmap, fb, c->d_plan, c->d_tiles_fast, P.ntiles_fast, total, W.d_counters + 1,
                     c->d_fast_coltab, c->d_fast_rowtab);
    }
    if (fork && !side_border && c->fork_early == 2) {     // blur starts behind FAST, next to compaction + selection
        ORB_CUDA(cudaEventRecord(W.ev_fork, s));
        ORB_CUDA(cudaStreamWaitEvent(W.aux_stream, W.ev_fork, 0));
    }
    if (fork && !side_border && (c->fork_early == 1 || c->fork_early == 2)) { launch_blur(W.aux_stream); ORB_CUDA(cudaEventRecord(W.ev_join, W.aux_stream)); }
    mark();
    if (c->compact_wide == 2 || (c->compact_wide && sm_compact))       // ORB_COMPACT_WIDE=2: for every call size (A/B timing)
        launch_k(pdl, k_cell_compact_wide, dim3(P.ncells, nimg), 256, 0, s, W.d_work, W.d_bitmap, fb, c->d_plan, c->d_cells, W.d_cand, W.d_ntotal);
    else
    {
        CompactGeom cg;
        memset(&cg, 0, sizeof cg);
        for (int l = 0; l < P.nlevels; l++) { cg.plane_off[l] = P.L[l].plane_off; cg.stride[l] = P.L[l].stride; cg.bm_off[l] = P.L[l].bm_off; cg.bm_pitch[l] = P.L[l].bm_pitch; }
        cg.ncells = P.ncells; cg.bm_total = P.bm_total; cg.cand_total = P.cand_total; cg.fast_th = P.fast_th;
        launch_k(pdl, k_cell_compact, dim3((P.ncells + 7) / 8, nimg), 256, 0, s, W.d_work, W.d_bitmap, fb, c->d_plan, c->d_cells, W.d_cand, W.d_ntotal, cg);
    }
    if (fork && !side_border && (c->fork_early == 0 || c->fork_early == 3)) {          // blur starts when compaction is done, i.e. next to the selection kernel
        ORB_CUDA(cudaEventRecord(W.ev_fork, s));
        ORB_CUDA(cudaStreamWaitEvent(W.aux_stream, W.ev_fork, 0));
        if (c->fork_early == 0) { launch_blur(W.aux_stream); ORB_CUDA(cudaEventRecord(W.ev_join, W.aux_stream)); }
    }
    mark();
    // a handful of frames: 32 warps per (frame, level) CTA instead of 8 (one cell per warp in one trip; k_select was 36 of the 130 us of a
    // single 640x480 frame), as long as the CTAs still fit the machine in one wave and the wider staging area fits its shared memory
    auto sel_bytes = [&](int warps) { return (size_t)P.sel_list_cap * 8 + (size_t)warps * SEL_WCAP * 6 + (size_t)P.sel_cells_cap * 14 + 16; };
    const bool sel_wide = c->select_wide && nimg * P.nlevels <= c->num_sms && sel_bytes(SEL_WARPS_WIDE) + 1024 <= 227 * 1024;
    const size_t sel_smem = sel_bytes(sel_wide ? SEL_WARPS_WIDE : SEL_WARPS);
    uint8_t* sel_spare = (size_t)P.cand_total * 2 <= fb ? W.d_work : nullptr;
    if (P.harris) {
        launch_k(pdl, k_harris, dim3((P.ncells + 7) / 8, nimg), 256, 0, s, W.d_planes, fb, c->d_plan, c->d_cells, W.d_cand, W.d_ntotal, W.d_cand64);
        if (c->select_serial)
            launch_k(pdl, k_select<true>, dim3(P.nlevels, nimg), 128, (size_t)P.sel_list_cap * 8 + SEL_STAGE * 4, s, c->d_plan, c->d_cells, W.d_cand, W.d_cand64, W.d_ntotal, W.d_lvl, W.d_nkept, c->d_status);
        else if (sel_wide)
            launch_k(pdl, k_select_fast<true, SEL_WARPS_WIDE>, dim3(P.nlevels, nimg), SEL_WARPS_WIDE * 32, sel_smem, s, c->d_plan, c->d_cells, W.d_cand, W.d_cand64, W.d_ntotal, W.d_lvl, W.d_nkept, c->d_status, sel_spare, fb);
        else
            launch_k(pdl, k_select_fast<true, SEL_WARPS>, dim3(P.nlevels, nimg), SEL_WARPS * 32, sel_smem, s, c->d_plan, c->d_cells, W.d_cand, W.d_cand64, W.d_ntotal, W.d_lvl, W.d_nkept, c->d_status, sel_spare, fb);
        launches++;
    } else if (!ORB_SKIP(2)) {
        if (c->select_serial)
            launch_k(pdl, k_select<false>, dim3(P.nlevels, nimg), 128, (size_t)P.sel_list_cap * 8 + SEL_STAGE * 4, s, c->d_plan, c->d_cells, W.d_cand, nullptr, W.d_ntotal, W.d_lvl, W.d_nkept, c->d_status);
        else if (sel_wide)
            launch_k(pdl, k_select_fast<false, SEL_WARPS_WIDE>, dim3(P.nlevels, nimg), SEL_WARPS_WIDE * 32, sel_smem, s, c->d_plan, c->d_cells, W.d_cand, nullptr, W.d_ntotal, W.d_lvl, W.d_nkept, c->d_status, sel_spare, fb);
        else
            launch_k(pdl, k_select_fast<false, SEL_WARPS>, dim3(P.nlevels, nimg), SEL_WARPS * 32, sel_smem, s, c->d_plan, c->d_cells, W.d_cand, nullptr, W.d_ntotal, W.d_lvl, W.d_nkept, c->d_status, sel_spare, fb);
    }
    if (fork && !side_border && c->fork_early == 3) {      // selection first: the latency-bound kernel takes the residency it needs, blur fills the rest
        launch_blur(W.aux_stream);
        ORB_CUDA(cudaEventRecord(W.ev_join, W.aux_stream));
    }
    mark();
    if (fork) ORB_CUDA(cudaStreamWaitEvent(s, W.ev_join, 0));
    else launch_blur(s);
    mark();
    const int slots = std::min(cap, P.kp_cap);
    if (!ORB_SKIP(4)) launch_k(pdl, k_describe, dim3((std::max(slots, 1) + 7) / 8, nimg), 256, 0, s, W.d_planes, W.d_blur, fb, c->d_plan, W.d_lvl, W.d_nkept,
                                                                    d_kps, d_desc, cap, d_counts);
    mark();
    launches += 5;
    c->last_launches = launches;
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is one value per kernel and device, shared by every context of the process: contexts
// of different shapes (ORB-SLAM keeps a 1000- and a 2000-feature extractor alive, src/Tracking.cc:111,126) must only ever RAISE it.
static int raise_dyn_smem(const void* fn, int slot, int bytes)
{
    static std::mutex mu;
    static int cur[11][64] = {};
    int dev = 0;
    ORB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && bytes <= cur[slot][dev]) return ORB_OK;
    ORB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (dev >= 0 && dev < 64) cur[slot][dev] = bytes;
    return ORB_OK;
}

int orb_resize_smem_setup(int max_bytes)
{
    static_assert(sizeof(PyrParams) <= 1024, "k_pyramid parameter block");
    int rc = raise_dyn_smem((const void*)k_resize, 0, max_bytes);
    if (!rc) rc = raise_dyn_smem((const void*)k_resize_u<8>, 6, max_bytes);
    if (!rc) rc = raise_dyn_smem((const void*)k_resize_u<4>, 9, max_bytes);
    if (!rc) rc = raise_dyn_smem((const void*)k_resize_u<2>, 10, max_bytes);
    return rc ? rc : raise_dyn_smem((const void*)k_pyramid, 5, max_bytes);
}

int orb_select_smem_setup(int list_cap, int cells_cap)     // the largest per-level keypoint list (u64 records) and cell grid of the plan
{
    const int serial = list_cap * 8 + SEL_STAGE * 4 + 1024, fast = list_cap * 8 + SEL_WARPS * SEL_WCAP * 6 + cells_cap * 14 + 1024;
    int rc = raise_dyn_smem((const void*)k_select<false>, 1, serial);
    if (!rc) rc = raise_dyn_smem((const void*)k_select<true>, 2, serial);
    if (!rc) rc = raise_dyn_smem((const void*)k_select_fast<true, SEL_WARPS>, 4, fast);
    if (!rc) rc = raise_dyn_smem((const void*)k_select_fast<false, SEL_WARPS>, 3, fast);
    const int wide = fast + (SEL_WARPS_WIDE - SEL_WARPS) * SEL_WCAP * 6;
    if (wide <= 227 * 1024) {
        if (!rc) rc = raise_dyn_smem((const void*)k_select_fast<true, SEL_WARPS_WIDE>, 7, wide);
        if (!rc) rc = raise_dyn_smem((const void*)k_select_fast<false, SEL_WARPS_WIDE>, 8, wide);
    }
    return rc;
}
