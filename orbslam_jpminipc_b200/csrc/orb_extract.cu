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
#include "introselect.h"

namespace {

__device__ __forceinline__ int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

// ------------------------------------------------------------------ K0
__global__ void __launch_bounds__(256)
k_level0(const uint8_t* __restrict__ src, int w, int h, int sstride, size_t spitch,
         uint8_t* __restrict__ planes, size_t fbytes, int pstride, int prows)
{
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int f = blockIdx.z;
    if (x4 >= pstride || y >= prows) return;
    const uint8_t* S = src + (size_t)f * spitch + (size_t)reflect101(y - ORB_EDGE, h) * sstride;
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int x = x4 + k;
        const uint32_t px = (x < w + 2 * ORB_EDGE) ? __ldg(S + reflect101(x - ORB_EDGE, w)) : 0u;
        v |= px << (8 * k);
    }
    *reinterpret_cast<uint32_t*>(planes + (size_t)f * fbytes + (size_t)y * pstride + x4) = v;
}

// ------------------------------------------------------------------ K1
// Each thread produces 4 horizontally adjacent pixels of the PADDED destination plane; border
// pixels evaluate the resize at their reflected coordinate, which is what copyMakeBorder copies.
__global__ void __launch_bounds__(256)
k_resize(uint8_t* __restrict__ planes, size_t fbytes, LevelGeom S, LevelGeom D,
         const int2* __restrict__ xtab, const int2* __restrict__ ytab)
{
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int f = blockIdx.z;
    if (x4 >= D.stride || y >= D.prows) return;
    uint8_t* base = planes + (size_t)f * fbytes;
    const uint8_t* sroi = base + S.plane_off + (size_t)ORB_EDGE * S.stride + ORB_EDGE;
    const int2 yt = __ldg(ytab + D.ytab_off + reflect101(y - ORB_EDGE, D.h));
    const uint8_t* S0 = sroi + (size_t)(yt.x & 0xffff) * S.stride;
    const uint8_t* S1 = sroi + (size_t)(yt.x >> 16) * S.stride;
    const int b0 = (short)(yt.y & 0xffff), b1 = (short)(yt.y >> 16);
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int x = x4 + k;
        uint32_t px = 0;
        if (x < D.w + 2 * ORB_EDGE) {
            const int2 xt = __ldg(xtab + D.xtab_off + reflect101(x - ORB_EDGE, D.w));
            const int sx0 = xt.x & 0xffff, sx1 = xt.x >> 16;
            const int a0 = (short)(xt.y & 0xffff), a1 = (short)(xt.y >> 16);
            const int r0 = S0[sx0] * a0 + S0[sx1] * a1;
            const int r1 = S1[sx0] * a0 + S1[sx1] * a1;
            int o = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
            px = (uint32_t)min(max(o, 0), 255);
        }
        v |= px << (8 * k);
    }
    *reinterpret_cast<uint32_t*>(base + D.plane_off + (size_t)y * D.stride + x4) = v;
}

// ------------------------------------------------------------------ K2
constexpr int FT_W = ORB_TILE_W, FT_H = ORB_TILE_H;
constexpr int FI_W = FT_W + 8, FI_H = FT_H + 8;       // image tile: 3 (ring) + 1 (NMS) px halo
constexpr int FS_W = FT_W + 4, FS_H = FT_H + 2;       // score tile: 1 px halo (row pitch padded to 68)

// FAST-9/16 corner strength at p: max over the 16 arcs of 9 contiguous ring pixels of
// min(v - ring) (dark arc) and min(ring - v) (bright arc).  corner at threshold t <=> result > t;
// OpenCV's response is result-1.  Returns <= th when the quick test proves "not a corner".
__device__ __forceinline__ int fast_strength(const uint8_t* p, int th)
{
    const int v = p[0];
    int d[16];
    d[0] = v - p[3 * FI_W];      d[8] = v - p[-3 * FI_W];
    d[4] = v - p[3];             d[12] = v - p[-3];
    // every 9-arc contains one of each opposite pair: need |d| > th on both tested pairs
    if (max(abs(d[0]), abs(d[8])) <= th || max(abs(d[4]), abs(d[12])) <= th) return 0;
    d[1] = v - p[3 * FI_W + 1];  d[2] = v - p[2 * FI_W + 2];   d[3] = v - p[FI_W + 3];
    d[5] = v - p[-FI_W + 3];     d[6] = v - p[-2 * FI_W + 2];  d[7] = v - p[-3 * FI_W + 1];
    d[9] = v - p[-3 * FI_W - 1]; d[10] = v - p[-2 * FI_W - 2]; d[11] = v - p[-FI_W - 3];
    d[13] = v - p[FI_W - 3];     d[14] = v - p[2 * FI_W - 2];  d[15] = v - p[3 * FI_W - 1];
    int mn2[16], mx2[16], mn4[16], mx4[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { mn2[k] = min(d[k], d[(k + 1) & 15]); mx2[k] = max(d[k], d[(k + 1) & 15]); }
#pragma unroll
    for (int k = 0; k < 16; k++) { mn4[k] = min(mn2[k], mn2[(k + 2) & 15]); mx4[k] = max(mx2[k], mx2[(k + 2) & 15]); }
    int best = -256, worst = 256;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int mn9 = min(min(mn4[k], mn4[(k + 4) & 15]), d[(k + 8) & 15]);
        const int mx9 = max(max(mx4[k], mx4[(k + 4) & 15]), d[(k + 8) & 15]);
        best = max(best, mn9);
        worst = min(worst, mx9);
    }
    return max(best, -worst);
}

__global__ void __launch_bounds__(256)
k_fast_nms(const uint8_t* __restrict__ planes, uint8_t* __restrict__ nms, size_t fbytes,
           const Plan* __restrict__ plan, const Tile* __restrict__ tiles)
{
    __shared__ __align__(16) uint8_t img[FI_H * FI_W];
    __shared__ __align__(16) uint8_t sc[FS_H * FS_W];
    __shared__ short colcell[FS_W], rowcell[FS_H];
    const Tile t = tiles[blockIdx.x];
    const LevelGeom& L = plan->L[t.level];
    const int f = blockIdx.y, tid = threadIdx.x;
    const uint8_t* plane = planes + (size_t)f * fbytes + L.plane_off;
    const int th = plan->th_lo;

    // image tile, origin (x0-4, y0-4) in ROI coords = (+16,+16) in padded coords; 32-bit loads
    {
        const int px0 = t.x0 - 4 + ORB_EDGE, py0 = t.y0 - 4 + ORB_EDGE;
        for (int i = tid; i < FI_H * (FI_W / 4); i += 256) {
            const int r = i / (FI_W / 4), cw = i - r * (FI_W / 4);
            const int py = py0 + r, px = px0 + cw * 4;
            uint32_t v = 0;
            if (py < L.prows && px + 3 < L.stride) v = __ldg(reinterpret_cast<const uint32_t*>(plane + (size_t)py * L.stride + px));
            reinterpret_cast<uint32_t*>(img)[i] = v;
        }
    }
    // cell id of each score-tile column / row (-1: outside every detection rectangle)
    if (tid < FS_W) {
        const int x = t.x0 - 1 + tid;
        int c = -1;
        if (x >= ORB_EDGE && tid < FT_W + 2) {
            c = (x - ORB_EDGE) / L.cellW;
            if (c >= L.cols - 1) { c = L.cols - 1; if (x >= L.w - ORB_EDGE) c = -1; }
        }
        colcell[tid] = (short)c;
    } else if (tid >= 128 && tid < 128 + FS_H) {
        const int i = tid - 128, y = t.y0 - 1 + i;
        int c = -1;
        if (y >= ORB_EDGE) {
            c = (y - ORB_EDGE) / L.cellH;
            if (c >= L.rows - 1) { c = L.rows - 1; if (y >= L.h - ORB_EDGE) c = -1; }
        }
        rowcell[i] = (short)c;
    }
    __syncthreads();
    // corner strength for the (FT_H+2) x (FT_W+2) positions
    for (int i = tid; i < FS_H * (FT_W + 2); i += 256) {
        const int r = i / (FT_W + 2), c = i - r * (FT_W + 2);
        int s = 0;
        if (colcell[c] >= 0 && rowcell[r] >= 0) {
            s = fast_strength(img + (r + 3) * FI_W + (c + 3), th);
            s = s > th ? s - 1 : 0;
        }
        sc[r * FS_W + c] = (uint8_t)s;
    }
    __syncthreads();
    // NMS restricted to the pixel's own cell; 4 pixels per thread, one 32-bit store
    uint8_t* out = nms + (size_t)f * fbytes + L.plane_off;
    for (int i = tid; i < FT_H * (FT_W / 4); i += 256) {
        const int r = i / (FT_W / 4), c4 = (i - r * (FT_W / 4)) * 4;
        uint32_t v = 0;
        const int rc = rowcell[r + 1];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int c = c4 + k;
            const uint8_t* q = sc + (r + 1) * FS_W + (c + 1);
            const int s = q[0];
            if (s == 0) continue;
            const int cc = colcell[c + 1];
            const bool l = colcell[c] == cc, rt = colcell[c + 2] == cc;
            const bool u = rowcell[r] == rc, dn = rowcell[r + 2] == rc;
            bool keep = true;
            keep = keep && !(l && q[-1] >= s) && !(rt && q[1] >= s);
            keep = keep && !(u && (q[-FS_W] >= s || (l && q[-FS_W - 1] >= s) || (rt && q[-FS_W + 1] >= s)));
            keep = keep && !(dn && (q[FS_W] >= s || (l && q[FS_W - 1] >= s) || (rt && q[FS_W + 1] >= s)));
            if (keep) v |= (uint32_t)s << (8 * k);
        }
        const int py = t.y0 + r + ORB_EDGE, px = t.x0 + c4 + ORB_EDGE;
        if (py < L.prows && px + 3 < L.stride)
            *reinterpret_cast<uint32_t*>(out + (size_t)py * L.stride + px) = v;
    }
}

// ------------------------------------------------------------------ K3
// One warp per (frame, cell): scans the cell's detection rectangle of the NMS map in raster
// order and appends survivors with ballot/popc prefix sums, so the list order is exactly the
// order cv::FAST emits (y, then x).  Then applies the reference's fallback: if fewer than 4
// survive at fastTh the cell is re-detected at threshold 7 (src/ORBextractor.cc:609-614) —
// both sets are sub-sequences of the th_lo list ("one-pass fallback").
// record = score<<24 | y_local<<12 | x_local   (cell-image coordinates, as cv::FAST reports)
__global__ void __launch_bounds__(256)
k_cell_compact(const uint8_t* __restrict__ nms, size_t fbytes, const Plan* __restrict__ plan,
               const CellGeom* __restrict__ cells, uint32_t* __restrict__ cand, int* __restrict__ ntotal)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= plan->ncells) return;
    const int f = blockIdx.y;
    const CellGeom g = cells[warp];
    const LevelGeom& L = plan->L[g.level];
    const uint8_t* map = nms + (size_t)f * fbytes + L.plane_off + (size_t)ORB_EDGE * L.stride + ORB_EDGE;
    uint32_t* out = cand + (size_t)f * plan->cand_total + g.cand_off;
    const int thP = plan->fast_th;
    int count = 0, nP = 0, n7 = 0;
    const uint32_t lt = (1u << lane) - 1;
    for (int y = g.y0; y < g.y1; y++) {
        const uint8_t* row = map + (size_t)y * L.stride;
        for (int xb = g.x0; xb < g.x1; xb += 32) {
            const int x = xb + lane;
            const int s = x < g.x1 ? row[x] : 0;
            const uint32_t m = __ballot_sync(0xffffffffu, s > 0);
            if (s > 0) out[count + __popc(m & lt)] = ((uint32_t)s << 24) | ((uint32_t)(y - g.iniy) << 12) | (uint32_t)(x - g.inix);
            count += __popc(m);
            nP += __popc(__ballot_sync(0xffffffffu, s >= thP));
            n7 += __popc(__ballot_sync(0xffffffffu, s >= 7));
        }
    }
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
    if (lane == 0) ntotal[(size_t)f * plan->ncells + warp] = count;
}

// ------------------------------------------------------------------ K4
// One CTA per (frame, level).  Thread 0 replays the quota redistribution loop (:622-670); every
// cell then runs retainBest (= libstdc++ introselect, first n survivors, see introselect.h) on
// its own list in parallel; the survivors are concatenated in cell order, converted to level
// coordinates, and capped to nDesired by a second introselect (:697-701).
// level record = score<<32 | y<<16 | x  (level ROI coordinates)
__global__ void __launch_bounds__(128)
k_select(const Plan* __restrict__ plan, const CellGeom* __restrict__ cells, uint32_t* __restrict__ cand,
         const int* __restrict__ ntotal, unsigned long long* __restrict__ lvl, int* __restrict__ nkept,
         int* __restrict__ status)
{
    extern __shared__ unsigned long long s_list[];
    __shared__ int s_total[ORB_MAX_CELLS_LEVEL], s_retain[ORB_MAX_CELLS_LEVEL], s_off[ORB_MAX_CELLS_LEVEL + 1];
    const int level = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const LevelGeom& L = plan->L[level];
    const int nCells = L.ncells;
    const CellGeom* cg = cells + L.cell_base;
    const int* nt = ntotal + (size_t)f * plan->ncells + L.cell_base;
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
        int o = 0;
        for (int c = 0; c < nCells; c++) { s_off[c] = o; o += s_retain[c]; }
        s_off[nCells] = o;
        if (o > L.lvl_cap) { atomicExch(status, ORB_ERR_CAPACITY); s_off[nCells] = -1; }
    }
    __syncthreads();
    int total = s_off[nCells];
    if (total < 0) { if (tid == 0) nkept[f * plan->nlevels + level] = 0; return; }
    for (int c = tid; c < nCells; c += blockDim.x) {
        const int n = s_total[c], keep = s_retain[c];
        uint32_t* v = cand + (size_t)f * plan->cand_total + cg[c].cand_off;
        if (n > keep && keep > 0) orbsel::nth_element(v, n, keep - 1, orbsel::KeyGreater<uint32_t, 24>());
        const int ix = cg[c].inix, iy = cg[c].iniy;
        for (int k = 0; k < keep; k++) {
            const uint32_t r = v[k];
            const unsigned long long x = (r & 0xfff) + ix, y = ((r >> 12) & 0xfff) + iy;
            s_list[s_off[c] + k] = ((unsigned long long)(r >> 24) << 32) | (y << 16) | x;
        }
    }
    __syncthreads();
    if (total > L.nDesired) {
        if (tid == 0) orbsel::nth_element(s_list, total, L.nDesired - 1, orbsel::KeyGreater<unsigned long long, 32>());
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
constexpr int BT_W = ORB_TILE_W, BT_H = ORB_TILE_H;
constexpr int BI_W = BT_W + 8, BI_H = BT_H + 6;      // input tile: 3 px halo (4 on x for alignment)

__global__ void __launch_bounds__(256)
k_blur(const uint8_t* __restrict__ planes, uint8_t* __restrict__ blurred, size_t fbytes,
       const Plan* __restrict__ plan, const Tile* __restrict__ tiles)
{
    __shared__ __align__(16) uint8_t img[BI_H * BI_W];
    __shared__ float rowp[BI_H * BT_W];
    const float k0 = __uint_as_float(0x3d8fafb1u), k1 = __uint_as_float(0x3e06387eu),
                k2 = __uint_as_float(0x3e434a39u), k3 = __uint_as_float(0x3e5d4ae0u);
    const Tile t = tiles[blockIdx.x];
    const LevelGeom& L = plan->L[t.level];
    const int f = blockIdx.y, tid = threadIdx.x;
    const uint8_t* plane = planes + (size_t)f * fbytes + L.plane_off;
    {
        const int px0 = t.x0 - 4 + ORB_EDGE, py0 = t.y0 - 3 + ORB_EDGE;
        for (int i = tid; i < BI_H * (BI_W / 4); i += 256) {
            const int r = i / (BI_W / 4), cw = i - r * (BI_W / 4);
            const int py = py0 + r, px = px0 + cw * 4;
            uint32_t v = 0;
            if (py < L.prows && px + 3 < L.stride) v = __ldg(reinterpret_cast<const uint32_t*>(plane + (size_t)py * L.stride + px));
            reinterpret_cast<uint32_t*>(img)[i] = v;
        }
    }
    __syncthreads();
    for (int i = tid; i < BI_H * BT_W; i += 256) {
        const int r = i / BT_W, c = i - r * BT_W;
        const uint8_t* p = img + r * BI_W + c + 1;          // p[0] = column c-3
        float s = __fmul_rn((float)p[0], k0);
        s = fmaf((float)p[1], k1, s); s = fmaf((float)p[2], k2, s); s = fmaf((float)p[3], k3, s);
        s = fmaf((float)p[4], k2, s); s = fmaf((float)p[5], k1, s); s = fmaf((float)p[6], k0, s);
        rowp[i] = s;
    }
    __syncthreads();
    uint8_t* out = blurred + (size_t)f * fbytes + L.plane_off;
    for (int i = tid; i < BT_H * (BT_W / 4); i += 256) {
        const int r = i / (BT_W / 4), c4 = (i - r * (BT_W / 4)) * 4;
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float* q = rowp + (r + 3) * BT_W + c4 + k;
            float s = __fmul_rn(k3, q[0]);
            s = fmaf(__fadd_rn(q[BT_W], q[-BT_W]), k2, s);
            s = fmaf(__fadd_rn(q[2 * BT_W], q[-2 * BT_W]), k1, s);
            s = fmaf(__fadd_rn(q[3 * BT_W], q[-3 * BT_W]), k0, s);
            const int o = min(max(__float2int_rn(s), 0), 255);
            v |= (uint32_t)o << (8 * k);
        }
        const int y = t.y0 + r, x = t.x0 + c4;
        if (y < L.h && x < L.w)      // pixels past the ROI edge land in the border, which is never read from here
            *reinterpret_cast<uint32_t*>(out + (size_t)(y + ORB_EDGE) * L.stride + x + ORB_EDGE) = v;
    }
}

// ------------------------------------------------------------------ K6
__constant__ int8_t c_pattern[1024] = {
#include "orb_pattern.inc"
};
__constant__ int c_umax[16];

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

// One warp per output keypoint slot.  Lanes 0..30 own patch column u = lane-15 for the moments;
// lane i then owns descriptor byte i (8 tests, 16 rotated samples).
__global__ void __launch_bounds__(256)
k_describe(const uint8_t* __restrict__ planes, const uint8_t* __restrict__ blurred, size_t fbytes,
           const Plan* __restrict__ plan, const unsigned long long* __restrict__ lvl, const int* __restrict__ nkept,
           orb_keypoint* __restrict__ kps, uint8_t* __restrict__ desc, int cap, int32_t* __restrict__ counts)
{
    __shared__ int8_t s_pat[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_pat[i] = c_pattern[i];
    __syncthreads();
    const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.y;
    const int nl = plan->nlevels;
    const int* nk = nkept + f * nl;
    int level = -1, idx = 0, acc = 0;
    for (int l = 0; l < nl; l++) {
        const int n = nk[l];
        if (level < 0 && slot < acc + n) { level = l; idx = slot - acc; }
        acc += n;
    }
    if (slot == 0 && lane == 0) counts[f] = acc;        // > cap means the caller's buffers truncated the output
    if (level < 0 || slot >= cap) return;
    const LevelGeom& L = plan->L[level];
    const unsigned long long rec = lvl[(size_t)f * plan->lvl_total + L.lvl_base + idx];
    const int x = (int)(rec & 0xffff), y = (int)((rec >> 16) & 0xffff), score = (int)(rec >> 32);
    const uint8_t* roi = planes + (size_t)f * fbytes + L.plane_off + (size_t)ORB_EDGE * L.stride + ORB_EDGE;
    const uint8_t* broi = blurred + (size_t)f * fbytes + L.plane_off + (size_t)ORB_EDGE * L.stride + ORB_EDGE;
    const uint8_t* center = roi + (size_t)y * L.stride + x;

    // IC_Angle (:124-151): m10 = sum u*I, m01 = sum v*I over the radius-15 disc
    int m10 = 0, m01 = 0;
    const int u = lane - 15;
    if (lane < 31) {
        const int au = abs(u);
#pragma unroll
        for (int v = -15; v <= 15; v++) {
            if (au <= c_umax[v < 0 ? -v : v]) {
                const int val = center[v * L.stride + u];
                m10 += u * val; m01 += v * val;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, o); m01 += __shfl_xor_sync(0xffffffffu, m01, o); }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // computeOrbDescriptor (:155-194); cos/sin pinned to correctly rounded FP32 via double
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    const float arad = __fmul_rn(angle, factorPI);
    float a, b;
    if (lane == 0) { a = (float)cos((double)arad); b = (float)sin((double)arad); }
    a = __shfl_sync(0xffffffffu, a, 0); b = __shfl_sync(0xffffffffu, b, 0);
    const int8_t* pat = s_pat + lane * 32;
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int t[2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const float px = (float)pat[4 * k + 2 * e], py = (float)pat[4 * k + 2 * e + 1];
            const int iy = __float2int_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)));
            const int ix = __float2int_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)));
            const int sx = x + ix, sy = y + iy;
            // the in-place blur only rewrites the ROI: samples that fall into the 16-px border
            // read the un-blurred reflected pixels (:760)
            const bool inside = sx >= 0 && sx < L.w && sy >= 0 && sy < L.h;
            const uint8_t* base = inside ? broi : roi;
            t[e] = base[(ptrdiff_t)sy * L.stride + sx];
        }
        val |= (t[0] < t[1]) << k;
    }
    desc[((size_t)f * cap + slot) * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        orb_keypoint kp;
        float fx = (float)x, fy = (float)y;
        if (level != 0) { fx = __fmul_rn(fx, L.scale); fy = __fmul_rn(fy, L.scale); }   // :769-775
        kp.x = fx; kp.y = fy; kp.size = (float)L.patch_size; kp.angle = angle;
        kp.response = (float)score; kp.octave = level; kp.class_id = -1;
        kps[(size_t)f * cap + slot] = kp;
    }
}

} // namespace

int orb_upload_constants(const int* umax)
{
    ORB_CUDA(cudaMemcpyToSymbol(c_umax, umax, sizeof(int) * 16));
    return ORB_OK;
}

int orb_launch_extract(orb_ctx* c, const uint8_t* d_imgs, int nimg, int w, int h, int stride, size_t frame_pitch,
                       orb_keypoint* d_kps, uint8_t* d_desc, int cap, int32_t* d_counts, cudaStream_t s)
{
    const Plan& P = c->plan;
    const size_t fb = (size_t)P.frame_bytes;
    int launches = 0;
    const dim3 blk(64, 4);
    auto mark = [&]() {           // stage boundary event (profiling mode only)
        if (!c->profile) return;
        cudaEvent_t e;
        if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); }
        else if (cudaEventCreate(&e) != cudaSuccess) return;
        cudaEventRecord(e, s);
        c->prof_events.push_back(e);
    };
    mark();
    {
        const LevelGeom& L = P.L[0];
        dim3 grid((L.stride / 4 + 63) / 64, (L.prows + 3) / 4, nimg);
        k_level0<<<grid, blk, 0, s>>>(d_imgs, w, h, stride, frame_pitch, c->d_planes, fb, L.stride, L.prows);
        launches++;
    }
    mark();
    for (int l = 1; l < P.nlevels; l++) {
        const LevelGeom& D = P.L[l];
        dim3 grid((D.stride / 4 + 63) / 64, (D.prows + 3) / 4, nimg);
        k_resize<<<grid, blk, 0, s>>>(c->d_planes, fb, P.L[l - 1], D, c->d_xtab, c->d_ytab);
        launches++;
    }
    mark();
    k_fast_nms<<<dim3(P.ntiles_fast, nimg), 256, 0, s>>>(c->d_planes, c->d_work, fb, c->d_plan, c->d_tiles_fast);
    mark();
    k_cell_compact<<<dim3((P.ncells + 7) / 8, nimg), 256, 0, s>>>(c->d_work, fb, c->d_plan, c->d_cells, c->d_cand, c->d_ntotal);
    mark();
    int maxcap = 0;
    for (int l = 0; l < P.nlevels; l++) maxcap = std::max(maxcap, P.L[l].lvl_cap);
    k_select<<<dim3(P.nlevels, nimg), 128, (size_t)maxcap * 8, s>>>(c->d_plan, c->d_cells, c->d_cand, c->d_ntotal, c->d_lvl, c->d_nkept, c->d_status);
    mark();
    k_blur<<<dim3(P.ntiles_blur, nimg), 256, 0, s>>>(c->d_planes, c->d_work, fb, c->d_plan, c->d_tiles_blur);
    mark();
    const int slots = std::min(cap, P.kp_cap);
    k_describe<<<dim3((std::max(slots, 1) + 7) / 8, nimg), 256, 0, s>>>(c->d_planes, c->d_work, fb, c->d_plan, c->d_lvl, c->d_nkept,
                                                                    d_kps, d_desc, cap, d_counts);
    mark();
    launches += 5;
    c->last_launches = launches;
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

int orb_select_smem_setup(int max_bytes)
{
    ORB_CUDA(cudaFuncSetAttribute(k_select, cudaFuncAttributeMaxDynamicSharedMemorySize, max_bytes));
    return ORB_OK;
}
