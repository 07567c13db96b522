// orb_match.cu — sm_100a kernels for the ORBmatcher hot loops (reference src/ORBmatcher.cc).
//
//   K7 k_knn2            DescriptorDistance (:1794-1810) + best/second-best scan (:197-222) over
//                        every DB row: queries in registers, DB tiles staged in shared memory by
//                        the bulk-copy engine (cp.async.bulk + mbarrier), XOR/POPC/min in the
//                        integer pipes.  No tensor cores: this is popcount work.
//      k_knn2_merge      exact merge of per-chunk / per-shard (idx1,d1,d2) triples
//      k_match_ratio     acceptance test (:224-226)
//   K8 k_grid_*          Frame grid (src/Frame.cc:109-123,:267-277)
//      k_sbp_*           SearchByProjection(Frame&,const Frame&,th) (:1507-1620)
//   K9 k_bow_*           SearchByBoW candidate scoring (:155-284)
#include "orb_internal.h"
#include <algorithm>
#include <climits>

namespace {

// ------------------------------------------------------------------ helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared (TMA engine; SASS UBLKCP), completion counted on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------ K7
constexpr int KNN_THREADS = 256;      // one query per thread
constexpr int KNN_TILE = 256;         // DB rows per shared-memory stage (8 KB)
constexpr int KNN_STAGES = 3;
constexpr int KNN_KEY_SHIFT = 22;     // packed key = distance << 22 | row within the chunk: chunks hold at most 2^22 rows (distance <= 256 fits above)

__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xe8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

struct Knn2Args {
    const uint8_t* q; const uint8_t* db;
    int nq; long long ndb;           // per pair
    int rows_per_chunk, nchunks;
    int32_t idx_base;
    int32_t* out;                    // nchunks > 1: partials [pair][chunk][3][nq]; else final idx1
    int32_t* o_idx1; int32_t* o_d1; int32_t* o_d2;
};

__global__ void __launch_bounds__(KNN_THREADS)
k_knn2(Knn2Args A)
{
    __shared__ __align__(128) uint4 tile[KNN_STAGES][KNN_TILE * 2];
    __shared__ __align__(8) uint64_t bar[KNN_STAGES];
    const int tid = threadIdx.x;
    const int chunk = blockIdx.x, qt = blockIdx.y, pair = blockIdx.z;
    const long long row0 = (long long)chunk * A.rows_per_chunk;
    const int nrows = (int)min((long long)A.rows_per_chunk, A.ndb - row0);
    const uint8_t* db = A.db + ((size_t)pair * A.ndb + row0) * 32;
    const int qi = qt * KNN_THREADS + tid;
    uint32_t qw[8];
    {
        const uint4* qp = reinterpret_cast<const uint4*>(A.q + ((size_t)pair * A.nq + min(qi, A.nq - 1)) * 32);
        const uint4 a = __ldg(qp), b = __ldg(qp + 1);
        qw[0] = a.x; qw[1] = a.y; qw[2] = a.z; qw[3] = a.w; qw[4] = b.x; qw[5] = b.y; qw[6] = b.z; qw[7] = b.w;
    }
    if (tid == 0) {
        for (int s = 0; s < KNN_STAGES; s++) mbar_init(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int ntiles = (nrows + KNN_TILE - 1) / KNN_TILE;
    auto issue = [&](int t) {
        const int s = t % KNN_STAGES;
        const int rows = min(KNN_TILE, nrows - t * KNN_TILE);
        mbar_expect_tx(&bar[s], (uint32_t)rows * 32);
        bulk_g2s(&tile[s][0], db + (size_t)t * KNN_TILE * 32, (uint32_t)rows * 32, &bar[s]);
    };
    if (tid == 0) for (int t = 0; t < min(KNN_STAGES - 1, ntiles); t++) issue(t);

#ifdef ORB_KNN_PLAIN      // the straightforward form: 8 XOR + 8 POPC + 7 adds per pair, POPC-pipe bound (16 lanes/clk/SM); kept for A/B timing
    int d1 = INT_MAX, d2 = INT_MAX, i1 = -1;
    for (int t = 0; t < ntiles; t++) {
        const int s = t % KNN_STAGES;
        // the stage refilled now was consumed in iteration t-1; the barrier below orders that
        __syncthreads();
        if (tid == 0 && t + KNN_STAGES - 1 < ntiles) issue(t + KNN_STAGES - 1);
        mbar_wait(&bar[s], (uint32_t)((t / KNN_STAGES) & 1));
        const int rows = min(KNN_TILE, nrows - t * KNN_TILE);
        const uint4* tp = &tile[s][0];
        const int jbase = t * KNN_TILE;
#pragma unroll 4
        for (int r = 0; r < rows; r++) {
            const uint4 a = tp[2 * r], b = tp[2 * r + 1];            // broadcast reads
            int d = __popc(qw[0] ^ a.x) + __popc(qw[1] ^ a.y) + __popc(qw[2] ^ a.z) + __popc(qw[3] ^ a.w)
                  + __popc(qw[4] ^ b.x) + __popc(qw[5] ^ b.y) + __popc(qw[6] ^ b.z) + __popc(qw[7] ^ b.w);
            // strict '<' scan in ascending row order: first minimum wins, d2 = 2nd of the multiset
            const bool lt = d < d1;
            d2 = lt ? d1 : min(d2, d);
            i1 = lt ? jbase + r : i1;
            d1 = min(d1, d);
        }
    }
#else
    // POPC runs at a quarter of the ALU rate (16 vs 64 lanes/clk/SM), so the eight XOR words of a pair are first compressed by three
    // carry-save adders (LOP3 0x96 = sum, 0xE8 = majority): x0..x7 -> two words of weight 1 (s3, x7) and three of weight 2
    // (c1, c2, c3).  5 POPC + 14 LOP3 instead of 8 POPC + 8 LOP3 — the two pipes end up equally loaded (~40 clk per warp and row),
    // exact like any adder tree.  The running best / second-best are two packed keys  distance << 22 | row-in-chunk  (strict '<'
    // scan == minimum key: lowest row among equal distances; second order statistic == second smallest key):
    //     m2 = min(m2, max(m1, key)),  m1 = min(m1, key)
    // three min/max instead of compare + two selects + two min, and the key is assembled by two IMADs on the otherwise idle FMA pipe.
    int m1 = INT_MAX, m2 = INT_MAX;
    for (int t = 0; t < ntiles; t++) {
        const int s = t % KNN_STAGES;
        // the stage refilled now was consumed in iteration t-1; the barrier below orders that
        __syncthreads();
        if (tid == 0 && t + KNN_STAGES - 1 < ntiles) issue(t + KNN_STAGES - 1);
        mbar_wait(&bar[s], (uint32_t)((t / KNN_STAGES) & 1));
        const int rows = min(KNN_TILE, nrows - t * KNN_TILE);
        const uint4* tp = &tile[s][0];
        const int jbase = t * KNN_TILE;
#pragma unroll 4
        for (int r = 0; r < rows; r++) {
            const uint4 a = tp[2 * r], b = tp[2 * r + 1];            // broadcast reads
            const uint32_t x0 = qw[0] ^ a.x, x1 = qw[1] ^ a.y, x2 = qw[2] ^ a.z, x3 = qw[3] ^ a.w;
            const uint32_t x4 = qw[4] ^ b.x, x5 = qw[5] ^ b.y, x6 = qw[6] ^ b.z, x7 = qw[7] ^ b.w;
            const uint32_t s1 = xor3(x0, x1, x2), c1 = maj3(x0, x1, x2);
            const uint32_t s2 = xor3(x3, x4, x5), c2 = maj3(x3, x4, x5);
            const uint32_t s3 = xor3(s1, s2, x6), c3 = maj3(s1, s2, x6);
            const int ones = __popc(s3) + __popc(x7);
            const int twos = __popc(c1) + __popc(c2) + __popc(c3);
            const int key = twos * (1 << (KNN_KEY_SHIFT + 1)) + (ones * (1 << KNN_KEY_SHIFT) + (jbase + r));
            m2 = min(m2, max(m1, key));
            m1 = min(m1, key);
        }
    }
    const int d1 = m1 == INT_MAX ? INT_MAX : m1 >> KNN_KEY_SHIFT, d2 = m2 == INT_MAX ? INT_MAX : m2 >> KNN_KEY_SHIFT;
    const int i1 = m1 == INT_MAX ? -1 : m1 & ((1 << KNN_KEY_SHIFT) - 1);
#endif
    if (qi < A.nq) {
        const int gi = i1 < 0 ? -1 : (int)(row0 + i1) + A.idx_base;
        if (A.nchunks > 1) {
            int32_t* o = A.out + ((size_t)(pair * A.nchunks + chunk) * 3) * A.nq;
            o[qi] = gi; o[A.nq + qi] = d1; o[2 * A.nq + qi] = d2;
        } else {
            const size_t o = (size_t)pair * A.nq + qi;
            A.o_idx1[o] = gi; A.o_d1[o] = d1; A.o_d2[o] = d2;
        }
    }
}

// parts[(p*3+k)*nq + i] for part p; parts are in ascending global-index order.
__global__ void k_knn2_merge(const int32_t* __restrict__ parts, int nparts, int nq, int npairs,
                             int32_t* __restrict__ idx1, int32_t* __restrict__ d1, int32_t* __restrict__ d2)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int pair = blockIdx.y;
    if (i >= nq || pair >= npairs) return;
    const int32_t* P = parts + (size_t)pair * nparts * 3 * nq;
    int b1 = INT_MAX, b2 = INT_MAX, bi = -1;
    for (int p = 0; p < nparts; p++) {
        const int pi = P[(size_t)(p * 3) * nq + i], pd1 = P[(size_t)(p * 3 + 1) * nq + i], pd2 = P[(size_t)(p * 3 + 2) * nq + i];
        if (pi < 0) continue;
        if (pd1 < b1) { b2 = b1; b1 = pd1; bi = pi; } else if (pd1 < b2) b2 = pd1;
        if (pd2 < b2) b2 = pd2;
    }
    const size_t o = (size_t)pair * nq + i;
    idx1[o] = bi; d1[o] = b1; d2[o] = b2;
}

// the same merge over parts that are separate buffers (one per rank; peer-device memory is read in place over NVLink)
struct MergePtrs { const int32_t* p[ORB_COMM_MAX_RANKS]; };
__global__ void k_knn2_merge_ptrs(MergePtrs parts, int nparts, int nq, int32_t* __restrict__ idx1, int32_t* __restrict__ d1, int32_t* __restrict__ d2)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    int b1 = INT_MAX, b2 = INT_MAX, bi = -1;
#pragma unroll 1
    for (int p = 0; p < nparts; p++) {
        const int32_t* P = parts.p[p];
        const int pi = P[i], pd1 = P[(size_t)nq + i], pd2 = P[(size_t)2 * nq + i];
        if (pi < 0) continue;
        if (pd1 < b1) { b2 = b1; b1 = pd1; bi = pi; } else if (pd1 < b2) b2 = pd1;
        if (pd2 < b2) b2 = pd2;
    }
    idx1[i] = bi; d1[i] = b1; d2[i] = b2;
}

__global__ void k_match_ratio(const int32_t* __restrict__ idx1, const int32_t* __restrict__ d1, const int32_t* __restrict__ d2,
                              int nq, float nnratio, int th, int32_t* __restrict__ match, int* __restrict__ count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false;
    if (i < nq) {
        ok = idx1[i] >= 0 && d1[i] <= th && (float)d1[i] < __fmul_rn(nnratio, (float)d2[i]);
        match[i] = ok ? idx1[i] : -1;
    }
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, __popc(m));
}

// POPC-pipe peak: register-resident dependent-free popcounts
__global__ void k_popc_bench(uint32_t* out, int iters)
{
    uint32_t a = threadIdx.x * 2654435761u + 1, b = a ^ 0x9e3779b9u, c = a + 77, d = b + 1234567;
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            s0 += __popc(a); s1 += __popc(b); s2 += __popc(c); s3 += __popc(d);
            a ^= s0; b ^= s1; c ^= s2; d ^= s3;
        }
    }
    if ((s0 ^ s1 ^ s2 ^ s3) == 0x12345u) out[0] = s0;
}

// ------------------------------------------------------------------ K8: Frame grid
// PosInGrid (src/Frame.cc:267-277): round() half away from zero; cells outside the grid drop
// the keypoint.  CSR order inside a cell = ascending keypoint index (push_back order, :116-123).
__global__ void __launch_bounds__(1024)
k_grid_build(const orb_keypoint* __restrict__ kps, int n, int min_x, int max_x, int min_y, int max_y,
             int32_t* __restrict__ cell_start, int32_t* __restrict__ cell_items)
{
    extern __shared__ unsigned short s_cell[];            // n entries
    __shared__ int s_cnt[ORB_GRID_COLS * ORB_GRID_ROWS];  // histogram, then running fill position
    __shared__ int s_warp[32];
    const int NC = ORB_GRID_COLS * ORB_GRID_ROWS;
    const int tid = threadIdx.x;
    const float invW = __fdiv_rn((float)ORB_GRID_COLS, (float)(max_x - min_x));
    const float invH = __fdiv_rn((float)ORB_GRID_ROWS, (float)(max_y - min_y));
    for (int c = tid; c < NC; c += blockDim.x) s_cnt[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
        const int px = (int)roundf(__fmul_rn(__fsub_rn(kps[i].x, (float)min_x), invW));
        const int py = (int)roundf(__fmul_rn(__fsub_rn(kps[i].y, (float)min_y), invH));
        const bool ok = !(px < 0 || px >= ORB_GRID_COLS || py < 0 || py >= ORB_GRID_ROWS);
        const int c = px * ORB_GRID_ROWS + py;
        s_cell[i] = ok ? (unsigned short)c : (unsigned short)0xffff;
        if (ok) atomicAdd(&s_cnt[c], 1);
    }
    __syncthreads();
    // block exclusive scan of the histogram (each thread owns 3 consecutive cells)
    int cnt[3];
#pragma unroll
    for (int k = 0; k < 3; k++) cnt[k] = s_cnt[tid * 3 + k];
    const int mine = cnt[0] + cnt[1] + cnt[2];
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += t; }
    if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        int w = s_warp[tid];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (tid >= o) w += t; }
        s_warp[tid] = w;
    }
    __syncthreads();
    int base = incl - mine + ((tid >> 5) ? s_warp[(tid >> 5) - 1] : 0);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int c = tid * 3 + k;
        cell_start[c] = base; s_cnt[c] = base;
        base += cnt[k];
    }
    if (tid == 1023) cell_start[NC] = base;
    __syncthreads();
    // Placement in ascending keypoint index inside every cell (the reference push_backs in index order).  Usual case — no cell holds
    // more than 32 keypoints (1000-2000 keypoints over 3072 cells): every thread drops its keypoints into their cells with an
    // atomic counter, in whatever order, and the owner of a cell (three cells per thread, as in the scan above) then sorts the
    // cell's few entries by index; ascending index IS insertion order.  The one-warp walk below did this in n / 32 serial trips of
    // match_any (13 of the 18-32 us this kernel took) and remains for frames with a crowded cell.
    __shared__ int s_big;
    if (tid == 0) s_big = 0;
    __syncthreads();
    if (max(cnt[0], max(cnt[1], cnt[2])) > 32) s_big = 1;
    __syncthreads();
    if (!s_big) {
        for (int i = tid; i < n; i += blockDim.x) {
            const unsigned c = s_cell[i];
            if (c != 0xffffu) cell_items[atomicAdd(&s_cnt[c], 1)] = i;
        }
        __syncthreads();
        int start = incl - mine + ((tid >> 5) ? s_warp[(tid >> 5) - 1] : 0);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int m = cnt[k];
            if (m > 1) {                                   // insertion sort of at most 32 indices, in registers' reach of L1
                int32_t* v = cell_items + start;
                for (int a = 1; a < m; a++) {
                    const int x = v[a];
                    int bpos = a - 1;
                    while (bpos >= 0 && v[bpos] > x) { v[bpos + 1] = v[bpos]; bpos--; }
                    v[bpos + 1] = x;
                }
            }
            start += m;
        }
        return;
    }
    // crowded cell: one warp walks the keypoints 32 at a time; lanes of the same cell rank themselves with match_any
    if (tid < 32) {
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + tid;
            const unsigned c = i < n ? s_cell[i] : 0xffffu;
            const unsigned same = __match_any_sync(0xffffffffu, c);
            if (c != 0xffffu) {
                const int rank = __popc(same & ((1u << tid) - 1));
                const int pos = s_cnt[c] + rank;
                cell_items[pos] = i;
                __syncwarp(same);
                if (rank == 0) s_cnt[c] += __popc(same);
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------ K8: SearchByProjection
constexpr int HISTO_LENGTH = 30;      // src/ORBmatcher.cc:42
constexpr int TH_HIGH = 100, TH_LOW = 50;

struct SbpArgs {
    orb_frame_view cur, last;         // device pointers inside
    const uint8_t* has_mp; const uint8_t* outlier; const float* xyz;
    float T[16];
    float sf[ORB_MAX_LEVELS];         // CurrentFrame.mvScaleFactors (src/Frame.cc:95-103)
    float th;
    int cap;                          // candidate slots per query
    uint32_t* list; int* cnt;         // list[i*cap + pos] = dist<<22 | i2 ; cnt[i] = candidates found (-1: not searched)
    const int32_t* prematch;          // match_cur as the caller passed it: keypoints that already carry a map point
};

__device__ __forceinline__ int hamming256(const uint32_t* q, const uint8_t* row)
{
    const uint4* p = reinterpret_cast<const uint4*>(row);
    const uint4 a = __ldg(p), b = __ldg(p + 1);
    return __popc(q[0] ^ a.x) + __popc(q[1] ^ a.y) + __popc(q[2] ^ a.z) + __popc(q[3] ^ a.w)
         + __popc(q[4] ^ b.x) + __popc(q[5] ^ b.y) + __popc(q[6] ^ b.z) + __popc(q[7] ^ b.w);
}

// rotation-histogram bin (src/ORBmatcher.cc:1583-1588, :234-239); factor is 1/HISTO_LENGTH as in the reference
__device__ __forceinline__ int rot_bin(float a_from, float a_to)
{
    const float factor = 1.0f / HISTO_LENGTH;
    float rot = __fsub_rn(a_from, a_to);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, factor));
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

// ComputeThreeMaxima, src/ORBmatcher.cc:1748-1789
__device__ void three_maxima(const int* hs, int& ind1, int& ind2, int& ind3)
{
    int max1 = 0, max2 = 0, max3 = 0;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < HISTO_LENGTH; i++) {
        const int s = hs[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) ind3 = -1;
}

// Frame::GetFeaturesInArea (src/Frame.cc:200-265) for a whole warp, in the reference's scan order (ix outer, iy inner, insertion order).
// The grid's CSR is indexed by cell id = ix * 48 + iy, so the cells iy = y0 .. y1 of ONE column ix are one contiguous run of cell_items:
// the lanes fetch the run bounds of up to 32 columns at once (a cell-by-cell walk paid two dependent loads for each of the up to
// 11 x 11 mostly empty cells of a window: 35 us per SearchByProjection), a warp prefix concatenates the runs and the items are visited 32 at
// a time.  visit(valid, id) is called by all lanes together (it ballots); valid lanes carry consecutive items of the scan order.
template <typename F>
__device__ __forceinline__ void warp_window_walk(const orb_frame_view& fr, int x0, int x1, int y0, int y1, int lane, F visit)
{
    for (int c0 = x0; c0 <= x1; c0 += 32) {
        const int c = c0 + lane;
        int b = 0, e = 0;
        if (c <= x1) { const int base = c * ORB_GRID_ROWS; b = fr.cell_start[base + y0]; e = fr.cell_start[base + y1 + 1]; }
        const int len = e - b;
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int tv = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += tv; }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const int start = b - (incl - len);                      // item t of this trip's concatenation sits at start + t for the run that holds it
        for (int t0 = 0; t0 < total; t0 += 32) {
            const int t = t0 + lane;
            int col = 0;                                         // number of runs that end at or before t = index of the run holding t
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) { const int v = __shfl_sync(0xffffffffu, incl, col + step - 1); if (v <= t) col += step; }
            const int st = __shfl_sync(0xffffffffu, start, col & 31);
            const bool valid = t < total;
            visit(valid, valid ? fr.cell_items[st + t] : 0);
        }
    }
}

// Sorts a warp's candidate list (dist << 22 | id, in scan order) by (distance, scan position), in place: rank by counting.  With the
// list in that order "the best candidate nobody earlier has claimed" is simply the FIRST unclaimed entry, and the second best of the
// ratio tests the second one (the reference's strict '<' keeps the earlier of two equal distances, :1568, :466-473), so a round of
// the resolution pass touches one or two entries per query instead of walking the whole list (27 -> 8 us per SearchByProjection).
// Lists longer than SORT_MAX stay in scan order; the resolution pass walks those in full.
constexpr int SORT_MAX = 256;
__device__ __forceinline__ void warp_sort_candidates(uint32_t* __restrict__ list, int n, int lane)
{
    if (n < 2 || n > SORT_MAX) return;
    __syncwarp();
    uint32_t e[SORT_MAX / 32]; int rank[SORT_MAX / 32];
#pragma unroll
    for (int k = 0; k < SORT_MAX / 32; k++) { const int p = lane + 32 * k; e[k] = p < n ? list[p] : 0u; rank[k] = 0; }
    for (int q = 0; q < n; q++) {
        const uint32_t kq = ((list[q] >> 22) << 10) | (uint32_t)q;          // broadcast load (L1)
#pragma unroll
        for (int k = 0; k < SORT_MAX / 32; k++) rank[k] += kq < (((e[k] >> 22) << 10) | (uint32_t)(lane + 32 * k));
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < SORT_MAX / 32; k++) if (lane + 32 * k < n) list[rank[k]] = e[k];
}

// Pass 1, one warp per last-frame feature: project its map point (:1529-1537), gate on the image
// bounds (:1539-1542) and enumerate GetFeaturesInArea(u, v, th*scale, oct-1, oct+1) in the
// reference's scan order (ix outer, iy inner, insertion order; src/Frame.cc:233-259) together with
// the Hamming distance of every candidate.  Claims are NOT looked at here (pass 2 does that).
__global__ void __launch_bounds__(256)
k_sbp_candidates(SbpArgs A)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= A.last.n) return;
    int total = -1;
    if (A.has_mp[i] && !A.outlier[i]) {
        const float X = A.xyz[3 * i], Y = A.xyz[3 * i + 1], Z = A.xyz[3 * i + 2];
        float c[3];
#pragma unroll
        for (int r = 0; r < 3; r++) {
            // cv::gemm small-matrix branch: FP32 sum of the three products, translation added in double
            const float t0 = __fadd_rn(__fadd_rn(__fmul_rn(A.T[4 * r], X), __fmul_rn(A.T[4 * r + 1], Y)), __fmul_rn(A.T[4 * r + 2], Z));
            c[r] = __double2float_rn((double)t0 + (double)A.T[4 * r + 3]);
        }
        const float invzc = __double2float_rn(1.0 / (double)c[2]);
        const float u = __fadd_rn(__fmul_rn(__fmul_rn(A.cur.fx, c[0]), invzc), A.cur.cx);
        const float v = __fadd_rn(__fmul_rn(__fmul_rn(A.cur.fy, c[1]), invzc), A.cur.cy);
        const bool inb = !(u < (float)A.cur.min_x || u > (float)A.cur.max_x) && !(v < (float)A.cur.min_y || v > (float)A.cur.max_y);
        if (inb) {
            total = 0;
            const int oct = A.last.kps[i].octave;
            const float r = __fmul_rn(A.th, A.sf[oct]);
            const int minLevel = oct - 1, maxLevel = oct + 1;
            const float invW = __fdiv_rn((float)ORB_GRID_COLS, (float)(A.cur.max_x - A.cur.min_x));
            const float invH = __fdiv_rn((float)ORB_GRID_ROWS, (float)(A.cur.max_y - A.cur.min_y));
            const float ux = __fsub_rn(u, (float)A.cur.min_x), vy = __fsub_rn(v, (float)A.cur.min_y);
            int x0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(ux, r), invW)));
            int x1 = min(ORB_GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(ux, r), invW)));
            int y0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(vy, r), invH)));
            int y1 = min(ORB_GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(vy, r), invH)));
            if (x0 >= ORB_GRID_COLS || x1 < 0 || y0 >= ORB_GRID_ROWS || y1 < 0) { x1 = -1; x0 = 0; }
            uint32_t q[8];
            {
                const uint4* qp = reinterpret_cast<const uint4*>(A.last.desc + (size_t)i * 32);
                const uint4 a = __ldg(qp), b = __ldg(qp + 1);
                q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
            }
            const bool sameLevel = minLevel == maxLevel;   // never true here; kept for the general rule of :225-253
            const bool checkLevels = !(minLevel == -1 && maxLevel == -1);
            uint32_t* out = A.list + (size_t)i * A.cap;
            const uint32_t lt = (1u << lane) - 1;
            if (x1 >= x0 && y1 >= y0)
            warp_window_walk(A.cur, x0, x1, y0, y1, lane, [&](bool valid, int id) {
                bool ok = false;
                if (valid) {
                    const orb_keypoint kp = A.cur.kps[id];
                    ok = true;
                    if (checkLevels && !sameLevel) { if (kp.octave < minLevel || kp.octave > maxLevel) ok = false; }
                    else if (sameLevel) { if (kp.octave != minLevel) ok = false; }
                    if (fabsf(__fsub_rn(kp.x, u)) > r || fabsf(__fsub_rn(kp.y, v)) > r) ok = false;
                    // CurrentFrame.mvpMapPoints[i2] set before the call (:1562): the keypoint can never be taken, whatever the claims
                    // of pass 2 turn out to be, so it does not enter the list (the array is only written after the last round of pass 2)
                    if (ok && A.prematch[id] >= 0) ok = false;
                }
                const uint32_t m = __ballot_sync(0xffffffffu, ok);
                if (ok) {
                    const int pos = total + __popc(m & lt);
                    if (pos < A.cap) out[pos] = ((uint32_t)hamming256(q, A.cur.desc + (size_t)id * 32) << 22) | (uint32_t)id;
                }
                total += __popc(m);
            });
            if (total <= A.cap) warp_sort_candidates(out, total, lane);
        }
    }
    if (lane == 0) A.cnt[i] = total;
}

// Pass 2, one CTA.  The reference walks the last-frame features in index order and lets each take its best
// still-unclaimed candidate (first in scan order on ties, :1559-1574; accepted at <= TH_HIGH and then claimed,
// :1576-1579).  That greedy order is reproduced without a serial walk by iterating to the unique fixed point of
//     choice[i] = best candidate of i that no j < i currently chooses
// (owner[c] = min{ j : choice[j] == c } is rebuilt every round with atomicMin).  Feature 0 is final after round 1
// and feature i one round after all j < i are, so the fixed point IS the sequential result; collisions are rare
// and it takes 2-4 rounds in practice.  Then the rotation histogram filter (:1581-1617).
// result[0] = nmatches, result[1] = error flag.
__global__ void __launch_bounds__(1024)
k_sbp_resolve(SbpArgs A, int check_ori, int32_t* __restrict__ match_cur, int8_t* __restrict__ bin_of, int* owner,
              int* choice, int* __restrict__ result, int use_smem_owner)
{
    // owner[] lives in shared memory when the frame's keypoints fit (the launch passes n2 * 4 bytes): every candidate of every round looks it up
    extern __shared__ int s_owner[];
    __shared__ int hist[HISTO_LENGTH];
    __shared__ int s_changed, s_err, s_cnt, s_rounds;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int* cnt = A.cnt;
    if (use_smem_owner) {           // s_owner[n2] | choice[n1] | cnt[n1]: everything the rounds touch per query, except the lists themselves
        owner = s_owner; choice = s_owner + A.cur.n; int* sc = choice + A.last.n;
        for (int i = tid; i < A.last.n; i += nt) sc[i] = A.cnt[i];
        cnt = sc;
    }
    const int n1 = A.last.n, n2 = A.cur.n;
    if (tid < HISTO_LENGTH) hist[tid] = 0;
    if (tid == 0) { s_err = 0; s_cnt = 0; s_rounds = 0; }
    for (int i = tid; i < n1; i += nt) choice[i] = -1;
    for (int k = tid; k < n2; k += nt) bin_of[k] = -1;
    __syncthreads();
    for (int round = 0; round <= n1; round++) {
        for (int k = tid; k < n2; k += nt) owner[k] = INT_MAX;
        if (tid == 0) s_changed = 0;
        __syncthreads();
        for (int i = tid; i < n1; i += nt) { const int c = choice[i]; if (c >= 0) atomicMin(&owner[c], i); }
        __syncthreads();
        for (int i = tid; i < n1; i += nt) {
            const int n = cnt[i];
            if (n <= 0) continue;
            if (n > A.cap) { s_err = 1; continue; }
            const uint32_t* L = A.list + (size_t)i * A.cap;
            uint32_t best = 0xffffffffu;          // dist<<22 | id ; scan position is the iteration order
            const bool sorted = n <= SORT_MAX;    // list ordered by (distance, scan position): the first unclaimed entry is the answer
            for (int p = 0; p < n; p++) {
                const uint32_t e = L[p];
                const int id = (int)(e & 0x3fffff);
                if (owner[id] < i) continue;                      // claimed by an earlier feature
                if (sorted) { best = e; break; }
                if ((e >> 22) < (best >> 22)) best = e;           // strict '<': first in scan order wins
            }
            const int c = (best != 0xffffffffu && (int)(best >> 22) <= TH_HIGH) ? (int)(best & 0x3fffff) : -1;
            if (c != choice[i]) { choice[i] = c; s_changed = 1; }
        }
        __syncthreads();
        const int ch = s_changed;
        __syncthreads();
        if (!ch) break;
    }
    // commit the claims, build the rotation histogram
    int mine = 0;
    for (int i = tid; i < n1; i += nt) {
        const int c = choice[i];
        if (c < 0) continue;
        match_cur[c] = i;
        mine++;
        if (check_ori) {
            const int b = rot_bin(A.last.kps[i].angle, A.cur.kps[c].angle);
            bin_of[c] = (int8_t)b;
            atomicAdd(&hist[b], 1);
        }
    }
    atomicAdd(&s_cnt, mine);
    __syncthreads();
    if (check_ori) {
        int i1, i2, i3;
        three_maxima(hist, i1, i2, i3);
        int removed = 0;
        for (int k = tid; k < n2; k += nt) {
            const int b = bin_of[k];
            if (b >= 0 && b != i1 && b != i2 && b != i3) { match_cur[k] = -1; removed++; }
        }
        atomicSub(&s_cnt, removed);
        __syncthreads();
    }
    if (tid == 0) { result[0] = s_cnt; result[1] = s_err; }
}

// ------------------------------------------------------------------ generic windowed greedy search
// The remaining projection / window searches of ORBmatcher (src/ORBmatcher.cc:49-125, :409-516, :519-594) are the
// same loop with different window sources and acceptance rules: queries in index order, candidates from
// Frame::GetFeaturesInArea, already-claimed keypoints skipped, best (and second best) Hamming distance, accept, claim.
enum { WIN_ACCEPT_BEST = 0,          // bestDist <= th_dist                                              (:1576, :1701)
       WIN_ACCEPT_RATIO = 1,         // (float)best <= (float)second*nnratio && best <= th_dist          (:476, :585)
       WIN_ACCEPT_LEVEL_RATIO = 2 }; // best <= th_dist && !(bestLevel==secondLevel && best > nnratio*second)  (:114-117)

struct WinArgs {
    orb_frame_view tgt;                         // frame that is searched (device pointers)
    int nq;
    const uint8_t* active; const uint8_t* qdesc;
    const float* u; const float* v;             // explicit window centres (project == 0)
    const float* xyz; float T[16]; int project; // or world points projected with T and the target intrinsics
    int check_bounds;
    const float* radius; float radius_const;    // per query (radius != NULL) or one value
    const int32_t* minl; const int32_t* maxl;
    const float* qangle;                        // query keypoint angles for the rotation histogram (may be NULL)
    int accept; float nnratio; int th_dist; int histogram;
    int cap; uint32_t* list; int* cnt;
    int sort_lists;                             // k_win_resolve only: lists of at most SORT_MAX entries ordered by (distance, scan position)
    const int32_t* prematch;                    // k_win_resolve only: the match array as passed in (NULL: no pre-existing matches to skip)
};

__global__ void __launch_bounds__(256)
k_win_candidates(WinArgs A)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= A.nq) return;
    int total = -1;
    if (A.active[i]) {
        float u, v;
        if (A.project) {
            const float X = A.xyz[3 * i], Y = A.xyz[3 * i + 1], Z = A.xyz[3 * i + 2];
            float c[3];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const float t0 = __fadd_rn(__fadd_rn(__fmul_rn(A.T[4 * r], X), __fmul_rn(A.T[4 * r + 1], Y)), __fmul_rn(A.T[4 * r + 2], Z));
                c[r] = __double2float_rn((double)t0 + (double)A.T[4 * r + 3]);
            }
            const float invzc = __double2float_rn(1.0 / (double)c[2]);
            u = __fadd_rn(__fmul_rn(__fmul_rn(A.tgt.fx, c[0]), invzc), A.tgt.cx);
            v = __fadd_rn(__fmul_rn(__fmul_rn(A.tgt.fy, c[1]), invzc), A.tgt.cy);
        } else { u = A.u[i]; v = A.v[i]; }
        bool inb = true;
        if (A.check_bounds)
            inb = !(u < (float)A.tgt.min_x || u > (float)A.tgt.max_x) && !(v < (float)A.tgt.min_y || v > (float)A.tgt.max_y);
        if (inb) {
            total = 0;
            const float r = A.radius ? A.radius[i] : A.radius_const;
            const int minLevel = A.minl[i], maxLevel = A.maxl[i];
            const float invW = __fdiv_rn((float)ORB_GRID_COLS, (float)(A.tgt.max_x - A.tgt.min_x));
            const float invH = __fdiv_rn((float)ORB_GRID_ROWS, (float)(A.tgt.max_y - A.tgt.min_y));
            const float ux = __fsub_rn(u, (float)A.tgt.min_x), vy = __fsub_rn(v, (float)A.tgt.min_y);
            int x0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(ux, r), invW)));
            int x1 = min(ORB_GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(ux, r), invW)));
            int y0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(vy, r), invH)));
            int y1 = min(ORB_GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(vy, r), invH)));
            if (x0 >= ORB_GRID_COLS || x1 < 0 || y0 >= ORB_GRID_ROWS || y1 < 0) { x1 = -1; x0 = 0; }
            uint32_t q[8];
            {
                const uint4* qp = reinterpret_cast<const uint4*>(A.qdesc + (size_t)i * 32);
                const uint4 a = __ldg(qp), b = __ldg(qp + 1);
                q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
            }
            const bool checkLevels = !(minLevel == -1 && maxLevel == -1);          // src/Frame.cc:225-231
            const bool sameLevel = checkLevels && minLevel == maxLevel;
            uint32_t* out = A.list + (size_t)i * A.cap;
            const uint32_t lt = (1u << lane) - 1;
            if (x1 >= x0 && y1 >= y0)
            warp_window_walk(A.tgt, x0, x1, y0, y1, lane, [&](bool valid, int id) {
                bool ok = false;
                if (valid) {
                    const orb_keypoint kp = A.tgt.kps[id];
                    ok = true;
                    if (checkLevels && !sameLevel) { if (kp.octave < minLevel || kp.octave > maxLevel) ok = false; }
                    else if (sameLevel) { if (kp.octave != minLevel) ok = false; }
                    if (fabsf(__fsub_rn(kp.x, u)) > r || fabsf(__fsub_rn(kp.y, v)) > r) ok = false;
                    if (ok && A.prematch && A.prematch[id] >= 0) ok = false;       // keypoint already carries a map point (k_win_resolve's rule, round-invariant)
                }
                const uint32_t m = __ballot_sync(0xffffffffu, ok);
                if (ok) {
                    const int pos = total + __popc(m & lt);
                    if (pos < A.cap) out[pos] = ((uint32_t)hamming256(q, A.tgt.desc + (size_t)id * 32) << 22) | (uint32_t)id;
                }
                total += __popc(m);
            });
            if (A.sort_lists && total <= A.cap) warp_sort_candidates(out, total, lane);
        }
    }
    if (lane == 0) A.cnt[i] = total;
}

// Same parallel fixed-point resolution as k_sbp_resolve, with the acceptance rules above.
__global__ void __launch_bounds__(1024)
k_win_resolve(WinArgs A, int32_t* __restrict__ match, int8_t* __restrict__ bin_of, int* owner,
              int* choice, int* __restrict__ result, int use_smem_owner)
{
    // owner[] lives in shared memory when the frame's keypoints fit (the launch passes n2 * 4 bytes): every candidate of every round looks it up
    extern __shared__ int s_owner[];
    __shared__ int hist[HISTO_LENGTH];
    __shared__ int s_changed, s_err, s_cnt;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int* cnt = A.cnt;
    if (use_smem_owner) {           // s_owner[n2] | choice[n1] | cnt[n1]: everything the rounds touch per query, except the lists themselves
        owner = s_owner; choice = s_owner + A.tgt.n; int* sc = choice + A.nq;
        for (int i = tid; i < A.nq; i += nt) sc[i] = A.cnt[i];
        cnt = sc;
    }
    const int n1 = A.nq, n2 = A.tgt.n;
    if (tid < HISTO_LENGTH) hist[tid] = 0;
    if (tid == 0) { s_err = 0; s_cnt = 0; }
    for (int i = tid; i < n1; i += nt) choice[i] = -1;
    for (int k = tid; k < n2; k += nt) bin_of[k] = -1;
    __syncthreads();
    for (int round = 0; round <= n1; round++) {
        for (int k = tid; k < n2; k += nt) owner[k] = INT_MAX;
        if (tid == 0) s_changed = 0;
        __syncthreads();
        for (int i = tid; i < n1; i += nt) { const int c = choice[i]; if (c >= 0) atomicMin(&owner[c], i); }
        __syncthreads();
        for (int i = tid; i < n1; i += nt) {
            const int n = cnt[i];
            if (n <= 0) continue;
            if (n > A.cap) { s_err = 1; continue; }
            const uint32_t* L = A.list + (size_t)i * A.cap;
            int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx = -1, bestLevel = -1, bestLevel2 = -1;
            const bool sorted = A.sort_lists && n <= SORT_MAX;    // ordered by (distance, scan position): the first two unclaimed entries are best and second best
            for (int p = 0; p < n; p++) {
                const uint32_t e = L[p];
                const int id = (int)(e & 0x3fffff), dist = (int)(e >> 22);
                if (owner[id] < i) continue;                      // claimed by an earlier query
                if (sorted && bestIdx >= 0) {
                    bestDist2 = dist;
                    if (A.accept == WIN_ACCEPT_LEVEL_RATIO) bestLevel2 = A.tgt.kps[id].octave;
                    break;
                }
                if (dist < bestDist) {
                    bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestIdx = id;
                    if (A.accept == WIN_ACCEPT_LEVEL_RATIO) bestLevel = A.tgt.kps[id].octave;
                } else if (dist < bestDist2) {
                    bestDist2 = dist;
                    if (A.accept == WIN_ACCEPT_LEVEL_RATIO) bestLevel2 = A.tgt.kps[id].octave;
                }
            }
            bool ok = bestIdx >= 0 && bestDist <= A.th_dist;
            if (ok && A.accept == WIN_ACCEPT_RATIO) ok = (float)bestDist <= __fmul_rn((float)bestDist2, A.nnratio);
            if (ok && A.accept == WIN_ACCEPT_LEVEL_RATIO) ok = !(bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(A.nnratio, (float)bestDist2));
            const int c = ok ? bestIdx : -1;
            if (c != choice[i]) { choice[i] = c; s_changed = 1; }
        }
        __syncthreads();
        const int ch = s_changed;
        __syncthreads();
        if (!ch) break;
    }
    int mine = 0;
    for (int i = tid; i < n1; i += nt) {
        const int c = choice[i];
        if (c < 0) continue;
        match[c] = i;
        mine++;
        if (A.histogram) {
            const int b = rot_bin(A.qangle[i], A.tgt.kps[c].angle);
            bin_of[c] = (int8_t)b;
            atomicAdd(&hist[b], 1);
        }
    }
    atomicAdd(&s_cnt, mine);
    __syncthreads();
    if (A.histogram) {
        int i1, i2, i3;
        three_maxima(hist, i1, i2, i3);
        int removed = 0;
        for (int k = tid; k < n2; k += nt) {
            const int b = bin_of[k];
            if (b >= 0 && b != i1 && b != i2 && b != i3) { match[k] = -1; removed++; }
        }
        atomicSub(&s_cnt, removed);
        __syncthreads();
    }
    if (tid == 0) { result[0] = s_cnt; result[1] = s_err; }
}

// ------------------------------------------------------------------ SearchForInitialization (:598-713)
// Here a later query may STEAL a keypoint: candidate c is skipped by query i only if an earlier query took it with a
// distance <= dist(i,c) (vMatchedDistance, :637), so
//     blocked(i,c)  <=>  exists j < i with choice[j] == c and dist(j,c) <= dist(i,c)
// still only looks at lower indices and the same fixed-point iteration applies.  Accepted distances are <= TH_LOW, so per
// keypoint a 51-entry table "lowest query index that chose me with distance <= d" answers the test in one load; it is
// rebuilt every round for the keypoints that have choosers.  The last taker (highest index) keeps the match; the rotation
// histogram counts every acceptance, stolen ones included, exactly like the reference's rotHist (:666-676,:688-703).
constexpr int INIT_TAB = TH_LOW + 2;       // 52 ints per F2 keypoint (entries 0..50 used)

__global__ void __launch_bounds__(1024)
k_init_resolve(WinArgs A, int32_t* __restrict__ matches12, float* __restrict__ prev_matched, int* __restrict__ tab /* n2 x INIT_TAB */,
               int* __restrict__ has /* n2 */, int* __restrict__ choice, int* __restrict__ cdist, int* __restrict__ result)
{
    __shared__ int hist[HISTO_LENGTH];
    __shared__ int s_changed, s_err, s_cnt;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int n1 = A.nq, n2 = A.tgt.n;
    if (tid < HISTO_LENGTH) hist[tid] = 0;
    if (tid == 0) { s_err = 0; s_cnt = 0; }
    for (int i = tid; i < n1; i += nt) { choice[i] = -1; cdist[i] = INT_MAX; }
    for (int k = tid; k < n2; k += nt) has[k] = 0;
    for (size_t k = tid; k < (size_t)n2 * INIT_TAB; k += nt) tab[k] = INT_MAX;
    __syncthreads();
    for (int round = 0; round <= n1; round++) {
        if (tid == 0) s_changed = 0;
        // (1) wipe the tables of last round's chosen keypoints (a warp per keypoint)
        for (int k = warp; k < n2; k += nwarps) {
            if (!has[k]) continue;
            for (int d = lane; d < INIT_TAB; d += 32) tab[(size_t)k * INIT_TAB + d] = INT_MAX;
            if (lane == 0) has[k] = 0;
        }
        __syncthreads();
        // (2) lowest chooser index per (keypoint, exact distance)
        for (int i = tid; i < n1; i += nt) {
            const int c = choice[i];
            if (c < 0) continue;
            atomicMin(&tab[(size_t)c * INIT_TAB + cdist[i]], i);
            has[c] = 1;
        }
        __syncthreads();
        // (3) running minimum over the distance axis: tab[c][d] = lowest index that chose c with distance <= d
        for (int k = warp; k < n2; k += nwarps) {
            if (!has[k]) continue;
            int* T = tab + (size_t)k * INIT_TAB;
            int v0 = T[lane], v1 = lane + 32 < INIT_TAB ? T[lane + 32] : INT_MAX;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t0 = __shfl_up_sync(0xffffffffu, v0, o), t1 = __shfl_up_sync(0xffffffffu, v1, o);
                if (lane >= o) { v0 = min(v0, t0); v1 = min(v1, t1); }
            }
            v1 = min(v1, __shfl_sync(0xffffffffu, v0, 31));
            T[lane] = v0;
            if (lane + 32 < INIT_TAB) T[lane + 32] = v1;
        }
        __syncthreads();
        // (4) every query rescans its candidate list against the tables
        for (int i = tid; i < n1; i += nt) {
            const int n = A.cnt[i];
            if (n <= 0) continue;
            if (n > A.cap) { s_err = 1; continue; }
            const uint32_t* L = A.list + (size_t)i * A.cap;
            int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx = -1;
            for (int p = 0; p < n; p++) {
                const uint32_t e = L[p];
                const int id = (int)(e & 0x3fffff), dist = (int)(e >> 22);
                if (has[id] && tab[(size_t)id * INIT_TAB + min(dist, TH_LOW)] < i) continue;      // vMatchedDistance[i2] <= dist (:637)
                if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx = id; }
                else if (dist < bestDist2) bestDist2 = dist;
            }
            const bool ok = bestIdx >= 0 && bestDist <= TH_LOW && (float)bestDist < __fmul_rn((float)bestDist2, A.nnratio);   // :652-654
            const int c = ok ? bestIdx : -1;
            if (c != choice[i]) { choice[i] = c; s_changed = 1; }
            cdist[i] = ok ? bestDist : INT_MAX;
        }
        __syncthreads();
        const int ch = s_changed;
        __syncthreads();
        if (!ch) break;
    }
    // converged: the highest index among a keypoint's takers keeps it (:656-662); has[] is reused as vnMatches21
    for (int k = tid; k < n2; k += nt) has[k] = -1;
    for (int i = tid; i < n1; i += nt) matches12[i] = -1;
    __syncthreads();
    for (int i = tid; i < n1; i += nt) {
        const int c = choice[i];
        if (c < 0) continue;
        atomicMax(&has[c], i);
        if (A.histogram) atomicAdd(&hist[rot_bin(A.qangle[i], A.tgt.kps[c].angle)], 1);
    }
    __syncthreads();
    int mine = 0;
    for (int k = tid; k < n2; k += nt)
        if (has[k] >= 0) { matches12[has[k]] = k; mine++; }
    atomicAdd(&s_cnt, mine);
    __syncthreads();
    if (A.histogram) {
        int i1, i2, i3;
        three_maxima(hist, i1, i2, i3);
        int removed = 0;
        for (int i = tid; i < n1; i += nt) {
            const int c = choice[i];
            if (c < 0) continue;
            const int b = rot_bin(A.qangle[i], A.tgt.kps[c].angle);
            if (b != i1 && b != i2 && b != i3 && matches12[i] >= 0) { matches12[i] = -1; removed++; }
        }
        atomicSub(&s_cnt, removed);
        __syncthreads();
    }
    for (int i = tid; i < n1; i += nt) {                       // update prev matched (:707-710)
        const int c = matches12[i];
        if (c >= 0) { prev_matched[2 * i] = A.tgt.kps[c].x; prev_matched[2 * i + 1] = A.tgt.kps[c].y; }
    }
    if (tid == 0) { result[0] = s_cnt; result[1] = s_err; }
}

// ------------------------------------------------------------------ K9: SearchByBoW scoring
struct BowArgs {
    orb_featvec_view kf, f;
    const uint8_t* kf_desc; const orb_keypoint* kf_kps; const uint8_t* kf_valid;
    const uint8_t* f_desc; const orb_keypoint* f_kps; int n_f;
    float nnratio; int check_ori;
    int32_t* match_f; int8_t* bin_of; int* hist; int* result;   // hist[30], result[0]=nmatches, result[1]=overlap flag
    int* seen;                                                   // n_f counters for the disjointness check
    const uint8_t* f_valid;                                      // KF-KF form (:769): second frame's feature needs a live map point (NULL: all)
    int strict_low;                                              // KF-KF form (:791): bestDist1 < TH_LOW instead of <=
    int32_t* match12; int n_kf;                                  // KF-KF form: output indexed by the first keyframe's features
};

// A frame feature that appears under two vocabulary nodes would couple the nodes through the claim
// array; DBoW2 never produces that, but if the caller's CSR does, fall back to one warp in node order.
__global__ void k_bow_check(BowArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = A.f.start[A.f.nnodes];
    if (j < total) { if (atomicAdd(&A.seen[A.f.items[j]], 1) > 0) A.result[1] = 1; }
}

__global__ void __launch_bounds__(256)
k_bow_match(BowArgs A)
{
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const bool serial = A.result[1] != 0;
    if (serial && gw != 0) return;
    int local_matches = 0;
    for (int a = serial ? 0 : gw; a < A.kf.nnodes; a += serial ? 1 : nw) {
        // merge-join of the two ascending node lists (:176-260) == binary search of this KF node in F
        const int node = A.kf.node_id[a];
        int lo = 0, hi = A.f.nnodes;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (A.f.node_id[mid] < node) lo = mid + 1; else hi = mid; }
        if (lo >= A.f.nnodes || A.f.node_id[lo] != node) continue;
        const int fb = A.f.start[lo], fe = A.f.start[lo + 1];
        for (int ik = A.kf.start[a]; ik < A.kf.start[a + 1]; ik++) {
            const int realIdxKF = A.kf.items[ik];
            if (!A.kf_valid[realIdxKF]) continue;
            uint32_t q[8];
            {
                const uint4* qp = reinterpret_cast<const uint4*>(A.kf_desc + (size_t)realIdxKF * 32);
                const uint4 x = __ldg(qp), y = __ldg(qp + 1);
                q[0] = x.x; q[1] = x.y; q[2] = x.z; q[3] = x.w; q[4] = y.x; q[5] = y.y; q[6] = y.z; q[7] = y.w;
            }
            unsigned long long best = ~0ull;      // (dist, scan position, id)
            int d1 = INT_MAX, d2 = INT_MAX;
            for (int jf = fb + lane; jf < fe; jf += 32) {
                const int realIdxF = A.f.items[jf];
                if (((volatile int32_t*)A.match_f)[realIdxF] >= 0) continue;      // :205 / vbMatched2 :769
                if (A.f_valid && !A.f_valid[realIdxF]) continue;
                const int dist = hamming256(q, A.f_desc + (size_t)realIdxF * 32);
                const unsigned long long key = ((unsigned long long)dist << 44) | ((unsigned long long)(jf - fb) << 22) | (unsigned long long)realIdxF;
                best = key < best ? key : best;
                if (dist < d1) { d2 = d1; d1 = dist; } else if (dist < d2) d2 = dist;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
                best = t < best ? t : best;
                const int e1 = __shfl_xor_sync(0xffffffffu, d1, o), e2 = __shfl_xor_sync(0xffffffffu, d2, o);
                const int n1 = min(d1, e1), n2 = min(max(d1, e1), min(d2, e2));   // two smallest of the multiset union
                d1 = n1; d2 = n2;
            }
            if (best == ~0ull) continue;
            const int bestIdxF = (int)(best & 0x3fffff);
            if ((A.strict_low ? d1 < TH_LOW : d1 <= TH_LOW) && (float)d1 < __fmul_rn(A.nnratio, (float)d2)) {    // :224-226 / :791-793
                if (lane == 0) {
                    A.match_f[bestIdxF] = realIdxKF;
                    if (A.check_ori) {
                        const int b = rot_bin(A.kf_kps[realIdxKF].angle, A.f_kps[bestIdxF].angle);
                        A.bin_of[bestIdxF] = (int8_t)b;
                        atomicAdd(&A.hist[b], 1);
                    }
                }
                local_matches++;
                __threadfence_block();
                __syncwarp();
            }
        }
    }
    if (lane == 0 && local_matches) atomicAdd(&A.result[0], local_matches);
}

__global__ void __launch_bounds__(256)
k_bow_orientation(BowArgs A)
{
    __shared__ int s_removed;
    if (threadIdx.x == 0) s_removed = 0;
    __syncthreads();
    int i1, i2, i3;
    three_maxima(A.hist, i1, i2, i3);
    int removed = 0;
    for (int k = threadIdx.x; k < A.n_f; k += blockDim.x) {
        const int b = A.bin_of[k];
        if (b >= 0 && b != i1 && b != i2 && b != i3) { A.match_f[k] = -1; removed++; }
    }
    if (removed) atomicAdd(&s_removed, removed);
    __syncthreads();
    if (threadIdx.x == 0) A.result[0] -= s_removed;
    if (A.match12) {          // KF-KF form: vpMatches12[idx1] = feature of the second keyframe
        for (int k = threadIdx.x; k < A.n_kf; k += blockDim.x) A.match12[k] = -1;
        __syncthreads();
        for (int k = threadIdx.x; k < A.n_f; k += blockDim.x) { const int first = A.match_f[k]; if (first >= 0) A.match12[first] = k; }
    }
}

} // namespace

int orb_launch_knn2(orb_ctx* c, const uint8_t* d_q, int nq, const uint8_t* d_db, int64_t ndb, int npairs, int32_t idx_base,
                    int32_t* d_idx1, int32_t* d_d1, int32_t* d_d2, cudaStream_t s)
{
    const int qtiles = (nq + KNN_THREADS - 1) / KNN_THREADS;
    // enough CTAs for a few waves over 148 SMs, chunks a multiple of the tile
    long long want_chunks = std::max<long long>(1, (148LL * 8 + (long long)qtiles * npairs - 1) / ((long long)qtiles * npairs));
    long long rows = (ndb + want_chunks - 1) / want_chunks;
    rows = std::max<long long>(rows, 4 * KNN_TILE);
    rows = ((rows + KNN_TILE - 1) / KNN_TILE) * KNN_TILE;
    if (rows > (1 << KNN_KEY_SHIFT)) rows = 1 << KNN_KEY_SHIFT;       // the packed (distance, row) key of k_knn2
    const int nchunks = (int)((ndb + rows - 1) / rows);
    if (nchunks == 0) {          // empty DB: idx1 = -1, d1 = d2 = INT_MAX for every query
        k_knn2_merge<<<dim3((nq + 127) / 128, npairs), 128, 0, s>>>(nullptr, 0, nq, npairs, d_idx1, d_d1, d_d2);
        c->last_launches = 1;
        ORB_CUDA(cudaGetLastError());
        return ORB_OK;
    }
    if (c->knn_engine == ORB_KNN_TENSOR) return orb_launch_knn2_tc(c, d_q, nq, d_db, ndb, npairs, idx_base, d_idx1, d_d1, d_d2, s);
    if (nchunks > 65535 || npairs > 65535 || qtiles > 65535) return ORB_ERR_CAPACITY;     // grid limits, before anything is allocated
    // global row indices are int32: a shard whose last row does not fit is refused instead of wrapping to negative indices
    // (which the merge would then drop as "empty")
    if (idx_base < 0) return ORB_ERR_INVALID;
    if (ndb + (int64_t)idx_base > (int64_t)INT_MAX) return ORB_ERR_CAPACITY;
    Knn2Args A;
    A.q = d_q; A.db = d_db; A.nq = nq; A.ndb = ndb; A.rows_per_chunk = (int)rows; A.nchunks = nchunks;
    A.idx_base = idx_base; A.out = nullptr; A.o_idx1 = d_idx1; A.o_d1 = d_d1; A.o_d2 = d_d2;
    int32_t* part = nullptr;
    if (nchunks > 1) {
        // per-call partials from the context's stream-ordered pool: calls on different streams (or from different threads) never
        // share scratch, and nothing synchronises
        const size_t need = (size_t)npairs * nchunks * 3 * nq * sizeof(int32_t);
        ORB_CUDA(cudaMallocFromPoolAsync((void**)&part, need, c->pool, s));
        A.out = part;
    }
    k_knn2<<<dim3(nchunks, qtiles, npairs), KNN_THREADS, 0, s>>>(A);
    c->last_launches = 1;
    if (nchunks > 1) {
        k_knn2_merge<<<dim3((nq + 127) / 128, npairs), 128, 0, s>>>(part, nchunks, nq, npairs, d_idx1, d_d1, d_d2);
        c->last_launches = 2;
        ORB_CUDA(cudaFreeAsync(part, s));
    }
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

int orb_launch_knn2_merge(const int32_t* d_parts, int nparts, int nq, int32_t* d_idx1, int32_t* d_d1, int32_t* d_d2, cudaStream_t s)
{
    k_knn2_merge<<<dim3((nq + 127) / 128, 1), 128, 0, s>>>(d_parts, nparts, nq, 1, d_idx1, d_d1, d_d2);
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

int orb_launch_knn2_merge_pairs(const int32_t* d_parts, int nparts, int nq, int npairs, int32_t* d_idx1, int32_t* d_d1, int32_t* d_d2, cudaStream_t s)
{
    k_knn2_merge<<<dim3((nq + 127) / 128, npairs), 128, 0, s>>>(d_parts, nparts, nq, npairs, d_idx1, d_d1, d_d2);
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

int orb_launch_knn2_merge_ptrs(const int32_t* const* d_parts, int nparts, int nq, int32_t* d_idx1, int32_t* d_d1, int32_t* d_d2, cudaStream_t s)
{
    if (nparts < 1 || nparts > ORB_COMM_MAX_RANKS) return ORB_ERR_CAPACITY;
    MergePtrs P;
    for (int i = 0; i < ORB_COMM_MAX_RANKS; i++) P.p[i] = i < nparts ? d_parts[i] : nullptr;
    k_knn2_merge_ptrs<<<(nq + 127) / 128, 128, 0, s>>>(P, nparts, nq, d_idx1, d_d1, d_d2);
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

int orb_launch_match_ratio(const int32_t* idx1, const int32_t* d1, const int32_t* d2, int nq, float nnratio, int th,
                           int32_t* match, int* d_count, cudaStream_t s)
{
    ORB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), s));
    k_match_ratio<<<(nq + 255) / 256, 256, 0, s>>>(idx1, d1, d2, nq, nnratio, th, match, d_count);
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

int orb_launch_grid_build(const orb_keypoint* kps, int n, int min_x, int max_x, int min_y, int max_y,
                          int32_t* cell_start, int32_t* cell_items, cudaStream_t s)
{
    const size_t sm = (size_t)std::max(n, 1) * sizeof(unsigned short);
    if (sm > 160 * 1024) return ORB_ERR_CAPACITY;
    if (sm > 48 * 1024) ORB_CUDA(cudaFuncSetAttribute(k_grid_build, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));   // per device, no caching
    k_grid_build<<<1, 1024, sm, s>>>(kps, n, min_x, max_x, min_y, max_y, cell_start, cell_items);
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

int orb_launch_popc_bench(double* gpopc, cudaStream_t s)
{
    uint32_t* d = nullptr;
    ORB_CUDA(cudaMalloc((void**)&d, 4));
    cudaEvent_t e0, e1;
    ORB_CUDA(cudaEventCreate(&e0)); ORB_CUDA(cudaEventCreate(&e1));
    const int iters = 4096, blocks = 148 * 8, threads = 256;
    k_popc_bench<<<blocks, threads, 0, s>>>(d, 64);
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        ORB_CUDA(cudaEventRecord(e0, s));
        k_popc_bench<<<blocks, threads, 0, s>>>(d, iters);
        ORB_CUDA(cudaEventRecord(e1, s));
        ORB_CUDA(cudaEventSynchronize(e1));
        float ms = 0; ORB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    *gpopc = (double)blocks * threads * iters * 32.0 / (best * 1e-3) / 1e9;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    return ORB_OK;
}

int orb_launch_search_by_projection(orb_ctx* c, const orb_frame_view* cur, const orb_frame_view* last,
                                    const uint8_t* last_has_mp, const uint8_t* last_outlier, const float* last_xyz,
                                    const float* T16_host, float th, int check_ori, int32_t* match_cur, int* d_result,
                                    uint8_t* scratch, size_t scratch_bytes, cudaStream_t s)
{
    (void)c;
    SbpArgs A;
    A.cur = *cur; A.last = *last; A.has_mp = last_has_mp; A.outlier = last_outlier; A.xyz = last_xyz;
    for (int i = 0; i < 16; i++) A.T[i] = T16_host[i];
    if (cur->nlevels < 1 || cur->nlevels > ORB_MAX_LEVELS) return ORB_ERR_INVALID;
    A.sf[0] = 1.0f;
    for (int i = 1; i < ORB_MAX_LEVELS; i++) A.sf[i] = i < cur->nlevels ? A.sf[i - 1] * cur->scale_factor : A.sf[i - 1];
    A.th = th; A.prematch = match_cur;
    if (cur->n >= (1 << 22)) return ORB_ERR_CAPACITY;
    // scratch: cnt[last.n] | choice[last.n] | owner[cur.n] | bin_of[cur.n] | list[last.n * cap]
    size_t off = 0;
    A.cnt = (int*)(scratch + off); off += ((size_t)last->n * 4 + 255) & ~(size_t)255;
    int* choice = (int*)(scratch + off); off += ((size_t)last->n * 4 + 255) & ~(size_t)255;
    int* owner = (int*)(scratch + off); off += ((size_t)cur->n * 4 + 255) & ~(size_t)255;
    int8_t* bin_of = (int8_t*)(scratch + off); off += ((size_t)cur->n + 255) & ~(size_t)255;
    const size_t avail = scratch_bytes > off ? (scratch_bytes - off) / 4 : 0;
    A.cap = (int)std::min<size_t>((size_t)cur->n, last->n ? avail / (size_t)last->n : 0);
    A.cap = std::min(A.cap, 1 << 20);
    A.list = (uint32_t*)(scratch + off);
    if (A.cap < 1) return ORB_ERR_CAPACITY;
    k_sbp_candidates<<<(last->n * 32 + 255) / 256, 256, 0, s>>>(A);
    const size_t rs_bytes = ((size_t)cur->n + 2 * (size_t)last->n) * 4;
    const int smem_owner = rs_bytes <= 40 * 1024;
    k_sbp_resolve<<<1, 1024, smem_owner ? rs_bytes : 0, s>>>(A, check_ori, match_cur, bin_of, owner, choice, d_result, smem_owner);
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

size_t orb_sbp_scratch_bytes(int n_cur, int n_last)
{
    const size_t cap = (size_t)std::min(n_cur, 1024);
    return 2 * (((size_t)n_last * 4 + 255) & ~(size_t)255) + (((size_t)n_cur * 4 + 255) & ~(size_t)255) +
           (((size_t)n_cur + 255) & ~(size_t)255) + (size_t)n_last * cap * 4 + 256;
}

int orb_launch_search_by_bow(orb_ctx* c, const orb_featvec_view* kf_fv, const uint8_t* kf_desc, const orb_keypoint* kf_kps,
                             const uint8_t* kf_mp_valid, const orb_featvec_view* f_fv, const uint8_t* f_desc,
                             const orb_keypoint* f_kps, int n_f, int f_items_total, float nnratio, int check_ori, int32_t* match_f,
                             uint8_t* scratch, cudaStream_t s, const uint8_t* f_valid, int32_t* match12, int n_kf)
{
    (void)c;
    if (n_f >= (1 << 22)) return ORB_ERR_CAPACITY;
    BowArgs A;
    A.kf = *kf_fv; A.f = *f_fv; A.kf_desc = kf_desc; A.kf_kps = kf_kps; A.kf_valid = kf_mp_valid;
    A.f_desc = f_desc; A.f_kps = f_kps; A.n_f = n_f; A.nnratio = nnratio; A.check_ori = check_ori; A.match_f = match_f;
    A.f_valid = f_valid; A.strict_low = match12 != nullptr; A.match12 = match12; A.n_kf = n_kf;
    // scratch: result[2] + hist[30] (256 B) | seen[n_f] | bin_of[n_f]
    A.result = (int*)scratch; A.hist = (int*)scratch + 2;
    A.seen = (int*)(scratch + 256);
    A.bin_of = (int8_t*)(scratch + 256 + (((size_t)n_f * 4 + 255) & ~(size_t)255));
    ORB_CUDA(cudaMemsetAsync(scratch, 0, 256 + (((size_t)n_f * 4 + 255) & ~(size_t)255), s));
    ORB_CUDA(cudaMemsetAsync(A.bin_of, 0xff, (size_t)std::max(n_f, 1), s));
    ORB_CUDA(cudaMemsetAsync(match_f, 0xff, (size_t)std::max(n_f, 1) * 4, s));     // :159
    if (f_items_total > 0) k_bow_check<<<(f_items_total + 255) / 256, 256, 0, s>>>(A);
    if (kf_fv->nnodes > 0) k_bow_match<<<std::max(1, std::min(148, (kf_fv->nnodes + 7) / 8)), 256, 0, s>>>(A);
    if (check_ori || match12) k_bow_orientation<<<1, 256, 0, s>>>(A);
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

// ------------------------------------------------------------------ ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:852-1014)
// Same vocabulary-node join as SearchByBoW, between the features of two keyframes that have no map point yet.  Per feature of
// KF1 the reference sorts the unclaimed candidates with dist <= TH_LOW by (dist, index), and takes the first one within
// 2 * BestDist that passes the epipolar test (:136-153); that is the minimum of (dist, index) over the passing candidates, so a warp
// needs two sweeps (smallest distance, then smallest passing key) and no sort.  Claims couple only features of one node.
struct TriArgs {
    orb_featvec_view fv1, fv2;
    const uint8_t* desc1; const orb_keypoint* kps1; const uint8_t* has_mp1; int n1;
    const uint8_t* desc2; const orb_keypoint* kps2; const uint8_t* has_mp2; int n2;
    float F[9]; float sigma2[ORB_MAX_LEVELS];
    int check_ori;
    int32_t* match12; int32_t* matched2; int8_t* bin_of; int* hist; int* result; int* seen;
};

__device__ __forceinline__ bool epipolar_ok(const TriArgs& A, const orb_keypoint& k1, const orb_keypoint& k2)
{
    const float a = __fadd_rn(__fadd_rn(__fmul_rn(k1.x, A.F[0]), __fmul_rn(k1.y, A.F[3])), A.F[6]);
    const float b = __fadd_rn(__fadd_rn(__fmul_rn(k1.x, A.F[1]), __fmul_rn(k1.y, A.F[4])), A.F[7]);
    const float c = __fadd_rn(__fadd_rn(__fmul_rn(k1.x, A.F[2]), __fmul_rn(k1.y, A.F[5])), A.F[8]);
    const float num = __fadd_rn(__fadd_rn(__fmul_rn(a, k2.x), __fmul_rn(b, k2.y)), c);
    const float den = __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
    if (den == 0.f) return false;
    const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
    const int oct = min(max(k2.octave, 0), ORB_MAX_LEVELS - 1);
    return (double)dsqr < __dmul_rn(3.84, (double)A.sigma2[oct]);
}

__global__ void k_tri_check(TriArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = A.fv2.start[A.fv2.nnodes];
    if (j < total) { if (atomicAdd(&A.seen[A.fv2.items[j]], 1) > 0) A.result[1] = 1; }
}

__global__ void __launch_bounds__(256)
k_tri_match(TriArgs A)
{
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const bool serial = A.result[1] != 0;
    if (serial && gw != 0) return;
    int local_matches = 0;
    for (int a = serial ? 0 : gw; a < A.fv1.nnodes; a += serial ? 1 : nw) {
        const int node = A.fv1.node_id[a];
        int lo = 0, hi = A.fv2.nnodes;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (A.fv2.node_id[mid] < node) lo = mid + 1; else hi = mid; }
        if (lo >= A.fv2.nnodes || A.fv2.node_id[lo] != node) continue;
        const int fb = A.fv2.start[lo], fe = A.fv2.start[lo + 1];
        for (int ik = A.fv1.start[a]; ik < A.fv1.start[a + 1]; ik++) {
            const int idx1 = A.fv1.items[ik];
            if (A.has_mp1[idx1]) continue;                                        // :891-893
            uint32_t q[8];
            {
                const uint4* qp = reinterpret_cast<const uint4*>(A.desc1 + (size_t)idx1 * 32);
                const uint4 x = __ldg(qp), y = __ldg(qp + 1);
                q[0] = x.x; q[1] = x.y; q[2] = x.z; q[3] = x.w; q[4] = y.x; q[5] = y.y; q[6] = y.z; q[7] = y.w;
            }
            int best = INT_MAX;                                                   // BestDist over the admissible candidates (:901-925)
            for (int jf = fb + lane; jf < fe; jf += 32) {
                const int idx2 = A.fv2.items[jf];
                if (((volatile int32_t*)A.matched2)[idx2] >= 0 || A.has_mp2[idx2]) continue;
                best = min(best, hamming256(q, A.desc2 + (size_t)idx2 * 32));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
            if (best > TH_LOW) continue;
            const int dist_th = 2 * best;                                         // round(2*BestDist), :927
            const orb_keypoint k1 = A.kps1[idx1];
            uint32_t key = 0xffffffffu;                                           // (dist, idx2): first passing entry of the sorted list
            for (int jf = fb + lane; jf < fe; jf += 32) {
                const int idx2 = A.fv2.items[jf];
                if (((volatile int32_t*)A.matched2)[idx2] >= 0 || A.has_mp2[idx2]) continue;
                const int dist = hamming256(q, A.desc2 + (size_t)idx2 * 32);
                if (dist > TH_LOW || dist > dist_th) continue;
                if (!epipolar_ok(A, k1, A.kps2[idx2])) continue;
                key = min(key, ((uint32_t)dist << 22) | (uint32_t)idx2);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
            if (key == 0xffffffffu) continue;
            const int idx2 = (int)(key & 0x3fffff);
            if (lane == 0) {
                A.matched2[idx2] = idx1;
                A.match12[idx1] = idx2;
                if (A.check_ori) {
                    const int b = rot_bin(k1.angle, A.kps2[idx2].angle);
                    A.bin_of[idx1] = (int8_t)b;
                    atomicAdd(&A.hist[b], 1);
                }
            }
            local_matches++;
            __threadfence_block();
            __syncwarp();
        }
    }
    if (lane == 0 && local_matches) atomicAdd(&A.result[0], local_matches);
}

__global__ void __launch_bounds__(256)
k_tri_orientation(TriArgs A)
{
    __shared__ int s_removed;
    if (threadIdx.x == 0) s_removed = 0;
    __syncthreads();
    int i1, i2, i3;
    three_maxima(A.hist, i1, i2, i3);
    int removed = 0;
    for (int k = threadIdx.x; k < A.n1; k += blockDim.x) {
        const int b = A.bin_of[k];
        if (b >= 0 && b != i1 && b != i2 && b != i3) { A.match12[k] = -1; removed++; }
    }
    if (removed) atomicAdd(&s_removed, removed);
    __syncthreads();
    if (threadIdx.x == 0) A.result[0] -= s_removed;
}

size_t orb_tri_scratch_bytes(int n1, int n2)
{
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    return 256 + al((size_t)std::max(n2, 1) * 4) * 2 + al((size_t)std::max(n1, 1)) + 256;
}

// scratch: result[2] + hist[30] (256 B) | seen[n2] | matched2[n2] | bin_of[n1]
int orb_launch_search_for_triangulation(const orb_featvec_view* fv1, const uint8_t* desc1, const orb_keypoint* kps1, const uint8_t* has_mp1,
                                        int n1, const orb_featvec_view* fv2, const uint8_t* desc2, const orb_keypoint* kps2,
                                        const uint8_t* has_mp2, int n2, int items2_total, const float* F12, const float* sigma2, int nlevels,
                                        int check_ori, int32_t* match12, uint8_t* scratch, cudaStream_t s)
{
    if (n2 >= (1 << 22)) return ORB_ERR_CAPACITY;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    TriArgs A;
    A.fv1 = *fv1; A.fv2 = *fv2; A.desc1 = desc1; A.kps1 = kps1; A.has_mp1 = has_mp1; A.n1 = n1;
    A.desc2 = desc2; A.kps2 = kps2; A.has_mp2 = has_mp2; A.n2 = n2; A.check_ori = check_ori; A.match12 = match12;
    for (int i = 0; i < 9; i++) A.F[i] = F12[i];
    for (int i = 0; i < ORB_MAX_LEVELS; i++) A.sigma2[i] = nlevels > 0 ? sigma2[std::min(i, nlevels - 1)] : 1.f;
    const size_t b2 = al((size_t)std::max(n2, 1) * 4);
    A.result = (int*)scratch; A.hist = (int*)scratch + 2;
    A.seen = (int*)(scratch + 256);
    A.matched2 = (int32_t*)(scratch + 256 + b2);
    A.bin_of = (int8_t*)(scratch + 256 + 2 * b2);
    ORB_CUDA(cudaMemsetAsync(scratch, 0, 256 + b2, s));
    ORB_CUDA(cudaMemsetAsync(A.matched2, 0xff, b2, s));
    ORB_CUDA(cudaMemsetAsync(A.bin_of, 0xff, (size_t)std::max(n1, 1), s));
    ORB_CUDA(cudaMemsetAsync(match12, 0xff, (size_t)std::max(n1, 1) * 4, s));
    if (items2_total > 0) k_tri_check<<<(items2_total + 255) / 256, 256, 0, s>>>(A);
    if (fv1->nnodes > 0 && n2 > 0) k_tri_match<<<std::max(1, std::min(148, (fv1->nnodes + 7) / 8)), 256, 0, s>>>(A);
    if (check_ori) k_tri_orientation<<<1, 256, 0, s>>>(A);
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

size_t orb_bow_scratch_bytes(int n_f)
{
    return 256 + (((size_t)n_f * 4 + 255) & ~(size_t)255) + (((size_t)n_f + 255) & ~(size_t)255) + 256;
}

int orb_launch_search_window(orb_ctx* c, const orb_frame_view* tgt, const orb_window_query_set* q, int accept, float nnratio, int th_dist,
                             int histogram, int32_t* match, int* d_result, uint8_t* scratch, size_t scratch_bytes, cudaStream_t s)
{
    (void)c;
    WinArgs A;
    A.tgt = *tgt; A.nq = q->n; A.active = q->active; A.qdesc = q->desc; A.u = q->u; A.v = q->v; A.xyz = q->xyz;
    A.project = q->xyz != nullptr && q->u == nullptr;
    for (int i = 0; i < 16; i++) A.T[i] = (A.project && q->Tcw16) ? q->Tcw16[i] : 0.f;
    A.check_bounds = q->check_bounds; A.radius = q->radius; A.radius_const = q->radius_const;
    A.minl = q->min_level; A.maxl = q->max_level; A.qangle = q->angle;
    A.accept = accept; A.nnratio = nnratio; A.th_dist = th_dist; A.histogram = histogram && q->angle;
    A.prematch = match; A.sort_lists = 1;
    if (tgt->n >= (1 << 22)) return ORB_ERR_CAPACITY;
    size_t off = 0;
    A.cnt = (int*)(scratch + off); off += ((size_t)q->n * 4 + 255) & ~(size_t)255;
    int* choice = (int*)(scratch + off); off += ((size_t)q->n * 4 + 255) & ~(size_t)255;
    int* owner = (int*)(scratch + off); off += ((size_t)tgt->n * 4 + 255) & ~(size_t)255;
    int8_t* bin_of = (int8_t*)(scratch + off); off += ((size_t)tgt->n + 255) & ~(size_t)255;
    const size_t avail = scratch_bytes > off ? (scratch_bytes - off) / 4 : 0;
    A.cap = (int)std::min<size_t>((size_t)tgt->n, q->n ? avail / (size_t)q->n : 0);
    A.cap = std::min(A.cap, 1 << 20);
    A.list = (uint32_t*)(scratch + off);
    if (A.cap < 1) return ORB_ERR_CAPACITY;
    k_win_candidates<<<(q->n * 32 + 255) / 256, 256, 0, s>>>(A);
    const size_t rs_bytes = ((size_t)tgt->n + 2 * (size_t)q->n) * 4;
    const int smem_owner = rs_bytes <= 40 * 1024;
    k_win_resolve<<<1, 1024, smem_owner ? rs_bytes : 0, s>>>(A, match, bin_of, owner, choice, d_result, smem_owner);
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

// Best candidate per query WITHOUT claims: the scoring loops of ORBmatcher::Fuse (src/ORBmatcher.cc:1016-1134, :1136-1265) and of
// both directions of SearchBySim3 (:1267-1505) only pick "the most similar keypoint in the radius"; what is done with it afterwards
// (replace / add observations, agreement test) does not feed back into the search.  Strict '<' in scan order like the reference.
__global__ void __launch_bounds__(256)
k_win_best(WinArgs A, int32_t* __restrict__ best_idx, int32_t* __restrict__ best_dist, int* __restrict__ result)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.nq) return;
    const int n = A.cnt[i];
    int bd = INT_MAX, bi = -1;
    if (n > A.cap) atomicExch(result + 1, 1);
    else {
        const uint32_t* L = A.list + (size_t)i * A.cap;
        for (int p = 0; p < n; p++) {
            const uint32_t e = L[p];
            const int dist = (int)(e >> 22);
            if (dist < bd) { bd = dist; bi = (int)(e & 0x3fffff); }
        }
    }
    best_idx[i] = bi; best_dist[i] = bd;
}

int orb_launch_search_window_best(orb_ctx* c, const orb_frame_view* tgt, const orb_window_query_set* q, int32_t* best_idx, int32_t* best_dist,
                                  int* d_result, uint8_t* scratch, size_t scratch_bytes, cudaStream_t s)
{
    (void)c;
    WinArgs A;
    A.tgt = *tgt; A.nq = q->n; A.active = q->active; A.qdesc = q->desc; A.u = q->u; A.v = q->v; A.xyz = q->xyz;
    A.project = q->xyz != nullptr && q->u == nullptr;
    for (int i = 0; i < 16; i++) A.T[i] = (A.project && q->Tcw16) ? q->Tcw16[i] : 0.f;
    A.check_bounds = q->check_bounds; A.radius = q->radius; A.radius_const = q->radius_const;
    A.minl = q->min_level; A.maxl = q->max_level; A.qangle = nullptr;
    A.accept = 0; A.nnratio = 0.f; A.th_dist = 256; A.histogram = 0; A.prematch = nullptr; A.sort_lists = 0;
    if (tgt->n >= (1 << 22)) return ORB_ERR_CAPACITY;
    size_t off = 0;
    A.cnt = (int*)(scratch + off); off += ((size_t)q->n * 4 + 255) & ~(size_t)255;
    const size_t avail = scratch_bytes > off ? (scratch_bytes - off) / 4 : 0;
    A.cap = (int)std::min<size_t>((size_t)tgt->n, q->n ? avail / (size_t)q->n : 0);
    A.cap = std::min(A.cap, 1 << 20);
    A.list = (uint32_t*)(scratch + off);
    if (A.cap < 1) return ORB_ERR_CAPACITY;
    ORB_CUDA(cudaMemsetAsync(d_result, 0, 8, s));
    k_win_candidates<<<(q->n * 32 + 255) / 256, 256, 0, s>>>(A);
    k_win_best<<<(q->n + 255) / 256, 256, 0, s>>>(A, best_idx, best_dist, d_result);
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

// ------------------------------------------------------------------ MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:185-250)
// One CTA of 64 threads per map point.  Row i's median is the k-th smallest (k = (int)(0.5*(N-1)), :236) of its N distances, found
// by bisection on the value (distances are 0..256): 9 counting passes that recompute the distances from the descriptors in shared
// memory, so N is unbounded and nothing but the descriptors is stored.  BestIdx = first row with the smallest median (:238).
constexpr int DD_THREADS = 64, DD_SMEM_DESC = 512;          // descriptors staged in shared memory; larger groups read global memory
__global__ void __launch_bounds__(DD_THREADS)
k_distinctive(const uint8_t* __restrict__ desc, const int32_t* __restrict__ start, int32_t* __restrict__ best_idx,
              int32_t* __restrict__ best_median)
{
    __shared__ uint4 sd[DD_SMEM_DESC * 2];
    __shared__ unsigned long long s_best;
    const int p = blockIdx.x, tid = threadIdx.x;
    const int s0 = start[p], N = start[p + 1] - s0;
    if (N <= 0) { if (tid == 0) { best_idx[p] = -1; best_median[p] = INT_MAX; } return; }
    const uint4* gd = reinterpret_cast<const uint4*>(desc + (size_t)s0 * 32);
    const bool in_smem = N <= DD_SMEM_DESC;
    if (in_smem) for (int i = tid; i < 2 * N; i += DD_THREADS) sd[i] = __ldg(gd + i);
    if (tid == 0) s_best = ~0ull;
    __syncthreads();
    const uint4* D = in_smem ? sd : gd;
    const int k = (int)(0.5 * (double)(N - 1));
    unsigned long long mine = ~0ull;
    for (int i = tid; i < N; i += DD_THREADS) {
        const uint4 a = D[2 * i], b = D[2 * i + 1];
        int lo = 0, hi = 256;                                   // smallest v with #(d <= v) > k
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            int cnt = 0;
            for (int j = 0; j < N; j++) {
                const uint4 c = D[2 * j], e = D[2 * j + 1];
                const int dist = __popc(a.x ^ c.x) + __popc(a.y ^ c.y) + __popc(a.z ^ c.z) + __popc(a.w ^ c.w) +
                                 __popc(b.x ^ e.x) + __popc(b.y ^ e.y) + __popc(b.z ^ e.z) + __popc(b.w ^ e.w);
                cnt += dist <= mid;
            }
            if (cnt > k) hi = mid; else lo = mid + 1;
        }
        mine = min(mine, ((unsigned long long)lo << 32) | (unsigned)i);       // rows ascend per thread: the first minimum survives
    }
    atomicMin(&s_best, mine);
    __syncthreads();
    if (tid == 0) { best_idx[p] = (int32_t)(s_best & 0xffffffffu); best_median[p] = (int32_t)(s_best >> 32); }
}

int orb_launch_distinctive(const uint8_t* d_desc, const int32_t* d_start, int npoints, int32_t* d_best_idx, int32_t* d_best_median,
                           cudaStream_t s)
{
    if (npoints <= 0) return ORB_OK;
    k_distinctive<<<npoints, DD_THREADS, 0, s>>>(d_desc, d_start, d_best_idx, d_best_median);
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

size_t orb_init_scratch_bytes(int n1, int n2)
{
    const size_t cap = (size_t)std::min(n2, 1024);
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    return al((size_t)n1 * 4) * 7 + al((size_t)n1) + al((size_t)n2 * 4) + al((size_t)n2 * INIT_TAB * 4) + (size_t)n1 * cap * 4 + 1024;
}

// window centres = vbPrevMatched, level range [0,0], only octave-0 features of F1 search (:613-618)
__global__ void k_init_prepare(const orb_keypoint* __restrict__ kps1, const float* __restrict__ prev, int n1, float* __restrict__ u,
                               float* __restrict__ v, int32_t* __restrict__ lv, uint8_t* __restrict__ active, float* __restrict__ angle)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1) return;
    const orb_keypoint kp = kps1[i];
    u[i] = prev[2 * i]; v[i] = prev[2 * i + 1]; lv[i] = 0; active[i] = kp.octave <= 0; angle[i] = kp.angle;
}

int orb_launch_search_for_initialization(orb_ctx* c, const orb_frame_view* f1, const orb_frame_view* f2, float* d_prev, int window,
                                         float nnratio, int check_ori, int32_t* d_matches12, int* d_result, uint8_t* scratch,
                                         size_t scratch_bytes, cudaStream_t s)
{
    (void)c;
    if (f2->n >= (1 << 22)) return ORB_ERR_CAPACITY;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t n1 = (size_t)f1->n, n2 = (size_t)f2->n;
    size_t off = 0;
    auto take = [&](size_t bytes) { uint8_t* p = scratch + off; off += al(bytes); return p; };
    float* d_u = (float*)take(n1 * 4); float* d_v = (float*)take(n1 * 4); float* d_angle = (float*)take(n1 * 4);
    int32_t* d_lv = (int32_t*)take(n1 * 4); uint8_t* d_active = take(n1);
    WinArgs A;
    A.tgt = *f2; A.nq = f1->n; A.active = d_active; A.qdesc = f1->desc; A.u = d_u; A.v = d_v; A.xyz = nullptr; A.project = 0;
    for (int i = 0; i < 16; i++) A.T[i] = 0.f;
    A.check_bounds = 0; A.radius = nullptr; A.radius_const = (float)window; A.minl = d_lv; A.maxl = d_lv; A.qangle = d_angle;
    A.accept = 0; A.nnratio = nnratio; A.th_dist = TH_LOW; A.histogram = check_ori; A.prematch = nullptr; A.sort_lists = 0;
    A.cnt = (int*)take(n1 * 4);
    int* choice = (int*)take(n1 * 4);
    int* cdist = (int*)take(n1 * 4);
    int* has = (int*)take(n2 * 4);
    int* tab = (int*)take(n2 * INIT_TAB * 4);
    const size_t avail = scratch_bytes > off ? (scratch_bytes - off) / 4 : 0;
    A.cap = (int)std::min<size_t>(n2, n1 ? avail / n1 : 0);
    A.list = (uint32_t*)(scratch + off);
    if (A.cap < 1) return ORB_ERR_CAPACITY;
    k_init_prepare<<<(f1->n + 255) / 256, 256, 0, s>>>(f1->kps, d_prev, f1->n, d_u, d_v, d_lv, d_active, d_angle);
    k_win_candidates<<<(f1->n * 32 + 255) / 256, 256, 0, s>>>(A);
    k_init_resolve<<<1, 1024, 0, s>>>(A, d_matches12, d_prev, tab, has, choice, cdist, d_result);
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}
