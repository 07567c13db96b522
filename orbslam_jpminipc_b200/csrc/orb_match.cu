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
    __shared__ int s_cnt[ORB_GRID_COLS * ORB_GRID_ROWS];
    __shared__ int s_warp[32];
    const int NC = ORB_GRID_COLS * ORB_GRID_ROWS;
    const int tid = threadIdx.x;
    const float invW = __fdiv_rn((float)ORB_GRID_COLS, (float)(max_x - min_x));
    const float invH = __fdiv_rn((float)ORB_GRID_ROWS, (float)(max_y - min_y));
    for (int i = tid; i < n; i += blockDim.x) {
        const int px = (int)roundf(__fmul_rn(__fsub_rn(kps[i].x, (float)min_x), invW));
        const int py = (int)roundf(__fmul_rn(__fsub_rn(kps[i].y, (float)min_y), invH));
        const bool ok = !(px < 0 || px >= ORB_GRID_COLS || py < 0 || py >= ORB_GRID_ROWS);
        s_cell[i] = ok ? (unsigned short)(px * ORB_GRID_ROWS + py) : (unsigned short)0xffff;
    }
    __syncthreads();
    // 3 cells per thread: count
    int cnt[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int c = tid * 3 + k;
        int v = 0;
        if (c < NC) for (int i = 0; i < n; i++) v += (s_cell[i] == c);
        cnt[k] = v;
    }
    // block exclusive scan over threads (each owns 3 consecutive cells)
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
        if (c < NC) { cell_start[c] = base; s_cnt[c] = base; }
        base += cnt[k];
    }
    if (tid == 1023) cell_start[NC] = base;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int c = tid * 3 + k;
        if (c < NC && cnt[k]) {
            int o = s_cnt[c];
            for (int i = 0; i < n; i++) if (s_cell[i] == c) cell_items[o++] = i;
        }
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
    if (rows > (1 << 30)) rows = 1 << 30;
    const int nchunks = (int)((ndb + rows - 1) / rows);
    Knn2Args A;
    A.q = d_q; A.db = d_db; A.nq = nq; A.ndb = ndb; A.rows_per_chunk = (int)rows; A.nchunks = nchunks;
    A.idx_base = idx_base; A.out = nullptr; A.o_idx1 = d_idx1; A.o_d1 = d_d1; A.o_d2 = d_d2;
    if (nchunks > 1) {
        const size_t need = (size_t)npairs * nchunks * 3 * nq * sizeof(int32_t);
        if (need > c->knn_part_bytes) {
            ORB_CUDA(cudaStreamSynchronize(s));
            if (c->d_knn_part) cudaFree(c->d_knn_part);
            c->d_knn_part = nullptr; c->knn_part_bytes = 0;
            ORB_CUDA(cudaMalloc((void**)&c->d_knn_part, need));
            c->knn_part_bytes = need;
        }
        A.out = c->d_knn_part;
    }
    if (nchunks > 65535 || npairs > 65535) return ORB_ERR_CAPACITY;
    k_knn2<<<dim3(nchunks, qtiles, npairs), KNN_THREADS, 0, s>>>(A);
    c->last_launches = 1;
    if (nchunks > 1) {
        k_knn2_merge<<<dim3((nq + 127) / 128, npairs), 128, 0, s>>>(c->d_knn_part, nchunks, nq, npairs, d_idx1, d_d1, d_d2);
        c->last_launches = 2;
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
    static bool attr_set = false;
    if (!attr_set) { ORB_CUDA(cudaFuncSetAttribute(k_grid_build, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); attr_set = true; }
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
