// orb_match_tc.cu — EXPERIMENT: brute-force Hamming kNN-2 on the 5th-generation tensor cores.
//
// k_knn2 (orb_match.cu) is bound by the POPC / ALU pipes.  The same numbers come out of an exact integer contraction: map every
// descriptor bit b to the int8 value 2b-1 in {-1,+1}; then for two 256-bit descriptors
//       dot(q', d') = (#equal bits) - (#different bits) = 256 - 2 * Hamming(q, d)
// so   Hamming = (256 - dot) / 2,  smallest distance = largest dot.  tcgen05.mma kind::i8 (s8 x s8 -> s32, exact) computes a
// 128-query x 256-row block of dots per tile (8 instructions of K = 32), the accumulator lives in TMEM, and the CUDA cores only
//   (1) expand the packed bits of the database tile into the canonical K-major shared-memory operand layout (one 2 KB lookup table:
//       byte -> eight +-1 bytes), and
//   (2) run the best / second-best scan on the accumulator read back with tcgen05.ld: per pair one IMAD on the FMA pipe (16-bit key =
//       distance*128 + column, two columns packed per register) and 1.25 16-bit SIMD min/max on the ALU pipe — against 5 POPC + 14 LOP3
//       + ... in k_knn2 (-DORB_TC_SCAN32 keeps the earlier 32-bit keys: one IMAD + 2.5 min/max per pair).
// The MMA of tile t runs asynchronously (tcgen05.commit -> mbarrier) while the CUDA cores scan tile t-1 out of the other TMEM stage.
// Results are bit-identical to k_knn2 (same scan semantics: lowest index wins, d2 = second order statistic); selected with
// orb_set_knn_engine(ctx, ORB_KNN_TENSOR) or ORB_KNN_ENGINE=tensor, never by default (north_star pins the POPC path).
#include "orb_internal.h"
#include <algorithm>
#include <climits>

namespace {

constexpr int TC_SCAN_WARPS = 8;      // warps 0..7: best / second-best scan of the accumulator (TMEM lane quarter = warp & 3, column half = warp >> 2)
constexpr int TC_PROD_WARPS = 4;      // warps 8..11: expand database bits into the int8 operand tile
constexpr int TC_THREADS = (TC_SCAN_WARPS + TC_PROD_WARPS + 1) * 32;      // + warp 12: TMEM allocation and the MMA issuer
constexpr int TC_M = 128;             // queries per M-tile = TMEM lanes
constexpr int TC_MT = 2;              // M-tiles per CTA: one expanded database tile feeds two accumulators (256 queries)
constexpr int TC_N = 128;             // database rows per tile = accumulator columns per (stage, M-tile): 2 stages x 2 M-tiles x 128 = all 512 TMEM columns
constexpr int TC_KBYTES = 256;        // one int8 per descriptor bit
constexpr int TC_KEYMUL = 300;        // scan key = (dot + 256) * 300 + (255 - column); see the scanner
constexpr int TC_LBO = 128;           // bytes between the two 16-byte K chunks of one core-matrix pair (K-major, no swizzle)
constexpr int TC_SBO = 16 * 128;      // bytes between 8-row groups: 16 K-chunks of 128 bytes each
constexpr int TC_A_BYTES = TC_M * TC_KBYTES;          // 32 KB per M-tile
constexpr int TC_B_BYTES = TC_N * TC_KBYTES;          // 32 KB per stage
constexpr int TC_LUT_COPIES = 16;                     // lane pair j reads copy j: at most 2-way bank conflicts on the table
constexpr int TC_LUT_BYTES = 256 * TC_LUT_COPIES * 8; // 32 KB
constexpr int TC_SMEM = TC_MT * TC_A_BYTES + 2 * TC_B_BYTES + TC_LUT_BYTES + 1024 /* barriers */ + 1024 /* alignment slack */;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits [0,14), leading byte offset >> 4 in [16,30),
// stride byte offset >> 4 in [32,46), version 1 in [46,48), layout type SWIZZLE_NONE = 0 in [61,64)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(TC_LBO >> 4) << 16) | ((uint64_t)(TC_SBO >> 4) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = S32 (c_format 2), A = B = signed 8 bit (format 1), both K-major,
// N >> 3 in [17,23), M >> 4 in [24,29)
constexpr uint32_t TC_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(TC_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ bool tc_mbar_test(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "TC_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra TC_DONE;\n\t"
        "bra TC_WAIT;\n\t"
        "TC_DONE:\n\t}" ::"r"(s_u32(bar)), "r"(parity) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp receives lane (base + i), columns c .. c+31.  Asynchronous: the
// registers are valid after tmem_ld_wait().
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
// the loaded registers pass THROUGH the wait as in/out operands, so that no use of them can be scheduled ahead of it
__device__ __forceinline__ void tmem_ld_wait(int (&v)[32])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
        : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
          "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
          "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
          "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
        :: "memory");
}

struct Knn2TcArgs {
    const uint8_t* q; const uint8_t* db;
    int nq; long long ndb;
    int rows_per_chunk, nchunks;
    int32_t idx_base;
    int32_t* out;                    // nchunks > 1: partials [pair][chunk][3][nq]
    int32_t* o_idx1; int32_t* o_d1; int32_t* o_d2;
};

// one descriptor row (32 bytes in two uint4) -> 256 int8 in the canonical operand layout: chunk c (16 bytes = 16 bits of the row)
// of row r lives at (r / 8) * SBO + c * LBO + (r % 8) * 16.  lutc = this lane's copy of the byte -> eight +-1 bytes table.
__device__ __forceinline__ void expand_row(uint8_t* op, int r, const uint4& lo, const uint4& hi, const uint8_t* lutc)
{
    uint8_t* base = op + (r >> 3) * TC_SBO + (r & 7) * 16;
    const uint32_t w[8] = { lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w };
    constexpr int ES = TC_LUT_COPIES * 8;                            // bytes between consecutive table entries of one copy
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint2 e0 = *reinterpret_cast<const uint2*>(lutc + (w[i] & 0xffu) * ES);
        const uint2 e1 = *reinterpret_cast<const uint2*>(lutc + ((w[i] >> 8) & 0xffu) * ES);
        const uint2 e2 = *reinterpret_cast<const uint2*>(lutc + ((w[i] >> 16) & 0xffu) * ES);
        const uint2 e3 = *reinterpret_cast<const uint2*>(lutc + (w[i] >> 24) * ES);
        *reinterpret_cast<uint4*>(base + (2 * i) * TC_LBO) = make_uint4(e0.x, e0.y, e1.x, e1.y);
        *reinterpret_cast<uint4*>(base + (2 * i + 1) * TC_LBO) = make_uint4(e2.x, e2.y, e3.x, e3.y);
    }
}

// Warp-specialised, three asynchronous stages joined by mbarriers (tile t uses stage s = t & 1, k-th use of a stage has parity k & 1):
//   producers : wait full[s] of tile t-2 (its MMAs have read the stage)  -> expand tile t -> fence.proxy.async -> arrive ready[s]
//   MMA issuer: wait ready[s], wait tfree[s] of tile t-2 (scanners done with the TMEM stage) -> 8 x tcgen05.mma -> commit -> full[s]
//   scanners  : wait full[s] -> tcgen05.ld + min/max scan -> arrive tfree[s]
__global__ void __launch_bounds__(TC_THREADS, 1)
k_knn2_tc(Knn2TcArgs A)
{
    extern __shared__ uint8_t tc_raw[];
    uint8_t* sm = reinterpret_cast<uint8_t*>(((uintptr_t)tc_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = sm;                                                                   // [TC_MT][32 KB]
    uint8_t* sB = sm + TC_MT * TC_A_BYTES;                                              // [2 stages][32 KB]
    uint8_t* lut = sB + 2 * TC_B_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(lut + TC_LUT_BYTES);
    uint64_t *bar_ready = bars, *bar_full = bars + 2, *bar_tfree = bars + 4;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 6);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunk = blockIdx.x, mtp = blockIdx.y, pair = blockIdx.z;
    const long long row0 = (long long)chunk * A.rows_per_chunk;
    const int nrows = (int)min((long long)A.rows_per_chunk, A.ndb - row0);
    const uint8_t* db = A.db + ((size_t)pair * A.ndb + row0) * 32;
    const int ntiles = (nrows + TC_N - 1) / TC_N;
    const bool scanner = warp < TC_SCAN_WARPS, producer = warp >= TC_SCAN_WARPS && warp < TC_SCAN_WARPS + TC_PROD_WARPS;

    // ---- one-time setup: TMEM (all 512 columns), barriers, lookup table, the two query operand tiles ----
    if (warp == TC_SCAN_WARPS + TC_PROD_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s_u32(s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < 2; i++) { tc_mbar_init(&bar_ready[i], TC_PROD_WARPS); tc_mbar_init(&bar_full[i], 1); tc_mbar_init(&bar_tfree[i], TC_SCAN_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 256 * TC_LUT_COPIES; i += TC_THREADS) {
        const int b = i / TC_LUT_COPIES;
        uint32_t lo = 0, hi = 0;                                    // byte value b -> eight int8: bit k set -> +1, clear -> -1
#pragma unroll
        for (int k = 0; k < 4; k++) {
            lo |= (((b >> k) & 1) ? 0x01u : 0xffu) << (8 * k);
            hi |= (((b >> (k + 4)) & 1) ? 0x01u : 0xffu) << (8 * k);
        }
        reinterpret_cast<uint2*>(lut)[i] = make_uint2(lo, hi);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *s_tmem;
    const uint8_t* lutc = lut + (lane >> 1) * 8;
    if (scanner) {                                                  // 256 threads = 256 query rows (rows past nq repeat the last query, never written back)
        const int m = tid >> 7, r = tid & 127;
        const int qi = min((mtp * TC_MT + m) * TC_M + r, A.nq - 1);
        const uint4* qp = reinterpret_cast<const uint4*>(A.q + ((size_t)pair * A.nq + qi) * 32);
        expand_row(sA + m * TC_A_BYTES, r, __ldg(qp), __ldg(qp + 1), lutc);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (producer) {
        const int p = tid - TC_SCAN_WARPS * 32;                     // row p of every tile
        uint4 rlo = make_uint4(0, 0, 0, 0), rhi = rlo;
        auto fetch = [&](int t) {
            const long long a = (long long)t * TC_N + p;
            if (a < nrows) { const uint4* g = reinterpret_cast<const uint4*>(db + (size_t)a * 32); rlo = __ldg(g); rhi = __ldg(g + 1); }
        };
        fetch(0);
        for (int t = 0; t < ntiles; t++) {
            const int s = t & 1;
            uint8_t* Bs = sB + s * TC_B_BYTES;
            if (t >= 2) tc_mbar_wait(&bar_full[s], (uint32_t)(((t - 2) >> 1) & 1));     // the MMAs of tile t-2 have consumed this stage
            if ((long long)t * TC_N + p < nrows) expand_row(Bs, p, rlo, rhi, lutc);
            if (t + 1 < ntiles) fetch(t + 1);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                // generic-proxy stores -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(&bar_ready[s]);
        }
    } else if (warp == TC_SCAN_WARPS + TC_PROD_WARPS) {
        if (lane == 0) {
            const uint64_t a_desc0 = smem_desc(s_u32(sA)), a_desc1 = smem_desc(s_u32(sA + TC_A_BYTES));
            for (int t = 0; t < ntiles; t++) {
                const int s = t & 1;
                tc_mbar_wait(&bar_ready[s], (uint32_t)((t >> 1) & 1));
                if (t >= 2) tc_mbar_wait(&bar_tfree[s], (uint32_t)(((t - 2) >> 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t b_desc = smem_desc(s_u32(sB + s * TC_B_BYTES));
#pragma unroll
                for (int m = 0; m < TC_MT; m++)
#pragma unroll
                    for (int k = 0; k < TC_KBYTES / 32; k++)            // K = 32 int8 per instruction = two 16-byte chunks = 2 * LBO bytes
                        mma_i8(tmem + (uint32_t)((s * TC_MT + m) * TC_N), (m ? a_desc1 : a_desc0) + (uint64_t)((k * 2 * TC_LBO) >> 4),
                               b_desc + (uint64_t)((k * 2 * TC_LBO) >> 4), k > 0);
                mma_commit(&bar_full[s]);
            }
        }
    } else {
        // scanner warp w: M-tile w >> 2, TMEM lane quarter w & 3; one query row per thread, all 128 columns of its accumulator
        const int m = warp >> 2, q4 = warp & 3;
        const uint32_t t_lane = tmem + ((uint32_t)(q4 * 32) << 16);
#ifndef ORB_TC_SCAN32
        // ---- 16-bit keys, two columns per register (default) ----
        // key = distance * 128 + column = -64 * dot + 16384 + column fits a 16-bit lane (distance <= 256, 128 columns per tile), smaller is
        // better and the lower column wins a tie.  Two columns are packed by two IMADs on the FMA pipe (no PRMT):
        //   K = dot[j] * -64 + dot[j+1] * (-64 << 16) + constant          (modulo 2^32; every lane stays inside [0, 32895])
        // and the scan keeps smallest / second smallest per lane (even / odd columns) with 5 ALU operations per FOUR keys,
        //   lo = min(A, B), hi = max(A, B);  p2 = min3(max(p1, lo), p2, hi);  p1 = min(p1, lo)
        // (second smallest of {p1 <= p2, lo <= hi}: lo < p1 -> min(p1, hi); lo >= p1 -> min(lo, p2)), against 5 per two 32-bit keys before.
        int R1 = INT_MAX, R2 = INT_MAX, RI = -1;                    // running best distance, second-best distance, best row (chunk-relative)
        int va[32], vb[32];                                          // two register buffers: a load is in flight while the other buffer is scanned
        bool early = false;                                         // the first load of this tile was already issued at the end of the previous one
        for (int t = 0; t < ntiles; t++) {
            const int s = t & 1;
            const int valid = min(TC_N, nrows - t * TC_N);              // columns >= valid belong to rows past the chunk
            const uint32_t tcol = t_lane + (uint32_t)((s * TC_MT + m) * TC_N);
            if (!early) {
                tc_mbar_wait(&bar_full[s], (uint32_t)((t >> 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                tmem_ld32(tcol, va);
            }
            uint32_t p1 = 0xffffffffu, p2 = 0xffffffffu;
            auto pack2 = [](int d0, int d1, int col) {               // keys of columns col, col + 1
                const uint32_t cst = (uint32_t)(16384 + col) + ((uint32_t)(16384 + col + 1) << 16);
                return (uint32_t)d1 * (uint32_t)(-64 * 65536) + ((uint32_t)d0 * (uint32_t)(-64) + cst);
            };
            auto scan32 = [&](const int (&v)[32], int cbase) {
                if (cbase + 32 <= valid) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const uint32_t A2 = pack2(v[j], v[j + 1], cbase + j), B2 = pack2(v[j + 2], v[j + 3], cbase + j + 2);
                        const uint32_t lo = __vminu2(A2, B2), hi = __vmaxu2(A2, B2);
                        p2 = __vimin3_u16x2(__vmaxu2(p1, lo), p2, hi);
                        p1 = __vminu2(p1, lo);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const uint32_t k0 = (cbase + j < valid) ? (uint32_t)(-64 * v[j] + 16384 + cbase + j) : 0xffffu;
                        const uint32_t k1 = (cbase + j + 1 < valid) ? (uint32_t)(-64 * v[j + 1] + 16384 + cbase + j + 1) : 0xffffu;
                        const uint32_t A2 = k0 | (k1 << 16);
                        p2 = __vminu2(p2, __vmaxu2(p1, A2));
                        p1 = __vminu2(p1, A2);
                    }
                }
            };
            tmem_ld_wait(va);
            tmem_ld32(tcol + 32, vb);
            scan32(va, 0);
            tmem_ld_wait(vb);
            tmem_ld32(tcol + 64, va);
            scan32(vb, 32);
            tmem_ld_wait(va);
            tmem_ld32(tcol + 96, vb);
            scan32(va, 64);
            tmem_ld_wait(vb);
            // va is free: if the NEXT tile's accumulator is already complete (its MMAs do not depend on this scan), start its first load
            // now so that its latency hides behind the last 32 columns of this tile.  The test must be warp-uniform: tcgen05.ld is .aligned.
            early = false;
            if (t + 1 < ntiles) {
                const bool ready = tc_mbar_test(&bar_full[s ^ 1], (uint32_t)(((t + 1) >> 1) & 1));
                if (__all_sync(0xffffffffu, ready)) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    tmem_ld32(t_lane + (uint32_t)(((s ^ 1) * TC_MT + m) * TC_N), va);
                    early = true;
                }
            }
            scan32(vb, 96);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(&bar_tfree[s]);
            // merge the even- and the odd-column lane, then fold the tile into the running result: earlier tiles hold lower rows
            const uint32_t a1 = p1 & 0xffffu, b1 = p1 >> 16, a2 = p2 & 0xffffu, b2 = p2 >> 16;
            const uint32_t best = min(a1, b1), second = min(max(a1, b1), min(a2, b2));
            if (best != 0xffffu) {
                const int t1 = (int)(best >> 7), ti = t * TC_N + (int)(best & 127u);
                const int t2 = second == 0xffffu ? INT_MAX : (int)(second >> 7);
                if (t1 < R1) { R2 = min(R1, t2); R1 = t1; RI = ti; }
                else R2 = min(R2, t1);
            }
        }
        const int d1f = R1, d2f = R2;
#else
        int R1 = INT_MIN, R2 = INT_MIN, RI = -1;                    // running best dot, second-best dot, best row (chunk-relative)
        int va[32], vb[32];                                          // two register buffers: a load is in flight while the other buffer is scanned
        bool early = false;                                         // the first load of this tile was already issued at the end of the previous one
        for (int t = 0; t < ntiles; t++) {
            const int s = t & 1;
            const int valid = min(TC_N, nrows - t * TC_N);              // columns >= valid belong to rows past the chunk
            const uint32_t tcol = t_lane + (uint32_t)((s * TC_MT + m) * TC_N);
            if (!early) {
                tc_mbar_wait(&bar_full[s], (uint32_t)((t >> 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                tmem_ld32(tcol, va);
            }
            // key = (dot + 256) * TC_KEYMUL + (255 - col) >= 0: larger dot first, then lower column.  The multiplier is deliberately
            // NOT a power of two: dot * 300 + constant is one IMAD on the (otherwise idle) FMA pipe, where dot * 256 + constant
            // becomes a shift-add on the ALU pipe, which the scan saturates (ncu: ALU 84 %, FMA 3 %).  Two keys per step:
            //   hi = max(a, b), lo = min(a, b);  m2 = max3(min(m1, hi), m2, lo);  m1 = max(m1, hi)      (5 ALU operations per 2 pairs)
            // second largest of {m1 >= m2, hi >= lo}: hi <= m1 -> max(hi, m2) (lo <= hi adds nothing); hi > m1 -> max(m1, lo) (m2 <= m1).
            int m1 = -1, m2 = -1;
            auto scan32 = [&](const int (&v)[32], int cbase) {
                if (cbase + 32 <= valid) {
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const int a = v[j] * TC_KEYMUL + (256 * TC_KEYMUL + 255 - (cbase + j));
                        const int b = v[j + 1] * TC_KEYMUL + (256 * TC_KEYMUL + 255 - (cbase + j + 1));
                        const int hi = max(a, b), lo = min(a, b);
                        m2 = __vimax3_s32(min(m1, hi), m2, lo);
                        m1 = max(m1, hi);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const int key = (cbase + j < valid) ? v[j] * TC_KEYMUL + (256 * TC_KEYMUL + 255 - (cbase + j)) : -1;
                        m2 = max(m2, min(m1, key));
                        m1 = max(m1, key);
                    }
                }
            };
            tmem_ld_wait(va);
            tmem_ld32(tcol + 32, vb);
            scan32(va, 0);
            tmem_ld_wait(vb);
            tmem_ld32(tcol + 64, va);
            scan32(vb, 32);
            tmem_ld_wait(va);
            tmem_ld32(tcol + 96, vb);
            scan32(va, 64);
            tmem_ld_wait(vb);
            // va is free: if the NEXT tile's accumulator is already complete (its MMAs do not depend on this scan), start its first load
            // now so that its latency hides behind the last 32 columns of this tile.  The test must be warp-uniform: tcgen05.ld is .aligned.
            early = false;
            if (t + 1 < ntiles) {
                const bool ready = tc_mbar_test(&bar_full[s ^ 1], (uint32_t)(((t + 1) >> 1) & 1));
                if (__all_sync(0xffffffffu, ready)) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    tmem_ld32(t_lane + (uint32_t)(((s ^ 1) * TC_MT + m) * TC_N), va);
                    early = true;
                }
            }
            scan32(vb, 96);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(&bar_tfree[s]);
            if (m1 >= 0) {                                              // fold the tile into the running result: earlier tiles hold lower rows
                const int q1 = m1 / TC_KEYMUL, t1 = q1 - 256, ti = t * TC_N + (255 - (m1 - q1 * TC_KEYMUL));
                const int t2 = m2 < 0 ? INT_MIN : m2 / TC_KEYMUL - 256;
                if (t1 > R1) { R2 = max(R1, t2); R1 = t1; RI = ti; }
                else R2 = max(R2, t1);
            }
        }
        const int d1f = RI < 0 ? INT_MAX : (256 - R1) >> 1, d2f = R2 == INT_MIN ? INT_MAX : (256 - R2) >> 1;
#endif
        // ---- dots -> distances, write ----
        const int qi = (mtp * TC_MT + m) * TC_M + q4 * 32 + lane;
        if (qi < A.nq) {
            const int d1 = d1f, d2 = d2f;
            const int gi = RI < 0 ? -1 : (int)(row0 + RI) + A.idx_base;
            if (A.nchunks > 1) {
                int32_t* o = A.out + ((size_t)(pair * A.nchunks + chunk) * 3) * A.nq;
                o[qi] = gi; o[A.nq + qi] = d1; o[2 * A.nq + qi] = d2;
            } else {
                const size_t o = (size_t)pair * A.nq + qi;
                A.o_idx1[o] = gi; A.o_d1[o] = d1; A.o_d2[o] = d2;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == TC_SCAN_WARPS + TC_PROD_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

} // namespace

int orb_launch_knn2_tc(orb_ctx* c, const uint8_t* d_q, int nq, const uint8_t* d_db, int64_t ndb, int npairs, int32_t idx_base,
                       int32_t* d_idx1, int32_t* d_d1, int32_t* d_d2, cudaStream_t s)
{
    if (ndb <= 0) return ORB_ERR_INVALID;                      // the empty database is handled by the caller (orb_launch_knn2)
    static int smem_set[64] = { 0 };
    int dev = 0;
    ORB_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !smem_set[dev]) {
        ORB_CUDA(cudaFuncSetAttribute(k_knn2_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
        smem_set[dev] = 1;
    }
    const int mtiles = (nq + TC_MT * TC_M - 1) / (TC_MT * TC_M);         // CTAs along the queries: 256 queries each
    // one CTA per SM (shared memory + all of TMEM): as many chunks as fill one wave of SMs, chunks a multiple of the tile
    long long want_chunks = std::max<long long>(1, (long long)c->num_sms / ((long long)mtiles * npairs));
    long long rows = (ndb + want_chunks - 1) / want_chunks;
    rows = std::max<long long>(rows, 8 * TC_N);
    rows = ((rows + TC_N - 1) / TC_N) * TC_N;
    if (rows > (1 << 22)) rows = 1 << 22;
    const int nchunks = (int)((ndb + rows - 1) / rows);
    if (nchunks > 65535 || npairs > 65535 || mtiles > 65535) return ORB_ERR_CAPACITY;
    if (idx_base < 0) return ORB_ERR_INVALID;
    if (ndb + (int64_t)idx_base > (int64_t)INT_MAX) return ORB_ERR_CAPACITY;
    Knn2TcArgs A;
    A.q = d_q; A.db = d_db; A.nq = nq; A.ndb = ndb; A.rows_per_chunk = (int)rows; A.nchunks = nchunks;
    A.idx_base = idx_base; A.out = nullptr; A.o_idx1 = d_idx1; A.o_d1 = d_d1; A.o_d2 = d_d2;
    int32_t* part = nullptr;
    if (nchunks > 1) {
        ORB_CUDA(cudaMallocFromPoolAsync((void**)&part, (size_t)npairs * nchunks * 3 * nq * sizeof(int32_t), c->pool, s));
        A.out = part;
    }
    k_knn2_tc<<<dim3(nchunks, mtiles, npairs), TC_THREADS, TC_SMEM, s>>>(A);
    c->last_launches = 1;
    if (nchunks > 1) {
        int rc = orb_launch_knn2_merge_pairs(part, nchunks, nq, npairs, d_idx1, d_d1, d_d2, s);
        if (rc) return rc;
        c->last_launches = 2;
        ORB_CUDA(cudaFreeAsync(part, s));
    }
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}
