// orb_comm.cu — the multi-GPU half of the C ABI (include/orb_b200.h, "multi-GPU"): SURVEY.md §8e behind plain C calls.
//
//   * keyframe-descriptor database sharded by contiguous row ranges, queries replicated, per-shard (idx1, d1, d2) from k_knn2 with
//     GLOBAL row indices, ONE exchange of 12 bytes per query and rank, exact merge (k_knn2_merge: lexicographic min of (d1, idx1),
//     second = 2nd smallest of the multiset union);
//   * frames sharded over the devices in contiguous blocks with no data-path collective.
//
// Two ways to own the ranks:
//   orb_comm_init(ngpus)                  ONE process drives devices 0..ngpus-1 (what the reference is: a single process,
//                                         src/main.cc:165-212) — one context + stream per device;
//   orb_comm_init_rank(ctx, n, r, id)     one process per GPU (torchrun / MPI style), the NCCL unique id travels through the host
//                                         application (orb_comm_unique_id on rank 0).
// Two transports for the exchange:
//   "nccl"  ncclAllGather on the rank's own stream (grouped over the local ranks), NVLink / NVSwitch underneath.  libnccl.so.2 is
//           opened at run time (dlopen) — no link-time dependency, and inside a process that already maps NCCL (PyTorch) the same
//           copy is reused;
//   "p2p"   single-process only: peer access is enabled between the devices and the merge kernel on rank 0 LOADS every rank's
//           partials straight over NVLink (no gather step, one kernel).  Default when NCCL cannot be opened; ORB_COMM_TRANSPORT=p2p|nccl.
#include "orb_internal.h"
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    bool ok = false;
};

NcclApi& nccl()
{
    static NcclApi A;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = { getenv("ORB_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
        for (const char* n : names) {
            if (!n || !*n) continue;
            A.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (A.h) break;
        }
        if (!A.h) return;
#define SYM(f) *(void**)&A.f = dlsym(A.h, "nccl" #f)
        SYM(GetUniqueId); SYM(CommInitRank); SYM(CommInitAll); SYM(CommDestroy); SYM(AllGather); SYM(GroupStart); SYM(GroupEnd);
        SYM(GetErrorString); SYM(GetVersion);
#undef SYM
        A.ok = A.GetUniqueId && A.CommInitRank && A.CommInitAll && A.CommDestroy && A.AllGather && A.GroupStart && A.GroupEnd;
    });
    return A;
}

int nccl_fail(ncclResult_t r, const char* what)
{
    g_last_cuda_error = std::string(what) + ": NCCL: " + (nccl().GetErrorString ? nccl().GetErrorString(r) : "error") + " (" + std::to_string((int)r) + ")";
    return ORB_ERR_CUDA;
}
#define ORB_NCCL(x) do { ncclResult_t r__ = (x); if (r__ != ncclSuccess) return nccl_fail(r__, #x); } while (0)

struct LocalRank {
    int device = 0, rank = 0;
    orb_ctx* ctx = nullptr; bool own_ctx = false;
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev = nullptr;
    // exchange buffers: part = this rank's (idx1, d1, d2) [3][nq]; all = every rank's, rank-major (NCCL transport only)
    int32_t* d_part = nullptr; int32_t* d_all = nullptr; int32_t* d_out = nullptr; uint8_t* d_q = nullptr;
    size_t cap_q = 0;                     // queries the buffers above hold
    // database shard
    const uint8_t* d_rows = nullptr; uint8_t* d_rows_owned = nullptr; size_t rows_cap = 0;
    int64_t nrows = 0, row_base = 0;
    long long ticket = -1;
};

} // namespace

struct orb_comm {
    int nranks = 0;
    bool single_process = true;
    bool use_nccl = false;
    std::vector<LocalRank> local;         // every rank (single process) or this process's one rank
    std::mutex mu;                        // a communicator is driven by one thread at a time
};

namespace {

int grow_exchange(orb_comm* m, LocalRank& R, int nq)
{
    if ((size_t)nq <= R.cap_q && R.d_part) return ORB_OK;
    ORB_CUDA(cudaSetDevice(R.device));
    ORB_CUDA(cudaStreamSynchronize(R.stream));
    if (m->single_process) for (LocalRank& o : m->local) { ORB_CUDA(cudaSetDevice(o.device)); ORB_CUDA(cudaStreamSynchronize(o.stream)); }   // rank 0 may still be loading our partials
    ORB_CUDA(cudaSetDevice(R.device));
    for (void* p : { (void*)R.d_part, (void*)R.d_all, (void*)R.d_out, (void*)R.d_q }) if (p) cudaFree(p);
    R.d_part = R.d_all = R.d_out = nullptr; R.d_q = nullptr; R.cap_q = 0;
    const size_t n = (size_t)std::max(nq, 256);
    ORB_CUDA(cudaMalloc((void**)&R.d_part, 3 * n * sizeof(int32_t)));
    ORB_CUDA(cudaMalloc((void**)&R.d_all, (size_t)m->nranks * 3 * n * sizeof(int32_t)));
    ORB_CUDA(cudaMalloc((void**)&R.d_out, 3 * n * sizeof(int32_t)));
    ORB_CUDA(cudaMalloc((void**)&R.d_q, n * 32));
    R.cap_q = n;
    return ORB_OK;
}

void free_rank(LocalRank& R)
{
    cudaSetDevice(R.device);
    if (R.stream) cudaStreamSynchronize(R.stream);
    if (R.comm && nccl().ok) nccl().CommDestroy(R.comm);
    for (void* p : { (void*)R.d_part, (void*)R.d_all, (void*)R.d_out, (void*)R.d_q, (void*)R.d_rows_owned }) if (p) cudaFree(p);
    if (R.ev) cudaEventDestroy(R.ev);
    if (R.stream) cudaStreamDestroy(R.stream);
    if (R.own_ctx && R.ctx) orb_destroy(R.ctx);
}

bool want_nccl(bool single_process)
{
    const char* e = getenv("ORB_COMM_TRANSPORT");
    if (e && !strcmp(e, "p2p") && single_process) return false;
    return nccl().ok;
}

// kNN-2 of one rank over its shard, the exchange, the merge — everything stream-ordered on R.stream (or the caller's stream)
int rank_knn2(orb_comm* m, LocalRank& R, const uint8_t* d_q, int nq, const uint8_t* d_rows, int64_t nrows, int64_t row_base,
              int32_t* part, cudaStream_t s)
{
    if (row_base < 0 || row_base + nrows > (int64_t)INT_MAX) return ORB_ERR_CAPACITY;
    return orb_launch_knn2(R.ctx, d_q, nq, d_rows, nrows, 1, (int32_t)row_base, part, part + nq, part + 2 * (size_t)nq, s);
}

} // namespace

extern "C" {

int orb_comm_unique_id(void* id128)
{
    if (!id128) return ORB_ERR_INVALID;
    if (!nccl().ok) { g_last_cuda_error = "orb_comm_unique_id: libnccl.so.2 could not be opened"; return ORB_ERR_CUDA; }
    static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id is 128 bytes");
    ncclUniqueId id;
    ORB_NCCL(nccl().GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return ORB_OK;
}

orb_comm* orb_comm_init_rank(orb_ctx* ctx, int nranks, int rank, const void* id128)
{
    if (!ctx || nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !id128)) { g_last_cuda_error = "orb_comm_init_rank: invalid argument"; return nullptr; }
    if (nranks > 1 && !nccl().ok) { g_last_cuda_error = "orb_comm_init_rank: libnccl.so.2 could not be opened (one process per GPU needs NCCL)"; return nullptr; }
    orb_comm* m = new orb_comm;
    m->nranks = nranks; m->single_process = false; m->use_nccl = nranks > 1;
    m->local.resize(1);
    LocalRank& R = m->local[0];
    R.device = ctx->device; R.rank = rank; R.ctx = ctx;
    bool ok = cudaSetDevice(R.device) == cudaSuccess && cudaStreamCreateWithFlags(&R.stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&R.ev, cudaEventDisableTiming) == cudaSuccess;
    if (ok && nranks > 1) {
        ncclUniqueId id;
        memcpy(&id, id128, sizeof id);
        const ncclResult_t r = nccl().CommInitRank(&R.comm, nranks, id, rank);
        if (r != ncclSuccess) { nccl_fail(r, "ncclCommInitRank"); ok = false; }
    } else if (!ok) orb_cuda_fail(cudaGetLastError(), "orb_comm_init_rank");
    if (!ok) { free_rank(R); delete m; return nullptr; }
    return m;
}

orb_comm* orb_comm_init(int ngpus)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { g_last_cuda_error = "orb_comm_init: no usable CUDA device"; cudaGetLastError(); return nullptr; }
    if (ngpus <= 0) ngpus = ndev;
    if (ngpus > ndev) { g_last_cuda_error = "orb_comm_init: more ranks than visible devices"; return nullptr; }
    orb_comm* m = new orb_comm;
    m->nranks = ngpus; m->single_process = true;
    m->local.resize(ngpus);
    bool ok = true;
    for (int r = 0; r < ngpus && ok; r++) {
        LocalRank& R = m->local[r];
        R.device = r; R.rank = r;
        // a matcher-only context: nothing of the extraction pipeline is allocated until an extract call arrives
        R.ctx = orb_create(r, 1000, 1.2f, 8, ORB_FAST_SCORE, 20, 64, 64, 1);
        R.own_ctx = true;
        ok = R.ctx && cudaSetDevice(r) == cudaSuccess && cudaStreamCreateWithFlags(&R.stream, cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&R.ev, cudaEventDisableTiming) == cudaSuccess;
    }
    m->use_nccl = ok && ngpus > 1 && want_nccl(true);
    if (ok && ngpus > 1 && !m->use_nccl) {
        // p2p transport: rank 0 loads the other ranks' partials directly
        for (int r = 1; r < ngpus && ok; r++) {
            int can = 0;
            ok = cudaDeviceCanAccessPeer(&can, 0, r) == cudaSuccess && can;
        }
        if (ok) {
            cudaSetDevice(0);
            for (int r = 1; r < ngpus; r++) { const cudaError_t e = cudaDeviceEnablePeerAccess(r, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false; cudaGetLastError(); }
        }
        if (!ok) g_last_cuda_error = "orb_comm_init: neither NCCL nor peer access between the devices is available";
    }
    if (ok && m->use_nccl) {
        std::vector<ncclComm_t> comms(ngpus);
        std::vector<int> devs(ngpus);
        for (int r = 0; r < ngpus; r++) devs[r] = r;
        const ncclResult_t rr = nccl().CommInitAll(comms.data(), ngpus, devs.data());
        if (rr != ncclSuccess) { nccl_fail(rr, "ncclCommInitAll"); ok = false; }
        else for (int r = 0; r < ngpus; r++) m->local[r].comm = comms[r];
    }
    if (!ok) {
        if (g_last_cuda_error.empty()) orb_cuda_fail(cudaGetLastError(), "orb_comm_init");
        for (LocalRank& R : m->local) free_rank(R);
        delete m;
        return nullptr;
    }
    return m;
}

void orb_comm_destroy(orb_comm* m)
{
    if (!m) return;
    for (LocalRank& R : m->local) free_rank(R);
    delete m;
}

int orb_comm_size(const orb_comm* m) { return m ? m->nranks : 0; }
const char* orb_comm_transport(const orb_comm* m) { return !m ? "" : (m->nranks == 1 ? "none (1 rank)" : (m->use_nccl ? "nccl" : "p2p")); }
orb_ctx* orb_comm_context(orb_comm* m, int rank)
{
    if (!m) return nullptr;
    for (LocalRank& R : m->local) if (R.rank == rank) return R.ctx;
    return nullptr;
}

int orb_comm_set_extractor(orb_comm* m, int nfeatures, float scale_factor, int nlevels, int score_type, int fast_th, int max_w, int max_h, int max_batch)
{
    if (!m || !m->single_process) return ORB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(m->mu);
    for (LocalRank& R : m->local) {
        orb_ctx* n = orb_create(R.device, nfeatures, scale_factor, nlevels, score_type, fast_th, max_w, max_h, max_batch);
        if (!n) return ORB_ERR_CUDA;
        if (R.own_ctx && R.ctx) orb_destroy(R.ctx);
        R.ctx = n; R.own_ctx = true;
    }
    return ORB_OK;
}

/* ---- database shards ---- */
int orb_comm_db_upload(orb_comm* m, const uint8_t* db, int64_t ndb)
{
    if (!m || !m->single_process || ndb < 0 || (ndb > 0 && !db) || ndb > (int64_t)INT_MAX) return ORB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(m->mu);
    const int64_t base = ndb / m->nranks, rem = ndb % m->nranks;
    for (LocalRank& R : m->local) {
        const int64_t lo = R.rank * base + std::min<int64_t>(R.rank, rem), n = base + (R.rank < rem ? 1 : 0);
        ORB_CUDA(cudaSetDevice(R.device));
        if ((size_t)n * 32 > R.rows_cap) {
            ORB_CUDA(cudaStreamSynchronize(R.stream));
            if (R.d_rows_owned) cudaFree(R.d_rows_owned);
            R.d_rows_owned = nullptr; R.rows_cap = 0;
            ORB_CUDA(cudaMalloc((void**)&R.d_rows_owned, std::max<size_t>((size_t)n * 32, 256)));
            R.rows_cap = std::max<size_t>((size_t)n * 32, 256);
        }
        if (n) ORB_CUDA(cudaMemcpyAsync(R.d_rows_owned, db + (size_t)lo * 32, (size_t)n * 32, cudaMemcpyDefault, R.stream));
        R.d_rows = R.d_rows_owned; R.nrows = n; R.row_base = lo;
    }
    for (LocalRank& R : m->local) { ORB_CUDA(cudaSetDevice(R.device)); ORB_CUDA(cudaStreamSynchronize(R.stream)); }
    return ORB_OK;
}

int orb_comm_db_attach(orb_comm* m, int rank, const uint8_t* d_rows, int64_t nrows, int64_t row_base)
{
    if (!m || nrows < 0 || row_base < 0 || (nrows > 0 && !d_rows) || ((uintptr_t)d_rows & 15)) return ORB_ERR_INVALID;
    if (row_base + nrows > (int64_t)INT_MAX) return ORB_ERR_CAPACITY;
    std::lock_guard<std::mutex> lk(m->mu);
    for (LocalRank& R : m->local) if (R.rank == rank) { R.d_rows = d_rows; R.nrows = nrows; R.row_base = row_base; return ORB_OK; }
    return ORB_ERR_INVALID;
}

/* one process per GPU: this rank's shard is passed with the call; everything is enqueued on `stream` */
int orb_knn2_sharded_device(orb_comm* m, const uint8_t* d_q, int nq, const uint8_t* d_rows, int64_t nrows, int64_t row_base,
                            int32_t* d_idx1, int32_t* d_d1, int32_t* d_d2, void* stream)
{
    if (!m || m->local.size() != 1 || nq < 0 || nrows < 0 || !d_idx1 || !d_d1 || !d_d2) return ORB_ERR_INVALID;
    if (nq == 0) return ORB_OK;
    if (!d_q || (nrows > 0 && !d_rows) || (((uintptr_t)d_q | (uintptr_t)d_rows) & 15)) return ORB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(m->mu);
    LocalRank& R = m->local[0];
    cudaStream_t s = (cudaStream_t)stream;
    ORB_CUDA(cudaSetDevice(R.device));
    if (m->nranks == 1) {
        if (row_base < 0 || row_base + nrows > (int64_t)INT_MAX) return ORB_ERR_CAPACITY;
        return orb_launch_knn2(R.ctx, d_q, nq, d_rows, nrows, 1, (int32_t)row_base, d_idx1, d_d1, d_d2, s);
    }
    int rc = grow_exchange(m, R, nq);
    if (rc) return rc;
    if ((rc = rank_knn2(m, R, d_q, nq, d_rows, nrows, row_base, R.d_part, s))) return rc;
    ORB_NCCL(nccl().AllGather(R.d_part, R.d_all, (size_t)3 * nq, ncclInt32, R.comm, s));
    return orb_launch_knn2_merge(R.d_all, m->nranks, nq, d_idx1, d_d1, d_d2, s);
}

/* single process: shards were set with orb_comm_db_upload / orb_comm_db_attach; q and the outputs are host pointers or device
 * pointers on rank 0's device */
int orb_knn2_sharded(orb_comm* m, const uint8_t* q, int nq, int32_t* idx1, int32_t* d1, int32_t* d2)
{
    if (!m || !m->single_process || nq < 0 || !idx1 || !d1 || !d2) return ORB_ERR_INVALID;
    if (nq == 0) return ORB_OK;
    if (!q) return ORB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(m->mu);
    int rc;
    for (LocalRank& R : m->local) if ((rc = grow_exchange(m, R, nq))) return rc;
    LocalRank& R0 = m->local[0];
    cudaPointerAttributes pa;
    const bool out_dev = cudaPointerGetAttributes(&pa, idx1) == cudaSuccess && (pa.type == cudaMemoryTypeDevice || pa.type == cudaMemoryTypeManaged);
    cudaGetLastError();
    // queries replicated to every device (64 KB for 2000 queries)
    for (LocalRank& R : m->local) {
        ORB_CUDA(cudaSetDevice(R.device));
        ORB_CUDA(cudaMemcpyAsync(R.d_q, q, (size_t)nq * 32, cudaMemcpyDefault, R.stream));
        if ((rc = rank_knn2(m, R, R.d_q, nq, R.d_rows, R.nrows, R.row_base, R.d_part, R.stream))) return rc;
    }
    int32_t* o = R0.d_out;
    if (m->nranks == 1) {
        o = R0.d_part;
    } else if (m->use_nccl) {
        ORB_NCCL(nccl().GroupStart());
        for (LocalRank& R : m->local) {
            const ncclResult_t r = nccl().AllGather(R.d_part, R.d_all, (size_t)3 * nq, ncclInt32, R.comm, R.stream);
            if (r != ncclSuccess) { nccl().GroupEnd(); return nccl_fail(r, "ncclAllGather"); }
        }
        ORB_NCCL(nccl().GroupEnd());
        ORB_CUDA(cudaSetDevice(R0.device));
        if ((rc = orb_launch_knn2_merge(R0.d_all, m->nranks, nq, o, o + nq, o + 2 * (size_t)nq, R0.stream))) return rc;
    } else {
        // p2p: rank 0's merge kernel reads every rank's partials in place, over NVLink
        const int32_t* ptrs[ORB_COMM_MAX_RANKS];
        for (LocalRank& R : m->local) {
            ptrs[R.rank] = R.d_part;
            if (R.rank == 0) continue;
            ORB_CUDA(cudaSetDevice(R.device));
            ORB_CUDA(cudaEventRecord(R.ev, R.stream));
            ORB_CUDA(cudaStreamWaitEvent(R0.stream, R.ev, 0));
        }
        ORB_CUDA(cudaSetDevice(R0.device));
        if ((rc = orb_launch_knn2_merge_ptrs(ptrs, m->nranks, nq, o, o + nq, o + 2 * (size_t)nq, R0.stream))) return rc;
    }
    ORB_CUDA(cudaSetDevice(R0.device));
    const cudaMemcpyKind k = out_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    ORB_CUDA(cudaMemcpyAsync(idx1, o, (size_t)nq * 4, k, R0.stream));
    ORB_CUDA(cudaMemcpyAsync(d1, o + nq, (size_t)nq * 4, k, R0.stream));
    ORB_CUDA(cudaMemcpyAsync(d2, o + 2 * (size_t)nq, (size_t)nq * 4, k, R0.stream));
    for (LocalRank& R : m->local) { ORB_CUDA(cudaSetDevice(R.device)); ORB_CUDA(cudaStreamSynchronize(R.stream)); }
    return ORB_OK;
}

/* frames in contiguous blocks over the devices, every device running its own copy/compute pipeline (orb_extract_batch_async);
 * one host thread enqueues all of them, then waits for all of them */
int orb_extract_batch_multi(orb_comm* m, const uint8_t* imgs, int nimg, int w, int h, int stride, size_t frame_pitch,
                            orb_keypoint* kps, uint8_t* desc, int cap, int32_t* counts)
{
    if (!m || !m->single_process || nimg < 0 || !kps || !desc || !counts || cap < 1) return ORB_ERR_INVALID;
    if (nimg == 0) return ORB_OK;
    std::lock_guard<std::mutex> lk(m->mu);
    const int base = nimg / m->nranks, rem = nimg % m->nranks;
    int rc = ORB_OK;
    for (LocalRank& R : m->local) {
        const int lo = R.rank * base + std::min(R.rank, rem), n = base + (R.rank < rem ? 1 : 0);
        R.ticket = -1;
        if (n == 0) continue;
        const int r = orb_extract_batch_async(R.ctx, imgs ? imgs + (size_t)lo * frame_pitch : nullptr, n, w, h, stride, frame_pitch,
                                              kps + (size_t)lo * cap, desc + (size_t)lo * cap * 32, cap, counts + lo, &R.ticket);
        if (r != ORB_OK && rc == ORB_OK) rc = r;
    }
    for (LocalRank& R : m->local) {
        if (R.ticket < 0) continue;
        const int r = orb_wait(R.ctx, R.ticket);
        if (r != ORB_OK && rc == ORB_OK) rc = r;
    }
    return rc;
}

} // extern "C"
