// orb_api.cu — C ABI (include/orb_b200.h): context, device buffers, extraction entry points.
#include "orb_internal.h"
#include <atomic>
#include <climits>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

thread_local std::string g_last_cuda_error;

int orb_cuda_fail(cudaError_t e, const char* what)
{
    g_last_cuda_error = std::string(what) + ": " + cudaGetErrorString(e);
    return ORB_ERR_CUDA;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency)
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// (re)encode the per-level descriptors over the pyramid buffer: dims {stride, prows, frames}, box {bw, bh, 1}
int orb_build_tmaps(orb_ctx* c, WorkSet& W, int nframes)
{
    if (W.tm_base == W.d_planes && W.tm_frames == nframes && W.tm_w == c->plan.w && W.tm_h == c->plan.h) return ORB_OK;
    PFN_encodeTiled enc = get_encode();
    if (!enc) { g_last_cuda_error = "cuTensorMapEncodeTiled entry point not available"; return ORB_ERR_CUDA; }
    const Plan& P = c->plan;
    for (int l = 0; l < P.nlevels; l++) {
        const LevelGeom& L = P.L[l];
        cuuint64_t dims[3] = { (cuuint64_t)L.stride, (cuuint64_t)L.prows, (cuuint64_t)nframes };
        cuuint64_t strides[2] = { (cuuint64_t)L.stride, (cuuint64_t)P.frame_bytes };
        cuuint32_t estr[3] = { 1, 1, 1 };
        cuuint32_t box_fast[3] = { ORB_TILE_W + 32, ORB_TILE_H + 8, 1 };
        cuuint32_t box_blur[3] = { 96, ORB_BLUR_TILE_H + 6, 1 };
        void* base = W.d_planes + L.plane_off;
        CUresult r1 = enc(&W.tm_fast.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box_fast, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r2 = enc(&W.tm_blur.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box_blur, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        for (int v = 0; v < 2 && l + 1 < P.nlevels && r1 == CUDA_SUCCESS; v++) {       // source descriptors for the resize that produces level l+1 (both tilings)
            cuuint32_t box_rs[3] = { (cuuint32_t)c->rs_box_w[v][l + 1], (cuuint32_t)c->rs_box_h[v][l + 1], 1 };
            r1 = enc(&W.tm_resize[v].m[l + 1], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box_rs, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
            g_last_cuda_error = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r1) + "," + std::to_string((int)r2) + ")";
            return ORB_ERR_CUDA;
        }
    }
    W.tm_base = W.d_planes; W.tm_frames = nframes; W.tm_w = P.w; W.tm_h = P.h;
    return ORB_OK;
}

static std::atomic<long long> g_alloc_gen{0};        // bumped by every (re)allocation: captured graphs hold raw pointers and are keyed on it
template <typename T>
static int ensure(T*& p, size_t& cap, size_t bytes)
{
    if (bytes <= cap && p) return ORB_OK;
    g_alloc_gen++;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    ORB_CUDA(cudaMalloc((void**)&p, std::max<size_t>(bytes, 256)));
    cap = std::max<size_t>(bytes, 256);
    return ORB_OK;
}

static bool is_device_ptr(const void* p)
{
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// (re)build the geometry for this image shape and make sure every device buffer can hold nimg frames
static int prepare(orb_ctx* c, int w, int h)
{
    ORB_CUDA(cudaSetDevice(c->device));
    const bool rebuild = !(c->plan_valid && c->plan.w == w && c->plan.h == h);
    if (rebuild) {
        int rc = orb_build_plan(c, w, h);
        if (rc != ORB_OK) return rc;
        if (!c->d_plan) { size_t z = 0; rc = ensure(c->d_plan, z, sizeof(Plan)); if (rc) return rc; }
        rc = ensure(c->d_cells, c->cap_cells, c->cells.size() * sizeof(CellGeom)); if (rc) return rc;
        rc = ensure(c->d_tiles_fast, c->cap_tiles_fast, c->tiles_fast.size() * sizeof(Tile)); if (rc) return rc;
        rc = ensure(c->d_tiles_blur, c->cap_tiles_blur, c->tiles_blur.size() * sizeof(Tile)); if (rc) return rc;
        rc = ensure(c->d_xtab, c->cap_xtab, c->xtab.size() * sizeof(int2)); if (rc) return rc;
        rc = ensure(c->d_ytab, c->cap_ytab, c->ytab.size() * sizeof(int2)); if (rc) return rc;
        rc = ensure(c->d_fast_coltab, c->cap_fast_coltab, c->fast_coltab.size()); if (rc) return rc;
        rc = ensure(c->d_fast_rowtab, c->cap_fast_rowtab, c->fast_rowtab.size() * sizeof(int16_t)); if (rc) return rc;
        ORB_CUDA(cudaDeviceSynchronize());       // nothing in flight may still read the old tables
        ORB_CUDA(cudaMemcpy(c->d_plan, &c->plan, sizeof(Plan), cudaMemcpyHostToDevice));
        ORB_CUDA(cudaMemcpy(c->d_cells, c->cells.data(), c->cells.size() * sizeof(CellGeom), cudaMemcpyHostToDevice));
        ORB_CUDA(cudaMemcpy(c->d_tiles_fast, c->tiles_fast.data(), c->tiles_fast.size() * sizeof(Tile), cudaMemcpyHostToDevice));
        ORB_CUDA(cudaMemcpy(c->d_tiles_blur, c->tiles_blur.data(), c->tiles_blur.size() * sizeof(Tile), cudaMemcpyHostToDevice));
        if (!c->xtab.empty()) ORB_CUDA(cudaMemcpy(c->d_xtab, c->xtab.data(), c->xtab.size() * sizeof(int2), cudaMemcpyHostToDevice));
        if (!c->ytab.empty()) ORB_CUDA(cudaMemcpy(c->d_ytab, c->ytab.data(), c->ytab.size() * sizeof(int2), cudaMemcpyHostToDevice));
        ORB_CUDA(cudaMemcpy(c->d_fast_coltab, c->fast_coltab.data(), c->fast_coltab.size(), cudaMemcpyHostToDevice));
        ORB_CUDA(cudaMemcpy(c->d_fast_rowtab, c->fast_rowtab.data(), c->fast_rowtab.size() * sizeof(int16_t), cudaMemcpyHostToDevice));
        int maxcap = 0;
        for (int l = 0; l < c->plan.nlevels; l++) maxcap = std::max(maxcap, c->plan.L[l].lvl_cap);
        if ((size_t)maxcap * 8 + (size_t)c->plan.sel_cells_cap * 14 > 170 * 1024) return ORB_ERR_CAPACITY;
        rc = orb_select_smem_setup(maxcap, c->plan.sel_cells_cap); if (rc) return rc;
        size_t rsm = 1024;
        for (int v = 0; v < 2; v++)
            for (int l = 1; l < c->plan.nlevels; l++) rsm = std::max(rsm, (size_t)2 * (((size_t)c->rs_box_w[v][l] * c->rs_box_h[v][l] + 127) & ~(size_t)127) + 16);
        rc = orb_resize_smem_setup((int)rsm); if (rc) return rc;
        c->plan_valid = true;
        c->plan_gen++;
    }
    for (WorkSet& W : c->ws) W.tm_w = rebuild ? 0 : W.tm_w;      // a new plan invalidates the descriptors
    return ORB_OK;
}

// make sure work set W can hold nimg frames of the current plan
static int prepare_ws(orb_ctx* c, WorkSet& W, int nimg)
{
    const Plan& P = c->plan;
    const size_t B = (size_t)std::min(std::max(nimg, 1), c->max_batch);
    int rc;
    rc = ensure(W.d_planes, W.planes_bytes, B * P.frame_bytes); if (rc) return rc;
    rc = ensure(W.d_work, W.work_bytes, B * P.frame_bytes); if (rc) return rc;
    rc = ensure(W.d_blur, W.blur_bytes, B * P.frame_bytes); if (rc) return rc;
    rc = ensure(W.d_bitmap, W.bitmap_bytes, B * (size_t)std::max(P.bm_total, 256)); if (rc) return rc;
    rc = ensure(W.d_cand, W.cand_bytes, B * (size_t)P.cand_total * 4); if (rc) return rc;
    if (P.harris) { rc = ensure(W.d_cand64, W.cand64_bytes, B * (size_t)P.cand_total * 8); if (rc) return rc; }
    rc = ensure(W.d_ntotal, W.ntotal_bytes, B * (size_t)P.ncells * 4); if (rc) return rc;
    rc = ensure(W.d_lvl, W.lvl_bytes, B * (size_t)P.lvl_total * 8); if (rc) return rc;
    rc = ensure(W.d_nkept, W.nkept_bytes, B * ORB_MAX_LEVELS * sizeof(int)); if (rc) return rc;
    rc = ensure(W.d_counters, W.counters_bytes, (32 + B * ORB_MAX_LEVELS) * sizeof(int)); if (rc) return rc;     // tile queues + k_pyramid's (frame, level) counters
    if (!W.aux_stream) {
        ORB_CUDA(cudaStreamCreateWithFlags(&W.aux_stream, cudaStreamNonBlocking));
        ORB_CUDA(cudaEventCreateWithFlags(&W.ev_fork, cudaEventDisableTiming));
        ORB_CUDA(cudaEventCreateWithFlags(&W.ev_join, cudaEventDisableTiming));
    }
    // descriptors span the whole allocation (capacity in frames), so they survive smaller batches
    return orb_build_tmaps(c, W, (int)(W.planes_bytes / P.frame_bytes));
}

// One extraction pass = 1 memset + 14 kernels on two streams.  For small batches (the tracking thread's one frame per call) the host
// cost of those launches is most of the latency, so a pass whose shape AND buffers repeat is captured into a CUDA graph the second
// time it comes by and replayed afterwards.  Not used on the legacy default stream (capture is not allowed there), in profiling mode,
// or for large batches, where launch cost is noise.
static int launch_extract(orb_ctx* c, WorkSet& W, const uint8_t* d_in, int n, int w, int h, int stride, size_t pitch,
                          orb_keypoint* o_k, uint8_t* o_d, int cap, int32_t* o_c, cudaStream_t s)
{
    // Calls of at most pdl_frames frames also carry programmatic dependent launch edges (orb_extract.cu, launch_k), in the captured graph
    // as well as in plain launches.  Measured on B200 for one 640x480 frame through the blocking call, microseconds, pageable / pinned
    // / device buffers: graph replay alone 136 / 123 / 113, PDL launches alone 119 / 112 / 80, both 116 / 112 / 82.  (While the pass
    // still began with a memset node the combination was SLOWER than either, 157 / 146: a programmatic edge behind a memset node
    // costs a replayed graph more than all the edges gain — the counters are now cleared by the pass's first kernel.)
    c->pdl_call = c->use_pdl && n <= c->pdl_frames && (double)n * w * h <= 1e6;     // one 1080p frame: 236 us with the edges, 220 without (parked CTAs of its larger grids)
    const bool eligible = c->use_graph && !c->profile && s != nullptr && s != cudaStreamLegacy && s != cudaStreamPerThread &&
                          (double)n * w * h <= 12e6;
    if (!eligible) return orb_launch_extract(c, W, d_in, n, w, h, stride, pitch, o_k, o_d, cap, o_c, s);
    WorkSet::GraphKey key;
    key.in = d_in; key.kps = o_k; key.desc = o_d; key.counts = o_c; key.planes = W.d_planes;
    key.n = n; key.w = w; key.h = h; key.stride = stride; key.cap = cap; key.pitch = pitch; key.plan_gen = c->plan_gen * 1000003LL + g_alloc_gen;
    if (W.graph_exec && W.graph_key == key) {
        ORB_CUDA(cudaGraphLaunch(W.graph_exec, s));
        c->last_launches = W.graph_launches;
        return ORB_OK;
    }
    const bool repeat = W.last_key == key;
    W.last_key = key;
    if (!repeat) return orb_launch_extract(c, W, d_in, n, w, h, stride, pitch, o_k, o_d, cap, o_c, s);
    if (W.graph_exec) { cudaGraphExecDestroy(W.graph_exec); W.graph_exec = nullptr; }
    ORB_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    const int rc = orb_launch_extract(c, W, d_in, n, w, h, stride, pitch, o_k, o_d, cap, o_c, s);
    cudaGraph_t g = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(s, &g);
    if (rc != ORB_OK || ce != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        if (rc != ORB_OK) return rc;
        c->use_graph = 0;                                   // capture is not possible here: fall back to plain launches for good
        return orb_launch_extract(c, W, d_in, n, w, h, stride, pitch, o_k, o_d, cap, o_c, s);
    }
    const cudaError_t ci = cudaGraphInstantiate(&W.graph_exec, g, 0);
    cudaGraphDestroy(g);
    if (ci != cudaSuccess) {
        W.graph_exec = nullptr; cudaGetLastError(); c->use_graph = 0;
        return orb_launch_extract(c, W, d_in, n, w, h, stride, pitch, o_k, o_d, cap, o_c, s);
    }
    W.graph_key = key;
    W.graph_launches = c->last_launches;
    ORB_CUDA(cudaGraphLaunch(W.graph_exec, s));
    return ORB_OK;
}

extern "C" {

const char* orb_error_string(int s)
{
    switch (s) {
    case ORB_OK: return "ok";
    case ORB_ERR_INVALID: return "invalid argument";
    case ORB_ERR_GEOMETRY: return "cell grid geometry the reference cannot process";
    case ORB_ERR_CAPACITY: return "buffer or context capacity exceeded";
    case ORB_ERR_CUDA: return "CUDA error / no usable device";
    case ORB_ERR_UNSUPPORTED: return "unsupported option";
    default: return "unknown status";
    }
}
const char* orb_last_cuda_error(void) { return g_last_cuda_error.c_str(); }
int orb_abi_version(void) { return 1; }

orb_ctx* orb_create(int device, int nfeatures, float scale_factor, int nlevels, int score_type,
                    int fast_th, int max_w, int max_h, int max_batch)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        g_last_cuda_error = "orb_create: no usable CUDA device (this library has no CPU path)";
        cudaGetLastError();
        return nullptr;
    }
    if ((score_type != ORB_FAST_SCORE && score_type != ORB_HARRIS_SCORE) || max_w < 1 || max_h < 1 || max_batch < 1 || fast_th < 0 || fast_th > 254) {     // fastTh 0 is valid for cv::FAST (k_fast_nms: a corner of strength 1 behaves like none)
        g_last_cuda_error = "orb_create: invalid argument";
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) { orb_cuda_fail(cudaGetLastError(), "cudaSetDevice"); return nullptr; }
    orb_ctx* c = new orb_ctx;
    c->device = device; c->nfeatures = nfeatures; c->scale_factor_f = scale_factor; c->nlevels = nlevels;
    c->score_type = score_type; c->fast_th = fast_th; c->max_w = max_w; c->max_h = max_h; c->max_batch = max_batch;
    cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device);
    if (const char* e = getenv("ORB_GRAPH")) c->use_graph = atoi(e);
    if (const char* e = getenv("ORB_KNN_ENGINE")) c->knn_engine = !strcmp(e, "tensor") ? ORB_KNN_TENSOR : ORB_KNN_POPC;
#ifdef ORB_DEBUG                 // stage-skipping switch of tools/exposure.sh: only in a -DORB_DEBUG build, never in the shipped library
    if (const char* e = getenv("ORB_DEBUG_SKIP")) c->debug_skip = atoi(e);
#endif
    if (const char* e = getenv("ORB_PYR_FUSED")) c->pyr_fused = atoi(e);
    if (const char* e = getenv("ORB_FAST_ETILE")) c->fast_etile = atoi(e);
    if (const char* e = getenv("ORB_RESIZE_UNROLLED")) c->rs_unrolled = atoi(e);
    if (const char* e = getenv("ORB_RESIZE_FLEX")) c->rs_flex_width = atoi(e) != 0;
    if (const char* e = getenv("ORB_RESIZE_ROWS")) c->rs_rows_pref = std::max(1, std::min(atoi(e), 32));
    if (const char* e = getenv("ORB_RESIZE_ROWS_SMALL")) c->rs_rows_small = std::max(1, std::min(atoi(e), 32));
    if (const char* e = getenv("ORB_SMALL_CALL")) c->small_call_frames = std::max(0, atoi(e));
    if (const char* e = getenv("ORB_SELECT_SERIAL")) c->select_serial = atoi(e);
    if (const char* e = getenv("ORB_SELECT_WIDE")) c->select_wide = atoi(e);
    if (const char* e = getenv("ORB_COMPACT_WIDE")) c->compact_wide = atoi(e);
    if (const char* e = getenv("ORB_FAST_WIDE")) c->fast_wide = atoi(e);
    if (const char* e = getenv("ORB_SIDE_BORDER")) c->side_border = atoi(e);
    if (const char* e = getenv("ORB_PDL")) c->use_pdl = atoi(e);
    if (const char* e = getenv("ORB_PDL_FRAMES")) c->pdl_frames = std::max(0, atoi(e));
    if (const char* e = getenv("ORB_STAGE_SMALL")) c->stage_small = atoi(e);
    if (const char* e = getenv("ORB_FORK_EARLY")) c->fork_early = atoi(e);
    if (const char* e = getenv("ORB_FAST_CTAS_FORK")) c->fast_ctas = atoi(e);
    if (const char* e = getenv("ORB_BLUR_CTAS")) c->blur_ctas = atoi(e);
    if (const char* e = getenv("ORB_SPLIT_DEVICE")) c->split_device = atoi(e);
    if (const char* e = getenv("ORB_CHAIN_CHUNKS")) c->chain_chunks = atoi(e) != 0;
    if (orb_build_tables(c) != ORB_OK || orb_upload_constants(c->umax) != ORB_OK) { delete c; return nullptr; }
    bool ok = cudaMalloc((void**)&c->d_status, 32 * sizeof(int)) == cudaSuccess &&     // [0] status, [1..] work counters
              cudaMemset(c->d_status, 0, 32 * sizeof(int)) == cudaSuccess;
    for (int i = 0; i < 2 && ok; i++) {
        ok = cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaStreamCreateWithFlags(&c->out_streams[i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->ev_out_done[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->ev_free[i], cudaEventDisableTiming) == cudaSuccess;
    }
    ok = ok && cudaEventCreateWithFlags(&c->ev_user, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->ev_half[0], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->ev_half[1], cudaEventDisableTiming) == cudaSuccess;
    if (ok) {      // private stream-ordered pool that keeps its memory across synchronisations (k_knn2 partials of the *_device calls)
        cudaMemPoolProps pp = {};
        pp.allocType = cudaMemAllocationTypePinned; pp.handleTypes = cudaMemHandleTypeNone;
        pp.location.type = cudaMemLocationTypeDevice; pp.location.id = device;
        unsigned long long keep = ~0ull;
        ok = cudaMemPoolCreate(&c->pool, &pp) == cudaSuccess &&
             cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess;
    }
    if (!ok) { orb_cuda_fail(cudaGetLastError(), "orb_create allocations"); orb_destroy(c); return nullptr; }
    return c;
}

orb_ctx* orb_default_context(void)
{
    static std::mutex mu;
    static orb_ctx* ctxs[64] = {};
    int dev = 0;
    if (const char* e = getenv("ORB_B200_DEVICE")) dev = atoi(e);
    else if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
    if (dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    // matcher-only: nothing of the extraction pipeline is allocated until an extract call arrives (there never is one here)
    if (!ctxs[dev]) ctxs[dev] = orb_create(dev, 1000, 1.2f, 8, ORB_FAST_SCORE, 20, 64, 64, 1);
    return ctxs[dev];
}

void orb_destroy(orb_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (WorkSet& W : c->ws) {
        void* wp[] = { W.d_planes, W.d_work, W.d_blur, W.d_bitmap, W.d_cand, W.d_cand64, W.d_ntotal, W.d_lvl, W.d_nkept, W.d_counters };
        for (void* p : wp) if (p) cudaFree(p);
        if (W.graph_exec) cudaGraphExecDestroy(W.graph_exec);
        if (W.aux_stream) cudaStreamDestroy(W.aux_stream);
        if (W.ev_fork) cudaEventDestroy(W.ev_fork);
        if (W.ev_join) cudaEventDestroy(W.ev_join);
    }
    for (orb_ctx::Ticket& t : c->tickets) {
        if (t.done) cudaEventDestroy(t.done);
        if (t.a) cudaEventDestroy(t.a);
        if (t.b) cudaEventDestroy(t.b);
        if (t.h_status) cudaFreeHost(t.h_status);
        if (t.h_stage) cudaFreeHost(t.h_stage);
    }
    if (c->done_stream) cudaStreamDestroy(c->done_stream);
    if (c->ev_dev_done) cudaEventDestroy(c->ev_dev_done);
    if (c->ev_user) cudaEventDestroy(c->ev_user);
    for (cudaEvent_t e : c->ev_half) if (e) cudaEventDestroy(e);
    void* ptrs[] = { c->d_plan, c->d_cells, c->d_tiles_fast, c->d_tiles_blur, c->d_xtab, c->d_ytab, c->d_fast_coltab, c->d_fast_rowtab, c->d_status, c->d_src[0], c->d_src[1], c->d_kps[0],
                     c->d_kps[1], c->d_desc[0], c->d_desc[1], c->d_counts[0], c->d_counts[1], c->d_small[0], c->d_small[1] };
    for (int i = 0; i < c->nlanes; i++) {
        MatchLane& L = c->lanes[i];
        if (L.d_scratch) cudaFree(L.d_scratch);
        if (L.h_arena) cudaFreeHost(L.h_arena);
        if (L.stream) cudaStreamDestroy(L.stream);
    }
    if (c->pool) cudaMemPoolDestroy(c->pool);
    for (void* p : ptrs) if (p) cudaFree(p);
    for (cudaEvent_t e : c->prof_events) cudaEventDestroy(e);
    for (cudaEvent_t e : c->prof_pool) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) {
        if (c->streams[i]) cudaStreamDestroy(c->streams[i]);
        if (c->out_streams[i]) cudaStreamDestroy(c->out_streams[i]);
        if (c->ev_out_done[i]) cudaEventDestroy(c->ev_out_done[i]);
        if (c->ev_free[i]) cudaEventDestroy(c->ev_free[i]);
    }
    delete c;
}

int orb_nlevels(const orb_ctx* c) { return c ? c->nlevels : 0; }

int orb_set_descriptor_fma(orb_ctx* c, int on)
{
    if (!c) return ORB_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> ex_lock(c->ex_mu);
    if ((on != 0) != (c->desc_fma != 0)) { c->desc_fma = on != 0; c->plan_valid = false; }     // the plan carries the flag to the device
    return ORB_OK;
}
float orb_scale_factor(const orb_ctx* c) { return c ? (float)c->scaleFactor : 0.f; }
int orb_keypoint_capacity(const orb_ctx* c)
{
    if (!c) return 0;
    int s = 0;
    for (int v : c->mnFeaturesPerLevel) s += v;
    return s;
}
int orb_last_launch_count(const orb_ctx* c) { return c ? c->last_launches : 0; }

int orb_extract_batch_device(orb_ctx* c, const uint8_t* d_imgs, int nimg, int w, int h, int stride, size_t frame_pitch,
                             orb_keypoint* d_kps, uint8_t* d_desc, int cap, int32_t* d_counts, void* stream)
{
    if (!c || !d_kps || !d_desc || !d_counts || cap < 1 || nimg < 0) return ORB_ERR_INVALID;
    if (nimg == 0) return ORB_OK;
    std::lock_guard<std::recursive_mutex> ex_lock(c->ex_mu);
    cudaStream_t us = (cudaStream_t)stream;
    if (!d_imgs || w <= 0 || h <= 0) {              // empty image: no keypoints (src/ORBextractor.cc:721-722)
        ORB_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(int32_t) * nimg, us));
        return ORB_OK;
    }
    if (nimg > c->max_batch) return ORB_ERR_CAPACITY;
    if (stride < w) return ORB_ERR_INVALID;
    // the work sets are shared with the host-buffer calls: anything those still have in flight finishes first
    while (c->waited_seq + 1 < c->next_seq) { const int rw = orb_wait(c, c->waited_seq + 1); if (rw != ORB_OK) return rw; }
    int rc = prepare(c, w, h);
    if (rc != ORB_OK) return rc;
    if (!c->ev_dev_done) ORB_CUDA(cudaEventCreateWithFlags(&c->ev_dev_done, cudaEventDisableTiming));
    // a context owns ONE set of pyramid / candidate buffers: a second device call, possibly on another caller stream, is ordered
    // behind the previous one instead of racing it on those buffers (same-stream callers pay one no-op wait)
    if (c->dev_call_pending) ORB_CUDA(cudaStreamWaitEvent(us, c->ev_dev_done, 0));
    struct Done { orb_ctx* c; cudaStream_t s; ~Done() { if (cudaEventRecord(c->ev_dev_done, s) == cudaSuccess) c->dev_call_pending = true; } } done{ c, us };
    // Optional (ORB_SPLIT_DEVICE=1): two halves on the two internal streams, forked from / joined to the caller's
    // stream.  Measured on B200 at 256 frames: 3.19 ms split vs 3.03 ms unsplit, so it is off by default; the
    // host-buffer path always alternates the two work sets so that copies and kernels of neighbouring chunks overlap.
    const bool split = c->split_device && !c->profile && nimg >= 16;
    const int n0 = split ? (nimg + 1) / 2 : nimg, n1 = nimg - n0;
    if ((rc = prepare_ws(c, c->ws[0], n0))) return rc;
    if (n1 && (rc = prepare_ws(c, c->ws[1], n1))) return rc;
    c->last_n0 = n0; c->last_n1 = n1;
    if (!split) return launch_extract(c, c->ws[0], d_imgs, nimg, w, h, stride, frame_pitch, d_kps, d_desc, cap, d_counts, us);
    ORB_CUDA(cudaEventRecord(c->ev_user, us));
    c->pdl_call = false;
    int launches = 0;
    for (int k = 0; k < 2; k++) {
        cudaStream_t s = c->streams[k];
        const int f0 = k ? n0 : 0, n = k ? n1 : n0;
        ORB_CUDA(cudaStreamWaitEvent(s, c->ev_user, 0));
        rc = orb_launch_extract(c, c->ws[k], d_imgs + (size_t)f0 * frame_pitch, n, w, h, stride, frame_pitch,
                                d_kps + (size_t)f0 * cap, d_desc + (size_t)f0 * cap * 32, cap, d_counts + f0, s);
        if (rc != ORB_OK) return rc;
        launches += c->last_launches;
        ORB_CUDA(cudaEventRecord(c->ev_half[k], s));
        ORB_CUDA(cudaStreamWaitEvent(us, c->ev_half[k], 0));
    }
    c->last_launches = launches;
    return ORB_OK;
}

// Completion of ticket `seq`: wait for its event, surface kernel-side errors, clamp counts like the synchronous call.
int orb_wait(orb_ctx* c, long long seq)
{
    if (!c) return ORB_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> ex_lock(c->ex_mu);
    if (seq < 0 || seq >= c->next_seq) return ORB_ERR_INVALID;
    if (seq <= c->waited_seq && seq + orb_ctx::NTICKETS < c->next_seq) return ORB_OK;     // record recycled: long complete
    orb_ctx::Ticket& t = c->tickets[seq % orb_ctx::NTICKETS];
    if (t.seq != seq) return ORB_OK;                                                       // recycled by a later call
    ORB_CUDA(cudaSetDevice(c->device));
    ORB_CUDA(cudaEventSynchronize(t.done));
    c->waited_seq = std::max(c->waited_seq, seq);
    t.seq = -1;
    const int st = *t.h_status;
    if (st != 0) {
        ORB_CUDA(cudaDeviceSynchronize());
        ORB_CUDA(cudaMemset(c->d_status, 0, sizeof(int)));
        return st;
    }
    int rc = ORB_OK;                  // every entry is clamped before the error is reported: a caller that reads counts[i] rows stays inside its slots
    if (t.staged) {                   // [keypoints n x cap | descriptors n x cap x 32 | counts n] in pinned staging -> the caller's buffers, valid rows only
        const size_t cap = (size_t)t.cap;
        const orb_keypoint* sk = reinterpret_cast<const orb_keypoint*>(t.h_stage);
        const uint8_t* sd = t.h_stage + (size_t)t.nimg * cap * sizeof(orb_keypoint);
        const int32_t* sc = reinterpret_cast<const int32_t*>(sd + (size_t)t.nimg * cap * 32);
        for (int i = 0; i < t.nimg; i++) {
            const size_t n = (size_t)std::max(0, std::min(sc[i], t.cap));
            memcpy(t.u_kps + i * cap, sk + i * cap, n * sizeof(orb_keypoint));
            memcpy(t.u_desc + i * cap * 32, sd + i * cap * 32, n * 32);
            t.counts[i] = sc[i];
        }
        t.staged = false;
    }
    if (t.host_out) for (int i = 0; i < t.nimg; i++) if (t.counts[i] > t.cap) { t.counts[i] = t.cap; rc = ORB_ERR_CAPACITY; }
    return rc;
}

// Enqueue one host- or device-buffer batch without waiting for it.  Chunks alternate between the two work sets / streams and keep
// alternating across calls, so the H2D copy of one call overlaps the kernels of the previous one.  Buffers must stay valid until
// orb_wait(ticket) returns; at most NTICKETS-1 calls may be in flight.
int orb_extract_batch_async(orb_ctx* c, const uint8_t* imgs, int nimg, int w, int h, int stride, size_t frame_pitch,
                            orb_keypoint* kps, uint8_t* desc, int cap, int32_t* counts, long long* ticket)
{
    if (!c || !kps || !desc || !counts || !ticket || cap < 1 || nimg < 0) return ORB_ERR_INVALID;
    *ticket = -1;
    std::lock_guard<std::recursive_mutex> ex_lock(c->ex_mu);
    ORB_CUDA(cudaSetDevice(c->device));
    const bool dev_in = imgs && is_device_ptr(imgs), dev_out = is_device_ptr(kps);
    if (dev_out != is_device_ptr(desc) || dev_out != is_device_ptr(counts)) return ORB_ERR_INVALID;
    const bool empty = nimg == 0 || !imgs || w <= 0 || h <= 0;
    if (!empty && (stride < w || frame_pitch < (size_t)stride * (h - 1) + w)) return ORB_ERR_INVALID;
    // the oldest record is about to be recycled: make sure its call has completed
    if (c->next_seq >= orb_ctx::NTICKETS) { const int rc = orb_wait(c, c->next_seq - orb_ctx::NTICKETS); if (rc != ORB_OK && rc != ORB_ERR_INVALID) return rc; }
    orb_ctx::Ticket& t = c->tickets[c->next_seq % orb_ctx::NTICKETS];
    if (!t.done) {
        ORB_CUDA(cudaEventCreateWithFlags(&t.done, cudaEventDisableTiming));
        ORB_CUDA(cudaEventCreateWithFlags(&t.a, cudaEventDisableTiming));
        ORB_CUDA(cudaEventCreateWithFlags(&t.b, cudaEventDisableTiming));
        ORB_CUDA(cudaMallocHost((void**)&t.h_status, sizeof(int)));
    }
    if (!c->done_stream) ORB_CUDA(cudaStreamCreateWithFlags(&c->done_stream, cudaStreamNonBlocking));
    *t.h_status = 0;
    t.staged = false;
    int launches = 0, simple_slot = -1;
    if (empty) {                                   // empty image: no keypoints (src/ORBextractor.cc:721-722)
        if (dev_out) { if (nimg) ORB_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * nimg, c->streams[0])); }
        else for (int i = 0; i < nimg; i++) counts[i] = 0;
    } else {
        const bool idle = c->waited_seq + 1 == c->next_seq;          // nothing in flight
        if (c->dev_call_pending) {                                   // a device-pointer call on a caller stream used the same work sets
            ORB_CUDA(cudaStreamWaitEvent(c->streams[0], c->ev_dev_done, 0));
            ORB_CUDA(cudaStreamWaitEvent(c->streams[1], c->ev_dev_done, 0));
            c->dev_call_pending = false;
        }
        if (idle) c->chunk_parity = 0;                               // blocking callers always start on work set 0 (debug getters rely on it)
        int rc = prepare(c, w, h);
        if (rc != ORB_OK) return rc;
        const int B = std::min(nimg, c->max_batch);
        // A call of a few frames into PAGEABLE output buffers (std::vector / cv::Mat of the reference's call site): three copies into
        // pageable memory are three blocking, driver-staged transfers after the kernels; instead the slot's results form one device
        // block that returns in ONE copy to pinned staging owned by the ticket, and orb_wait hands the valid rows to the caller.
        size_t stage_bytes = 0;
        if (c->stage_small && !dev_out && nimg <= B && nimg <= c->stage_frames) {
            cudaPointerAttributes pa;
            if (cudaPointerGetAttributes(&pa, kps) == cudaSuccess && pa.type == cudaMemoryTypeUnregistered)
                stage_bytes = (size_t)nimg * cap * (sizeof(orb_keypoint) + 32) + (size_t)nimg * sizeof(int32_t);
            else cudaGetLastError();
        }
        if (stage_bytes) {
            if (t.stage_bytes < stage_bytes) {
                if (t.h_stage) { cudaFreeHost(t.h_stage); t.h_stage = nullptr; t.stage_bytes = 0; }
                ORB_CUDA(cudaMallocHost((void**)&t.h_stage, stage_bytes));
                t.stage_bytes = stage_bytes;
            }
            t.staged = true; t.u_kps = kps; t.u_desc = desc;
        }
        if ((rc = prepare_ws(c, c->ws[0], B))) return rc;
        if ((nimg > B || !idle || c->chunk_parity) && (rc = prepare_ws(c, c->ws[1], B))) return rc;
        const size_t src_chunk = (size_t)B * frame_pitch;
        for (int i = 0; i < 2; i++) {
            if (!dev_in) { rc = ensure(c->d_src[i], c->src_bytes[i], src_chunk); if (rc) return rc; }
            if (stage_bytes) { rc = ensure(c->d_small[i], c->small_bytes[i], stage_bytes); if (rc) return rc; }
            else if (!dev_out) {
                rc = ensure(c->d_kps[i], c->kps_bytes[i], (size_t)B * cap * sizeof(orb_keypoint)); if (rc) return rc;
                rc = ensure(c->d_desc[i], c->desc_bytes[i], (size_t)B * cap * 32); if (rc) return rc;
                rc = ensure(c->d_counts[i], c->counts_bytes[i], (size_t)B * sizeof(int32_t)); if (rc) return rc;
            }
        }
        // a short first chunk keeps the only copy that cannot overlap any kernel (the first H2D of an idle pipeline) small
        const int first = (idle && nimg > B) ? std::max(B / 4, 1) : B;
        int k = 0;
        for (int f0 = 0, n = 0; f0 < nimg; f0 += n, k++) {
            n = std::min(k == 0 ? first : B, nimg - f0);
            const int slot = c->chunk_parity;
            c->chunk_parity ^= 1;
            cudaStream_t s = c->streams[slot];
            const uint8_t* d_in = imgs + (size_t)f0 * frame_pitch;
            if (!dev_in) {
                ORB_CUDA(cudaMemcpyAsync(c->d_src[slot], d_in, (size_t)(n - 1) * frame_pitch + (size_t)stride * (h - 1) + w,
                                         cudaMemcpyHostToDevice, s));
                d_in = c->d_src[slot];
            }
            orb_keypoint* o_k = dev_out ? kps + (size_t)f0 * cap : c->d_kps[slot];
            uint8_t* o_d = dev_out ? desc + (size_t)f0 * cap * 32 : c->d_desc[slot];
            int32_t* o_c = dev_out ? counts + f0 : c->d_counts[slot];
            if (stage_bytes) {            // one chunk (nimg <= B): [keypoints | descriptors | counts]
                o_k = reinterpret_cast<orb_keypoint*>(c->d_small[slot]);
                o_d = c->d_small[slot] + (size_t)nimg * cap * sizeof(orb_keypoint);
                o_c = reinterpret_cast<int32_t*>(o_d + (size_t)nimg * cap * 32);
            }
            // a slot owns its staging and work buffers; stream order alone protects their reuse two chunks later.
            // Kernels of neighbouring LARGE chunks are chained (ORB_CHAIN_CHUNKS, default on): persistent grids of two streams that
            // become co-resident only steal each other's SMs, while the copies on either side still overlap freely.  Measured on
            // B200, 256 frames 752x480 per step, streaming: chunk 256 chained 86.9 K frames/s, free 81.6 K; chunk 64 chained 67.4 K,
            // free 78.9 K (small grids do not fill the GPU and gain from overlapping), hence the size test.
            if (c->chain_chunks && c->kernels_pending && (double)n * w * h >= 40e6) ORB_CUDA(cudaStreamWaitEvent(s, c->ev_free[slot ^ 1], 0));
            // the slot's output staging buffers are still being copied out by the chunk two back (its D2H runs on the slot's OUT
            // stream, below): the kernels wait for that copy, the H2D copy above did not have to
            if (!dev_out && c->out_pending[slot]) ORB_CUDA(cudaStreamWaitEvent(s, c->ev_out_done[slot], 0));
            rc = launch_extract(c, c->ws[slot], d_in, n, w, h, stride, frame_pitch, o_k, o_d, cap, o_c, s);
            if (rc != ORB_OK) return rc;
            launches += c->last_launches;
            ORB_CUDA(cudaEventRecord(c->ev_free[slot], s));
            c->kernels_pending = true;
            if (!dev_out) {
                // D2H on a stream of its own: on the slot's stream it sat between this chunk's kernels and the NEXT H2D copy of the
                // slot, which then finished too late for its kernels to start when the other slot's end (7.1 instead of 6.4 ms per
                // 1024 frames of 640x480: round 2, 148.6 K -> see DESIGN.md)
                cudaStream_t os = c->out_streams[slot];
                ORB_CUDA(cudaStreamWaitEvent(os, c->ev_free[slot], 0));
                if (stage_bytes) ORB_CUDA(cudaMemcpyAsync(t.h_stage, c->d_small[slot], stage_bytes, cudaMemcpyDeviceToHost, os));
                else {
                ORB_CUDA(cudaMemcpyAsync(kps + (size_t)f0 * cap, o_k, (size_t)n * cap * sizeof(orb_keypoint), cudaMemcpyDeviceToHost, os));
                ORB_CUDA(cudaMemcpyAsync(desc + (size_t)f0 * cap * 32, o_d, (size_t)n * cap * 32, cudaMemcpyDeviceToHost, os));
                ORB_CUDA(cudaMemcpyAsync(counts + f0, o_c, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, os));
                }
                ORB_CUDA(cudaEventRecord(c->ev_out_done[slot], os));
                c->out_pending[slot] = true;
            }
            if (slot == 0) { c->last_n0 = n; c->last_n1 = 0; } else c->last_n1 = n;
            if (idle && nimg <= B) simple_slot = slot;       // one chunk on an idle pipeline: its own stream order already covers everything the ticket stands for
        }
    }
    if (simple_slot >= 0) {
        // the blocking single-chunk call (one frame per call in the reference): status and completion event ride on the stream that
        // carries the call's last operation instead of a third stream behind three event waits
        cudaStream_t ls = dev_out ? c->streams[simple_slot] : c->out_streams[simple_slot];
        ORB_CUDA(cudaMemcpyAsync(t.h_status, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, ls));
        ORB_CUDA(cudaEventRecord(t.done, ls));
    } else {
    // completion record on a third stream, so that the two work streams never wait for each other
    ORB_CUDA(cudaEventRecord(t.a, c->streams[0]));
    ORB_CUDA(cudaEventRecord(t.b, c->streams[1]));
    ORB_CUDA(cudaStreamWaitEvent(c->done_stream, t.a, 0));
    ORB_CUDA(cudaStreamWaitEvent(c->done_stream, t.b, 0));
    for (int i = 0; i < 2; i++) if (c->out_pending[i]) ORB_CUDA(cudaStreamWaitEvent(c->done_stream, c->ev_out_done[i], 0));     // latest D2H of either slot
    ORB_CUDA(cudaMemcpyAsync(t.h_status, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, c->done_stream));
    ORB_CUDA(cudaEventRecord(t.done, c->done_stream));
    }
    t.counts = counts; t.nimg = nimg; t.cap = cap; t.host_out = !dev_out; t.seq = c->next_seq;
    c->last_launches = launches;
    *ticket = c->next_seq++;
    return ORB_OK;
}

int orb_extract_batch(orb_ctx* c, const uint8_t* imgs, int nimg, int w, int h, int stride, size_t frame_pitch,
                      orb_keypoint* kps, uint8_t* desc, int cap, int32_t* counts)
{
    long long ticket = -1;
    const int rc = orb_extract_batch_async(c, imgs, nimg, w, h, stride, frame_pitch, kps, desc, cap, counts, &ticket);
    if (rc != ORB_OK) return rc;
    const int launches = c->last_launches;
    const int rw = orb_wait(c, ticket);
    c->last_launches = launches;
    return rw;
}

int orb_extract(orb_ctx* c, const uint8_t* img, int w, int h, int stride,
                orb_keypoint* kps, uint8_t* desc, int cap, int* n)
{
    if (!n) return ORB_ERR_INVALID;
    *n = 0;
    if (!img || w <= 0 || h <= 0) return c ? ORB_OK : ORB_ERR_INVALID;
    if (is_device_ptr(kps)) return ORB_ERR_INVALID;      // single-image call reports n on the host
    int32_t cnt = 0;
    int rc = orb_extract_batch(c, img, 1, w, h, stride, (size_t)stride * h, kps, desc, cap, &cnt);
    *n = cnt;
    return rc;
}

int orb_profile_enable(orb_ctx* c, int on)
{
    if (!c) return ORB_ERR_INVALID;
    c->profile = on != 0;
    return ORB_OK;
}
const char* orb_profile_stage_name(int i)
{
    static const char* n[ORB_NSTAGES] = { "k_level0", "k_resize(x7)", "k_fast_nms", "k_cell_compact", "k_select", "k_blur", "k_describe" };
    return (i >= 0 && i < ORB_NSTAGES) ? n[i] : "";
}
int orb_profile_read(orb_ctx* c, double* ms, int* ncalls)
{
    if (!c || !ms || !ncalls) return ORB_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> ex_lock(c->ex_mu);
    ORB_CUDA(cudaSetDevice(c->device));
    ORB_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < ORB_NSTAGES; i++) ms[i] = 0;
    const size_t per = ORB_NSTAGES + 1;
    const size_t n = c->prof_events.size() / per;
    for (size_t k = 0; k < n; k++)
        for (int i = 0; i < ORB_NSTAGES; i++) {
            float t = 0;
            ORB_CUDA(cudaEventElapsedTime(&t, c->prof_events[k * per + i], c->prof_events[k * per + i + 1]));
            ms[i] += t;
        }
    for (cudaEvent_t e : c->prof_events) c->prof_pool.push_back(e);
    c->prof_events.clear();
    *ncalls = (int)n;
    return ORB_OK;
}

int orb_debug_level_info(orb_ctx* c, int frame, int level, int32_t* info)
{
    if (!c) return ORB_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> ex_lock(c->ex_mu);
    if (!c->plan_valid || level < 0 || level >= c->plan.nlevels || frame < 0 || frame >= c->last_n0 + c->last_n1) return ORB_ERR_INVALID;
    const LevelGeom& L = c->plan.L[level];
    const WorkSet& W = c->ws[frame < c->last_n0 ? 0 : 1];
    if (frame >= c->last_n0) frame -= c->last_n0;
    int nk = 0;
    ORB_CUDA(cudaSetDevice(c->device));
    ORB_CUDA(cudaDeviceSynchronize());
    ORB_CUDA(cudaMemcpy(&nk, W.d_nkept + frame * c->plan.nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
    int v[10] = { L.w, L.h, L.stride, L.nDesired, L.cols, L.rows, L.cellW, L.cellH, L.nfCell, nk };
    memcpy(info, v, sizeof v);
    return ORB_OK;
}

int orb_debug_level_plane(orb_ctx* c, int frame, int level, int which, uint8_t* out, size_t out_bytes)
{
    if (!c) return ORB_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> ex_lock(c->ex_mu);
    if (!c->plan_valid || level < 0 || level >= c->plan.nlevels || frame < 0 || frame >= c->last_n0 + c->last_n1) return ORB_ERR_INVALID;
    const LevelGeom& L = c->plan.L[level];
    const WorkSet& W = c->ws[frame < c->last_n0 ? 0 : 1];
    if (frame >= c->last_n0) frame -= c->last_n0;
    const size_t bytes = (size_t)L.stride * L.prows;
    if (out_bytes < bytes) return ORB_ERR_CAPACITY;
    ORB_CUDA(cudaSetDevice(c->device));
    ORB_CUDA(cudaDeviceSynchronize());
    const uint8_t* src = (which ? W.d_blur : W.d_planes) + (size_t)frame * c->plan.frame_bytes + L.plane_off;
    ORB_CUDA(cudaMemcpy(out, src, bytes, cudaMemcpyDeviceToHost));
    return ORB_OK;
}

/* ------------------------------------------------------------------ matching */
int orb_descriptor_distance(const uint8_t* a, const uint8_t* b)
{
    // ORBmatcher::DescriptorDistance is a static 32-byte helper on host rows (src/ORBmatcher.cc:1794-1810)
    int d = 0;
    for (int i = 0; i < 4; i++) {
        uint64_t x, y;
        memcpy(&x, a + 8 * i, 8); memcpy(&y, b + 8 * i, 8);
        d += __builtin_popcountll(x ^ y);
    }
    return d;
}

} // extern "C"

// ---- matcher lanes (orb_internal.h): per-call stream + scratch, so that Tracking / LocalMapping / LoopClosing threads may share a context ----
LaneGuard::LaneGuard(orb_ctx* ctx) : c(ctx), lane(nullptr)
{
    std::unique_lock<std::mutex> lk(c->lane_mu);
    for (;;) {
        for (int i = 0; i < c->nlanes; i++) if (!c->lanes[i].busy) { lane = &c->lanes[i]; lane->busy = true; return; }
        if (c->nlanes < orb_ctx::MAX_LANES) {
            MatchLane& L = c->lanes[c->nlanes];
            if (cudaSetDevice(c->device) != cudaSuccess || cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking) != cudaSuccess) {
                orb_cuda_fail(cudaGetLastError(), "matcher lane stream");
                return;
            }
            c->nlanes++;
            L.busy = true; lane = &L;
            return;
        }
        c->lane_cv.wait(lk);
    }
}
LaneGuard::~LaneGuard()
{
    if (!lane) return;
    { std::lock_guard<std::mutex> lk(c->lane_mu); lane->busy = false; }
    c->lane_cv.notify_one();
}

int orb_lane_scratch(MatchLane* L, size_t bytes, size_t arena_bytes)
{
    if (arena_bytes > L->arena_bytes || !L->h_arena) {
        ORB_CUDA(cudaStreamSynchronize(L->stream));          // the lane is ours: nothing else uses its buffers
        if (L->h_arena) cudaFreeHost(L->h_arena);
        L->h_arena = nullptr; L->arena_bytes = 0;
        const size_t want = std::max<size_t>(arena_bytes, 1 << 20);
        ORB_CUDA(cudaMallocHost((void**)&L->h_arena, want));
        L->arena_bytes = want;
    }
    if (bytes <= L->scratch_bytes && L->d_scratch) return ORB_OK;
    ORB_CUDA(cudaStreamSynchronize(L->stream));
    if (L->d_scratch) cudaFree(L->d_scratch);
    L->d_scratch = nullptr; L->scratch_bytes = 0;
    const size_t want = std::max<size_t>(bytes + bytes / 4, 1 << 20);     // head room: the next slightly larger frame does not reallocate
    ORB_CUDA(cudaMalloc(&L->d_scratch, want));
    L->scratch_bytes = want;
    return ORB_OK;
}
static inline size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" {

int orb_hamming_knn2_device(orb_ctx* c, const uint8_t* d_q, int nq, const uint8_t* d_db, int64_t ndb, int npairs,
                            int32_t idx_base, int32_t* d_idx1, int32_t* d_d1, int32_t* d_d2, void* stream)
{
    if (!c || nq < 0 || ndb < 0 || npairs < 1 || !d_idx1 || !d_d1 || !d_d2) return ORB_ERR_INVALID;
    if (nq == 0) return ORB_OK;
    if (!d_q || (ndb > 0 && !d_db)) return ORB_ERR_INVALID;
    if (((uintptr_t)d_q | (uintptr_t)d_db) & 15) return ORB_ERR_INVALID;     // bulk copies need 16-byte alignment
    ORB_CUDA(cudaSetDevice(c->device));
    return orb_launch_knn2(c, d_q, nq, d_db, ndb, npairs, idx_base, d_idx1, d_d1, d_d2, (cudaStream_t)stream);
}

int orb_hamming_knn2(orb_ctx* c, const uint8_t* q, int nq, const uint8_t* db, int64_t ndb,
                     int32_t* idx1, int32_t* d1, int32_t* d2)
{
    if (!c || nq < 0 || ndb < 0 || !idx1 || !d1 || !d2) return ORB_ERR_INVALID;
    if (nq == 0) return ORB_OK;
    ORB_CUDA(cudaSetDevice(c->device));
    const bool dq = is_device_ptr(q), ddb = is_device_ptr(db), dout = is_device_ptr(idx1);
    LaneGuard lg(c);
    if (!lg.lane) return ORB_ERR_CUDA;
    cudaStream_t s = lg.lane->stream;
    const size_t qb = al256((size_t)nq * 32), dbb = al256((size_t)ndb * 32), ob = al256((size_t)nq * 4);
    int rc = orb_lane_scratch(lg.lane, (dq ? 0 : qb) + (ddb ? 0 : dbb) + (dout ? 0 : 3 * ob) + 256);
    if (rc) return rc;
    uint8_t* p = (uint8_t*)lg.lane->d_scratch;
    const uint8_t* d_q = q; const uint8_t* d_db = db;
    if (!dq) { ORB_CUDA(cudaMemcpyAsync(p, q, (size_t)nq * 32, cudaMemcpyHostToDevice, s)); d_q = p; p += qb; }
    if (!ddb) { if (ndb) ORB_CUDA(cudaMemcpyAsync(p, db, (size_t)ndb * 32, cudaMemcpyHostToDevice, s)); d_db = p; p += dbb; }
    int32_t *o0 = idx1, *o1 = d1, *o2 = d2;
    if (!dout) { o0 = (int32_t*)p; o1 = (int32_t*)(p + ob); o2 = (int32_t*)(p + 2 * ob); }
    rc = orb_hamming_knn2_device(c, d_q, nq, d_db, ndb, 1, 0, o0, o1, o2, s);
    if (rc) return rc;
    if (!dout) {
        ORB_CUDA(cudaMemcpyAsync(idx1, o0, (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
        ORB_CUDA(cudaMemcpyAsync(d1, o1, (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
        ORB_CUDA(cudaMemcpyAsync(d2, o2, (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
    }
    ORB_CUDA(cudaStreamSynchronize(s));
    return ORB_OK;
}

int orb_set_knn_engine(orb_ctx* c, int engine)
{
    if (!c || (engine != ORB_KNN_POPC && engine != ORB_KNN_TENSOR)) return ORB_ERR_INVALID;
    c->knn_engine = engine;
    return ORB_OK;
}

int orb_knn2_merge_device(orb_ctx* c, const int32_t* d_parts, int nparts, int nq,
                          int32_t* d_idx1, int32_t* d_d1, int32_t* d_d2, void* stream)
{
    if (!c || !d_parts || nparts < 1 || nq < 0) return ORB_ERR_INVALID;
    if (nq == 0) return ORB_OK;
    ORB_CUDA(cudaSetDevice(c->device));
    return orb_launch_knn2_merge(d_parts, nparts, nq, d_idx1, d_d1, d_d2, (cudaStream_t)stream);
}

int orb_match_ratio(orb_ctx* c, const int32_t* idx1, const int32_t* d1, const int32_t* d2, int nq,
                    float nnratio, int th, int32_t* match, int* nmatches)
{
    if (!c || nq < 0 || !match || !nmatches) return ORB_ERR_INVALID;
    *nmatches = 0;
    if (nq == 0) return ORB_OK;
    ORB_CUDA(cudaSetDevice(c->device));
    const bool din = is_device_ptr(idx1), dout = is_device_ptr(match);
    LaneGuard lg(c);
    if (!lg.lane) return ORB_ERR_CUDA;
    cudaStream_t s = lg.lane->stream;
    const size_t ob = al256((size_t)nq * 4);
    int rc = orb_lane_scratch(lg.lane, 4 * ob + 256);
    if (rc) return rc;
    uint8_t* p = (uint8_t*)lg.lane->d_scratch;
    int* d_cnt = (int*)p; p += 256;
    const int32_t *i0 = idx1, *i1 = d1, *i2 = d2;
    if (!din) {
        ORB_CUDA(cudaMemcpyAsync(p, idx1, (size_t)nq * 4, cudaMemcpyHostToDevice, s));
        ORB_CUDA(cudaMemcpyAsync(p + ob, d1, (size_t)nq * 4, cudaMemcpyHostToDevice, s));
        ORB_CUDA(cudaMemcpyAsync(p + 2 * ob, d2, (size_t)nq * 4, cudaMemcpyHostToDevice, s));
        i0 = (int32_t*)p; i1 = (int32_t*)(p + ob); i2 = (int32_t*)(p + 2 * ob);
    }
    int32_t* om = dout ? match : (int32_t*)(p + 3 * ob);
    rc = orb_launch_match_ratio(i0, i1, i2, nq, nnratio, th, om, d_cnt, s);
    if (rc) return rc;
    if (!dout) ORB_CUDA(cudaMemcpyAsync(match, om, (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
    ORB_CUDA(cudaMemcpyAsync(nmatches, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, s));
    ORB_CUDA(cudaStreamSynchronize(s));
    return ORB_OK;
}

int orb_frame_grid_build(orb_ctx* c, const orb_keypoint* kps, int n, int min_x, int max_x, int min_y, int max_y,
                         int32_t* cell_start, int32_t* cell_items)
{
    if (!c || n < 0 || !cell_start || (n > 0 && (!kps || !cell_items)) || max_x <= min_x || max_y <= min_y) return ORB_ERR_INVALID;
    ORB_CUDA(cudaSetDevice(c->device));
    const int NC = ORB_GRID_COLS * ORB_GRID_ROWS + 1;
    LaneGuard lg(c);
    if (!lg.lane) return ORB_ERR_CUDA;
    cudaStream_t s = lg.lane->stream;
    if (is_device_ptr(cell_start)) {
        int rc = orb_launch_grid_build(kps, n, min_x, max_x, min_y, max_y, cell_start, cell_items, s);
        if (rc) return rc;
        ORB_CUDA(cudaStreamSynchronize(s));
        return ORB_OK;
    }
    const size_t kb = al256((size_t)std::max(n, 1) * sizeof(orb_keypoint)), sb = al256((size_t)NC * 4), ib = al256((size_t)std::max(n, 1) * 4);
    int rc = orb_lane_scratch(lg.lane, kb + sb + ib);
    if (rc) return rc;
    uint8_t* p = (uint8_t*)lg.lane->d_scratch;
    if (n) ORB_CUDA(cudaMemcpyAsync(p, kps, (size_t)n * sizeof(orb_keypoint), cudaMemcpyHostToDevice, s));
    rc = orb_launch_grid_build((orb_keypoint*)p, n, min_x, max_x, min_y, max_y, (int32_t*)(p + kb), (int32_t*)(p + kb + sb), s);
    if (rc) return rc;
    ORB_CUDA(cudaMemcpyAsync(cell_start, p + kb, (size_t)NC * 4, cudaMemcpyDeviceToHost, s));
    if (n) ORB_CUDA(cudaMemcpyAsync(cell_items, p + kb + sb, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    ORB_CUDA(cudaStreamSynchronize(s));
    return ORB_OK;
}

} // extern "C"

// bump allocator over the matcher scratch buffer; host arrays are uploaded, device arrays passed through
// Host inputs are first gathered in a pinned arena that mirrors the scratch layout and then go up in ONE copy.
struct Bump {
    uint8_t* base; uint8_t* hbase; size_t off = 0, cap, staged = 0;
    Bump(void* b, void* h, size_t c) : base((uint8_t*)b), hbase((uint8_t*)h), cap(c) {}
    void* take(size_t bytes) { void* p = base + off; off += al256(std::max<size_t>(bytes, 1)); return p; }
    int flush(cudaStream_t s)
    {
        if (staged) ORB_CUDA(cudaMemcpyAsync(base, hbase, staged, cudaMemcpyHostToDevice, s));
        return ORB_OK;
    }
};
template <typename T>
static int stage_in(Bump& b, bool dev, const T*& p, size_t count, cudaStream_t s)
{
    (void)s;
    if (dev || !p) return ORB_OK;
    const size_t o = b.off;
    T* d = (T*)b.take(count * sizeof(T));
    if (count) memcpy(b.hbase + o, p, count * sizeof(T));
    b.staged = b.off;
    p = d;
    return ORB_OK;
}
static size_t frame_view_bytes(const orb_frame_view* f)
{
    const size_t n = (size_t)std::max(f->n, 1);
    return al256(n * sizeof(orb_keypoint)) + al256(n * 32) + al256((ORB_GRID_COLS * ORB_GRID_ROWS + 1) * 4) + al256(n * 4);
}
static int stage_frame(Bump& b, bool dev, orb_frame_view& f, bool need_grid, cudaStream_t s)
{
    int rc;
    if ((rc = stage_in(b, dev, f.kps, (size_t)f.n, s))) return rc;
    if ((rc = stage_in(b, dev, f.desc, (size_t)f.n * 32, s))) return rc;
    if (need_grid) {
        if ((rc = stage_in(b, dev, f.cell_start, (size_t)ORB_GRID_COLS * ORB_GRID_ROWS + 1, s))) return rc;
        if ((rc = stage_in(b, dev, f.cell_items, (size_t)f.n, s))) return rc;
    }
    return ORB_OK;
}

extern "C" {

int orb_search_by_projection(orb_ctx* c, const orb_frame_view* cur, const orb_frame_view* last,
                             const uint8_t* last_has_mp, const uint8_t* last_outlier, const float* last_xyz,
                             const float* Tcw16, float th, int check_ori, int32_t* match_cur, int* nmatches)
{
    if (!c || !cur || !last || !Tcw16 || !nmatches || cur->n < 0 || last->n < 0) return ORB_ERR_INVALID;
    *nmatches = 0;
    if (cur->n == 0 || last->n == 0) return ORB_OK;
    if (!match_cur || !cur->kps || !cur->desc || !cur->cell_start || !cur->cell_items || !last->kps || !last->desc ||
        !last_has_mp || !last_outlier || !last_xyz || cur->max_x <= cur->min_x || cur->max_y <= cur->min_y) return ORB_ERR_INVALID;
    ORB_CUDA(cudaSetDevice(c->device));
    const bool dev = is_device_ptr(cur->kps);
    if (is_device_ptr(match_cur) != dev || is_device_ptr(last->kps) != dev) return ORB_ERR_INVALID;
    float T[16];
    if (is_device_ptr(Tcw16)) ORB_CUDA(cudaMemcpy(T, Tcw16, sizeof T, cudaMemcpyDeviceToHost)); else memcpy(T, Tcw16, sizeof T);
    LaneGuard lg(c);
    if (!lg.lane) return ORB_ERR_CUDA;
    cudaStream_t s = lg.lane->stream;
    const size_t work = orb_sbp_scratch_bytes(cur->n, last->n);
    const size_t in_bytes = dev ? 0 : frame_view_bytes(cur) + frame_view_bytes(last) + 2 * al256(last->n) + al256((size_t)last->n * 12) + al256((size_t)cur->n * 4);
    int rc = orb_lane_scratch(lg.lane, 256 + in_bytes + work, 256 + in_bytes + 4096);
    if (rc) return rc;
    Bump b(lg.lane->d_scratch, lg.lane->h_arena, lg.lane->scratch_bytes);
    int* d_result = (int*)b.take(8);
    orb_frame_view dc = *cur, dl = *last;
    if ((rc = stage_frame(b, dev, dc, true, s)) || (rc = stage_frame(b, dev, dl, false, s))) return rc;
    if ((rc = stage_in(b, dev, last_has_mp, (size_t)last->n, s)) || (rc = stage_in(b, dev, last_outlier, (size_t)last->n, s)) ||
        (rc = stage_in(b, dev, last_xyz, (size_t)last->n * 3, s))) return rc;
    int32_t* d_match = match_cur;
    if (!dev) { const int32_t* m = match_cur; if ((rc = stage_in(b, false, m, (size_t)cur->n, s))) return rc; d_match = (int32_t*)m; }
    uint8_t* wk = (uint8_t*)b.take(work);
    if ((rc = b.flush(s))) return rc;
    rc = orb_launch_search_by_projection(c, &dc, &dl, last_has_mp, last_outlier, last_xyz, T, th, check_ori, d_match, d_result, wk, work, s);
    if (rc) return rc;
    int res[2] = { 0, 0 };
    if (!dev) ORB_CUDA(cudaMemcpyAsync(match_cur, d_match, (size_t)cur->n * 4, cudaMemcpyDeviceToHost, s));
    ORB_CUDA(cudaMemcpyAsync(res, d_result, sizeof res, cudaMemcpyDeviceToHost, s));
    ORB_CUDA(cudaStreamSynchronize(s));
    *nmatches = res[0];
    return res[1] ? ORB_ERR_CAPACITY : ORB_OK;
}

int orb_search_window(orb_ctx* c, const orb_frame_view* target, const orb_window_query_set* q, int accept_mode,
                      float nnratio, int th_dist, int check_ori, int32_t* match_target, int* nmatches)
{
    if (!c || !target || !q || !nmatches || target->n < 0 || q->n < 0 || accept_mode < 0 || accept_mode > 2) return ORB_ERR_INVALID;
    *nmatches = 0;
    if (target->n == 0 || q->n == 0) return ORB_OK;
    const bool project = q->xyz != nullptr && q->u == nullptr;
    if (!match_target || !target->kps || !target->desc || !target->cell_start || !target->cell_items || !q->active || !q->desc ||
        !q->min_level || !q->max_level || (!project && (!q->u || !q->v)) || (project && !q->Tcw16) ||
        target->max_x <= target->min_x || target->max_y <= target->min_y) return ORB_ERR_INVALID;
    ORB_CUDA(cudaSetDevice(c->device));
    const bool dev = is_device_ptr(target->kps);
    if (is_device_ptr(match_target) != dev || is_device_ptr(q->desc) != dev) return ORB_ERR_INVALID;
    LaneGuard lg(c);
    if (!lg.lane) return ORB_ERR_CUDA;
    cudaStream_t s = lg.lane->stream;
    const size_t work = orb_sbp_scratch_bytes(target->n, q->n);
    const size_t nq = (size_t)q->n;
    const size_t in_bytes = dev ? 0 : frame_view_bytes(target) + al256(nq) + al256(nq * 32) + 7 * al256(nq * 4) + al256(nq * 12) + al256((size_t)target->n * 4);
    int rc = orb_lane_scratch(lg.lane, 256 + in_bytes + work, 256 + in_bytes + 4096);
    if (rc) return rc;
    Bump b(lg.lane->d_scratch, lg.lane->h_arena, lg.lane->scratch_bytes);
    int* d_result = (int*)b.take(8);
    orb_frame_view dt = *target;
    orb_window_query_set dq = *q;
    if ((rc = stage_frame(b, dev, dt, true, s))) return rc;
    if ((rc = stage_in(b, dev, dq.active, nq, s)) || (rc = stage_in(b, dev, dq.desc, nq * 32, s)) ||
        (rc = stage_in(b, dev, dq.u, nq, s)) || (rc = stage_in(b, dev, dq.v, nq, s)) || (rc = stage_in(b, dev, dq.xyz, nq * 3, s)) ||
        (rc = stage_in(b, dev, dq.radius, nq, s)) || (rc = stage_in(b, dev, dq.min_level, nq, s)) ||
        (rc = stage_in(b, dev, dq.max_level, nq, s)) || (rc = stage_in(b, dev, dq.angle, nq, s))) return rc;
    float T[16] = { 0 };
    if (project) { if (is_device_ptr(q->Tcw16)) ORB_CUDA(cudaMemcpy(T, q->Tcw16, sizeof T, cudaMemcpyDeviceToHost)); else memcpy(T, q->Tcw16, sizeof T); }
    dq.Tcw16 = T;
    int32_t* d_match = match_target;
    if (!dev) { const int32_t* m = match_target; if ((rc = stage_in(b, false, m, (size_t)target->n, s))) return rc; d_match = (int32_t*)m; }
    uint8_t* wk = (uint8_t*)b.take(work);
    if ((rc = b.flush(s))) return rc;
    rc = orb_launch_search_window(c, &dt, &dq, accept_mode, nnratio, th_dist, check_ori, d_match, d_result, wk, work, s);
    if (rc) return rc;
    int res[2] = { 0, 0 };
    if (!dev) ORB_CUDA(cudaMemcpyAsync(match_target, d_match, (size_t)target->n * 4, cudaMemcpyDeviceToHost, s));
    ORB_CUDA(cudaMemcpyAsync(res, d_result, sizeof res, cudaMemcpyDeviceToHost, s));
    ORB_CUDA(cudaStreamSynchronize(s));
    *nmatches = res[0];
    return res[1] ? ORB_ERR_CAPACITY : ORB_OK;
}

int orb_search_for_initialization(orb_ctx* c, const orb_frame_view* f1, const orb_frame_view* f2, float* prev_matched, int window_size,
                                  float nnratio, int check_ori, int32_t* matches12, int* nmatches)
{
    if (!c || !f1 || !f2 || !nmatches || f1->n < 0 || f2->n < 0) return ORB_ERR_INVALID;
    *nmatches = 0;
    if (f1->n == 0) return ORB_OK;
    if (!matches12 || !prev_matched || !f1->kps || !f1->desc) return ORB_ERR_INVALID;
    ORB_CUDA(cudaSetDevice(c->device));
    const bool dev = is_device_ptr(f1->kps);
    if (is_device_ptr(matches12) != dev || is_device_ptr(prev_matched) != dev) return ORB_ERR_INVALID;
    LaneGuard lg(c);
    if (!lg.lane) return ORB_ERR_CUDA;
    cudaStream_t s = lg.lane->stream;
    if (f2->n == 0) {                                       // no candidates anywhere: vnMatches12 = -1 (:601), prev untouched
        if (dev) ORB_CUDA(cudaMemsetAsync(matches12, 0xff, (size_t)f1->n * 4, s)); else for (int i = 0; i < f1->n; i++) matches12[i] = -1;
        if (dev) ORB_CUDA(cudaStreamSynchronize(s));
        return ORB_OK;
    }
    if (!f2->kps || !f2->desc || !f2->cell_start || !f2->cell_items || is_device_ptr(f2->kps) != dev ||
        f2->max_x <= f2->min_x || f2->max_y <= f2->min_y) return ORB_ERR_INVALID;
    const size_t n1 = (size_t)f1->n;
    const size_t work = orb_init_scratch_bytes(f1->n, f2->n);
    const size_t in_bytes = dev ? 0 : frame_view_bytes(f2) + al256(n1 * 28) + al256(n1 * 32) + al256(n1 * 8) + al256(n1 * 4);
    int rc = orb_lane_scratch(lg.lane, 256 + in_bytes + work, 256 + in_bytes + 4096);
    if (rc) return rc;
    Bump b(lg.lane->d_scratch, lg.lane->h_arena, lg.lane->scratch_bytes);
    int* d_result = (int*)b.take(8);
    orb_frame_view d1 = *f1, d2 = *f2;
    if ((rc = stage_frame(b, dev, d2, true, s))) return rc;
    if ((rc = stage_in(b, dev, d1.kps, n1, s)) || (rc = stage_in(b, dev, d1.desc, n1 * 32, s))) return rc;
    float* d_prev = prev_matched;
    int32_t* d_m12 = matches12;
    if (!dev) {
        const float* p = prev_matched;
        if ((rc = stage_in(b, false, p, n1 * 2, s))) return rc;
        d_prev = (float*)p;
        d_m12 = (int32_t*)b.take(n1 * 4);
    }
    uint8_t* wk = (uint8_t*)b.take(work);
    if ((rc = b.flush(s))) return rc;
    rc = orb_launch_search_for_initialization(c, &d1, &d2, d_prev, window_size, nnratio, check_ori, d_m12, d_result, wk, work, s);
    if (rc) return rc;
    int res[2] = { 0, 0 };
    if (!dev) {
        ORB_CUDA(cudaMemcpyAsync(matches12, d_m12, n1 * 4, cudaMemcpyDeviceToHost, s));
        ORB_CUDA(cudaMemcpyAsync(prev_matched, d_prev, n1 * 8, cudaMemcpyDeviceToHost, s));
    }
    ORB_CUDA(cudaMemcpyAsync(res, d_result, sizeof res, cudaMemcpyDeviceToHost, s));
    ORB_CUDA(cudaStreamSynchronize(s));
    *nmatches = res[0];
    return res[1] ? ORB_ERR_CAPACITY : ORB_OK;
}

int orb_search_window_best(orb_ctx* c, const orb_frame_view* target, const orb_window_query_set* q, int32_t* best_idx, int32_t* best_dist)
{
    if (!c || !target || !q || target->n < 0 || q->n < 0) return ORB_ERR_INVALID;
    if (q->n == 0) return ORB_OK;
    if (!best_idx || !best_dist) return ORB_ERR_INVALID;
    ORB_CUDA(cudaSetDevice(c->device));
    const bool dev = is_device_ptr(best_idx);
    if (is_device_ptr(best_dist) != dev) return ORB_ERR_INVALID;
    LaneGuard lg(c);
    if (!lg.lane) return ORB_ERR_CUDA;
    cudaStream_t s = lg.lane->stream;
    if (target->n == 0) {                                   // no keypoints: nothing is ever in the radius
        if (dev) { ORB_CUDA(cudaMemsetAsync(best_idx, 0xff, (size_t)q->n * 4, s)); ORB_CUDA(cudaMemsetAsync(best_dist, 0x7f, (size_t)q->n * 4, s)); ORB_CUDA(cudaStreamSynchronize(s)); }
        else for (int i = 0; i < q->n; i++) { best_idx[i] = -1; best_dist[i] = INT_MAX; }
        return ORB_OK;
    }
    const bool project = q->xyz != nullptr && q->u == nullptr;
    if (!target->kps || !target->desc || !target->cell_start || !target->cell_items || !q->active || !q->desc ||
        !q->min_level || !q->max_level || (!project && (!q->u || !q->v)) || (project && !q->Tcw16) ||
        target->max_x <= target->min_x || target->max_y <= target->min_y) return ORB_ERR_INVALID;
    if (is_device_ptr(target->kps) != dev || is_device_ptr(q->desc) != dev) return ORB_ERR_INVALID;
    const size_t work = orb_sbp_scratch_bytes(target->n, q->n);
    const size_t nq = (size_t)q->n;
    const size_t in_bytes = dev ? 0 : frame_view_bytes(target) + al256(nq) + al256(nq * 32) + 7 * al256(nq * 4) + al256(nq * 12) + 2 * al256(nq * 4);
    int rc = orb_lane_scratch(lg.lane, 256 + in_bytes + work, 256 + in_bytes + 4096);
    if (rc) return rc;
    Bump b(lg.lane->d_scratch, lg.lane->h_arena, lg.lane->scratch_bytes);
    int* d_result = (int*)b.take(8);
    orb_frame_view dt = *target;
    orb_window_query_set dq = *q;
    if ((rc = stage_frame(b, dev, dt, true, s))) return rc;
    if ((rc = stage_in(b, dev, dq.active, nq, s)) || (rc = stage_in(b, dev, dq.desc, nq * 32, s)) ||
        (rc = stage_in(b, dev, dq.u, nq, s)) || (rc = stage_in(b, dev, dq.v, nq, s)) || (rc = stage_in(b, dev, dq.xyz, nq * 3, s)) ||
        (rc = stage_in(b, dev, dq.radius, nq, s)) || (rc = stage_in(b, dev, dq.min_level, nq, s)) ||
        (rc = stage_in(b, dev, dq.max_level, nq, s))) return rc;
    dq.angle = nullptr;
    float T[16] = { 0 };
    if (project) { if (is_device_ptr(q->Tcw16)) ORB_CUDA(cudaMemcpy(T, q->Tcw16, sizeof T, cudaMemcpyDeviceToHost)); else memcpy(T, q->Tcw16, sizeof T); }
    dq.Tcw16 = T;
    int32_t *d_bi = best_idx, *d_bd = best_dist;
    if (!dev) { d_bi = (int32_t*)b.take(nq * 4); d_bd = (int32_t*)b.take(nq * 4); }
    uint8_t* wk = (uint8_t*)b.take(work);
    if ((rc = b.flush(s))) return rc;
    if ((rc = orb_launch_search_window_best(c, &dt, &dq, d_bi, d_bd, d_result, wk, work, s))) return rc;
    int res[2] = { 0, 0 };
    if (!dev) {
        ORB_CUDA(cudaMemcpyAsync(best_idx, d_bi, nq * 4, cudaMemcpyDeviceToHost, s));
        ORB_CUDA(cudaMemcpyAsync(best_dist, d_bd, nq * 4, cudaMemcpyDeviceToHost, s));
    }
    ORB_CUDA(cudaMemcpyAsync(res, d_result, sizeof res, cudaMemcpyDeviceToHost, s));
    ORB_CUDA(cudaStreamSynchronize(s));
    return res[1] ? ORB_ERR_CAPACITY : ORB_OK;
}

int orb_distinctive_descriptors(orb_ctx* c, const uint8_t* desc, const int32_t* start, int npoints, int32_t* best_idx, int32_t* best_median)
{
    if (!c || npoints < 0) return ORB_ERR_INVALID;
    if (npoints == 0) return ORB_OK;
    if (!start || !best_idx || !best_median) return ORB_ERR_INVALID;
    ORB_CUDA(cudaSetDevice(c->device));
    LaneGuard lg(c);
    if (!lg.lane) return ORB_ERR_CUDA;
    cudaStream_t s = lg.lane->stream;
    const bool dev = is_device_ptr(start);
    if (is_device_ptr(best_idx) != dev || is_device_ptr(best_median) != dev) return ORB_ERR_INVALID;
    int total = 0;
    if (dev) ORB_CUDA(cudaMemcpy(&total, start + npoints, 4, cudaMemcpyDeviceToHost)); else total = start[npoints];
    if (total < 0 || (total > 0 && (!desc || is_device_ptr(desc) != dev))) return ORB_ERR_INVALID;
    if (dev) {
        const int rc = orb_launch_distinctive(desc, start, npoints, best_idx, best_median, s);
        if (rc) return rc;
        ORB_CUDA(cudaStreamSynchronize(s));
        return ORB_OK;
    }
    for (int p = 0; p < npoints; p++) if (start[p + 1] < start[p]) return ORB_ERR_INVALID;
    const size_t P = (size_t)npoints, in_bytes = al256((size_t)total * 32 + 32) + al256((P + 1) * 4);
    int rc = orb_lane_scratch(lg.lane, 256 + in_bytes + 2 * al256(P * 4), 256 + in_bytes);
    if (rc) return rc;
    Bump b(lg.lane->d_scratch, lg.lane->h_arena, lg.lane->scratch_bytes);
    const uint8_t* d_desc = desc; const int32_t* d_start = start;
    if ((rc = stage_in(b, false, d_desc, (size_t)total * 32, s)) || (rc = stage_in(b, false, d_start, P + 1, s))) return rc;
    int32_t* d_bi = (int32_t*)b.take(P * 4);
    int32_t* d_bm = (int32_t*)b.take(P * 4);
    if ((rc = b.flush(s))) return rc;
    if ((rc = orb_launch_distinctive(total ? d_desc : nullptr, d_start, npoints, d_bi, d_bm, s))) return rc;
    ORB_CUDA(cudaMemcpyAsync(best_idx, d_bi, P * 4, cudaMemcpyDeviceToHost, s));
    ORB_CUDA(cudaMemcpyAsync(best_median, d_bm, P * 4, cudaMemcpyDeviceToHost, s));
    ORB_CUDA(cudaStreamSynchronize(s));
    return ORB_OK;
}

} // extern "C"

// shared body of SearchByBoW(KF, Frame) and SearchByBoW(KF, KF): the latter adds the validity of the second side's map
// points, the strict TH_LOW test and an output indexed by the first keyframe's features (match12)
static int search_by_bow_impl(orb_ctx* c, const orb_featvec_view* kf_fv, const uint8_t* kf_desc, const orb_keypoint* kf_kps,
                              const uint8_t* kf_mp_valid, int n_kf,
                              const orb_featvec_view* f_fv, const uint8_t* f_desc, const orb_keypoint* f_kps, int n_f,
                              float nnratio, int check_ori, int32_t* match_f, int* nmatches, const uint8_t* f_valid, int32_t* match12)
{
    if (!c || !kf_fv || !f_fv || !nmatches || n_kf < 0 || n_f < 0 || kf_fv->nnodes < 0 || f_fv->nnodes < 0) return ORB_ERR_INVALID;
    *nmatches = 0;
    if (n_f == 0) { if (match12 && !is_device_ptr(match12)) for (int i = 0; i < n_kf; i++) match12[i] = -1; return ORB_OK; }
    if ((!match_f && !match12) || !f_desc || !f_kps || (n_kf > 0 && (!kf_desc || !kf_kps || !kf_mp_valid))) return ORB_ERR_INVALID;
    ORB_CUDA(cudaSetDevice(c->device));
    const bool dev = is_device_ptr(f_desc);
    if (is_device_ptr(match12 ? (const void*)match12 : (const void*)match_f) != dev) return ORB_ERR_INVALID;
    LaneGuard lg(c);
    if (!lg.lane) return ORB_ERR_CUDA;
    cudaStream_t s = lg.lane->stream;
    // item totals live at start[nnodes]
    int kf_total = 0, f_total = 0;
    if (kf_fv->nnodes) { if (dev) ORB_CUDA(cudaMemcpy(&kf_total, kf_fv->start + kf_fv->nnodes, 4, cudaMemcpyDeviceToHost)); else kf_total = kf_fv->start[kf_fv->nnodes]; }
    if (f_fv->nnodes) { if (dev) ORB_CUDA(cudaMemcpy(&f_total, f_fv->start + f_fv->nnodes, 4, cudaMemcpyDeviceToHost)); else f_total = f_fv->start[f_fv->nnodes]; }
    const size_t work = orb_bow_scratch_bytes(n_f);
    size_t in_bytes = 0;
    if (!dev) in_bytes = al256((size_t)n_kf * 32) + al256((size_t)n_kf * 28) + al256(n_kf) + al256((size_t)n_f * 32) + al256((size_t)n_f * 28) +
                         al256((size_t)n_f * 4) + al256(n_f) + al256((size_t)std::max(n_kf, 1) * 4) + 2 * al256((size_t)(kf_fv->nnodes + 1) * 4) + al256((size_t)kf_total * 4 + 4) +
                         2 * al256((size_t)(f_fv->nnodes + 1) * 4) + al256((size_t)f_total * 4 + 4) + 4096;
    int rc = orb_lane_scratch(lg.lane, in_bytes + work + 256, in_bytes + 4096);
    if (rc) return rc;
    Bump b(lg.lane->d_scratch, lg.lane->h_arena, lg.lane->scratch_bytes);
    orb_featvec_view a = *kf_fv, f = *f_fv;
    static const int32_t zero_start[1] = { 0 };
    if (!a.start) a.start = zero_start;
    if (!f.start) f.start = zero_start;
    if ((rc = stage_in(b, dev, a.node_id, (size_t)a.nnodes, s)) || (rc = stage_in(b, dev, a.start, (size_t)a.nnodes + 1, s)) ||
        (rc = stage_in(b, dev, a.items, (size_t)kf_total, s)) || (rc = stage_in(b, dev, f.node_id, (size_t)f.nnodes, s)) ||
        (rc = stage_in(b, dev, f.start, (size_t)f.nnodes + 1, s)) || (rc = stage_in(b, dev, f.items, (size_t)f_total, s)) ||
        (rc = stage_in(b, dev, kf_desc, (size_t)n_kf * 32, s)) || (rc = stage_in(b, dev, kf_kps, (size_t)n_kf, s)) ||
        (rc = stage_in(b, dev, kf_mp_valid, (size_t)n_kf, s)) || (rc = stage_in(b, dev, f_desc, (size_t)n_f * 32, s)) ||
        (rc = stage_in(b, dev, f_kps, (size_t)n_f, s)) || (rc = stage_in(b, dev, f_valid, (size_t)n_f, s))) return rc;
    int32_t* d_match = (dev && match_f) ? match_f : (int32_t*)b.take((size_t)n_f * 4);
    int32_t* d_match12 = !match12 ? nullptr : (dev ? match12 : (int32_t*)b.take((size_t)std::max(n_kf, 1) * 4));
    uint8_t* wk = (uint8_t*)b.take(work);
    if ((rc = b.flush(s))) return rc;
    rc = orb_launch_search_by_bow(c, &a, kf_desc, kf_kps, kf_mp_valid, &f, f_desc, f_kps, n_f, f_total, nnratio, check_ori, d_match, wk, s,
                                  f_valid, d_match12, n_kf);
    if (rc) return rc;
    int res = 0;
    if (!dev && match_f) ORB_CUDA(cudaMemcpyAsync(match_f, d_match, (size_t)n_f * 4, cudaMemcpyDeviceToHost, s));
    if (!dev && match12 && n_kf) ORB_CUDA(cudaMemcpyAsync(match12, d_match12, (size_t)n_kf * 4, cudaMemcpyDeviceToHost, s));
    ORB_CUDA(cudaMemcpyAsync(&res, wk, sizeof res, cudaMemcpyDeviceToHost, s));
    ORB_CUDA(cudaStreamSynchronize(s));
    *nmatches = res;
    return ORB_OK;
}


extern "C" {

int orb_search_by_bow(orb_ctx* c, const orb_featvec_view* kf_fv, const uint8_t* kf_desc, const orb_keypoint* kf_kps,
                      const uint8_t* kf_mp_valid, int n_kf,
                      const orb_featvec_view* f_fv, const uint8_t* f_desc, const orb_keypoint* f_kps, int n_f,
                      float nnratio, int check_ori, int32_t* match_f, int* nmatches)
{
    if (!match_f) return ORB_ERR_INVALID;
    return search_by_bow_impl(c, kf_fv, kf_desc, kf_kps, kf_mp_valid, n_kf, f_fv, f_desc, f_kps, n_f, nnratio, check_ori, match_f, nmatches,
                              nullptr, nullptr);
}

int orb_search_by_bow_kf(orb_ctx* c, const orb_featvec_view* fv1, const uint8_t* desc1, const orb_keypoint* kps1, const uint8_t* valid1, int n1,
                         const orb_featvec_view* fv2, const uint8_t* desc2, const orb_keypoint* kps2, const uint8_t* valid2, int n2,
                         float nnratio, int check_ori, int32_t* match12, int* nmatches)
{
    if (!match12 || (n2 > 0 && !valid2)) return ORB_ERR_INVALID;
    return search_by_bow_impl(c, fv1, desc1, kps1, valid1, n1, fv2, desc2, kps2, n2, nnratio, check_ori, nullptr, nmatches, valid2, match12);
}

int orb_search_for_triangulation(orb_ctx* c, const orb_featvec_view* fv1, const uint8_t* desc1, const orb_keypoint* kps1,
                                 const uint8_t* has_mp1, int n1, const orb_featvec_view* fv2, const uint8_t* desc2,
                                 const orb_keypoint* kps2, const uint8_t* has_mp2, int n2, const float* F12, const float* level_sigma2,
                                 int nlevels, int check_ori, int32_t* match12, int* nmatches)
{
    if (!c || !fv1 || !fv2 || !nmatches || n1 < 0 || n2 < 0 || fv1->nnodes < 0 || fv2->nnodes < 0 || !F12 || !level_sigma2 ||
        nlevels < 1 || nlevels > ORB_MAX_LEVELS) return ORB_ERR_INVALID;
    *nmatches = 0;
    if (n1 == 0) return ORB_OK;
    if (!match12 || !desc1 || !kps1 || !has_mp1 || (n2 > 0 && (!desc2 || !kps2 || !has_mp2))) return ORB_ERR_INVALID;
    ORB_CUDA(cudaSetDevice(c->device));
    const bool dev = is_device_ptr(desc1);
    if (is_device_ptr(match12) != dev) return ORB_ERR_INVALID;
    LaneGuard lg(c);
    if (!lg.lane) return ORB_ERR_CUDA;
    cudaStream_t s = lg.lane->stream;
    int t1 = 0, t2 = 0;
    if (fv1->nnodes) { if (dev) ORB_CUDA(cudaMemcpy(&t1, fv1->start + fv1->nnodes, 4, cudaMemcpyDeviceToHost)); else t1 = fv1->start[fv1->nnodes]; }
    if (fv2->nnodes) { if (dev) ORB_CUDA(cudaMemcpy(&t2, fv2->start + fv2->nnodes, 4, cudaMemcpyDeviceToHost)); else t2 = fv2->start[fv2->nnodes]; }
    const size_t work = orb_tri_scratch_bytes(n1, n2);
    size_t in_bytes = 0;
    if (!dev) in_bytes = al256((size_t)n1 * 32) + al256((size_t)n1 * 28) + al256(n1) + al256((size_t)n2 * 32 + 32) + al256((size_t)n2 * 28 + 28) + al256(n2 + 1) +
                         2 * al256((size_t)(fv1->nnodes + 1) * 4) + al256((size_t)t1 * 4 + 4) + 2 * al256((size_t)(fv2->nnodes + 1) * 4) + al256((size_t)t2 * 4 + 4) +
                         al256((size_t)n1 * 4) + 4096;
    int rc = orb_lane_scratch(lg.lane, in_bytes + work + 256, in_bytes + 4096);
    if (rc) return rc;
    Bump b(lg.lane->d_scratch, lg.lane->h_arena, lg.lane->scratch_bytes);
    orb_featvec_view a = *fv1, f = *fv2;
    static const int32_t zero_start[1] = { 0 };
    if (!dev) { if (!a.start) a.start = zero_start; if (!f.start) f.start = zero_start; }
    if ((rc = stage_in(b, dev, a.node_id, (size_t)a.nnodes, s)) || (rc = stage_in(b, dev, a.start, (size_t)a.nnodes + 1, s)) ||
        (rc = stage_in(b, dev, a.items, (size_t)t1, s)) || (rc = stage_in(b, dev, f.node_id, (size_t)f.nnodes, s)) ||
        (rc = stage_in(b, dev, f.start, (size_t)f.nnodes + 1, s)) || (rc = stage_in(b, dev, f.items, (size_t)t2, s)) ||
        (rc = stage_in(b, dev, desc1, (size_t)n1 * 32, s)) || (rc = stage_in(b, dev, kps1, (size_t)n1, s)) ||
        (rc = stage_in(b, dev, has_mp1, (size_t)n1, s)) || (rc = stage_in(b, dev, desc2, (size_t)n2 * 32, s)) ||
        (rc = stage_in(b, dev, kps2, (size_t)n2, s)) || (rc = stage_in(b, dev, has_mp2, (size_t)n2, s))) return rc;
    int32_t* d_m12 = dev ? match12 : (int32_t*)b.take((size_t)n1 * 4);
    uint8_t* wk = (uint8_t*)b.take(work);
    if ((rc = b.flush(s))) return rc;
    float F[9], sg[ORB_MAX_LEVELS];
    if (is_device_ptr(F12)) ORB_CUDA(cudaMemcpy(F, F12, sizeof F, cudaMemcpyDeviceToHost)); else memcpy(F, F12, sizeof F);
    if (is_device_ptr(level_sigma2)) ORB_CUDA(cudaMemcpy(sg, level_sigma2, (size_t)nlevels * 4, cudaMemcpyDeviceToHost)); else memcpy(sg, level_sigma2, (size_t)nlevels * 4);
    rc = orb_launch_search_for_triangulation(&a, desc1, kps1, has_mp1, n1, &f, desc2, kps2, has_mp2, n2, t2, F, sg, nlevels, check_ori, d_m12, wk, s);
    if (rc) return rc;
    int res = 0;
    if (!dev) ORB_CUDA(cudaMemcpyAsync(match12, d_m12, (size_t)n1 * 4, cudaMemcpyDeviceToHost, s));
    ORB_CUDA(cudaMemcpyAsync(&res, wk, sizeof res, cudaMemcpyDeviceToHost, s));
    ORB_CUDA(cudaStreamSynchronize(s));
    *nmatches = res;
    return ORB_OK;
}

int orb_measure_popc_peak(orb_ctx* c, double* gpopc_per_s)
{
    if (!c || !gpopc_per_s) return ORB_ERR_INVALID;
    ORB_CUDA(cudaSetDevice(c->device));
    LaneGuard lg(c);
    if (!lg.lane) return ORB_ERR_CUDA;
    return orb_launch_popc_bench(gpopc_per_s, lg.lane->stream);
}

void* orb_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { orb_cuda_fail(cudaGetLastError(), "cudaMallocHost"); return nullptr; }
    return p;
}
void orb_host_free(void* p) { if (p) cudaFreeHost(p); }
/* page-locked AND write-combined: for input frames the CPU only ever writes (camera / decoder output).  The CPU caches are not
 * snooped during the host->device copy; CPU reads of such memory are very slow. */
void* orb_host_alloc_input(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocWriteCombined) != cudaSuccess) { orb_cuda_fail(cudaGetLastError(), "cudaHostAlloc"); return nullptr; }
    return p;
}

} // extern "C"
