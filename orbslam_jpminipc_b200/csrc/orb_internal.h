// orb_internal.h — shared host/device structures of liborb_b200 (not part of the public ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <condition_variable>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/orb_b200.h"

#define ORB_MAX_LEVELS 16
#define ORB_MAX_GRID 96            // max grid cols / rows per level
#define ORB_MAX_CELLS_LEVEL 1024   // cells of one level's grid (~5000 features on ONE level; 1000..2000 features over 8 levels need 45..90): shared arrays of the selection kernels
#define ORB_EDGE 16                // EDGE_THRESHOLD, reference src/ORBextractor.cc:77
// Width of the reflect-101 ring k_border actually writes around every level ROI.  The reference materialises all 16 pixels
// (copyMakeBorder, :806,:814) but nothing on the path reads further out than 3: FAST, IC_Angle and HarrisResponses stay inside the ROI
// (keypoints lie >= 16 px from its edge), GaussianBlur 7x7 reaches 3 px, and the rotated BRIEF pattern reaches
// ceil(sqrt(13^2 + 13^2)) = 19 px from a keypoint, i.e. at most 3 px past the edge.  The outer 12 px of the padded planes are
// never written and never read (orb_debug_level_plane returns them as they are).
#define ORB_RING 4
#ifndef ORB_TILE_W
#define ORB_TILE_W 64            // FAST tile width (multiple of 64); measured 64x32 1.38 ms, 64x64 1.28, 128x64 1.25, 64x128 1.24 per 256 frames
#endif
#ifndef ORB_RESIZE_THREADS
#define ORB_RESIZE_THREADS 256     // k_resize CTA size; a thread owns 4 output columns, so a tile is (4*threads/tile_w)*rows high
#define ORB_RESIZE_CTAS 4
#endif
#ifndef ORB_TILE_H
#define ORB_TILE_H 112           // with the half-lane tile (k_fast_nms<true>) 112 rows keep six CTAs per SM resident: 2.888 -> 2.814 ms per 1024 frames against 128
#endif
#define ORB_BLUR_TILE_W 64
#define ORB_BLUR_TILE_H 56

// Geometry of one pyramid level for one image shape (reference src/ORBextractor.cc:527-547,:786).
struct LevelGeom {
    int w, h;              // level ROI size
    int stride;            // padded row pitch in bytes (w+32 rounded up to 16)
    int prows;             // h + 32
    int plane_off;         // byte offset of the padded plane inside a frame block
    int nDesired, cols, rows, cellW, cellH, nfCell;
    int xend, yend;        // FAST detection region is [16,xend) x [16,yend) in ROI coordinates
    int cell_base;         // first cell of this level in the frame-wide cell arrays
    int ncells;
    int lvl_base, lvl_cap; // per-level keypoint list (u64 records) inside the frame's list block
    int patch_size;        // (int)(31*scale)
    float scale;           // mvScaleFactor[level]
    int xtab_off, ytab_off;// offsets into the resize coefficient tables (int2 entries)
    int kp_base;           // prefix of nDesired over levels (unused slots stay empty)
    int border_base;       // first k_border work item of this level
    int border_items;      // 32-bit words of this level's 16-px frame
    int ct_off, ct_len;    // k_fast_nms column masks: byte offset of this level's three arrays (in-region, left / right neighbour in the
                           // same cell; entry x + 4 for ROI column x) in the mask table, bytes per array (multiple of 4)
    int rt_off;            // k_fast_nms detection-cell row of ROI row y at rt_off + y + 1 of the row table (int16, -1 = outside)
    int ring;              // width of the reflect-101 ring k_border writes around the ROI: ORB_RING, or ORB_EDGE when a detection cell reaches past [16, size - 16)
    int bm_off, bm_pitch;  // NMS-survivor bitmap of this level: byte offset in the frame's bitmap block, row pitch in bytes (bit i = ROI x 16+i)
};

struct Plan {
    int nlevels, w, h;
    int frame_bytes;       // bytes of one frame's padded pyramid (multiple of 256)
    int ncells;            // cells per frame, all levels
    int cand_total;        // candidate slots per frame (u32 records)
    int lvl_total;         // level-list slots per frame (u64 records)
    int kp_cap;            // sum of nDesired
    int fast_th, th_lo;
    int harris;                  // HARRIS_SCORE: candidate responses are replaced by HarrisResponses before the selection (:616-620)
    int ntiles_fast, ntiles_blur;
    int bm_total;          // bitmap bytes per frame
    int sel_list_cap;      // max lvl_cap over levels (k_select shared-memory list)
    int sel_cells_cap;     // max ncells over levels, rounded up to a multiple of 4 (k_select_fast per-cell tables)
    int border_total;      // k_border work items (32-bit words of all frame regions) per image
    int desc_fma;          // descriptor rotation with the reference compiler's FMA contraction (orb_set_descriptor_fma)
    LevelGeom L[ORB_MAX_LEVELS];
};

// One FAST detection cell (reference src/ORBextractor.cc:560-599): detection rectangle in ROI
// coordinates, the cell image origin (iniX, iniY) and where its candidates live.
struct CellGeom {
    int x0, x1, y0, y1;    // detection rect [x0,x1) x [y0,y1)
    int inix, iniy;        // cell image origin (x0-3, y0-3 for every processed cell)
    int cand_off, cand_cap;
    int level, idx;        // idx = i*cols + j
    int skipped, pad;      // skipped: the reference `continue`s before FAST (:570,:594); its quota state stays open
};

struct Tile { int level, x0, y0, pad; };

// one TMA descriptor per pyramid level: u8 tensor (x = padded row bytes, y = padded rows, z = frame)
struct TmapSet { CUtensorMap m[ORB_MAX_LEVELS]; };

// per-stream work buffers of the extraction pipeline
struct WorkSet {
    uint8_t* d_planes = nullptr;  size_t planes_bytes = 0;     // un-blurred padded pyramids
    uint8_t* d_work = nullptr;    size_t work_bytes = 0;       // NMS score map
    uint8_t* d_blur = nullptr;    size_t blur_bytes = 0;       // blurred ROIs + un-blurred frame
    uint8_t* d_bitmap = nullptr;  size_t bitmap_bytes = 0;     // 1 bit per detection pixel: NMS survivor
    uint32_t* d_cand = nullptr;   size_t cand_bytes = 0;
    unsigned long long* d_cand64 = nullptr; size_t cand64_bytes = 0;   // HARRIS_SCORE only: (float response bits << 32) | y << 12 | x
    int* d_ntotal = nullptr;      size_t ntotal_bytes = 0;
    unsigned long long* d_lvl = nullptr; size_t lvl_bytes = 0;
    int* d_nkept = nullptr;       size_t nkept_bytes = 0;
    int* d_counters = nullptr;    size_t counters_bytes = 0;   // dynamic tile-queue counters (32 ints)
    TmapSet tm_fast{}, tm_blur{}, tm_resize[2]{};              // boxes: FAST tile / blur tile / resize source footprint (map l reads level l-1; [1] = the small-call tiling)
    const uint8_t* tm_base = nullptr; int tm_frames = 0, tm_w = 0, tm_h = 0;
    cudaStream_t aux_stream = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // CUDA-graph replay of one extraction pass (orb_api.cu, launch_extract): captured the second time the same call shape and
    // buffers come by, replayed while they stay the same
    struct GraphKey {
        const void *in = nullptr, *kps = nullptr, *desc = nullptr, *counts = nullptr, *planes = nullptr;
        int n = 0, w = 0, h = 0, stride = 0, cap = 0; size_t pitch = 0; long long plan_gen = -1;
        bool operator==(const GraphKey& o) const
        {
            return in == o.in && kps == o.kps && desc == o.desc && counts == o.counts && planes == o.planes && n == o.n && w == o.w && h == o.h &&
                   stride == o.stride && cap == o.cap && pitch == o.pitch && plan_gen == o.plan_gen;
        }
    };
    GraphKey graph_key{}, last_key{};
    cudaGraphExec_t graph_exec = nullptr;
    int graph_launches = 0;
};

// One matcher "lane": stream + device scratch + pinned staging arena of ONE host-pointer matcher / vocabulary / frame-plumbing call.
// The reference calls ORBmatcher from three threads (Tracking, LocalMapping, LoopClosing: src/main.cc:165,182,193), so a context
// keeps a small pool of lanes; every such call borrows one for its duration (LaneGuard) and never touches the extraction buffers.
struct MatchLane {
    cudaStream_t stream = nullptr;
    void* d_scratch = nullptr; size_t scratch_bytes = 0;
    uint8_t* h_arena = nullptr; size_t arena_bytes = 0;     // pinned mirror of the input part of the scratch: one H2D per call
    bool busy = false;
};

struct orb_ctx {
    int device = 0;
    // extraction entry points are serialised per context (the reference object is not re-entrant either, include/ORBextractor.h:74-75)
    std::recursive_mutex ex_mu;
    static constexpr int MAX_LANES = 8;
    std::mutex lane_mu; std::condition_variable lane_cv;
    MatchLane lanes[MAX_LANES]; int nlanes = 0;
    cudaMemPool_t pool = nullptr;                           // stream-ordered scratch of the *_device matcher calls (k_knn2 partials)
    int nfeatures = 0, nlevels = 0, score_type = 1, fast_th = 20;
    float scale_factor_f = 1.2f;
    double scaleFactor = 1.2;
    int max_w = 0, max_h = 0, max_batch = 0;
    std::vector<float> mvScaleFactor, mvInvScaleFactor;
    std::vector<int> mnFeaturesPerLevel;
    int umax[16];

    // current plan (rebuilt when the image shape changes)
    Plan plan{};
    bool plan_valid = false;
    std::vector<CellGeom> cells;
    std::vector<Tile> tiles_fast, tiles_blur;
    std::vector<int2> xtab, ytab;
    std::vector<uint8_t> fast_coltab;      // per level: in-region / left / right byte masks of every column (k_fast_nms)
    std::vector<int16_t> fast_rowtab;      // per level: detection-cell row of every row
    // two tilings of the resize cascade: [0] for batches (8 rows per thread), [1] for calls of a few frames (rs_rows_small rows per
    // thread: more, shorter CTAs — the seven dependent launches of ONE frame are bound by the run time of a single tile)
    int rs_box_w[2][ORB_MAX_LEVELS] = { { 0 } }, rs_box_h[2][ORB_MAX_LEVELS] = { { 0 } };   // k_resize TMA box (source footprint of one output tile)
    int rs_tile_w[2][ORB_MAX_LEVELS] = { { 0 } }, rs_rows[2][ORB_MAX_LEVELS] = { { 0 } };   // k_resize tile width / rows per thread

    // device tables (shared by both work sets)
    Plan* d_plan = nullptr;
    CellGeom* d_cells = nullptr;
    Tile* d_tiles_fast = nullptr; Tile* d_tiles_blur = nullptr;
    int2* d_xtab = nullptr; int2* d_ytab = nullptr;
    uint8_t* d_fast_coltab = nullptr; int16_t* d_fast_rowtab = nullptr;
    size_t cap_fast_coltab = 0, cap_fast_rowtab = 0;
    size_t cap_cells = 0, cap_tiles_fast = 0, cap_tiles_blur = 0, cap_xtab = 0, cap_ytab = 0;
    int* d_status = nullptr;          // error flag raised by kernels
    // two independent sets of work buffers: two chunks (or the two halves of one device batch) run on two streams
    // so that one half's latency-bound selection overlaps the other half's FAST / blur
    WorkSet ws[2];
    cudaEvent_t ev_user = nullptr, ev_half[2] = { nullptr, nullptr };
    int last_n0 = 0, last_n1 = 0;     // frames handled by ws[0] / ws[1] in the last launch (debug getters)
    // asynchronous host-buffer calls (orb_extract_batch_async / orb_wait): a ring of completion records
    static constexpr int NTICKETS = 8;
    struct Ticket { cudaEvent_t done = nullptr, a = nullptr, b = nullptr; int* h_status = nullptr; int32_t* counts = nullptr;
                    int nimg = 0, cap = 0; long long seq = -1; bool host_out = false;
                    // small call into pageable output buffers: the results come back as ONE block into pinned staging and orb_wait copies
                    // counts[i] rows per frame to the caller (three copies into pageable memory are three blocking, staged transfers)
                    uint8_t* h_stage = nullptr; size_t stage_bytes = 0; bool staged = false;
                    orb_keypoint* u_kps = nullptr; uint8_t* u_desc = nullptr; };
    Ticket tickets[NTICKETS];
    cudaStream_t done_stream = nullptr;
    long long next_seq = 0, waited_seq = -1;   // tickets issued / highest ticket known complete
    int chunk_parity = 0;             // work set / stream the next chunk goes to
    bool chain_chunks = true, kernels_pending = false;
    cudaEvent_t ev_dev_done = nullptr; bool dev_call_pending = false;   // last orb_extract_batch_device call on a caller stream
    // staging for host-pointer calls (two slots for copy/compute overlap)
    uint8_t* d_src[2] = { nullptr, nullptr };  size_t src_bytes[2] = { 0, 0 };
    size_t kps_bytes[2] = { 0, 0 }, desc_bytes[2] = { 0, 0 }, counts_bytes[2] = { 0, 0 };
    orb_keypoint* d_kps[2] = { nullptr, nullptr };
    uint8_t* d_desc[2] = { nullptr, nullptr };
    int32_t* d_counts[2] = { nullptr, nullptr };
    uint8_t* d_small[2] = { nullptr, nullptr };  size_t small_bytes[2] = { 0, 0 };   // [keypoints | descriptors | counts] of a staged small call
    int stage_small = 1;                                   // ORB_STAGE_SMALL=0: always copy straight into the caller's buffers (A/B timing)
    cudaStream_t streams[2] = { nullptr, nullptr };
    cudaStream_t out_streams[2] = { nullptr, nullptr };   // device->host copies of a slot's results (host-buffer calls)
    cudaEvent_t ev_out_done[2] = { nullptr, nullptr };
    bool out_pending[2] = { false, false };
    cudaEvent_t ev_free[2] = { nullptr, nullptr };
    int last_launches = 0;
    int num_sms = 148;
    int split_device = 0;
    int desc_fma = 0;                                      // orb_set_descriptor_fma
    int debug_skip = 0;                                    // ORB_DEBUG_SKIP, honoured only by a -DORB_DEBUG build (timing experiments, results are wrong): 1 no blur, 2 no selection, 4 no describe
    bool pdl_call = false;                                 // this call's kernels carry the PDL launch attribute (set by launch_extract)
    int use_pdl = 1, pdl_frames = 2;                       // ORB_PDL=0: plain stream-ordered launches (A/B timing); ORB_PDL_FRAMES: largest call launched with PDL (1 frame 93 against 102 us, 2: 108 / 111, 4: 149 / 139)
    int side_border = 1;                                   // ORB_SIDE_BORDER=0: k_border stays on the main stream for small calls too (A/B timing)
    int fast_wide = 1;                                     // ORB_FAST_WIDE=0: 128-thread k_fast_nms CTAs for small calls too (A/B timing)
    int compact_wide = 1;                                  // ORB_COMPACT_WIDE=0: k_cell_compact (warp per cell) for small calls too (A/B timing)
    int select_wide = 1;                                   // ORB_SELECT_WIDE=0: k_select_fast keeps 8 warps per CTA for small calls too (A/B timing)
    int rs_flex_width = 1;                                 // ORB_RESIZE_FLEX=0: fixed 128-column k_resize tiles (A/B timing)
    // ORB_RESIZE_ROWS_SMALL / ORB_SMALL_CALL: rows per thread of the small-call resize tiling; the largest call (frames) that still counts
    // as small (0 = no small-call forms at all).  The individual forms have their own limits (orb_extract.cu, orb_launch_extract).
    int rs_rows_small = 2, small_call_frames = 12;
    int stage_frames = 4;                                  // largest call whose pageable outputs return through the staged block (its rows are copied once more by the host)
    int rs_rows_pref = 8;                                  // ORB_RESIZE_ROWS: output rows per k_resize thread (tile height = 8 * rows at 128 columns)
    int fast_etile = 1;                                    // ORB_FAST_ETILE=0: k_fast_nms<false> (ring samples by PRMT from the raw tile) instead of the half-lane tile (A/B timing)
    int rs_unrolled = 1;                                   // ORB_RESIZE_UNROLLED=0: k_resize for every level instead of k_resize_u (A/B timing)
    int rs_packed[ORB_MAX_LEVELS] = { 0 };                 // every 4-column group's taps lie within 8 source bytes (k_resize_u's only form)
    int rs_xg_off[ORB_MAX_LEVELS] = { 0 };                 // k_resize_u column-group table of the level: offset into xtab (int2 entries, 4 per group)
    int pyr_fused = 0;                                     // ORB_PYR_FUSED=0: one k_resize launch per level instead of the single k_pyramid launch (A/B timing)
    int select_serial = 0;                                 // ORB_SELECT_SERIAL=1: the thread-per-cell selection kernel (A/B timing)
    int knn_engine = 0;                                    // ORB_KNN_POPC (default) / ORB_KNN_TENSOR: orb_set_knn_engine, ORB_KNN_ENGINE=tensor
    int use_graph = 1;                                     // ORB_GRAPH=0 switches the CUDA-graph replay off
    long long plan_gen = 0;                                // bumped whenever the plan (image shape) is rebuilt
    int fork_early = 0, fast_ctas = 6, blur_ctas = 8;      // stream-overlap tuning (ORB_FORK_EARLY / ORB_FAST_CTAS_FORK / ORB_BLUR_CTAS env)
    bool profile = false;
    std::vector<cudaEvent_t> prof_events;   // (ORB_NSTAGES+1) per profiled launch
    std::vector<cudaEvent_t> prof_pool;
    int last_nimg = 0;
};

// borrows a lane for the scope of one call (blocks while all MAX_LANES are in use); lane == nullptr after a CUDA failure
struct LaneGuard {
    orb_ctx* c; MatchLane* lane;
    explicit LaneGuard(orb_ctx* ctx);
    ~LaneGuard();
    LaneGuard(const LaneGuard&) = delete; LaneGuard& operator=(const LaneGuard&) = delete;
};
// make sure the lane's device scratch / pinned arena hold `bytes` / `arena_bytes`
int orb_lane_scratch(MatchLane* L, size_t bytes, size_t arena_bytes = 0);

// status helpers
extern thread_local std::string g_last_cuda_error;
int orb_cuda_fail(cudaError_t e, const char* what);
#define ORB_CUDA(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return orb_cuda_fail(e__, #x); } while (0)

// orb_plan.cu
int orb_build_tables(orb_ctx* c);
int orb_build_plan(orb_ctx* c, int w, int h);
// orb_extract.cu
int orb_launch_extract(orb_ctx* c, WorkSet& W, const uint8_t* d_imgs, int nimg, int w, int h, int stride, size_t frame_pitch,
                       orb_keypoint* d_kps, uint8_t* d_desc, int cap, int32_t* d_counts, cudaStream_t s);
int orb_upload_constants(const int* umax);
int orb_build_tmaps(orb_ctx* c, WorkSet& W, int nframes);
int orb_select_smem_setup(int list_cap, int cells_cap);
int orb_resize_smem_setup(int max_bytes);
// orb_match.cu
int orb_launch_knn2(orb_ctx* c, const uint8_t* d_q, int nq, const uint8_t* d_db, int64_t ndb, int npairs, int32_t idx_base,
                    int32_t* d_idx1, int32_t* d_d1, int32_t* d_d2, cudaStream_t s);
int orb_launch_knn2_tc(orb_ctx* c, const uint8_t* d_q, int nq, const uint8_t* d_db, int64_t ndb, int npairs, int32_t idx_base,
                       int32_t* d_idx1, int32_t* d_d1, int32_t* d_d2, cudaStream_t s);      // orb_match_tc.cu
int orb_launch_knn2_merge_pairs(const int32_t* d_parts, int nparts, int nq, int npairs, int32_t* d_idx1, int32_t* d_d1, int32_t* d_d2, cudaStream_t s);
int orb_launch_knn2_merge(const int32_t* d_parts, int nparts, int nq, int32_t* d_idx1, int32_t* d_d1, int32_t* d_d2, cudaStream_t s);
#define ORB_COMM_MAX_RANKS 16
// the same merge with one pointer per part (parts may live on peer devices: orb_comm.cu, "p2p" transport)
int orb_launch_knn2_merge_ptrs(const int32_t* const* d_parts, int nparts, int nq, int32_t* d_idx1, int32_t* d_d1, int32_t* d_d2, cudaStream_t s);
int orb_launch_match_ratio(const int32_t* idx1, const int32_t* d1, const int32_t* d2, int nq, float nnratio, int th,
                           int32_t* match, int* d_count, cudaStream_t s);
int orb_launch_grid_build(const orb_keypoint* kps, int n, int min_x, int max_x, int min_y, int max_y,
                          int32_t* cell_start, int32_t* cell_items, cudaStream_t s);
int orb_launch_search_by_projection(orb_ctx* c, const orb_frame_view* cur, const orb_frame_view* last,
                                    const uint8_t* last_has_mp, const uint8_t* last_outlier, const float* last_xyz,
                                    const float* T16_host, float th, int check_ori, int32_t* match_cur, int* d_result,
                                    uint8_t* scratch, size_t scratch_bytes, cudaStream_t s);
size_t orb_sbp_scratch_bytes(int n_cur, int n_last);
int orb_launch_search_by_bow(orb_ctx* c, const orb_featvec_view* kf_fv, const uint8_t* kf_desc, const orb_keypoint* kf_kps,
                             const uint8_t* kf_mp_valid, const orb_featvec_view* f_fv, const uint8_t* f_desc,
                             const orb_keypoint* f_kps, int n_f, int f_items_total, float nnratio, int check_ori, int32_t* match_f,
                             uint8_t* scratch, cudaStream_t s, const uint8_t* f_valid = nullptr, int32_t* match12 = nullptr, int n_kf = 0);
size_t orb_bow_scratch_bytes(int n_f);
int orb_launch_search_window(orb_ctx* c, const orb_frame_view* tgt, const orb_window_query_set* q, int accept, float nnratio, int th_dist,
                             int histogram, int32_t* match, int* d_result, uint8_t* scratch, size_t scratch_bytes, cudaStream_t s);
int orb_launch_search_window_best(orb_ctx* c, const orb_frame_view* tgt, const orb_window_query_set* q, int32_t* best_idx, int32_t* best_dist,
                                  int* d_result, uint8_t* scratch, size_t scratch_bytes, cudaStream_t s);
size_t orb_tri_scratch_bytes(int n1, int n2);
int orb_launch_search_for_triangulation(const orb_featvec_view* fv1, const uint8_t* desc1, const orb_keypoint* kps1, const uint8_t* has_mp1,
                                        int n1, const orb_featvec_view* fv2, const uint8_t* desc2, const orb_keypoint* kps2,
                                        const uint8_t* has_mp2, int n2, int items2_total, const float* F12, const float* sigma2, int nlevels,
                                        int check_ori, int32_t* match12, uint8_t* scratch, cudaStream_t s);
int orb_launch_distinctive(const uint8_t* d_desc, const int32_t* d_start, int npoints, int32_t* d_best_idx, int32_t* d_best_median,
                           cudaStream_t s);
size_t orb_init_scratch_bytes(int n1, int n2);
int orb_launch_search_for_initialization(orb_ctx* c, const orb_frame_view* f1, const orb_frame_view* f2, float* d_prev, int window,
                                         float nnratio, int check_ori, int32_t* d_matches12, int* d_result, uint8_t* scratch,
                                         size_t scratch_bytes, cudaStream_t s);
int orb_launch_popc_bench(double* gpopc, cudaStream_t s);
