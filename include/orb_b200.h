/*
 * orb_b200.h — C ABI of liborb_b200.so: the B200-native (sm_100a) ORB front end.
 *
 * Drop-in boundary for ONE path of caomw/ORBSLAM_jpMiniPC (ORB-SLAM v1): ORB extraction and
 * binary-descriptor matching.  The reference has no FFI layer of its own; the boundary is the
 * two C++ classes ORB_SLAM::ORBextractor (include/ORBextractor.h:32-77, called from
 * src/Frame.cc:60) and ORB_SLAM::ORBmatcher (include/ORBmatcher.h:37-107).  Each entry point
 * below names the reference symbol it replaces.  include/ORBextractor.h and
 * include/ORBmatcher.h in this repo are header-only C++ shims with the reference's class
 * names and call signatures on top of this ABI.
 *
 * Conventions: plain pointers and sizes, opaque handle, int status (0 = ok, <0 = error, see
 * orb_error_string), no exceptions cross the ABI, caller owns all output buffers.  Image,
 * keypoint and descriptor pointers may be HOST or DEVICE pointers unless a function says
 * otherwise (detected with cudaPointerGetAttributes); the *_device entry points take device
 * pointers only, enqueue on the given stream and do not synchronise.
 * Every function fails with ORB_ERR_CUDA when no CUDA device is usable: there is no CPU path.
 */
#ifndef ORB_B200_H
#define ORB_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* bit-compatible with cv::KeyPoint (28 bytes; also the record the reference fork serialises,
 * include/SaveLoadWorld.h:1408-1424) */
typedef struct orb_keypoint {
    float x, y;        /* pt, level-0 image coordinates (src/ORBextractor.cc:769-775) */
    float size;        /* (int)(31*scale[level])       (:675,692) */
    float angle;       /* degrees [0,360), IC_Angle    (:124-151) */
    float response;    /* FAST-9/16 score               (:607,613) */
    int32_t octave;    /* pyramid level                 (:691) */
    int32_t class_id;  /* -1 */
} orb_keypoint;

enum {
    ORB_OK = 0,
    ORB_ERR_INVALID = -1,      /* bad argument */
    ORB_ERR_GEOMETRY = -2,     /* cell grid the reference itself cannot process (it throws / divides by 0) */
    ORB_ERR_CAPACITY = -3,     /* caller buffer or context limit (max_w/max_h/max_batch/cap) too small */
    ORB_ERR_CUDA = -4,         /* CUDA runtime error or no device; details via orb_last_cuda_error */
    ORB_ERR_UNSUPPORTED = -5   /* an option outside the accelerated path (e.g. a tilted-sensor distortion model, a non-L1 vocabulary score) */
};
enum { ORB_HARRIS_SCORE = 0, ORB_FAST_SCORE = 1 };   /* include/ORBextractor.h:37 */

typedef struct orb_ctx orb_ctx;

const char* orb_error_string(int status);
const char* orb_last_cuda_error(void);
int         orb_abi_version(void);

/* ---------------------------------------------------------------- extraction ----------
 * ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, scoreType, fastTh)
 * (src/ORBextractor.cc:457-511).  A context is stateful and not re-entrant, like the
 * reference object (pyramid buffers are members, include/ORBextractor.h:74-75): use one
 * context per calling thread.  max_w/max_h/max_batch size the device buffers. */
orb_ctx* orb_create(int device, int nfeatures, float scale_factor, int nlevels, int score_type,
                    int fast_th, int max_w, int max_h, int max_batch);
void     orb_destroy(orb_ctx*);
/* A process-wide matcher-only context on the current CUDA device (ORB_B200_DEVICE overrides), created on first use and kept until the
 * process exits: what a reference-signature `ORBmatcher matcher(0.9, true);` on the stack (src/Tracking.cc:392,528,596,759,907,946,
 * src/LocalMapping.cc:232,421, src/LoopClosing.cc:252,578) runs on.  Matcher calls on one context are safe from several threads
 * (each call borrows its own stream and scratch), so Tracking, LocalMapping and LoopClosing share it.  NULL without a usable GPU. */
orb_ctx* orb_default_context(void);
int      orb_nlevels(const orb_ctx*);          /* ORBextractor::GetLevels()      include/ORBextractor.h:47 */
float    orb_scale_factor(const orb_ctx*);     /* ORBextractor::GetScaleFactor() include/ORBextractor.h:50 */
int      orb_keypoint_capacity(const orb_ctx*);/* rows to allocate per image: sum of per-level quotas */
/* Which reference BUILD the descriptors reproduce.  The rotation of the sampling pattern, `x*b + y*a` / `x*a - y*b`
 * (src/ORBextractor.cc:166-167), is written with two roundings per expression; the reference's own flags (-O3 -march=native,
 * CMakeLists.txt:12-13) let GCC contract it to fma(x, b, y*a) / fma(x, a, -(y*b)) on an FMA-capable host, which changes about 2
 * descriptor bits per 40 000 keypoints.  on = 0 (default): as written; on = 1: the contracted form.  Keypoints are unaffected.
 * The rotation factors themselves, `a = (float)cos(angle), b = (float)sin(angle)` (:160, libm's cosf / sinf on a float), are glibc's
 * (>= 2.28) cosf / sinf bit for bit on every float angle (csrc/orb_trig.h, tools/cpp/sincos_exhaustive.cu). */
int      orb_set_descriptor_fma(orb_ctx*, int on);

/* ORBextractor::operator()(image, mask, keypoints, descriptors), src/ORBextractor.cc:718-779.
 * The mask is a no-op in the reference (built at :791-810 but never handed to FAST, :601-607),
 * so it is not part of the ABI.  img: 8-bit gray, row stride in bytes.  Empty image
 * (img==NULL or w/h <= 0) -> *n = 0, ORB_OK (the reference returns silently, :721-722).
 * kps[cap], desc[cap*32]; *n receives the keypoint count. */
int orb_extract(orb_ctx*, const uint8_t* img, int w, int h, int stride,
                orb_keypoint* kps, uint8_t* desc, int cap, int* n);

/* nimg frames of the same shape, frame i at imgs + i*frame_pitch (bytes).  Outputs for frame i
 * at kps + i*cap and desc + i*cap*32; counts[i] = its keypoint count.  Frames are independent
 * (this is what is sharded across GPUs); internally processed in chunks of max_batch with
 * copies and kernels overlapped on two streams when the buffers are host memory. */
int orb_extract_batch(orb_ctx*, const uint8_t* imgs, int nimg, int w, int h, int stride, size_t frame_pitch,
                      orb_keypoint* kps, uint8_t* desc, int cap, int32_t* counts);
/* Asynchronous form of orb_extract_batch for streaming ingest: enqueues the batch and returns a ticket; consecutive calls keep
 * alternating the two internal work sets, so the host->device copy of one call overlaps the kernels of the previous one.  All
 * buffers of a call must stay valid (host buffers pinned for real overlap) until orb_wait(ticket) returns; at most 7 calls may
 * be in flight (an 8th waits for the oldest).  orb_wait returns what the synchronous call would have returned. */
int orb_extract_batch_async(orb_ctx*, const uint8_t* images, int nimg, int width, int height, int stride, size_t frame_pitch,
                            orb_keypoint* kps, uint8_t* desc, int cap, int32_t* counts, long long* ticket);
int orb_wait(orb_ctx*, long long ticket);

/* device pointers only, nimg <= max_batch, asynchronous on `stream` (a cudaStream_t).  Truncation contract of the device-output
 * form (nothing is read back, so no status can report it): d_counts[i] is the number of keypoints the frame HAS; when it exceeds
 * cap only the first cap rows of its slot were written, so a consumer reads min(d_counts[i], cap) rows.  (The host-output calls
 * clamp every counts[i] to cap and return ORB_ERR_CAPACITY instead.)  cap >= orb_keypoint_capacity(ctx) never truncates.
 * A context owns one set of work buffers: consecutive device calls on one context are ordered behind each other by an event even
 * when they are given different streams. */
int orb_extract_batch_device(orb_ctx*, const uint8_t* d_imgs, int nimg, int w, int h, int stride, size_t frame_pitch,
                             orb_keypoint* d_kps, uint8_t* d_desc, int cap, int32_t* d_counts, void* stream);
/* kernels launched by the last orb_extract* call on this context (for bench accounting) */
int orb_last_launch_count(const orb_ctx*);
/* per-stage device timing of the extraction pipeline (CUDA events recorded on the launching stream
 * between the stages).  orb_profile_read synchronises, returns the milliseconds summed per stage over
 * the extract launches since the last read (ms[ORB_NSTAGES]) and how many launches that was. */
enum { ORB_NSTAGES = 7 };   /* level0, resize, fast_nms, cell_compact, select, blur, describe */
int orb_profile_enable(orb_ctx*, int on);
int orb_profile_read(orb_ctx*, double* ms, int* ncalls);
const char* orb_profile_stage_name(int stage);

/* test / inspection hooks (valid after an extract call; frame < nimg of that call's last chunk):
 * info[0..9] = w,h,stride,nDesired,gridCols,gridRows,cellW,cellH,nfeaturesCell,nKept */
int orb_debug_level_info(orb_ctx*, int frame, int level, int32_t* info);
/* copies the padded (h+32) x stride plane to host memory `out`; which: 0 un-blurred, 1 blurred ROI
 * (its 16-px border is undefined: descriptors read the un-blurred border, src/ORBextractor.cc:760) */
int orb_debug_level_plane(orb_ctx*, int frame, int level, int which, uint8_t* out, size_t out_bytes);

/* ------------------------------------------------------------------ matching ---------- */
/* ORBmatcher::DescriptorDistance(a, b), src/ORBmatcher.cc:1794-1810 (two 32-byte host rows). */
int orb_descriptor_distance(const uint8_t* a, const uint8_t* b);

/* Brute-force best / second-best over ALL db rows with the reference's scan semantics
 * (src/ORBmatcher.cc:197-222): idx1 = lowest index attaining the minimum, d1 = that distance,
 * d2 = second order statistic of the distance multiset (ties with d1 count).  ndb == 0 ->
 * idx1 = -1, d1 = d2 = INT32_MAX.  q[nq*32], db[ndb*32]. */
int orb_hamming_knn2(orb_ctx*, const uint8_t* q, int nq, const uint8_t* db, int64_t ndb,
                     int32_t* idx1, int32_t* d1, int32_t* d2);
/* npairs independent (q, db) blocks laid out back to back (block p at q + p*nq*32, db + p*ndb*32);
 * device pointers, asynchronous.  idx_base is added to every idx1 >= 0 (global index of a DB shard). */
int orb_hamming_knn2_device(orb_ctx*, const uint8_t* d_q, int nq, const uint8_t* d_db, int64_t ndb, int npairs,
                            int32_t idx_base, int32_t* d_idx1, int32_t* d_d1, int32_t* d_d2, void* stream);
/* Which kernel computes the brute-force scan.  ORB_KNN_POPC (default, pinned by the path's contract): XOR + carry-save + POPC on the
 * integer pipes.  ORB_KNN_TENSOR (experiment): descriptor bits as +-1 int8, tcgen05.mma kind::i8 with the accumulator in TMEM,
 * Hamming = (256 - dot) / 2 — exact integer arithmetic, results bit-identical.  Also ORB_KNN_ENGINE=tensor|popc at orb_create. */
enum { ORB_KNN_POPC = 0, ORB_KNN_TENSOR = 1 };
int orb_set_knn_engine(orb_ctx*, int engine);
/* Exact merge of per-shard results: parts[s] = (idx1, d1, d2) of shard s, each
 * nq int32 laid out as parts[(s*3+k)*nq + i].  best = lexicographic min of (d1, idx1);
 * second = 2nd smallest of the multiset union {d1_s, d2_s}.  Device pointers, asynchronous. */
int orb_knn2_merge_device(orb_ctx*, const int32_t* d_parts, int nparts, int nq,
                          int32_t* d_idx1, int32_t* d_d1, int32_t* d_d2, void* stream);
/* acceptance test of src/ORBmatcher.cc:224-226: d1 <= th && (float)d1 < nnratio*(float)d2.
 * match[i] = idx1[i] or -1; returns the number of matches in *nmatches.  Host or device pointers. */
int orb_match_ratio(orb_ctx*, const int32_t* idx1, const int32_t* d1, const int32_t* d2, int nq,
                    float nnratio, int th, int32_t* match, int* nmatches);

/* The slice of ORB_SLAM::Frame the matcher reads (src/Frame.cc:56-128, include/Frame.h). */
typedef struct orb_frame_view {
    int32_t n;                     /* Frame::N */
    const orb_keypoint* kps;       /* mvKeysUn (== mvKeys: configs use zero distortion, src/Frame.cc:291-295) */
    const uint8_t* desc;           /* mDescriptors, n x 32 */
    float fx, fy, cx, cy;          /* Frame::fx.. (src/Frame.cc:80-83) */
    int32_t min_x, max_x, min_y, max_y;   /* mnMinX.. (src/Frame.cc:342-348) */
    int32_t nlevels; float scale_factor;  /* mnScaleLevels, mfScaleFactor (src/Frame.cc:92-93) */
    const int32_t* cell_start;     /* 64*48+1: CSR of mGrid, cell id = ix*48+iy (include/Frame.h:35-36,90) */
    const int32_t* cell_items;     /* keypoint indices in insertion order */
} orb_frame_view;
enum { ORB_GRID_COLS = 64, ORB_GRID_ROWS = 48 };

/* Grid assignment of src/Frame.cc:109-123 with PosInGrid :267-277.  Host or device pointers
 * (all of one kind).  cell_start[64*48+1], cell_items[n]. */
int orb_frame_grid_build(orb_ctx*, const orb_keypoint* kps, int n, int min_x, int max_x, int min_y, int max_y,
                         int32_t* cell_start, int32_t* cell_items);

/* ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, float th),
 * src/ORBmatcher.cc:1507-1620, constructed as ORBmatcher(nnratio, checkOri).  All pointers
 * inside the views and the arrays below are HOST pointers or all DEVICE pointers.
 * last_has_mp[i] != 0 <=> LastFrame.mvpMapPoints[i] != NULL; last_outlier = mvbOutlier;
 * last_xyz[3*i..] = pMP->GetWorldPos(); Tcw = CurrentFrame.mTcw row-major 4x4.
 * match_cur[i2] (in/out, cur->n entries) = index i of the last-frame feature whose map point is
 * assigned to CurrentFrame.mvpMapPoints[i2], or -1.  *nmatches = the reference's return value. */
int orb_search_by_projection(orb_ctx*, const orb_frame_view* cur, const orb_frame_view* last,
                             const uint8_t* last_has_mp, const uint8_t* last_outlier, const float* last_xyz,
                             const float* Tcw16, float th, int check_ori, int32_t* match_cur, int* nmatches);

/* ------------------------------------------------------------------------------------------------
 * Generic windowed greedy search: the loop shared by the other projection / window searches of ORBmatcher
 * (SURVEY.md §8f.1).  Queries are processed in index order like the reference; each takes the best keypoint of
 * `target` inside its window that no earlier query (and no pre-existing match) has claimed.
 *   ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*>&, th)  src/ORBmatcher.cc:49-125
 *       u,v = mTrackProjX/Y; radius = RadiusByViewingCos(viewCos)[*th]*scale[level]; levels [level-1, level];
 *       accept = ORB_ACCEPT_LEVEL_RATIO, th_dist = TH_HIGH, no histogram
 *   ORBmatcher::WindowSearch(F1, F2, windowSize, ...)                        src/ORBmatcher.cc:409-516
 *       u,v = F1 keypoint; radius_const = windowSize; levels [l, l]; accept = ORB_ACCEPT_RATIO, TH_HIGH, histogram
 *   ORBmatcher::SearchByProjection(F1, F2, windowSize, ...)                  src/ORBmatcher.cc:519-594
 *       xyz + Tcw16 (projected with the target's intrinsics, no bounds test); radius_const = windowSize;
 *       levels [l, l]; accept = ORB_ACCEPT_RATIO, TH_HIGH, no histogram; match_target pre-filled with F2's map points
 * All pointers inside the structures are HOST pointers or all DEVICE pointers. */
typedef struct orb_window_query_set {
    int32_t n;
    const uint8_t* active;          /* n: 0 = query skipped (no map point / bad / out of view / level filtered) */
    const uint8_t* desc;            /* n x 32 query descriptors */
    const float* u; const float* v; /* explicit window centres, or NULL ... */
    const float* xyz; const float* Tcw16;   /* ... then 3-D points (n x 3) projected with Tcw16 (4x4 row major, host memory) */
    int32_t check_bounds;           /* reject centres outside [min_x,max_x] x [min_y,max_y] (src/ORBmatcher.cc:1539-1542) */
    const float* radius; float radius_const;   /* per-query radius, or NULL and one value */
    const int32_t* min_level; const int32_t* max_level;   /* octave range per query (-1,-1: any) */
    const float* angle;             /* query keypoint angles (rotation histogram), may be NULL */
} orb_window_query_set;
enum { ORB_ACCEPT_BEST = 0, ORB_ACCEPT_RATIO = 1, ORB_ACCEPT_LEVEL_RATIO = 2 };
/* match_target[target->n] in/out: index of the query matched to each target keypoint, or -1 */
int orb_search_window(orb_ctx*, const orb_frame_view* target, const orb_window_query_set* queries, int accept_mode,
                      float nnratio, int th_dist, int check_ori, int32_t* match_target, int* nmatches);

/* ORBmatcher::SearchForInitialization(Frame &F1, Frame &F2, vector<cv::Point2f> &vbPrevMatched, vector<int> &vnMatches12,
 * int windowSize), src/ORBmatcher.cc:598-713 (monocular initialisation, src/Tracking.cc:393-401).  Octave-0 features of F1
 * search a window around prev_matched[i] in F2; a closer later feature steals an earlier one's match (:637,:656-660).
 * prev_matched: n1 x 2 floats in/out; matches12[n1] out = F2 index or -1.  f1 needs kps/desc only, f2 also its grid. */
int orb_search_for_initialization(orb_ctx*, const orb_frame_view* f1, const orb_frame_view* f2, float* prev_matched, int window_size,
                                  float nnratio, int check_ori, int32_t* matches12, int* nmatches);

/* Best candidate per query without claims ("match to the most similar keypoint in the radius"): the scoring loops of
 * ORBmatcher::Fuse (src/ORBmatcher.cc:1016-1134 and :1136-1265, candidates of KeyFrame::GetFeaturesInArea with octave in
 * [nPredictedLevel-1, nPredictedLevel]) and of both directions of ORBmatcher::SearchBySim3 (:1267-1505).  best_idx[i] = target
 * keypoint or -1, best_dist[i] = its distance or INT_MAX; the caller applies TH_LOW / TH_HIGH and the graph updates.
 * ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) (:286-407) is orb_search_window with ORB_ACCEPT_BEST,
 * th_dist = 50, octave range [nPredictedLevel-1, nPredictedLevel] and match_target pre-filled from vpMatched. */
int orb_search_window_best(orb_ctx*, const orb_frame_view* target, const orb_window_query_set* queries, int32_t* best_idx,
                           int32_t* best_dist);

/* DBoW2::FeatureVector as CSR (Thirdparty/DBoW2/DBoW2/FeatureVector.cpp:31-45): node ids ascending,
 * per node the feature indices in insertion order. */
typedef struct orb_featvec_view {
    int32_t nnodes; const int32_t* node_id; const int32_t* start; const int32_t* items;
} orb_featvec_view;
/* ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches),
 * src/ORBmatcher.cc:155-284 (candidate scoring; the vocabulary transform that produces the
 * FeatureVectors is out of scope).  kf_mp_valid[i] != 0 <=> KF feature i has a live map point.
 * match_f[n_f] out = KF feature index matched to frame feature, or -1. */
int orb_search_by_bow(orb_ctx*, const orb_featvec_view* kf_fv, const uint8_t* kf_desc, const orb_keypoint* kf_kps,
                      const uint8_t* kf_mp_valid, int n_kf,
                      const orb_featvec_view* f_fv, const uint8_t* f_desc, const orb_keypoint* f_kps, int n_f,
                      float nnratio, int check_ori, int32_t* match_f, int* nmatches);

/* ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12), src/ORBmatcher.cc:715-850
 * (loop closing).  valid1/valid2[i] != 0 <=> the feature has a live map point; match12[n1] out = feature index in
 * the second keyframe whose map point is assigned to vpMatches12[idx1], or -1.  Acceptance: bestDist1 < TH_LOW and
 * (float)bestDist1 < nnratio*(float)bestDist2 (:791-793). */
int orb_search_by_bow_kf(orb_ctx*, const orb_featvec_view* fv1, const uint8_t* desc1, const orb_keypoint* kps1,
                         const uint8_t* valid1, int n1,
                         const orb_featvec_view* fv2, const uint8_t* desc2, const orb_keypoint* kps2,
                         const uint8_t* valid2, int n2,
                         float nnratio, int check_ori, int32_t* match12, int* nmatches);

/* MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:185-250) for npoints map points at once: point p owns the observation
 * descriptors desc[start[p] .. start[p+1]) (32 bytes each, in the order of its observation map); best_idx[p] = index within the
 * group of the descriptor with the least median distance to the others (first on ties, -1 for an empty group), best_median[p] =
 * that median.  Host or device pointers. */
int orb_distinctive_descriptors(orb_ctx*, const uint8_t* desc, const int32_t* start, int npoints, int32_t* best_idx,
                                int32_t* best_median);

/* ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedKeys1, vMatchedKeys2, vMatchedPairs), src/ORBmatcher.cc:852-1014
 * (LocalMapping::CreateNewMapPoints, src/LocalMapping.cc:274): features of two keyframes without a map point, joined through the
 * shared vocabulary nodes, distance <= TH_LOW and within 2*BestDist, epipolar test CheckDistEpipolarLine (:136-153) with the
 * 3x3 CV_32F fundamental matrix F12 (row major) and pKF2's level_sigma2[nlevels] (mvLevelSigma2), rotation histogram.
 * has_mp1/2[i] != 0: the feature already has a map point.  match12[n1] out = index in KF2 or -1; vMatchedPairs are the pairs
 * (i, match12[i]) in ascending i. */
int orb_search_for_triangulation(orb_ctx*, const orb_featvec_view* fv1, const uint8_t* desc1, const orb_keypoint* kps1,
                                 const uint8_t* has_mp1, int n1, const orb_featvec_view* fv2, const uint8_t* desc2,
                                 const orb_keypoint* kps2, const uint8_t* has_mp2, int n2, const float* F12, const float* level_sigma2,
                                 int nlevels, int check_ori, int32_t* match12, int* nmatches);

/* ---- frame plumbing on either side of the extractor (reference src/Tracking.cc:200-212, src/Frame.cc:289-349) ---- */
enum { ORB_RGB = 0, ORB_BGR = 1 };
/* cvtColor(image, im, CV_RGB2GRAY / CV_BGR2GRAY) of Tracking::GrabImage (src/Tracking.cc:202-208) for nimg interleaved 8-bit
 * 3-channel frames: gray = (R*9798 + G*19235 + B*3735 + 2^14) >> 15 (OpenCV 4.x).  stride / pitch in bytes; host or device. */
int orb_cvt_gray(orb_ctx*, const uint8_t* src, int nimg, int width, int height, size_t stride, size_t frame_pitch, int order,
                 uint8_t* dst, size_t dst_stride, size_t dst_pitch);
/* the colour conversion followed by orb_extract_batch */
int orb_extract_batch_color(orb_ctx*, const uint8_t* images, int nimg, int width, int height, size_t stride, size_t frame_pitch,
                            int order, orb_keypoint* kps, uint8_t* desc, int cap, int32_t* counts);
/* Frame::UndistortKeyPoints (src/Frame.cc:289-320): cv::undistortPoints(pts, pts, mK, mDistCoef, Mat(), mK) on the keypoint
 * positions, everything else of the keypoint copied; dist = (k1,k2,p1,p2[,k3[,k4,k5,k6[,s1..s4]]]) as CV_32F like mDistCoef;
 * dist[0] == 0 means no distortion (the copy of :291-295).  In place allowed.  Host or device pointers. */
int orb_undistort_keypoints(orb_ctx*, const orb_keypoint* kps, int n, float fx, float fy, float cx, float cy, const float* dist,
                            int ndist, orb_keypoint* kps_un);
/* Frame::ComputeImageBounds (src/Frame.cc:322-349): bounds = { mnMinX, mnMaxX, mnMinY, mnMaxY } */
int orb_image_bounds(orb_ctx*, int width, int height, float fx, float fy, float cx, float cy, const float* dist, int ndist,
                     int32_t bounds[4]);

/* ---- the fork's binary dumps (include/SaveLoadWorld.h:1408-1459) as database files; host-side file I/O, no GPU involved ----
 * Descriptor file: per keyframe { 0xEB 0x90, int32 n, n x 32 bytes }; keypoint file: { 0xEB 0x90, size_t n, n x 28-byte KeyPoint }.
 * Readers concatenate the rows of all records; rec_start (may be NULL) receives the nrecords+1 row offsets of the keyframes.
 * Call with rows == NULL to size the buffers (returns the totals); ORB_ERR_CAPACITY if a buffer is too small (totals still set). */
int orb_db_read_descriptors(const char* path, uint8_t* desc, int64_t cap_rows, int32_t* rec_start, int cap_records, int64_t* nrows,
                            int32_t* nrecords);
int orb_db_write_descriptors(const char* path, const uint8_t* desc, const int32_t* rec_start, int nrecords);
int orb_db_read_keypoints(const char* path, orb_keypoint* kps, int64_t cap_rows, int32_t* rec_start, int cap_records, int64_t* nrows,
                          int32_t* nrecords);
int orb_db_write_keypoints(const char* path, const orb_keypoint* kps, const int32_t* rec_start, int nrecords);

/* ---- DBoW2 vocabulary tree (ORB descriptors): Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h ----
 * Nodes are given in the order of the reference's text format (loadFromTextFile, :1338-1425): node 0 is the root, node i >= 1
 * has parent[i] < i, a 32-byte descriptor and a weight; children keep file order, nodes without children are the words and get
 * word ids in node order (:328, :1408-1414).  scoring / weighting are DBoW2's ScoringType / WeightingType (BowVector.h:36-53);
 * Data/ORBvoc.txt of the reference is k=10, L=6, L1_NORM, TF_IDF. */
typedef struct orb_vocab orb_vocab;
enum { ORB_L1_NORM = 0, ORB_L2_NORM, ORB_CHI_SQUARE, ORB_KL, ORB_BHATTACHARYYA, ORB_DOT_PRODUCT };
enum { ORB_TF_IDF = 0, ORB_TF, ORB_IDF, ORB_BINARY };
int  orb_vocab_create(orb_ctx*, int k, int L, int scoring, int weighting, int nnodes, const int32_t* parent, const uint8_t* desc,
                      const double* weight, orb_vocab** out);
/* ORBVocabulary::loadFromTextFile(strVocFile), src/main.cc:85-97 */
int  orb_vocab_load_text(orb_ctx*, const char* path, orb_vocab** out);
void orb_vocab_destroy(orb_vocab*);
int  orb_vocab_info(const orb_vocab*, int* k, int* L, int* nnodes, int* nwords);
/* transform(feature, word_id, weight, nid, levelsup) for n features, :1218-1260.  All pointers host or all device. */
int  orb_vocab_transform_features(orb_ctx*, orb_vocab*, const uint8_t* desc, int n, int levelsup, int32_t* word, double* weight,
                                  int32_t* node);
/* transform(features, BowVector&, FeatureVector&, levelsup), :1127-1193, as called by Frame::ComputeBoW (src/Frame.cc:279-287) and
 * KeyFrame::ComputeBoW (src/KeyFrame.cc:56-65), for nframes frames in one launch.  Frame f has counts[f] descriptors starting at row
 * f*slot_rows of desc (the layout orb_extract_batch writes).  Outputs use fixed slots of cap entries per frame (cap >= slot_rows,
 * <= 8192): BowVector = bow_word (ascending) / bow_val with nbow[f] entries; FeatureVector = CSR fv_node (ascending) / fv_start
 * (cap+1 per frame) / fv_items with nfv[f] nodes.  All pointers host or all device. */
int  orb_vocab_transform_batch(orb_ctx*, orb_vocab*, const uint8_t* desc, int slot_rows, const int32_t* counts, int nframes, int levelsup,
                               int cap, int32_t* bow_word, double* bow_val, int32_t* nbow, int32_t* fv_node, int32_t* fv_start,
                               int32_t* fv_items, int32_t* nfv);
/* Shared-word count and L1 score (ScoringObject.cpp:22-64) of a query BowVector against nkf keyframe BowVectors stored as CSR: the
 * loops of KeyFrameDatabase::DetectRelocalisationCandidates (src/KeyFrameDatabase.cc:198-252) / DetectLoopCandidates (:75-135)
 * without the list and covisibility bookkeeping.  common[k] = shared words; score[k] = (float)score for keyframes with more than
 * (int)(max_common*0.8f) shared words (all keyframes sharing a word if score_all), else 0. */
int  orb_bow_score_db(orb_ctx*, orb_vocab*, const int32_t* qword, const double* qval, int nq, int nkf, const int32_t* kf_start,
                      const int32_t* kf_word, const double* kf_val, int score_all, int32_t* common, float* score, int* max_common);
/* The two retrieval queries complete: KeyFrameDatabase::DetectRelocalisationCandidates(Frame*) (src/KeyFrameDatabase.cc:198-308,
 * loop = 0) and DetectLoopCandidates(KeyFrame*, minScore) (:75-196, loop = 1) over nkf keyframes in the order they were add()ed.
 *   excluded[k] != 0 (loop only, may be NULL): keyframe k is connected to the query keyframe (pKF->GetConnectedKeyFrames(), :77,:95)
 *     and never enters the list.
 *   cov_start / cov_idx (CSR, may be NULL): GetBestCovisibilityKeyFrames(10) of every keyframe, in its order (:146,:267); at most
 *     the first ten entries of a row are read.
 *   kf_score (in / out): the mRelocScore / mLoopScore member of every keyframe.  Keyframes scored by this query are rewritten, the
 *     others keep their value: the reference adds the STALE score of a covisible keyframe that shares a word with the query without
 *     having been scored by it (:278-281), so the state has to travel with the caller like the member travels with the keyframe.
 *   common[k] (out): words shared with the query (0 for excluded keyframes).
 *   cand (out, room for nkf), *ncand: the returned keyframes in the reference's order (first nomination in list order, :170-190).
 * Loop rules: only keyframes with score >= min_score are nominators, neighbours need more than minCommonWords shared words, the best
 * accumulated score starts at min_score.  All array pointers host or all device. */
int  orb_bow_detect_candidates(orb_ctx*, orb_vocab*, const int32_t* qword, const double* qval, int nq, int nkf, const int32_t* kf_start,
                               const int32_t* kf_word, const double* kf_val, const uint8_t* excluded, int loop, float min_score,
                               const int32_t* cov_start, const int32_t* cov_idx, float* kf_score, int32_t* common, int32_t* cand, int* ncand);

/* ------------------------------------------------------------------ multi-GPU ----------
 * The path shards two ways: frames are independent (contiguous blocks of frames per GPU, no data-path collective), and the
 * relocalisation-sized kNN splits the keyframe-descriptor DATABASE by contiguous row ranges with the queries replicated; each
 * rank returns (idx1, d1, d2) with GLOBAL row indices, ONE exchange of 12 bytes per query and rank follows (ncclAllGather over
 * NVLink / NVSwitch on the rank's own stream; or, single process without NCCL, rank 0's merge kernel loading the peers' partials
 * in place), then the exact merge: best = lexicographic min of (d1, idx1), second = 2nd smallest of the multiset union.  The
 * result is bit-identical to orb_hamming_knn2 over the whole database.  The reference is a single process (src/main.cc:165-212),
 * hence orb_comm_init; one process per GPU (torchrun / MPI) uses orb_comm_unique_id + orb_comm_init_rank.
 * libnccl.so.2 is opened at run time (no link dependency); ORB_COMM_TRANSPORT=p2p|nccl, ORB_NCCL_LIB=<path> override.
 * A communicator is driven by one host thread at a time (calls are serialised by a mutex). */
typedef struct orb_comm orb_comm;
/* single process: ranks = devices 0..ngpus-1 (ngpus <= 0: every visible device), one matcher-only context + stream per device */
orb_comm* orb_comm_init(int ngpus);
/* one process per GPU: rank 0 calls orb_comm_unique_id (128 bytes) and the host application broadcasts it; ctx is the rank's
 * context (not owned) */
int       orb_comm_unique_id(void* id128);
orb_comm* orb_comm_init_rank(orb_ctx* ctx, int nranks, int rank, const void* id128);
void      orb_comm_destroy(orb_comm*);
int       orb_comm_size(const orb_comm*);
const char* orb_comm_transport(const orb_comm*);          /* "nccl", "p2p" or "none (1 rank)" */
orb_ctx*  orb_comm_context(orb_comm*, int rank);          /* the context of a local rank (e.g. for orb_keypoint_capacity) */
/* single process: size the per-device extractors (ORBextractor constructor arguments, as orb_create) */
int orb_comm_set_extractor(orb_comm*, int nfeatures, float scale_factor, int nlevels, int score_type, int fast_th,
                           int max_w, int max_h, int max_batch);
/* single process: split ndb rows (host or device memory) into contiguous ranges, rank r gets rows [lo_r, hi_r), and upload them */
int orb_comm_db_upload(orb_comm*, const uint8_t* db, int64_t ndb);
/* a shard that already lives on the rank's device (e.g. generated there, or read from a dump by orb_db_read_descriptors and copied):
 * rows are global rows [row_base, row_base + nrows); not owned */
int orb_comm_db_attach(orb_comm*, int rank, const uint8_t* d_rows, int64_t nrows, int64_t row_base);
/* single process: q[nq*32] and the outputs are host pointers, or device pointers on rank 0's device; blocking */
int orb_knn2_sharded(orb_comm*, const uint8_t* q, int nq, int32_t* idx1, int32_t* d1, int32_t* d2);
/* one process per GPU: this rank's shard and replicated queries as device pointers; kNN, exchange and merge are enqueued on
 * `stream`, every rank receives the full result; asynchronous */
int orb_knn2_sharded_device(orb_comm*, const uint8_t* d_q, int nq, const uint8_t* d_rows, int64_t nrows, int64_t row_base,
                            int32_t* d_idx1, int32_t* d_d1, int32_t* d_d2, void* stream);
/* single process: orb_extract_batch with the frames split into contiguous blocks over the devices (after orb_comm_set_extractor);
 * host buffers (pinned for real overlap); every device runs its own copy / compute pipeline, one host thread drives all of them */
int orb_extract_batch_multi(orb_comm*, const uint8_t* imgs, int nimg, int w, int h, int stride, size_t frame_pitch,
                            orb_keypoint* kps, uint8_t* desc, int cap, int32_t* counts);

/* pinned host memory helpers (page-locked buffers make the host<->device copies asynchronous) */
void* orb_host_alloc(size_t bytes);
void  orb_host_free(void* p);
/* the same, write-combined (cudaHostAllocWriteCombined): input frames that the CPU writes once and never reads */
void* orb_host_alloc_input(size_t bytes);

/* register-resident __popc micro-benchmark: measured POPC32 rate of this GPU in G ops/s
 * (the integer-pipe roofline denominator for the matcher) */
int orb_measure_popc_peak(orb_ctx*, double* gpopc_per_s);

#ifdef __cplusplus
}
#endif
#endif
