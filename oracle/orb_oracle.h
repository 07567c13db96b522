/*
 * orb_oracle.h — C API of the CPU ORACLE (test infrastructure, NOT product code).
 *
 * The oracle is a from-scratch CPU restatement of the reference's ORB front end
 * (caomw/ORBSLAM_jpMiniPC: src/ORBextractor.cc, src/ORBmatcher.cc, src/Frame.cc)
 * with the OpenCV primitives the reference calls re-implemented to OpenCV-4.13
 * semantics.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product (liborb_b200.so) never does.
 *
 * PARITY PIN: the reference has no tests or golden vectors of its own and cannot be
 * compiled in the build container (needs ROS + OpenCV C++).  The oracle is therefore
 * pinned against the real OpenCV (python cv2 4.13.0) primitive by primitive and
 * through cv2.ORB known-answer tests (tests/golden/, tests/test_oracle_vs_cv2.py).
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* bit-compatible with cv::KeyPoint (28 B) */
typedef struct orc_keypoint {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orc_keypoint;

enum { ORC_BLUR_F32_SEPFILTER = 0,   /* OpenCV 4.13, non-isolated sub-matrix (what the reference hits) */
       ORC_BLUR_FIXED_256 = 1,       /* OpenCV 4.x bit-exact fixed-point path (contiguous Mat)          */
       ORC_BLUR_FIXED_257 = 2 };     /* OpenCV 2.4-era 8-bit fixed-point taps (sum 257)                 */

typedef struct orc_extractor orc_extractor;

/* ORBextractor::ORBextractor, src/ORBextractor.cc:457-511 */
orc_extractor* orc_extractor_create(int nfeatures, float scale_factor, int nlevels,
                                    int score_type, int fast_th, int blur_variant);
void orc_extractor_destroy(orc_extractor*);
/* restated cosf / sinf of the descriptor rotation vs the host's libm on angle bit patterns lo, lo + step, .. <= hi: number of differing angles */
long long orc_trig_mismatches(uint32_t lo, uint32_t hi, uint32_t step);
/* descriptor rotation x*b + y*a (src/ORBextractor.cc:166-167) as GCC contracts it under the reference's own -O3 -march=native on an
 * FMA host: fma(x, b, y*a) / fma(x, a, -(y*b)).  Default 0 = as written (two roundings). */
void orc_extractor_set_descriptor_fma(orc_extractor*, int on);

/* ORBextractor::operator(), src/ORBextractor.cc:718-779.  Returns 0, or <0 on error
 * (-2: geometry the reference itself would throw on, -3: capacity). */
int orc_extract(orc_extractor*, const uint8_t* img, int w, int h, int stride,
                orc_keypoint* kps, uint8_t* desc, int cap, int* n);

/* constant tables / per-level geometry (valid after a call to orc_extract) */
int   orc_nlevels(const orc_extractor*);
float orc_scale_factor(const orc_extractor*, int level);
float orc_inv_scale_factor(const orc_extractor*, int level);
int   orc_features_per_level(const orc_extractor*, int level);
const int* orc_umax(const orc_extractor*);                 /* 16 ints */
/* info[0..9] = w,h,stride(padded),nDesired,gridCols,gridRows,cellW,cellH,nfeaturesCell,nKept */
int   orc_level_info(const orc_extractor*, int level, int* info);
/* padded plane (w+32)x(h+32); blurred=0: as FAST/IC_Angle saw it, 1: after the in-place blur */
const uint8_t* orc_level_plane(const orc_extractor*, int level, int blurred);
/* candidates that entered retainBest (after the th=7 fallback), cell row-major then raster:
 * arrays of capacity cap; returns count (or -needed).  x,y are cell-local. */
int   orc_level_candidates(const orc_extractor*, int level, int cap, int* cell, int* x, int* y, int* score);
/* per-cell nTotal / nToRetain, gridRows*gridCols ints each */
int   orc_level_quota(const orc_extractor*, int level, int* ntotal, int* nretain);

/* ---- primitives, exposed so each can be checked against cv2 on its own ---- */
void  orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride,
                           uint8_t* dst, int dw, int dh, int dstride);
void  orc_border_reflect101(uint8_t* plane, int w, int h, int stride, int border); /* plane = padded origin */
/* cv::FAST(img, kps, th, true) TYPE_9_16; returns count; x/y/score arrays capacity cap */
int   orc_fast9_nms(const uint8_t* img, int w, int h, int stride, int th, int cap, int* x, int* y, int* score);
float orc_fast_atan2(float y, float x);
void  orc_gaussian_blur7(const uint8_t* src_padded_roi, int w, int h, int stride, uint8_t* dst, int dstride, int variant);
/* std::nth_element(first, first+nth, last, response >) over (resp, idx) pairs — permutes both arrays */
void  orc_nth_element_desc(float* resp, int32_t* idx, int n, int nth);
/* KeyPointsFilter::retainBest + the reference's resize(n): returns new count */
int   orc_retain_best(float* resp, int32_t* idx, int n, int npoints);
/* HarrisResponses(img, pts, 7, 0.04f) for one point, src/ORBextractor.cc:79-120 */
float orc_harris_response(const uint8_t* img, int stride, int x, int y);
float orc_ic_angle(const uint8_t* center, int stride);
void  orc_rbrief(const uint8_t* center, int stride, float angle_deg, uint8_t* desc32);

/* ---- matcher ---- */
/* ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:1794-1810 */
int   orc_descriptor_distance(const uint8_t* a, const uint8_t* b);
/* best / second-best scan (pattern src/ORBmatcher.cc:197-222) over all db rows, no claims */
void  orc_knn2(const uint8_t* q, int nq, const uint8_t* db, int64_t ndb,
               int32_t* idx1, int32_t* d1, int32_t* d2, int use_popcnt);
/* ratio/threshold acceptance: SearchByBoW form (src/ORBmatcher.cc:224-226) */
int   orc_match_ratio(const int32_t* idx1, const int32_t* d1, const int32_t* d2, int nq,
                      float nnratio, int th, int32_t* match);

/* Frame grid: src/Frame.cc:109-123 (fill) + :267-277 (PosInGrid).  CSR over 64x48 cells,
 * cell id = ix*48+iy.  cell_start has 64*48+1 entries, cell_items capacity n. */
typedef struct orc_frame {
    int n;
    const orc_keypoint* kps;       /* mvKeysUn (== mvKeys, zero distortion) */
    const uint8_t* desc;           /* n x 32 */
    float fx, fy, cx, cy;
    int min_x, max_x, min_y, max_y;  /* mnMinX.. (0,w,0,h for zero distortion, src/Frame.cc:342-348) */
    int nlevels; float scale_factor; /* mnScaleLevels, mfScaleFactor (src/Frame.cc:92-103) */
    const int32_t* cell_start; const int32_t* cell_items;
} orc_frame;
void  orc_frame_grid(const orc_keypoint* kps, int n, int min_x, int max_x, int min_y, int max_y,
                     int32_t* cell_start, int32_t* cell_items);
/* Frame::GetFeaturesInArea, src/Frame.cc:200-265; returns count */
int   orc_features_in_area(const orc_frame* f, float x, float y, float r, int min_level, int max_level,
                           int32_t* out, int cap);
/* ORBmatcher::SearchByProjection(Frame&,const Frame&,float), src/ORBmatcher.cc:1507-1620.
 * last_has_mp / last_outlier: per last-frame feature; last_xyz: world position of its map point.
 * match_cur[i2] in/out: index of the last-frame feature whose map point claimed keypoint i2, or -1. */
int   orc_search_by_projection(const orc_frame* cur, const orc_frame* last, const uint8_t* last_has_mp,
                               const uint8_t* last_outlier, const float* last_xyz, const float* Tcw16,
                               float th, int check_ori, int32_t* match_cur);
/* ORBmatcher::SearchByBoW(KeyFrame*,Frame&,...), src/ORBmatcher.cc:155-284.  FeatureVectors as CSR:
 * node ids ascending, per node the feature indices in insertion (ascending) order. */
typedef struct orc_featvec { int nnodes; const int32_t* node_id; const int32_t* start; const int32_t* items; } orc_featvec;
int   orc_search_by_bow(const orc_featvec* kf_fv, const uint8_t* kf_desc, const orc_keypoint* kf_kps,
                        const uint8_t* kf_mp_valid, int n_kf,
                        const orc_featvec* f_fv, const uint8_t* f_desc, const orc_keypoint* f_kps, int n_f,
                        float nnratio, int check_ori, int32_t* match_f);
void  orc_three_maxima(const int* hist_sizes, int L, int* ind1, int* ind2, int* ind3);

/* ---- further ORBmatcher searches ---- */
/* ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*>&, th), src/ORBmatcher.cc:49-125.
 * Per map point: track_in_view (&& !isBad), mTrackProjX/Y, mnTrackScaleLevel, mTrackViewCos, descriptor.
 * match_f[idx] in/out = index of the map point assigned to F.mvpMapPoints[idx], or -1. */
int   orc_search_by_projection_mappoints(const orc_frame* f, int nmp, const uint8_t* in_view, const float* proj_x, const float* proj_y,
                                         const int32_t* level, const float* view_cos, const uint8_t* mp_desc, float th, float nnratio,
                                         int32_t* match_f);
/* ORBmatcher::WindowSearch(F1, F2, windowSize, vpMapPointMatches2, minLevel, maxLevel), src/ORBmatcher.cc:409-516.
 * f1_has_mp[i1] != 0 <=> F1.mvpMapPoints[i1] is a live map point.  match2[i2] out = i1 or -1. */
int   orc_window_search(const orc_frame* f1, const orc_frame* f2, const uint8_t* f1_has_mp, int window, int min_level, int max_level,
                        float nnratio, int check_ori, int32_t* match2);
/* ORBmatcher::SearchByProjection(F1, F2, windowSize, vpMapPointMatches2), src/ORBmatcher.cc:519-594.
 * f1_active[i1] != 0 <=> F1 map point is live and not already found in F2; f1_xyz its world position;
 * match2[i2] in/out: >= 0 (or the caller's marker) means F2.mvpMapPoints[i2] is already set. */
int   orc_search_by_projection_window(const orc_frame* f1, const orc_frame* f2, const uint8_t* f1_active, const float* f1_xyz,
                                      const float* Tc2w16, int window, float nnratio, int32_t* match2);

/* ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*>&, float th, int ORBdist),
 * src/ORBmatcher.cc:1622-1746.  active[i]: KF map point i is live and not already found; pred_level[i]: the level the
 * reference derives from dist3D/minDistance (:1663-1669, computed by the caller); kf_angle[i] = pKF->GetKeyPointUn(i).angle. */
int   orc_search_by_projection_kf(const orc_frame* cur, int nmp, const uint8_t* active, const float* xyz, const float* Tcw16,
                                  const int32_t* pred_level, const uint8_t* mp_desc, const float* kf_angle, float th, int orb_dist,
                                  int check_ori, int32_t* match_cur);

/* ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize), src/ORBmatcher.cc:598-713.
 * prev_matched: n1 x 2 floats in/out (vbPrevMatched); matches12[n1] out. */
int   orc_search_for_initialization(const orc_frame* f1, const orc_frame* f2, float* prev_matched, int window, float nnratio,
                                    int check_ori, int32_t* matches12);

/* ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*>&), src/ORBmatcher.cc:715-850.
 * valid1/valid2: feature has a live map point.  match12[idx1] out = idx2 or -1. */
int   orc_search_by_bow_kf(const orc_featvec* fv1, const uint8_t* desc1, const orc_keypoint* kps1, const uint8_t* valid1, int n1,
                           const orc_featvec* fv2, const uint8_t* desc2, const orc_keypoint* kps2, const uint8_t* valid2, int n2,
                           float nnratio, int check_ori, int32_t* match12);

/* ---- back-end searches of ORBmatcher on pre-projected map points ---- */
/* ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th), src/ORBmatcher.cc:286-407, from :350 on: u, v and the
 * predicted level come from the caller (Sim3 projection :300-349); radius = th * scaleFactor[level] (:352).  matched[idx] in/out:
 * >= 0 = vpMatched[idx] already set (skipped, :376), on return the index of the map point assigned to it. */
int   orc_search_by_projection_sim3(const orc_frame* kf, int nmp, const uint8_t* active, const float* u, const float* v,
                                    const int32_t* pred_level, const uint8_t* mp_desc, int th, int32_t* matched);
/* "Match to the most similar keypoint in the radius" without claims: the loops of ORBmatcher::Fuse (:1083-1115, :1211-1243) and
 * SearchBySim3 (:1356-1385, :1436-1465); best_idx -1 / best_dist INT_MAX when no candidate */
void  orc_window_best(const orc_frame* f, int nq, const uint8_t* active, const float* u, const float* v, const float* radius,
                      const int32_t* pred_level, const uint8_t* desc, int32_t* best_idx, int32_t* best_dist);

/* ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, ...), src/ORBmatcher.cc:852-1014 with CheckDistEpipolarLine (:136-153).
 * has_mp1/2: the feature already has a map point (skipped); F12: 3x3 CV_32F row major; level_sigma2 = pKF2's mvLevelSigma2.
 * match12[n1] out = idx2 or -1 (vMatchedPairs = the pairs (i, match12[i]) in ascending i). */
int   orc_search_for_triangulation(const orc_featvec* fv1, const uint8_t* desc1, const orc_keypoint* kps1, const uint8_t* has_mp1, int n1,
                                   const orc_featvec* fv2, const uint8_t* desc2, const orc_keypoint* kps2, const uint8_t* has_mp2, int n2,
                                   const float* F12, const float* level_sigma2, int check_ori, int32_t* match12);

/* ---- DBoW2 vocabulary, Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h ---- */
typedef struct orc_vocab orc_vocab;
enum { ORC_L1_NORM = 0, ORC_L2_NORM, ORC_CHI_SQUARE, ORC_KL, ORC_BHATTACHARYYA, ORC_DOT_PRODUCT };   /* BowVector.h:45-53 */
enum { ORC_TF_IDF = 0, ORC_TF, ORC_IDF, ORC_BINARY };                                                /* BowVector.h:36-42 */
/* Nodes in text-file order (loadFromTextFile, TemplatedVocabulary.h:1338-1425): node 0 is the root, rows i>=1 give parent,
 * 32-byte descriptor and weight; children keep file order; leaves (nodes without children, :328) get word ids in node order. */
orc_vocab* orc_vocab_create(int k, int L, int scoring, int weighting, int nnodes, const int32_t* parent, const uint8_t* desc,
                            const double* weight);
orc_vocab* orc_vocab_load_text(const char* path);
void  orc_vocab_destroy(orc_vocab*);
int   orc_vocab_nnodes(const orc_vocab*);
int   orc_vocab_nwords(const orc_vocab*);
/* transform(feature, word_id, weight, nid, levelsup), TemplatedVocabulary.h:1218-1260 */
void  orc_vocab_transform_feature(const orc_vocab*, const uint8_t* desc32, int levelsup, int32_t* word, double* weight, int32_t* node);
/* transform(features, BowVector&, FeatureVector&, levelsup), :1127-1193.  BowVector as (word ascending, value); FeatureVector as
 * CSR (fv_start has nfv+1 entries).  Capacities: n entries each (fv_start n+1). */
void  orc_vocab_transform(const orc_vocab*, const uint8_t* desc, int n, int levelsup, int32_t* bow_word, double* bow_val, int* nbow,
                          int32_t* fv_node, int32_t* fv_start, int32_t* fv_items, int* nfv);
/* L1Scoring::score, ScoringObject.cpp:22-64 */
double orc_bow_score_l1(const int32_t* w1, const double* v1, int n1, const int32_t* w2, const double* v2, int n2);
/* the data-parallel part of KeyFrameDatabase::DetectRelocalisationCandidates, src/KeyFrameDatabase.cc:198-252: words shared with
 * every keyframe (mnRelocWords), max, and the score of those with more than (int)(max*0.8f) shared words (others 0). */
void  orc_bow_score_db(const int32_t* qw, const double* qv, int nq, int nkf, const int32_t* kf_start, const int32_t* kf_word,
                       const double* kf_val, int32_t* common, float* score, int* max_common);
/* both retrieval queries complete (src/KeyFrameDatabase.cc:75-196, :198-308) on flat arrays; returns the number of candidates */
int   orc_bow_detect_candidates(const int32_t* qw, const double* qv, int nq, int nkf, const int32_t* kf_start, const int32_t* kf_word,
                                const double* kf_val, const uint8_t* excluded, int loop, float min_score, const int32_t* cov_start,
                                const int32_t* cov_idx, float* kf_score, int32_t* common, int32_t* cand);

/* ---- frame plumbing ---- */
/* cvtColor(CV_RGB2GRAY / CV_BGR2GRAY), src/Tracking.cc:202-208; order 0 = RGB, 1 = BGR; OpenCV 4.x 15-bit coefficients */
void  orc_cvt_gray(const uint8_t* src, int w, int h, int stride, int order, uint8_t* dst, int dstride);
/* cv::undistortPoints(pts, pts, K, D, Mat(), K) as used by Frame::UndistortKeyPoints / ComputeImageBounds (src/Frame.cc:289-349);
 * xy: n x 2 floats in/out; dist: ndist CV_32F coefficients */
void  orc_undistort_points(float* xy, int n, float fx, float fy, float cx, float cy, const float* dist, int ndist);
void  orc_undistort_keypoints(const orc_keypoint* in, int n, float fx, float fy, float cx, float cy, const float* dist, int ndist,
                              orc_keypoint* out);
void  orc_image_bounds(int w, int h, float fx, float fy, float cx, float cy, const float* dist, int ndist, int32_t bounds[4]);

/* MapPoint::ComputeDistinctiveDescriptors, src/MapPoint.cc:185-250, for npoints map points at once: point p owns the observation
 * descriptors desc[start[p] .. start[p+1]); best_idx[p] = BestIdx within its group (-1 for an empty group), best_median[p] = BestMedian */
void  orc_distinctive_descriptors(const uint8_t* desc, const int32_t* start, int npoints, int32_t* best_idx, int32_t* best_median);

#ifdef __cplusplus
}
#endif
#endif
