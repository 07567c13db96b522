/*
 * ref_glue.cpp — C entry points around the REFERENCE's own classes, compiled together with the reference's sources from
 * /root/reference (never copied) against oracle/refshim/ into oracle/_ref/libref_orbslam.so.  TEST INFRASTRUCTURE: it
 * validates the oracle restatement (tests/test_ref_build.py) and serves bench.py's `--impl reference` arm; the product
 * never loads it.  See oracle/refshim/minicv.hpp for what is real reference code and what is shim.
 */
#include <cstdint>
#include <cstring>
#include <vector>

#include "ORBextractor.h"            /* /root/reference/include */

extern "C" {

/* ORB_SLAM::ORBextractor, include/ORBextractor.h:32-77 */
void* ref_extractor_create(int nfeatures, float scale_factor, int nlevels, int score_type, int fast_th)
{
    return new ORB_SLAM::ORBextractor(nfeatures, scale_factor, nlevels, score_type, fast_th);
}
void ref_extractor_destroy(void* e) { delete (ORB_SLAM::ORBextractor*)e; }

/* ORBextractor::operator()(image, mask = Mat(), keypoints, descriptors) as Frame::Frame calls it (src/Frame.cc:60).
 * kps: 28-byte cv::KeyPoint records.  Returns 0, -3 when cap is too small (n still reports the count), -1 on an exception. */
int ref_extract(void* e, const uint8_t* img, int w, int h, int stride, void* kps, uint8_t* desc, int cap, int* n)
{
    try {
        cv::Mat image(h, w, CV_8UC1);
        for (int y = 0; y < h; y++) std::memcpy(image.ptr(y), img + (size_t)y * stride, (size_t)w);
        std::vector<cv::KeyPoint> keys;
        cv::Mat descriptors;
        (*(ORB_SLAM::ORBextractor*)e)(image, cv::Mat(), keys, descriptors);
        *n = (int)keys.size();
        if (*n > cap) return -3;
        static_assert(sizeof(cv::KeyPoint) == 28, "cv::KeyPoint layout");
        if (*n) {
            std::memcpy(kps, keys.data(), keys.size() * sizeof(cv::KeyPoint));
            for (int i = 0; i < *n; i++) std::memcpy(desc + (size_t)i * 32, descriptors.ptr(i), 32);
        }
        return 0;
    } catch (const std::exception& ex) {
        std::fprintf(stderr, "ref_extract: %s\n", ex.what());
        return -1;
    }
}

} // extern "C"

#ifndef REF_GLUE_EXTRACTOR_ONLY
/* =====================================================================================================================
 * Map classes and ORBmatcher: the reference's Frame / KeyFrame / MapPoint / Map / KeyFrameDatabase / ORBmatcher objects
 * built from flat arrays.  Every search below runs the reference's own member function (src/ORBmatcher.cc); the glue only
 * builds its inputs and turns MapPoint pointers in its outputs back into feature indices.
 * ===================================================================================================================== */
#include <map>
#include <set>

#include "Frame.h"
#include "KeyFrame.h"
#include "KeyFrameDatabase.h"
#include "Map.h"
#include "MapPoint.h"
#include "ORBVocabulary.h"
#include "ORBmatcher.h"
#include "Converter.h"               /* oracle/refshim/Converter.h */

namespace ORB_SLAM {
/* src/Converter.cc:28-36 needs Eigen + g2o for its other members; this is the one the front end calls */
std::vector<cv::Mat> Converter::toDescriptorVector(const cv::Mat& Descriptors)
{
    std::vector<cv::Mat> v;
    v.reserve(Descriptors.rows);
    for (int j = 0; j < Descriptors.rows; j++) v.push_back(Descriptors.row(j));
    return v;
}
}

using namespace ORB_SLAM;

namespace {
struct RefFrame {
    Frame f;
    KeyFrame* kf = nullptr;
    Map* map = nullptr;
    std::vector<MapPoint*> owned;                  /* every MapPoint created for this frame */
    int min_x, max_x, min_y, max_y;
    float fx, fy, cx, cy;
    void statics()                                 /* the camera lives in static members of Frame (include/Frame.h:62-65,128-131) */
    {
        Frame::mnMinX = min_x; Frame::mnMaxX = max_x; Frame::mnMinY = min_y; Frame::mnMaxY = max_y;
        Frame::fx = fx; Frame::fy = fy; Frame::cx = cx; Frame::cy = cy;
        Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(max_x - min_x);
        Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(max_y - min_y);
        Frame::mbInitialComputations = false;
    }
    KeyFrame* keyframe()
    {
        if (!kf) {
            statics();
            map = new Map();
            kf = new KeyFrame(f, map, nullptr);
        }
        return kf;
    }
    MapPoint* new_point(const float* xyz, const uint8_t* desc32)
    {
        cv::Mat pos(3, 1, CV_32F);
        for (int k = 0; k < 3; k++) pos.at<float>(k) = xyz ? xyz[k] : 0.f;
        MapPoint* p = new MapPoint(pos, keyframe(), map);
        if (desc32) { cv::Mat d(1, 32, CV_8U, (void*)desc32); p->SetDescriptor(d.clone()); }
        owned.push_back(p);
        return p;
    }
};
template <typename F> int guarded(const char* what, F fn)
{
    try { return fn(); }
    catch (const std::exception& ex) { std::fprintf(stderr, "%s: %s\n", what, ex.what()); return -1000; }
    catch (const std::string& s) { std::fprintf(stderr, "%s: %s\n", what, s.c_str()); return -1000; }
}
void index_of(const std::vector<MapPoint*>& pts, std::map<MapPoint*, int>& m)
{
    for (size_t i = 0; i < pts.size(); i++) if (pts[i]) m[pts[i]] = (int)i;
}
}

extern "C" {

/* A Frame from undistorted keypoints + descriptors (what Frame::Frame leaves behind after extraction and
 * UndistortKeyPoints, src/Frame.cc:56-128): scale tables as :92-107, the grid filled by the loop of :109-123 calling the
 * reference's own PosInGrid (:267-277). */
void* ref_frame_create(const void* kps_un, const uint8_t* desc, int n, int min_x, int max_x, int min_y, int max_y,
                       float fx, float fy, float cx, float cy, int nlevels, float scale_factor)
{
    RefFrame* r = new RefFrame();
    r->min_x = min_x; r->max_x = max_x; r->min_y = min_y; r->max_y = max_y;
    r->fx = fx; r->fy = fy; r->cx = cx; r->cy = cy;
    r->statics();
    Frame& F = r->f;
    F.mpORBvocabulary = nullptr; F.mpORBextractor = nullptr; F.mTimeStamp = 0; F.mpReferenceKF = nullptr;
    F.N = n;
    F.mvKeysUn.resize(n);
    if (n) std::memcpy((void*)F.mvKeysUn.data(), kps_un, (size_t)n * sizeof(cv::KeyPoint));
    F.mvKeys = F.mvKeysUn;
    F.mDescriptors.create(std::max(n, 1), 32, CV_8U);
    if (n) std::memcpy(F.mDescriptors.data, desc, (size_t)n * 32);
    if (n == 0) F.mDescriptors = F.mDescriptors.rowRange(0, 0);
    F.mvpMapPoints = std::vector<MapPoint*>(n, static_cast<MapPoint*>(nullptr));
    F.mvbOutlier = std::vector<bool>(n, false);
    F.mK = cv::Mat::eye(3, 3, CV_32F);
    F.mK.at<float>(0, 0) = fx; F.mK.at<float>(1, 1) = fy; F.mK.at<float>(0, 2) = cx; F.mK.at<float>(1, 2) = cy;
    F.mDistCoef = cv::Mat::zeros(4, 1, CV_32F);
    F.mTcw = cv::Mat::eye(4, 4, CV_32F);
    F.UpdatePoseMatrices();
    F.mnId = Frame::nNextId++;
    F.mnScaleLevels = nlevels;
    F.mfScaleFactor = scale_factor;
    F.mvScaleFactors.resize(nlevels);
    F.mvLevelSigma2.resize(nlevels);
    F.mvScaleFactors[0] = 1.0f;
    F.mvLevelSigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        F.mvScaleFactors[i] = F.mvScaleFactors[i - 1] * F.mfScaleFactor;
        F.mvLevelSigma2[i] = F.mvScaleFactors[i] * F.mvScaleFactors[i];
    }
    F.mvInvLevelSigma2.resize(nlevels);
    for (int i = 0; i < nlevels; i++) F.mvInvLevelSigma2[i] = 1 / F.mvLevelSigma2[i];
    for (size_t i = 0; i < F.mvKeysUn.size(); i++) {
        int gx, gy;
        if (F.PosInGrid(F.mvKeysUn[i], gx, gy)) F.mGrid[gx][gy].push_back(i);
    }
    return r;
}

/* Frame::Frame(im, timeStamp, extractor, voc, K, distCoef), src/Frame.cc:56-128, run as is: extraction, UndistortKeyPoints,
 * ComputeImageBounds, grid.  dist4 = k1 k2 p1 p2.  voc may be NULL. */
void* ref_frame_from_image(void* extractor, void* voc, const uint8_t* img, int w, int h, int stride, float fx, float fy, float cx, float cy,
                           const float* dist4)
{
    RefFrame* r = nullptr;
    int rc = guarded("ref_frame_from_image", [&] {
        cv::Mat image(h, w, CV_8UC1);
        for (int y = 0; y < h; y++) std::memcpy(image.ptr(y), img + (size_t)y * stride, (size_t)w);
        cv::Mat K = cv::Mat::eye(3, 3, CV_32F);
        K.at<float>(0, 0) = fx; K.at<float>(1, 1) = fy; K.at<float>(0, 2) = cx; K.at<float>(1, 2) = cy;
        cv::Mat D(4, 1, CV_32F);
        for (int i = 0; i < 4; i++) D.at<float>(i) = dist4[i];
        Frame::mbInitialComputations = true;               /* a new camera: recompute the image bounds */
        Frame F(image, 0.0, (ORBextractor*)extractor, (ORBVocabulary*)voc, K, D);
        r = new RefFrame();
        r->f = F;
        r->min_x = Frame::mnMinX; r->max_x = Frame::mnMaxX; r->min_y = Frame::mnMinY; r->max_y = Frame::mnMaxY;
        r->fx = fx; r->fy = fy; r->cx = cx; r->cy = cy;
        r->f.mTcw = cv::Mat::eye(4, 4, CV_32F);
        r->f.UpdatePoseMatrices();
        return 0;
    });
    return rc == 0 ? r : nullptr;
}
void ref_frame_destroy(void* h)
{
    RefFrame* r = (RefFrame*)h;
    for (MapPoint* p : r->owned) delete p;
    delete r->kf;
    delete r->map;
    delete r;
}
int ref_frame_n(void* h) { return ((RefFrame*)h)->f.N; }
/* keys (as extracted), undistorted keys, descriptors, bounds[4] = minX maxX minY maxY; any pointer may be NULL */
void ref_frame_get(void* h, void* keys, void* keys_un, uint8_t* desc, int32_t* bounds)
{
    RefFrame* r = (RefFrame*)h;
    const int n = r->f.N;
    if (keys && n) std::memcpy(keys, r->f.mvKeys.data(), (size_t)n * sizeof(cv::KeyPoint));
    if (keys_un && n) std::memcpy(keys_un, r->f.mvKeysUn.data(), (size_t)n * sizeof(cv::KeyPoint));
    if (desc) for (int i = 0; i < n; i++) std::memcpy(desc + (size_t)i * 32, r->f.mDescriptors.ptr(i), 32);
    if (bounds) { bounds[0] = r->min_x; bounds[1] = r->max_x; bounds[2] = r->min_y; bounds[3] = r->max_y; }
}
/* the grid as CSR, cell id = ix*48+iy, items in insertion order (src/Frame.cc:109-123) */
void ref_frame_grid(void* h, int32_t* cell_start, int32_t* cell_items)
{
    RefFrame* r = (RefFrame*)h;
    int k = 0;
    for (int ix = 0; ix < FRAME_GRID_COLS; ix++)
        for (int iy = 0; iy < FRAME_GRID_ROWS; iy++) {
            cell_start[ix * FRAME_GRID_ROWS + iy] = k;
            for (size_t v : r->f.mGrid[ix][iy]) cell_items[k++] = (int32_t)v;
        }
    cell_start[FRAME_GRID_COLS * FRAME_GRID_ROWS] = k;
}
/* Frame::GetFeaturesInArea, src/Frame.cc:200-265 */
int ref_features_in_area(void* h, float x, float y, float rad, int min_level, int max_level, int32_t* out, int cap)
{
    RefFrame* r = (RefFrame*)h;
    r->statics();
    std::vector<size_t> v = r->f.GetFeaturesInArea(x, y, rad, min_level, max_level);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int32_t)v[i];
    return (int)v.size();
}
void ref_frame_set_pose(void* h, const float* Tcw16)
{
    RefFrame* r = (RefFrame*)h;
    cv::Mat T(4, 4, CV_32F);
    std::memcpy(T.data, Tcw16, 64);
    r->f.mTcw = T;
    r->f.UpdatePoseMatrices();
    if (r->kf) r->kf->SetPose(T);
}
/* FeatureVector from CSR (node ids ascending, feature indices in insertion order); must precede anything that makes the KeyFrame */
void ref_frame_set_featvec(void* h, int nnodes, const int32_t* node_id, const int32_t* start, const int32_t* items)
{
    RefFrame* r = (RefFrame*)h;
    r->f.mFeatVec.clear();
    for (int k = 0; k < nnodes; k++)
        for (int j = start[k]; j < start[k + 1]; j++) r->f.mFeatVec.addFeature((DBoW2::NodeId)node_id[k], (unsigned)items[j]);
}
/* map points on the frame's features: has[i] != 0 -> a MapPoint at xyz[3i..] (NULL: origin) observed by this frame's KeyFrame at
 * feature i; its descriptor comes from MapPoint::ComputeDistinctiveDescriptors over that one observation.  outlier may be NULL. */
void ref_frame_set_mappoints(void* h, const uint8_t* has, const float* xyz, const uint8_t* outlier)
{
    RefFrame* r = (RefFrame*)h;
    KeyFrame* kf = r->keyframe();
    for (int i = 0; i < r->f.N; i++) {
        r->f.mvbOutlier[i] = outlier && outlier[i];
        if (!has[i]) { r->f.mvpMapPoints[i] = nullptr; continue; }
        MapPoint* p = r->new_point(xyz ? xyz + 3 * i : nullptr, nullptr);
        p->AddObservation(kf, i);
        p->ComputeDistinctiveDescriptors();
        kf->AddMapPoint(p, i);
        r->f.mvpMapPoints[i] = p;
    }
}

/* ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:1794-1810 */
int ref_descriptor_distance(const uint8_t* a, const uint8_t* b)
{
    cv::Mat ma(1, 32, CV_8U, (void*)a), mb(1, 32, CV_8U, (void*)b);
    return ORBmatcher::DescriptorDistance(ma, mb);
}

/* ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, float th), :1507-1620.
 * match_cur[i2] in: >= 0 -> CurrentFrame.mvpMapPoints[i2] already holds LastFrame's point of that feature; out: the same map. */
int ref_search_by_projection_ff(void* cur_, void* last_, float th, float nnratio, int check_ori, int32_t* match_cur)
{
    RefFrame *cur = (RefFrame*)cur_, *last = (RefFrame*)last_;
    return guarded("SearchByProjection(F,F)", [&] {
        cur->statics();
        std::map<MapPoint*, int> idx;
        index_of(last->f.mvpMapPoints, idx);
        for (int i = 0; i < cur->f.N; i++) cur->f.mvpMapPoints[i] = match_cur[i] >= 0 ? last->f.mvpMapPoints[match_cur[i]] : nullptr;
        ORBmatcher m(nnratio, check_ori != 0);
        const int n = m.SearchByProjection(cur->f, last->f, th);
        for (int i = 0; i < cur->f.N; i++) match_cur[i] = cur->f.mvpMapPoints[i] ? idx.at(cur->f.mvpMapPoints[i]) : -1;
        return n;
    });
}

/* ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*>&, th), :49-125, on nmp free-standing map points whose tracking
 * members (set by Frame::isInFrustum in the reference) are given.  match_f[idx] in/out = index of the point on feature idx or -1. */
int ref_search_by_projection_mappoints(void* f_, int nmp, const uint8_t* in_view, const float* proj_x, const float* proj_y, const int32_t* level,
                                       const float* view_cos, const uint8_t* mp_desc, float th, float nnratio, int32_t* match_f)
{
    RefFrame* f = (RefFrame*)f_;
    return guarded("SearchByProjection(F,MapPoints)", [&] {
        f->statics();
        std::vector<MapPoint*> pts(nmp);
        for (int i = 0; i < nmp; i++) {
            MapPoint* p = f->new_point(nullptr, mp_desc + (size_t)i * 32);
            p->mbTrackInView = in_view[i] != 0;
            p->mTrackProjX = proj_x[i]; p->mTrackProjY = proj_y[i];
            p->mnTrackScaleLevel = level[i]; p->mTrackViewCos = view_cos[i];
            pts[i] = p;
        }
        std::map<MapPoint*, int> idx;
        index_of(pts, idx);
        for (int i = 0; i < f->f.N; i++) f->f.mvpMapPoints[i] = match_f[i] >= 0 ? pts[match_f[i]] : nullptr;
        ORBmatcher m(nnratio, true);
        const int n = m.SearchByProjection(f->f, pts, th);
        for (int i = 0; i < f->f.N; i++) match_f[i] = f->f.mvpMapPoints[i] ? idx.at(f->f.mvpMapPoints[i]) : -1;
        return n;
    });
}

/* ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&), :155-284.  match_f[idxF] out = KeyFrame feature whose point matched. */
int ref_search_by_bow(void* kf_, void* f_, float nnratio, int check_ori, int32_t* match_f)
{
    RefFrame *k = (RefFrame*)kf_, *f = (RefFrame*)f_;
    return guarded("SearchByBoW(KF,F)", [&] {
        f->statics();
        std::map<MapPoint*, int> idx;
        index_of(k->keyframe()->GetMapPointMatches(), idx);
        std::vector<MapPoint*> out;
        ORBmatcher m(nnratio, check_ori != 0);
        const int n = m.SearchByBoW(k->keyframe(), f->f, out);
        for (int i = 0; i < f->f.N; i++) match_f[i] = out[i] ? idx.at(out[i]) : -1;
        return n;
    });
}

/* ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&), :715-850.  match12[idx1] out = idx2 or -1. */
int ref_search_by_bow_kf(void* k1_, void* k2_, float nnratio, int check_ori, int32_t* match12)
{
    RefFrame *k1 = (RefFrame*)k1_, *k2 = (RefFrame*)k2_;
    return guarded("SearchByBoW(KF,KF)", [&] {
        k1->statics();
        std::map<MapPoint*, int> idx;
        index_of(k2->keyframe()->GetMapPointMatches(), idx);
        std::vector<MapPoint*> out;
        ORBmatcher m(nnratio, check_ori != 0);
        const int n = m.SearchByBoW(k1->keyframe(), k2->keyframe(), out);
        for (int i = 0; i < k1->f.N; i++) match12[i] = out[i] ? idx.at(out[i]) : -1;
        return n;
    });
}

/* ORBmatcher::WindowSearch(F1, F2, windowSize, vpMapPointMatches2, minLevel, maxLevel), :409-516.  match2[i2] out = i1 or -1. */
int ref_window_search(void* f1_, void* f2_, int window, int min_level, int max_level, float nnratio, int check_ori, int32_t* match2)
{
    RefFrame *f1 = (RefFrame*)f1_, *f2 = (RefFrame*)f2_;
    return guarded("WindowSearch", [&] {
        f1->statics();
        std::map<MapPoint*, int> idx;
        index_of(f1->f.mvpMapPoints, idx);
        std::vector<MapPoint*> out;
        ORBmatcher m(nnratio, check_ori != 0);
        const int n = m.WindowSearch(f1->f, f2->f, window, out, min_level, max_level);
        for (int i = 0; i < f2->f.N; i++) match2[i] = out[i] ? idx.at(out[i]) : -1;
        return n;
    });
}

/* ORBmatcher::SearchByProjection(F1, F2, windowSize, vpMapPointMatches2), :519-594.  F2's pose must be set.  match2[i2] in: >= 0 ->
 * F2.mvpMapPoints[i2] is already set (to a point that F1 does not hold); out: i1 for new matches, the input value otherwise. */
int ref_search_by_projection_window(void* f1_, void* f2_, int window, float nnratio, int32_t* match2)
{
    RefFrame *f1 = (RefFrame*)f1_, *f2 = (RefFrame*)f2_;
    return guarded("SearchByProjection(F1,F2,window)", [&] {
        f1->statics();
        std::map<MapPoint*, int> idx;
        index_of(f1->f.mvpMapPoints, idx);
        for (int i = 0; i < f2->f.N; i++) f2->f.mvpMapPoints[i] = match2[i] >= 0 ? f2->new_point(nullptr, f2->f.mDescriptors.ptr(i)) : nullptr;
        std::vector<MapPoint*> out;
        ORBmatcher m(nnratio, true);
        const int n = m.SearchByProjection(f1->f, f2->f, window, out);
        for (int i = 0; i < f2->f.N; i++)
            if (out[i] && idx.count(out[i])) match2[i] = idx[out[i]];
            else if (!out[i]) match2[i] = -1;
        return n;
    });
}

/* ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize), :598-713 */
int ref_search_for_initialization(void* f1_, void* f2_, float* prev_matched, int window, float nnratio, int check_ori, int32_t* matches12)
{
    RefFrame *f1 = (RefFrame*)f1_, *f2 = (RefFrame*)f2_;
    return guarded("SearchForInitialization", [&] {
        f1->statics();
        std::vector<cv::Point2f> prev(f1->f.N);
        for (int i = 0; i < f1->f.N; i++) prev[i] = cv::Point2f(prev_matched[2 * i], prev_matched[2 * i + 1]);
        std::vector<int> m12;
        ORBmatcher m(nnratio, check_ori != 0);
        const int n = m.SearchForInitialization(f1->f, f2->f, prev, m12, window);
        for (int i = 0; i < f1->f.N; i++) { matches12[i] = m12[i]; prev_matched[2 * i] = prev[i].x; prev_matched[2 * i + 1] = prev[i].y; }
        return n;
    });
}

/* ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, ...), :852-1014.  match12[n1] out = idx2 of the pair (idx1, idx2) or -1;
 * returns the reference's return value, npairs the length of vMatchedPairs. */
int ref_search_for_triangulation(void* k1_, void* k2_, const float* F12, float nnratio, int check_ori, int32_t* match12, int* npairs)
{
    RefFrame *k1 = (RefFrame*)k1_, *k2 = (RefFrame*)k2_;
    return guarded("SearchForTriangulation", [&] {
        k1->statics();
        cv::Mat F(3, 3, CV_32F);
        std::memcpy(F.data, F12, 36);
        std::vector<cv::KeyPoint> a, b;
        std::vector<std::pair<size_t, size_t> > pairs;
        ORBmatcher m(nnratio, check_ori != 0);
        const int n = m.SearchForTriangulation(k1->keyframe(), k2->keyframe(), F, a, b, pairs);
        for (int i = 0; i < k1->f.N; i++) match12[i] = -1;
        for (auto& p : pairs) match12[p.first] = (int32_t)p.second;
        *npairs = (int)pairs.size();
        return n;
    });
}

/* MapPoint::ComputeDistinctiveDescriptors, src/MapPoint.cc:185-250, for a point observed at features obs[0..nobs) of this frame's
 * KeyFrame... one KeyFrame can hold a point once, so each observation gets its own single-feature KeyFrame built from desc rows. */
int ref_distinctive_descriptor(const uint8_t* desc, int nobs, uint8_t* out32)
{
    return guarded("ComputeDistinctiveDescriptors", [&] {
        std::vector<RefFrame*> frames;
        cv::KeyPoint kp(10.f, 10.f, 31.f, 0.f, 1.f, 0, -1);
        MapPoint* p = nullptr;
        for (int i = 0; i < nobs; i++) {
            RefFrame* r = (RefFrame*)ref_frame_create(&kp, desc + (size_t)i * 32, 1, 0, 640, 0, 480, 500, 500, 320, 240, 8, 1.2f);
            frames.push_back(r);
            if (!p) p = r->new_point(nullptr, nullptr);
            p->AddObservation(r->keyframe(), 0);
        }
        int rc = -1;
        if (p) {
            p->ComputeDistinctiveDescriptors();
            cv::Mat d = p->GetDescriptor();
            if (!d.empty()) { std::memcpy(out32, d.data, 32); rc = 0; }
        }
        for (RefFrame* r : frames) ref_frame_destroy(r);
        return rc;
    });
}

/* ---- DBoW2 (Thirdparty/DBoW2): ORBVocabulary = TemplatedVocabulary<FORB::TDescriptor, FORB> ---- */
void* ref_vocab_load_text(const char* path)
{
    ORBVocabulary* v = new ORBVocabulary();
    bool ok = false;
    guarded("loadFromTextFile", [&] { ok = v->loadFromTextFile(path); return 0; });
    if (!ok) { delete v; return nullptr; }
    return v;
}
void ref_vocab_destroy(void* v) { delete (ORBVocabulary*)v; }
int ref_vocab_nwords(void* v) { return (int)((ORBVocabulary*)v)->size(); }
/* transform(features, BowVector&, FeatureVector&, levelsup), TemplatedVocabulary.h:1127-1193 */
int ref_vocab_transform(void* v_, const uint8_t* desc, int n, int levelsup, int32_t* bow_word, double* bow_val, int* nbow,
                        int32_t* fv_node, int32_t* fv_start, int32_t* fv_items, int* nfv)
{
    return guarded("transform", [&] {
        ORBVocabulary* v = (ORBVocabulary*)v_;
        cv::Mat D(std::max(n, 1), 32, CV_8U);
        if (n) std::memcpy(D.data, desc, (size_t)n * 32);
        std::vector<cv::Mat> feats = Converter::toDescriptorVector(n ? D : D.rowRange(0, 0));
        DBoW2::BowVector bv;
        DBoW2::FeatureVector fv;
        v->transform(feats, bv, fv, levelsup);
        int k = 0;
        for (auto& e : bv) { bow_word[k] = (int32_t)e.first; bow_val[k] = e.second; k++; }
        *nbow = k;
        int nn = 0, ni = 0;
        for (auto& e : fv) {
            fv_node[nn] = (int32_t)e.first; fv_start[nn] = ni; nn++;
            for (unsigned f : e.second) fv_items[ni++] = (int32_t)f;
        }
        fv_start[nn] = ni;
        *nfv = nn;
        return 0;
    });
}
/* transform(feature, word_id, weight, nid, levelsup), :1218-1260, through the public single-feature transform (:1197-1205) */
int ref_vocab_transform_feature(void* v_, const uint8_t* desc32, int32_t* word)
{
    ORBVocabulary* v = (ORBVocabulary*)v_;
    cv::Mat d(1, 32, CV_8U, (void*)desc32);
    *word = (int32_t)v->transform(d);
    return 0;
}
/* ORBVocabulary::score -> L1Scoring::score, ScoringObject.cpp:22-64 */
double ref_vocab_score(void* v_, const int32_t* w1, const double* v1, int n1, const int32_t* w2, const double* v2, int n2)
{
    ORBVocabulary* v = (ORBVocabulary*)v_;
    DBoW2::BowVector a, b;
    for (int i = 0; i < n1; i++) a.insert(a.end(), std::make_pair((DBoW2::WordId)w1[i], v1[i]));
    for (int i = 0; i < n2; i++) b.insert(b.end(), std::make_pair((DBoW2::WordId)w2[i], v2[i]));
    return v->score(a, b);
}

/* BowVector of a frame from (word ascending, value) arrays; must precede anything that makes the KeyFrame (it copies F.mBowVec) */
void ref_frame_set_bowvec(void* h, int n, const int32_t* word, const double* val)
{
    RefFrame* r = (RefFrame*)h;
    r->f.mBowVec.clear();
    for (int i = 0; i < n; i++) r->f.mBowVec.insert(r->f.mBowVec.end(), std::make_pair((DBoW2::WordId)word[i], val[i]));
}

/* KeyFrameDatabase::add for every keyframe, then KeyFrameDatabase::DetectRelocalisationCandidates(F), src/KeyFrameDatabase.cc:198-308.
 * common[k] = pKF->mnRelocWords, score[k] = pKF->mRelocScore (0 where the reference does not compute it), candidate[k] = returned. */
int ref_detect_relocalisation_candidates(void* voc_, void* query_, void** kfs, int nkf, int32_t* common, float* score, int32_t* candidate)
{
    return guarded("DetectRelocalisationCandidates", [&] {
        ORBVocabulary* voc = (ORBVocabulary*)voc_;
        RefFrame* q = (RefFrame*)query_;
        if (q->f.mnId == 0) q->f.mnId = Frame::nNextId++;     /* the reference compares mnRelocQuery (initially 0) with F->mnId */
        KeyFrameDatabase db(*voc);
        std::map<KeyFrame*, int> idx;
        for (int k = 0; k < nkf; k++) {
            KeyFrame* kf = ((RefFrame*)kfs[k])->keyframe();
            kf->mnRelocWords = 0; kf->mRelocScore = 0.f; kf->mnRelocQuery = 0;
            db.add(kf);
            idx[kf] = k;
        }
        std::vector<KeyFrame*> cand = db.DetectRelocalisationCandidates(&q->f);
        for (int k = 0; k < nkf; k++) {
            KeyFrame* kf = ((RefFrame*)kfs[k])->keyframe();
            common[k] = kf->mnRelocQuery == q->f.mnId ? kf->mnRelocWords : 0;
            score[k] = kf->mRelocScore;
            candidate[k] = 0;
        }
        for (KeyFrame* kf : cand) candidate[idx.at(kf)] = 1;
        return (int)cand.size();
    });
}

/* Both retrieval queries of the reference with a covisibility graph: KeyFrameDatabase::add for every keyframe (in order),
 * KeyFrame::AddConnection for nedges weighted edges (edge e joins keyframes ea[e], eb[e] with weight ew[e], both directions, as
 * KeyFrame::UpdateConnections does; -1 in ea = the QUERY keyframe, used by the loop query's GetConnectedKeyFrames), member scores preset
 * from score_io, then DetectRelocalisationCandidates(&F) (loop = 0, src/KeyFrameDatabase.cc:198-308) or DetectLoopCandidates(pKF,
 * min_score) (loop = 1, :75-196).  Outputs: the returned keyframes IN ORDER (cand, returns their number), common[k] = mnRelocWords /
 * mnLoopWords of the keyframes the query marked (0 otherwise), score_io[k] = mRelocScore / mLoopScore afterwards, and
 * best10[k*10..] = GetBestCovisibilityKeyFrames(10) as indices (-1 padded) so that the flat-array callers use the reference's own lists. */
int ref_detect_candidates(void* voc_, void* query_, void** kfs, int nkf, int nedges, const int32_t* ea, const int32_t* eb, const int32_t* ew,
                          int loop, float min_score, float* score_io, int32_t* common, int32_t* cand, int32_t* best10)
{
    return guarded("DetectCandidates", [&] {
        ORBVocabulary* voc = (ORBVocabulary*)voc_;
        RefFrame* q = (RefFrame*)query_;
        if (q->f.mnId == 0) q->f.mnId = Frame::nNextId++;
        KeyFrameDatabase db(*voc);
        std::map<KeyFrame*, int> idx;
        std::vector<KeyFrame*> kf(nkf);
        for (int k = 0; k < nkf; k++) kf[k] = ((RefFrame*)kfs[k])->keyframe();
        KeyFrame* qkf = loop ? q->keyframe() : nullptr;
        if (qkf && qkf->mnId == 0) qkf->mnId = KeyFrame::nNextId++;   /* the reference compares mnLoopQuery (initially 0) with pKF->mnId */
        for (int k = 0; k < nkf; k++) {
            kf[k]->mnRelocWords = 0; kf[k]->mnRelocQuery = 0; kf[k]->mnLoopWords = 0; kf[k]->mnLoopQuery = 0;
            kf[k]->mRelocScore = score_io[k]; kf[k]->mLoopScore = score_io[k];
            db.add(kf[k]);
            idx[kf[k]] = k;
        }
        for (int e = 0; e < nedges; e++) {
            KeyFrame* a = ea[e] < 0 ? qkf : kf[ea[e]];
            KeyFrame* b = kf[eb[e]];
            if (!a) continue;
            a->AddConnection(b, ew[e]);
            b->AddConnection(a, ew[e]);
        }
        for (int k = 0; k < nkf; k++) {
            std::vector<KeyFrame*> nb = kf[k]->GetBestCovisibilityKeyFrames(10);
            for (int j = 0; j < 10; j++) best10[k * 10 + j] = j < (int)nb.size() && idx.count(nb[j]) ? idx[nb[j]] : -1;
        }
        std::vector<KeyFrame*> out = loop ? db.DetectLoopCandidates(qkf, min_score) : db.DetectRelocalisationCandidates(&q->f);
        const long unsigned qid = loop ? qkf->mnId : q->f.mnId;
        for (int k = 0; k < nkf; k++) {
            common[k] = loop ? (kf[k]->mnLoopQuery == qid ? kf[k]->mnLoopWords : 0) : (kf[k]->mnRelocQuery == qid ? kf[k]->mnRelocWords : 0);
            score_io[k] = loop ? kf[k]->mLoopScore : kf[k]->mRelocScore;
        }
        int n = 0;
        for (KeyFrame* p : out) cand[n++] = idx.at(p);
        return n;
    });
}

/* MapPoint::UpdateNormalAndDepth (src/MapPoint.cc:271-313) on every point of this frame's KeyFrame: fills the normal and the scale
 * invariance distances the back-end searches read.  Call after the pose is set. */
void ref_frame_update_points(void* h)
{
    RefFrame* r = (RefFrame*)h;
    for (MapPoint* p : r->f.mvpMapPoints) if (p) p->UpdateNormalAndDepth();
}

/* The call sequence of Tracking::TrackWithMotionModel, src/Tracking.cc:594-606 (the pose optimisation that follows needs g2o and is
 * not part of the path): matcher on the stack, motion-model pose, map points cleared, SearchByProjection(CurrentFrame, LastFrame, 15).
 * velocity16 = mVelocity (4x4 row major).  match_cur[i2] out = last-frame feature whose map point the current keypoint received. */
int ref_track_with_motion_model(void* cur_, void* last_, const float* velocity16, int32_t* match_cur)
{
    RefFrame *cur = (RefFrame*)cur_, *last = (RefFrame*)last_;
    return guarded("TrackWithMotionModel", [&] {
        cur->statics();
        Frame &mCurrentFrame = cur->f, &mLastFrame = last->f;
        cv::Mat mVelocity(4, 4, CV_32F);
        std::memcpy(mVelocity.data, velocity16, 64);
        ORBmatcher matcher(0.9, true);
        mCurrentFrame.mTcw = mVelocity * mLastFrame.mTcw;
        std::fill(mCurrentFrame.mvpMapPoints.begin(), mCurrentFrame.mvpMapPoints.end(), static_cast<MapPoint*>(NULL));
        int nmatches = matcher.SearchByProjection(mCurrentFrame, mLastFrame, 15);
        std::map<MapPoint*, int> idx;
        index_of(mLastFrame.mvpMapPoints, idx);
        for (int i = 0; i < mCurrentFrame.N; i++) match_cur[i] = mCurrentFrame.mvpMapPoints[i] ? idx.at(mCurrentFrame.mvpMapPoints[i]) : -1;
        return nmatches;
    });
}
/* which implementation of ORBextractor / ORBmatcher / ORBVocabulary this library was compiled against */
const char* ref_glue_flavour(void)
{
#ifdef REF_GLUE_DROPIN
    return "dropin: reference call sites + include/ORBextractor.h, ORBmatcher.h, ORBVocabulary.h of this repository (liborb_b200.so)";
#else
    return "reference: src/ORBextractor.cc, src/ORBmatcher.cc, DBoW2";
#endif
}

/* ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, th, ORBdist), :1622-1746.
 * already_found[i] != 0 puts the KeyFrame's point i into sAlreadyFound.  match_cur[i2] in: >= 0 -> the keypoint already carries some
 * other map point; out: KeyFrame feature index of the point assigned, the input value where it was occupied, else -1.
 * pred_level[i] out: the level the reference derives for point i (:1662-1669, recomputed here with the same member calls) or -1 when
 * the point is absent; the C ABI of the CUDA path takes that level from its caller. */
int ref_search_by_projection_kf(void* cur_, void* kf_, const uint8_t* already_found, float th, int orb_dist, float nnratio, int check_ori,
                                int32_t* match_cur, int32_t* pred_level)
{
    RefFrame *cur = (RefFrame*)cur_, *k = (RefFrame*)kf_;
    return guarded("SearchByProjection(F,KF)", [&] {
        cur->statics();
        KeyFrame* kf = k->keyframe();
        std::vector<MapPoint*> pts = kf->GetMapPointMatches();
        std::map<MapPoint*, int> idx;
        index_of(pts, idx);
        std::set<MapPoint*> found;
        const cv::Mat Rcw = cur->f.mTcw.rowRange(0, 3).colRange(0, 3);
        const cv::Mat tcw = cur->f.mTcw.rowRange(0, 3).col(3);
        const cv::Mat Ow = -Rcw.t() * tcw;
        for (size_t i = 0; i < pts.size(); i++) {
            pred_level[i] = -1;
            if (!pts[i]) continue;
            if (already_found[i]) found.insert(pts[i]);
            cv::Mat x3Dw = pts[i]->GetWorldPos();
            float minDistance = pts[i]->GetMinDistanceInvariance();
            cv::Mat PO = x3Dw - Ow;
            float dist3D = cv::norm(PO);
            float ratio = dist3D / minDistance;
            std::vector<float>::iterator it = std::lower_bound(cur->f.mvScaleFactors.begin(), cur->f.mvScaleFactors.end(), ratio);
            pred_level[i] = std::min(static_cast<int>(it - cur->f.mvScaleFactors.begin()), cur->f.mnScaleLevels - 1);
        }
        std::vector<MapPoint*> other(cur->f.N, nullptr);
        for (int i = 0; i < cur->f.N; i++) {
            other[i] = match_cur[i] >= 0 ? cur->new_point(nullptr, cur->f.mDescriptors.ptr(i)) : nullptr;
            cur->f.mvpMapPoints[i] = other[i];
        }
        ORBmatcher m(nnratio, check_ori != 0);
        const int n = m.SearchByProjection(cur->f, kf, found, th, orb_dist);
        for (int i = 0; i < cur->f.N; i++) {
            MapPoint* p = cur->f.mvpMapPoints[i];
            if (!p) match_cur[i] = -1;
            else if (p != other[i]) match_cur[i] = idx.at(p);
        }
        return n;
    });
}

} // extern "C"
namespace {
/* What the back-end searches compute for every candidate point before their radius search, restated from
 * ORBmatcher::SearchByProjection(KeyFrame*, Scw, ...) :316-362 and ORBmatcher::Fuse :1036-1070 with the same member calls: the C ABI
 * of the CUDA path takes (active, u, v, predicted level) from its caller, so the tests hand these to the oracle and compare what
 * follows (window, level filter, best descriptor, threshold) with what the reference's own function did. */
void project_candidates(KeyFrame* kf, const cv::Mat& Rcw, const cv::Mat& tcw, const cv::Mat& Ow, const std::vector<MapPoint*>& pts,
                        const std::set<MapPoint*>& skip, bool skip_in_kf, uint8_t* active, float* pu, float* pv, int32_t* level)
{
    const float fx = kf->fx, fy = kf->fy, cx = kf->cx, cy = kf->cy;
    const int nMaxLevel = kf->GetScaleLevels() - 1;
    std::vector<float> vfScaleFactors = kf->GetScaleFactors();
    for (size_t i = 0; i < pts.size(); i++) {
        active[i] = 0; pu[i] = pv[i] = 0.f; level[i] = 0;
        MapPoint* pMP = pts[i];
        if (!pMP) continue;
        if (pMP->isBad() || skip.count(pMP) || (skip_in_kf && pMP->IsInKeyFrame(kf))) continue;
        cv::Mat p3Dw = pMP->GetWorldPos();
        cv::Mat p3Dc = Rcw * p3Dw + tcw;
        if (p3Dc.at<float>(2) < 0.0f) continue;
        const float invz = 1 / p3Dc.at<float>(2);
        const float x = p3Dc.at<float>(0) * invz;
        const float y = p3Dc.at<float>(1) * invz;
        const float u = fx * x + cx;
        const float v = fy * y + cy;
        if (!kf->IsInImage(u, v)) continue;
        const float maxDistance = pMP->GetMaxDistanceInvariance();
        const float minDistance = pMP->GetMinDistanceInvariance();
        cv::Mat PO = p3Dw - Ow;
        const float dist = cv::norm(PO);
        if (dist < minDistance || dist > maxDistance) continue;
        cv::Mat Pn = pMP->GetNormal();
        if (PO.dot(Pn) < 0.5 * dist) continue;
        const float ratio = dist / minDistance;
        std::vector<float>::iterator it = std::lower_bound(vfScaleFactors.begin(), vfScaleFactors.end(), ratio);
        active[i] = 1; pu[i] = u; pv[i] = v;
        level[i] = std::min(static_cast<int>(it - vfScaleFactors.begin()), nMaxLevel);
    }
}
}

extern "C" {

/* ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, vpPoints, vpMatched, th), :286-407.  The candidates are the map points
 * of `src` (its feature i <-> point i).  matched[idx] in: >= 0 -> vpMatched[idx] already holds some other point; out: source index of
 * the point assigned, the input where it was occupied, else -1.  active/u/v/level: see project_candidates. */
int ref_search_by_projection_sim3(void* kf_, void* src_, const float* Scw16, int th, int32_t* matched, uint8_t* active, float* u, float* v,
                                  int32_t* level)
{
    RefFrame *k = (RefFrame*)kf_, *src = (RefFrame*)src_;
    return guarded("SearchByProjection(KF,Scw)", [&] {
        k->statics();
        KeyFrame* kf = k->keyframe();
        cv::Mat Scw(4, 4, CV_32F);
        std::memcpy(Scw.data, Scw16, 64);
        std::vector<MapPoint*> pts = src->f.mvpMapPoints;
        std::map<MapPoint*, int> idx;
        index_of(pts, idx);
        std::vector<MapPoint*> cand;                       /* vpPoints holds no NULLs in the reference (:313 dereferences) */
        std::vector<int> cand_src;
        for (size_t i = 0; i < pts.size(); i++) if (pts[i]) { cand.push_back(pts[i]); cand_src.push_back((int)i); }
        std::vector<MapPoint*> vpMatched(k->f.N, nullptr), other(k->f.N, nullptr);
        for (int i = 0; i < k->f.N; i++) if (matched[i] >= 0) vpMatched[i] = other[i] = k->new_point(nullptr, k->f.mDescriptors.ptr(i));
        {
            cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
            const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
            cv::Mat Rcw = sRcw / scw;
            cv::Mat tcw = Scw.rowRange(0, 3).col(3) / scw;
            cv::Mat Ow = -Rcw.t() * tcw;
            std::set<MapPoint*> found(vpMatched.begin(), vpMatched.end());
            found.erase(nullptr);
            std::vector<uint8_t> a(cand.size()); std::vector<float> uu(cand.size()), vv(cand.size()); std::vector<int32_t> ll(cand.size());
            project_candidates(kf, Rcw, tcw, Ow, cand, found, false, a.data(), uu.data(), vv.data(), ll.data());
            for (size_t i = 0; i < pts.size(); i++) { active[i] = 0; u[i] = v[i] = 0.f; level[i] = 0; }
            for (size_t c = 0; c < cand.size(); c++) { const int i = cand_src[c]; active[i] = a[c]; u[i] = uu[c]; v[i] = vv[c]; level[i] = ll[c]; }
        }
        ORBmatcher m(0.6f, true);
        const int n = m.SearchByProjection(kf, Scw, cand, vpMatched, th);
        for (int i = 0; i < k->f.N; i++) {
            if (!vpMatched[i]) matched[i] = -1;
            else if (vpMatched[i] != other[i]) matched[i] = idx.at(vpMatched[i]);
        }
        return n;
    });
}

/* ORBmatcher::Fuse(KeyFrame* pKF, vector<MapPoint*>&, th), :1016-1134, called with one candidate at a time (the choice of keypoint
 * does not depend on the bookkeeping of earlier candidates: Fuse keeps no claims), so that the keypoint each point fused with can be
 * read back: fused[i] = keypoint of pKF that source point i was merged into (Replace) or attached to (AddObservation), else -1. */
int ref_fuse(void* kf_, void* src_, float th, int32_t* fused, uint8_t* active, float* u, float* v, int32_t* level)
{
    RefFrame *k = (RefFrame*)kf_, *src = (RefFrame*)src_;
    return guarded("Fuse", [&] {
        k->statics();
        KeyFrame* kf = k->keyframe();
        KeyFrame* skf = src->keyframe();
        std::vector<MapPoint*> pts = src->f.mvpMapPoints;
        project_candidates(kf, kf->GetRotation(), kf->GetTranslation(), kf->GetCameraCenter(), pts, std::set<MapPoint*>(), true,
                           active, u, v, level);
        ORBmatcher m(0.6f, true);
        int total = 0;
        for (size_t i = 0; i < pts.size(); i++) {
            fused[i] = -1;
            if (!pts[i]) continue;
            std::vector<MapPoint*> one(1, pts[i]);
            const int n = m.Fuse(kf, one, th);
            if (n != 1) continue;
            total++;
            if (!pts[i]->isBad()) fused[i] = pts[i]->GetIndexInKeyFrame(kf);          /* AddObservation(pKF, bestIdx) */
            else {                                                                    /* Replace(pMPinKF): it took over the observation (skf, i) */
                std::vector<MapPoint*> in_kf = kf->GetMapPointMatches();
                for (size_t j = 0; j < in_kf.size(); j++)
                    if (in_kf[j] && in_kf[j] != pts[i] && in_kf[j]->IsInKeyFrame(skf) && in_kf[j]->GetIndexInKeyFrame(skf) == (int)i) { fused[i] = (int)j; break; }
            }
        }
        return total;
    });
}

} // extern "C"
extern "C" {

/* ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th), :1267-1505.  matches12[i1] in: >= 0 -> vpMatches12[i1] already
 * holds pKF2's point of that feature; out: feature of pKF2 whose point was matched, else -1.  act/u/v/level 12: the points of pKF1
 * projected into pKF2 (:1316-1353), 21: the other way round (:1398-1433), restated with the same expressions. */
int ref_search_by_sim3(void* k1_, void* k2_, float s12, const float* R12_9, const float* t12_3, float th, int32_t* matches12,
                       uint8_t* act12, float* u12, float* v12, int32_t* l12, uint8_t* act21, float* u21, float* v21, int32_t* l21)
{
    RefFrame *a = (RefFrame*)k1_, *b = (RefFrame*)k2_;
    return guarded("SearchBySim3", [&] {
        a->statics();
        KeyFrame *pKF1 = a->keyframe(), *pKF2 = b->keyframe();
        cv::Mat R12(3, 3, CV_32F), t12(3, 1, CV_32F);
        std::memcpy(R12.data, R12_9, 36);
        std::memcpy(t12.data, t12_3, 12);
        std::vector<MapPoint*> vp1 = pKF1->GetMapPointMatches(), vp2 = pKF2->GetMapPointMatches();
        const int N1 = (int)vp1.size(), N2 = (int)vp2.size();
        std::map<MapPoint*, int> idx2;
        index_of(vp2, idx2);
        std::vector<MapPoint*> vpMatches12(N1, nullptr);
        std::vector<bool> done1(N1, false), done2(N2, false);
        for (int i = 0; i < N1; i++)
            if (matches12[i] >= 0 && vp2[matches12[i]]) { vpMatches12[i] = vp2[matches12[i]]; done1[i] = true; done2[matches12[i]] = true; }
        {
            const float fx = pKF1->fx, fy = pKF1->fy, cx = pKF1->cx, cy = pKF1->cy;
            cv::Mat R1w = pKF1->GetRotation(), t1w = pKF1->GetTranslation(), R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
            cv::Mat sR12 = s12 * R12;
            cv::Mat sR21 = (1.0 / s12) * R12.t();
            cv::Mat t21 = -sR21 * t12;
            for (int dir = 0; dir < 2; dir++) {
                const std::vector<MapPoint*>& src = dir == 0 ? vp1 : vp2;
                KeyFrame* dst = dir == 0 ? pKF2 : pKF1;
                const std::vector<bool>& done = dir == 0 ? done1 : done2;
                uint8_t* act = dir == 0 ? act12 : act21; float* pu = dir == 0 ? u12 : u21; float* pv = dir == 0 ? v12 : v21; int32_t* pl = dir == 0 ? l12 : l21;
                const int nMaxLevel = dst->GetScaleLevels() - 1;
                std::vector<float> vfScaleFactors = dst->GetScaleFactors();
                for (size_t i = 0; i < src.size(); i++) {
                    act[i] = 0; pu[i] = pv[i] = 0.f; pl[i] = 0;
                    MapPoint* pMP = src[i];
                    if (!pMP || done[i]) continue;
                    if (pMP->isBad()) continue;
                    cv::Mat p3Dw = pMP->GetWorldPos();
                    cv::Mat p3Dc;
                    if (dir == 0) { cv::Mat p3Dc1 = R1w * p3Dw + t1w; p3Dc = sR21 * p3Dc1 + t21; }
                    else { cv::Mat p3Dc2 = R2w * p3Dw + t2w; p3Dc = sR12 * p3Dc2 + t12; }
                    if (p3Dc.at<float>(2) < 0.0) continue;
                    float invz = 1.0 / p3Dc.at<float>(2);
                    float x = p3Dc.at<float>(0) * invz;
                    float y = p3Dc.at<float>(1) * invz;
                    float u = fx * x + cx;
                    float v = fy * y + cy;
                    if (!dst->IsInImage(u, v)) continue;
                    float maxDistance = pMP->GetMaxDistanceInvariance();
                    float minDistance = pMP->GetMinDistanceInvariance();
                    float dist3D = cv::norm(p3Dc);
                    if (dist3D < minDistance || dist3D > maxDistance) continue;
                    float ratio = dist3D / minDistance;
                    std::vector<float>::iterator it = std::lower_bound(vfScaleFactors.begin(), vfScaleFactors.end(), ratio);
                    act[i] = 1; pu[i] = u; pv[i] = v;
                    pl[i] = std::min(static_cast<int>(it - vfScaleFactors.begin()), nMaxLevel);
                }
            }
        }
        ORBmatcher m(0.6f, true);
        const int n = m.SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th);
        for (int i = 0; i < N1; i++) matches12[i] = vpMatches12[i] ? idx2.at(vpMatches12[i]) : -1;
        return n;
    });
}

} // extern "C"
extern "C" {

/* ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>&, th), :1136-1265 (loop correction), one candidate per call
 * like ref_fuse; here the keyframe's point is replaced BY the candidate (:1249), so the candidate always ends up observing pKF at the
 * keypoint it fused with. */
int ref_fuse_sim3(void* kf_, void* src_, const float* Scw16, float th, int32_t* fused, uint8_t* active, float* u, float* v, int32_t* level)
{
    RefFrame *k = (RefFrame*)kf_, *src = (RefFrame*)src_;
    return guarded("Fuse(Scw)", [&] {
        k->statics();
        KeyFrame* kf = k->keyframe();
        cv::Mat Scw(4, 4, CV_32F);
        std::memcpy(Scw.data, Scw16, 64);
        std::vector<MapPoint*> pts = src->f.mvpMapPoints;
        {
            cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
            const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
            cv::Mat Rcw = sRcw / scw;
            cv::Mat tcw = Scw.rowRange(0, 3).col(3) / scw;
            cv::Mat Ow = -Rcw.t() * tcw;
            project_candidates(kf, Rcw, tcw, Ow, pts, kf->GetMapPoints(), false, active, u, v, level);
        }
        ORBmatcher m(0.6f, true);
        int total = 0;
        for (size_t i = 0; i < pts.size(); i++) {
            fused[i] = -1;
            if (!pts[i]) continue;
            std::vector<MapPoint*> one(1, pts[i]);
            if (m.Fuse(kf, Scw, one, th) != 1) continue;
            total++;
            fused[i] = pts[i]->GetIndexInKeyFrame(kf);
        }
        return total;
    });
}

} // extern "C"
#endif /* REF_GLUE_EXTRACTOR_ONLY */
