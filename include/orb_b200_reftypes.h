/*
 * orb_b200_reftypes.h — the REFERENCE-SIGNATURE ORB_SLAM::ORBmatcher on top of liborb_b200.so.
 *
 * Included by include/ORBmatcher.h when ORB_B200_WITH_REFERENCE_TYPES is defined.  With this repository's include/ directory
 * placed BEFORE the reference's on the include path, the reference's own translation units (src/Frame.cc, src/KeyFrame.cc,
 * src/MapPoint.cc, src/Tracking.cc, src/LocalMapping.cc, src/LoopClosing.cc ...) compile UNCHANGED against it:
 *   ORBmatcher matcher(0.9, true);                                  // include/ORBmatcher.h:41 — no context argument
 *   matcher.SearchByProjection(mCurrentFrame, mLastFrame, 15);      // src/Tracking.cc:605
 *   matcher.SearchByBoW(pKF, mCurrentFrame, vvpMapPointMatches[i]); // src/Tracking.cc:925
 *   ORBmatcher::DescriptorDistance(vDescriptors[i], vDescriptors[j]) // src/MapPoint.cc:224
 * Each method flattens the slice of Frame / KeyFrame / MapPoint the reference implementation reads into plain arrays (exactly the
 * reads of src/ORBmatcher.cc, cited per method), calls the C ABI, and writes the MapPoint* results back where the reference does.
 * The matcher object holds no GPU state: it runs on orb_default_context(), which serves concurrent callers (one lane per call).
 * oracle/Makefile target _ref/libref_dropin.so builds the reference's unmodified sources this way and tests/test_gpu_dropin.py
 * compares it with the reference's own ORBextractor.cc / ORBmatcher.cc (oracle/_ref/libref_orbslam.so) call for call.
 *
 * All fourteen public methods of the reference class are provided.  For the four back-end searches whose per-point preamble is host
 * arithmetic on cv::Mat in the reference (pose / Sim3 transform, depth, image, distance and viewing-angle tests, level prediction:
 * SearchByProjection(Frame&, KeyFrame*, ...), SearchByProjection(KeyFrame*, Scw, ...), SearchBySim3, Fuse x 2) that preamble is
 * restated here with the SAME cv:: expressions (`Rcw*p3Dw+tcw`, `1/z` vs `1.0/z`, `cv::norm(PO)`, `PO.dot(Pn) < 0.5*dist` as each
 * method writes them), so it rounds like the reference in whatever OpenCV the node links; radius search, level filter and Hamming
 * arg-min run on the GPU; Replace / AddObservation / AddMapPoint are applied on the host in the reference's order.
 */
#ifndef ORB_B200_REFTYPES_H
#define ORB_B200_REFTYPES_H

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <set>

#include "MapPoint.h"
#include "KeyFrame.h"
#include "Frame.h"

namespace ORB_SLAM
{
namespace b200
{

inline void keysTo(const std::vector<cv::KeyPoint>& k, std::vector<orb_keypoint>& o)
{
    static_assert(sizeof(cv::KeyPoint) == sizeof(orb_keypoint), "cv::KeyPoint must be the 28-byte record of orb_keypoint");
    o.resize(k.size());
    if (!k.empty()) std::memcpy(static_cast<void*>(o.data()), static_cast<const void*>(k.data()), k.size() * sizeof(orb_keypoint));
}
inline void descTo(const cv::Mat& D, std::vector<unsigned char>& o)
{
    o.resize((size_t)D.rows * 32 + 32);                      // never empty: the ABI wants a pointer even for n = 0
    for (int i = 0; i < D.rows; i++) std::memcpy(&o[(size_t)i * 32], D.ptr<unsigned char>(i), 32);
}
// mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS] -> CSR, cell id = ix * 48 + iy, items in insertion order (src/Frame.cc:109-123)
template <typename Grid> inline void gridTo(const Grid& g, size_t n, std::vector<int32_t>& start, std::vector<int32_t>& items)
{
    start.assign(ORB_GRID_COLS * ORB_GRID_ROWS + 1, 0);
    items.assign(n ? n : 1, 0);
    int k = 0;
    for (int ix = 0; ix < ORB_GRID_COLS; ix++)
        for (int iy = 0; iy < ORB_GRID_ROWS; iy++) {
            start[ix * ORB_GRID_ROWS + iy] = k;
            for (size_t j = 0; j < g[ix][iy].size(); j++) items[k++] = (int32_t)g[ix][iy][j];
        }
    start[ORB_GRID_COLS * ORB_GRID_ROWS] = k;
}
inline void flatten(const Frame& F, FrameArrays& A, bool grid)
{
    keysTo(F.mvKeysUn, A.mvKeysUn);
    descTo(F.mDescriptors, A.mDescriptors);
    A.fx = Frame::fx; A.fy = Frame::fy; A.cx = Frame::cx; A.cy = Frame::cy;                      // static camera, include/Frame.h:62-65
    A.mnMinX = Frame::mnMinX; A.mnMaxX = Frame::mnMaxX; A.mnMinY = Frame::mnMinY; A.mnMaxY = Frame::mnMaxY;
    A.mnScaleLevels = F.mnScaleLevels; A.mfScaleFactor = F.mfScaleFactor;
    if (grid) gridTo(F.mGrid, F.mvKeysUn.size(), A.cellStart, A.cellItems);
}
inline void flatten(KeyFrame* pKF, FrameArrays& A, bool grid = false)
{
    keysTo(pKF->GetKeyPointsUn(), A.mvKeysUn);
    if (grid) gridTo(pKF->GetmGrid(), A.mvKeysUn.size(), A.cellStart, A.cellItems);       // the copy of Frame::mGrid the keyframe keeps
    descTo(pKF->GetDescriptors(), A.mDescriptors);
    A.fx = pKF->fx; A.fy = pKF->fy; A.cx = pKF->cx; A.cy = pKF->cy;
    const std::vector<int> b = pKF->GetMinMaxXY();                                                // minX, minY, maxX, maxY
    A.mnMinX = b[0]; A.mnMinY = b[1]; A.mnMaxX = b[2]; A.mnMaxY = b[3];
    A.mnScaleLevels = pKF->GetScaleLevels(); A.mfScaleFactor = A.mnScaleLevels > 1 ? pKF->GetScaleFactor(1) : 1.0f;
}
struct FeatVec {
    std::vector<int32_t> node, start, items;
    explicit FeatVec(const DBoW2::FeatureVector& fv)
    {
        start.push_back(0);
        for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
            node.push_back((int32_t)it->first);
            for (size_t j = 0; j < it->second.size(); j++) items.push_back((int32_t)it->second[j]);
            start.push_back((int32_t)items.size());
        }
        if (items.empty()) items.push_back(0);
        if (node.empty()) node.push_back(0);
    }
    orb_featvec_view view() const
    {
        orb_featvec_view v = { (int32_t)start.size() - 1, node.data(), start.data(), items.data() };
        return v;
    }
};
inline void validTo(const std::vector<MapPoint*>& pts, std::vector<unsigned char>& valid, bool skipBad)
{
    valid.assign(pts.size() + 1, 0);
    for (size_t i = 0; i < pts.size(); i++) valid[i] = pts[i] && !(skipBad && pts[i]->isBad());
}

// nPredictedLevel of the reference loops: lower_bound of dist / minDistance in the scale factors, capped at the last level
// (src/ORBmatcher.cc:346-347, :1076-1077, :1200-1201, :1357-1358, :1664-1665)
inline int predictLevel(const std::vector<float>& scaleFactors, float ratio, int nMaxLevel)
{
    const int l = (int)(std::lower_bound(scaleFactors.begin(), scaleFactors.end(), ratio) - scaleFactors.begin());
    return l < nMaxLevel ? l : nMaxLevel;
}
// what the scoring loops need of a set of projected map points
struct Projected {
    std::vector<unsigned char> active, desc;
    std::vector<float> u, v;
    std::vector<int32_t> level;
    explicit Projected(size_t n) : active(n + 1, 0), desc(n * 32 + 32, 0), u(n + 1, 0.f), v(n + 1, 0.f), level(n + 1, 0) {}
    void set(size_t i, float u_, float v_, int l, MapPoint* pMP)
    {
        active[i] = 1; u[i] = u_; v[i] = v_; level[i] = l;
        const cv::Mat d = pMP->GetDescriptor();
        std::memcpy(&desc[i * 32], d.ptr<unsigned char>(), 32);
    }
    void trim(size_t n) { active.resize(n); u.resize(n); v.resize(n); level.resize(n); desc.resize(n * 32 + 32); }
};

} // namespace b200

class ORBmatcher
{
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

    // Computes the Hamming distance between two ORB descriptors (src/ORBmatcher.cc:1794-1810)
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) { return orb_descriptor_distance(a.ptr<unsigned char>(), b.ptr<unsigned char>()); }

    // Search matches between Frame keypoints and projected MapPoints (tracking the local map), src/ORBmatcher.cc:49-125
    int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3)
    {
        const size_t n = vpMapPoints.size();
        if (n == 0 || F.mvKeysUn.empty()) return 0;
        std::vector<unsigned char> inView(n), desc(n * 32);
        std::vector<float> px(n), py(n), vc(n);
        std::vector<int32_t> lv(n);
        for (size_t i = 0; i < n; i++) {
            MapPoint* pMP = vpMapPoints[i];
            inView[i] = pMP->mbTrackInView && !pMP->isBad();                                      // :57-61
            if (!inView[i]) continue;
            px[i] = pMP->mTrackProjX; py[i] = pMP->mTrackProjY; lv[i] = pMP->mnTrackScaleLevel; vc[i] = pMP->mTrackViewCos;
            const cv::Mat d = pMP->GetDescriptor();
            std::memcpy(&desc[i * 32], d.ptr<unsigned char>(), 32);
        }
        FrameArrays A;
        b200::flatten(F, A, true);
        std::vector<int32_t> match(F.mvpMapPoints.size());
        for (size_t i = 0; i < match.size(); i++) match[i] = F.mvpMapPoints[i] ? 0 : -1;         // :88-89: taken keypoints are skipped
        std::vector<unsigned char> pre(match.size());
        for (size_t i = 0; i < match.size(); i++) pre[i] = match[i] >= 0;
        ORBmatcherArrays m(orb_default_context(), mfNNratio, mbCheckOrientation);
        const int nm = m.SearchByProjection(A, inView, px, py, lv, vc, desc, th, match);
        for (size_t i = 0; i < match.size(); i++) if (!pre[i] && match[i] >= 0) F.mvpMapPoints[i] = vpMapPoints[match[i]];
        return nm;
    }

    // Project MapPoints tracked in last frame into the current frame and search matches, src/ORBmatcher.cc:1507-1620
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, float th)
    {
        const size_t nl = LastFrame.mvpMapPoints.size(), nc = CurrentFrame.mvpMapPoints.size();
        if (nl == 0 || nc == 0) return 0;
        std::vector<unsigned char> has(nl), outl(nl);
        std::vector<float> xyz(nl * 3, 0.f);
        for (size_t i = 0; i < nl; i++) {
            MapPoint* pMP = LastFrame.mvpMapPoints[i];
            has[i] = pMP != NULL; outl[i] = LastFrame.mvbOutlier[i];                              // :1523-1527
            if (!pMP || outl[i]) continue;
            const cv::Mat p = pMP->GetWorldPos();
            xyz[3 * i] = p.at<float>(0); xyz[3 * i + 1] = p.at<float>(1); xyz[3 * i + 2] = p.at<float>(2);
        }
        FrameArrays C, L;
        b200::flatten(CurrentFrame, C, true);
        b200::flatten(LastFrame, L, false);
        // nPredictedOctave = LastFrame.mvKeys[i].octave (:1546) and the rotation test reads mvKeysUn angles (:1585): same values
        float T[16];
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) T[4 * r + c] = CurrentFrame.mTcw.at<float>(r, c);
        std::vector<int32_t> match(nc);
        std::vector<unsigned char> pre(nc);
        for (size_t i = 0; i < nc; i++) { pre[i] = CurrentFrame.mvpMapPoints[i] != NULL; match[i] = pre[i] ? 0 : -1; }   // :1561-1562
        ORBmatcherArrays m(orb_default_context(), mfNNratio, mbCheckOrientation);
        const int nm = m.SearchByProjection(C, L, has, outl, xyz, T, th, match);
        for (size_t i = 0; i < nc; i++) if (!pre[i]) CurrentFrame.mvpMapPoints[i] = match[i] >= 0 ? LastFrame.mvpMapPoints[match[i]] : static_cast<MapPoint*>(NULL);
        return nm;
    }

    // Brute force constrained to ORB that belong to the same vocabulary node, src/ORBmatcher.cc:155-284
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches)
    {
        const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
        vpMapPointMatches = std::vector<MapPoint*>(F.mvpMapPoints.size(), static_cast<MapPoint*>(NULL));
        if (F.mvKeys.empty() || vpMapPointsKF.empty()) return 0;
        FrameArrays K;
        b200::flatten(pKF, K);
        std::vector<unsigned char> valid, fdesc;
        b200::validTo(vpMapPointsKF, valid, true);                                                // :187-191
        std::vector<orb_keypoint> fkeys;
        b200::keysTo(F.mvKeys, fkeys);                                                            // :232 reads F.mvKeys[...].angle
        b200::descTo(F.mDescriptors, fdesc);
        const b200::FeatVec fvK(pKF->GetFeatureVector()), fvF(F.mFeatVec);
        std::vector<int32_t> matchF;
        ORBmatcherArrays m(orb_default_context(), mfNNratio, mbCheckOrientation);
        const int nm = m.SearchByBoW(fvK.view(), K.mDescriptors.data(), K.mvKeysUn.data(), valid.data(), (int)K.mvKeysUn.size(),
                                     fvF.view(), fdesc.data(), fkeys.data(), (int)fkeys.size(), matchF);
        for (size_t i = 0; i < matchF.size() && i < vpMapPointMatches.size(); i++) if (matchF[i] >= 0) vpMapPointMatches[i] = vpMapPointsKF[matchF[i]];
        return nm;
    }

    // SearchByBoW between two keyframes (loop closing), src/ORBmatcher.cc:715-850
    int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12)
    {
        const std::vector<MapPoint*> vp1 = pKF1->GetMapPointMatches(), vp2 = pKF2->GetMapPointMatches();
        vpMatches12 = std::vector<MapPoint*>(vp1.size(), static_cast<MapPoint*>(NULL));
        if (vp1.empty()) return 0;
        FrameArrays A, B;
        b200::flatten(pKF1, A);
        b200::flatten(pKF2, B);
        std::vector<unsigned char> v1, v2;
        b200::validTo(vp1, v1, true);                                                             // :752-756
        b200::validTo(vp2, v2, true);                                                             // :770-776
        const b200::FeatVec f1(pKF1->GetFeatureVector()), f2(pKF2->GetFeatureVector());
        std::vector<int32_t> m12;
        ORBmatcherArrays m(orb_default_context(), mfNNratio, mbCheckOrientation);
        const int nm = m.SearchByBoW(f1.view(), A.mDescriptors.data(), A.mvKeysUn.data(), v1.data(), (int)A.mvKeysUn.size(),
                                     f2.view(), B.mDescriptors.data(), B.mvKeysUn.data(), v2.data(), (int)B.mvKeysUn.size(), m12);
        for (size_t i = 0; i < m12.size(); i++) if (m12[i] >= 0) vpMatches12[i] = vp2[m12[i]];
        return nm;
    }

    // Search MapPoints tracked in Frame1 in Frame2 in a window centered at their position in Frame1, src/ORBmatcher.cc:409-516
    int WindowSearch(Frame& F1, Frame& F2, int windowSize, std::vector<MapPoint*>& vpMapPointMatches2, int minOctave = -1, int maxOctave = INT_MAX)
    {
        vpMapPointMatches2 = std::vector<MapPoint*>(F2.mvpMapPoints.size(), static_cast<MapPoint*>(NULL));
        if (F1.mvKeysUn.empty() || F2.mvKeysUn.empty()) return 0;
        std::vector<unsigned char> has;
        b200::validTo(F1.mvpMapPoints, has, true);                                                // :424-429
        has.resize(F1.mvKeysUn.size());
        FrameArrays A, B;
        b200::flatten(F1, A, false);
        b200::flatten(F2, B, true);
        std::vector<int32_t> m21;
        ORBmatcherArrays m(orb_default_context(), mfNNratio, mbCheckOrientation);
        const int nm = m.WindowSearch(A, B, windowSize, has, m21, minOctave, maxOctave);
        for (size_t i = 0; i < m21.size(); i++) if (m21[i] >= 0) vpMapPointMatches2[i] = F1.mvpMapPoints[m21[i]];
        return nm;
    }

    // Refined matching when we have a guess of Frame 2 pose, src/ORBmatcher.cc:519-594
    int SearchByProjection(Frame& F1, Frame& F2, int windowSize, std::vector<MapPoint*>& vpMapPointMatches2)
    {
        vpMapPointMatches2 = F2.mvpMapPoints;
        const size_t n1 = F1.mvpMapPoints.size();
        if (n1 == 0 || F2.mvKeysUn.empty()) return 0;
        const std::set<MapPoint*> found(vpMapPointMatches2.begin(), vpMapPointMatches2.end());
        std::vector<unsigned char> active(n1);
        std::vector<float> xyz(n1 * 3, 0.f);
        for (size_t i = 0; i < n1; i++) {
            MapPoint* p = F1.mvpMapPoints[i];
            active[i] = p && !p->isBad() && !found.count(p);                                      // :533-538
            if (!active[i]) continue;
            const cv::Mat w = p->GetWorldPos();
            xyz[3 * i] = w.at<float>(0); xyz[3 * i + 1] = w.at<float>(1); xyz[3 * i + 2] = w.at<float>(2);
        }
        FrameArrays A, B;
        b200::flatten(F1, A, false);
        b200::flatten(F2, B, true);
        float T[16];
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) T[4 * r + c] = F2.mTcw.at<float>(r, c);
        std::vector<int32_t> m2(vpMapPointMatches2.size());
        std::vector<unsigned char> pre(m2.size());
        for (size_t i = 0; i < m2.size(); i++) { pre[i] = vpMapPointMatches2[i] != NULL; m2[i] = pre[i] ? 0 : -1; }
        ORBmatcherArrays m(orb_default_context(), mfNNratio, mbCheckOrientation);
        const int nm = m.SearchByProjection(A, B, windowSize, active, xyz, T, m2);
        for (size_t i = 0; i < m2.size(); i++) if (!pre[i] && m2[i] >= 0) vpMapPointMatches2[i] = F1.mvpMapPoints[m2[i]];
        return nm;
    }

    // Matching for the Map Initialization, src/ORBmatcher.cc:598-713
    int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10)
    {
        vnMatches12 = std::vector<int>(F1.mvKeysUn.size(), -1);
        if (F1.mvKeysUn.empty()) return 0;
        FrameArrays A, B;
        b200::flatten(F1, A, false);
        b200::flatten(F2, B, true);
        std::vector<float> prev(2 * F1.mvKeysUn.size(), 0.f);
        for (size_t i = 0; i < vbPrevMatched.size() && i < F1.mvKeysUn.size(); i++) { prev[2 * i] = vbPrevMatched[i].x; prev[2 * i + 1] = vbPrevMatched[i].y; }
        std::vector<int32_t> m12;
        ORBmatcherArrays m(orb_default_context(), mfNNratio, mbCheckOrientation);
        const int nm = m.SearchForInitialization(A, B, prev, m12, windowSize);
        for (size_t i = 0; i < m12.size(); i++) vnMatches12[i] = m12[i];
        for (size_t i = 0; i < vbPrevMatched.size() && i < F1.mvKeysUn.size(); i++) { vbPrevMatched[i].x = prev[2 * i]; vbPrevMatched[i].y = prev[2 * i + 1]; }   // :703-706
        return nm;
    }

    // Matching to triangulate new MapPoints. Check Epipolar Constraint, src/ORBmatcher.cc:852-1014
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<cv::KeyPoint>& vMatchedKeys1,
                               std::vector<cv::KeyPoint>& vMatchedKeys2, std::vector<std::pair<size_t, size_t> >& vMatchedPairs)
    {
        const std::vector<MapPoint*> vp1 = pKF1->GetMapPointMatches(), vp2 = pKF2->GetMapPointMatches();
        const std::vector<cv::KeyPoint> k1 = pKF1->GetKeyPointsUn(), k2 = pKF2->GetKeyPointsUn();
        vMatchedKeys1.clear(); vMatchedKeys2.clear(); vMatchedPairs.clear();
        if (k1.empty() || k2.empty()) return 0;
        FrameArrays A, B;
        b200::flatten(pKF1, A);
        b200::flatten(pKF2, B);
        std::vector<unsigned char> has1, has2;
        b200::validTo(vp1, has1, false);                                                          // :890-894: any map point, bad or not
        b200::validTo(vp2, has2, false);
        has1.resize(k1.size() + 1); has2.resize(k2.size() + 1);
        float F[9];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) F[3 * r + c] = F12.at<float>(r, c);
        const b200::FeatVec f1(pKF1->GetFeatureVector()), f2(pKF2->GetFeatureVector());
        ORBmatcherArrays m(orb_default_context(), mfNNratio, mbCheckOrientation);
        const int nm = m.SearchForTriangulation(f1.view(), A, has1, f2.view(), B, has2, F, pKF2->GetVectorScaleSigma2(), vMatchedPairs);
        for (size_t i = 0; i < vMatchedPairs.size(); i++) {                                       // :1002-1010
            vMatchedKeys1.push_back(k1[vMatchedPairs[i].first]);
            vMatchedKeys2.push_back(k2[vMatchedPairs[i].second]);
        }
        return nm;
    }

    // Project MapPoints seen in KeyFrame into the Frame and search matches (relocalisation), src/ORBmatcher.cc:1622-1746.
    // The pose projection and the bounds test run on the device (same FP32 / FP64 mix as :1648-1661); the level prediction of
    // :1663-1669 is host arithmetic on cv::Mat exactly as written there.
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, float th, int ORBdist)
    {
        const std::vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();
        const size_t n = vpMPs.size();
        if (n == 0 || CurrentFrame.mvKeysUn.empty()) return 0;
        const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
        const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
        const cv::Mat Ow = -Rcw.t() * tcw;
        std::vector<unsigned char> active(n, 0), desc(n * 32 + 32, 0);
        std::vector<float> xyz(n * 3, 0.f), kfAngle(n, 0.f);
        std::vector<int32_t> lv(n, 0);
        const std::vector<cv::KeyPoint> kfKeys = pKF->GetKeyPointsUn();
        for (size_t i = 0; i < n; i++) {
            MapPoint* pMP = vpMPs[i];
            if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;                       // :1641-1645
            active[i] = 1;
            cv::Mat x3Dw = pMP->GetWorldPos();
            xyz[3 * i] = x3Dw.at<float>(0); xyz[3 * i + 1] = x3Dw.at<float>(1); xyz[3 * i + 2] = x3Dw.at<float>(2);
            float minDistance = pMP->GetMinDistanceInvariance();
            cv::Mat PO = x3Dw - Ow;
            float dist3D = cv::norm(PO);
            float ratio = dist3D / minDistance;
            lv[i] = b200::predictLevel(CurrentFrame.mvScaleFactors, ratio, CurrentFrame.mnScaleLevels - 1);
            const cv::Mat d = pMP->GetDescriptor();
            std::memcpy(&desc[i * 32], d.ptr<unsigned char>(), 32);
            kfAngle[i] = kfKeys[i].angle;                                                         // :1705
        }
        FrameArrays C;
        b200::flatten(CurrentFrame, C, true);
        float T[16];
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) T[4 * r + c] = CurrentFrame.mTcw.at<float>(r, c);
        const size_t nc = CurrentFrame.mvpMapPoints.size();
        std::vector<int32_t> match(nc);
        std::vector<unsigned char> pre(nc);
        for (size_t i = 0; i < nc; i++) { pre[i] = CurrentFrame.mvpMapPoints[i] != NULL; match[i] = pre[i] ? 0 : -1; }     // :1683-1684
        ORBmatcherArrays m(orb_default_context(), mfNNratio, mbCheckOrientation);
        const int nm = m.SearchByProjection(C, active, xyz, T, lv, desc, kfAngle, th, ORBdist, match);
        for (size_t i = 0; i < nc; i++) if (!pre[i]) CurrentFrame.mvpMapPoints[i] = match[i] >= 0 ? vpMPs[match[i]] : static_cast<MapPoint*>(NULL);
        return nm;
    }

    // Project MapPoints using a Similarity Transformation and search matches (loop detection), src/ORBmatcher.cc:286-407
    int SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched, int th)
    {
        const float fx = pKF->fx, fy = pKF->fy, cx = pKF->cx, cy = pKF->cy;
        const int nMaxLevel = pKF->GetScaleLevels() - 1;
        const std::vector<float> vfScaleFactors = pKF->GetScaleFactors();
        cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);                                         // :296-302, as written
        const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
        cv::Mat Rcw = sRcw / scw;
        cv::Mat tcw = Scw.rowRange(0, 3).col(3) / scw;
        cv::Mat Ow = -Rcw.t() * tcw;
        std::set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
        spAlreadyFound.erase(static_cast<MapPoint*>(NULL));
        const size_t n = vpPoints.size();
        if (n == 0 || vpMatched.empty()) return 0;
        b200::Projected P(n);
        for (size_t i = 0; i < n; i++) {
            MapPoint* pMP = vpPoints[i];
            if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
            cv::Mat p3Dw = pMP->GetWorldPos();
            cv::Mat p3Dc = Rcw * p3Dw + tcw;
            if (p3Dc.at<float>(2) < 0.0) continue;
            const float invz = 1 / p3Dc.at<float>(2);
            const float x = p3Dc.at<float>(0) * invz;
            const float y = p3Dc.at<float>(1) * invz;
            const float u = fx * x + cx;
            const float v = fy * y + cy;
            if (!pKF->IsInImage(u, v)) continue;
            const float maxDistance = pMP->GetMaxDistanceInvariance();
            const float minDistance = pMP->GetMinDistanceInvariance();
            cv::Mat PO = p3Dw - Ow;
            const float dist = cv::norm(PO);
            if (dist < minDistance || dist > maxDistance) continue;
            cv::Mat Pn = pMP->GetNormal();
            if (PO.dot(Pn) < 0.5 * dist) continue;
            P.set(i, u, v, b200::predictLevel(vfScaleFactors, dist / minDistance, nMaxLevel), pMP);
        }
        P.trim(n);
        FrameArrays K;
        b200::flatten(pKF, K, true);
        std::vector<int32_t> matched(vpMatched.size());
        std::vector<unsigned char> pre(vpMatched.size());
        for (size_t i = 0; i < matched.size(); i++) { pre[i] = vpMatched[i] != NULL; matched[i] = pre[i] ? 0 : -1; }      // :376-377
        ORBmatcherArrays m(orb_default_context(), mfNNratio, mbCheckOrientation);
        const int nm = m.SearchByProjection(K, P.active, P.u, P.v, P.level, P.desc, th, matched);
        for (size_t i = 0; i < matched.size(); i++) if (!pre[i] && matched[i] >= 0) vpMatched[i] = vpPoints[matched[i]];
        return nm;
    }

    // Search matches between MapPoints seen in KF1 and KF2 transforming by a Sim3 [s12*R12|t12], src/ORBmatcher.cc:1267-1505
    int SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12, const cv::Mat& t12, float th)
    {
        const float fx = pKF1->fx, fy = pKF1->fy, cx = pKF1->cx, cy = pKF1->cy;
        cv::Mat R1w = pKF1->GetRotation(), t1w = pKF1->GetTranslation();
        cv::Mat R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
        cv::Mat sR12 = s12 * R12;                                                                 // :1283-1285, as written
        cv::Mat sR21 = (1.0 / s12) * R12.t();
        cv::Mat t21 = -sR21 * t12;
        const int nMaxLevel1 = pKF1->GetScaleLevels() - 1, nMaxLevel2 = pKF2->GetScaleLevels() - 1;
        const std::vector<float> sf1 = pKF1->GetScaleFactors(), sf2 = pKF2->GetScaleFactors();
        const std::vector<MapPoint*> vp1 = pKF1->GetMapPointMatches(), vp2 = pKF2->GetMapPointMatches();
        const int N1 = (int)vp1.size(), N2 = (int)vp2.size();
        if (N1 == 0 || N2 == 0) return 0;
        std::vector<bool> done1(N1, false), done2(N2, false);
        for (int i = 0; i < N1; i++) {                                                            // :1301-1311
            MapPoint* pMP = vpMatches12[i];
            if (!pMP) continue;
            done1[i] = true;
            const int idx2 = pMP->GetIndexInKeyFrame(pKF2);
            if (idx2 >= 0 && idx2 < N2) done2[idx2] = true;
        }
        b200::Projected P12(N1), P21(N2);
        for (int i1 = 0; i1 < N1; i1++) {                                                         // KF1's points into KF2, :1317-1360
            MapPoint* pMP = vp1[i1];
            if (!pMP || done1[i1] || pMP->isBad()) continue;
            cv::Mat p3Dw = pMP->GetWorldPos();
            cv::Mat p3Dc1 = R1w * p3Dw + t1w;
            cv::Mat p3Dc2 = sR21 * p3Dc1 + t21;
            if (p3Dc2.at<float>(2) < 0.0) continue;
            float invz = 1.0 / p3Dc2.at<float>(2);
            float x = p3Dc2.at<float>(0) * invz;
            float y = p3Dc2.at<float>(1) * invz;
            float u = fx * x + cx;
            float v = fy * y + cy;
            if (!pKF2->IsInImage(u, v)) continue;
            float maxDistance = pMP->GetMaxDistanceInvariance();
            float minDistance = pMP->GetMinDistanceInvariance();
            float dist3D = cv::norm(p3Dc2);
            if (dist3D < minDistance || dist3D > maxDistance) continue;
            P12.set(i1, u, v, b200::predictLevel(sf2, dist3D / minDistance, nMaxLevel2), pMP);
        }
        for (int i2 = 0; i2 < N2; i2++) {                                                         // KF2's points into KF1, :1402-1445
            MapPoint* pMP = vp2[i2];
            if (!pMP || done2[i2] || pMP->isBad()) continue;
            cv::Mat p3Dw = pMP->GetWorldPos();
            cv::Mat p3Dc2 = R2w * p3Dw + t2w;
            cv::Mat p3Dc1 = sR12 * p3Dc2 + t12;
            if (p3Dc1.at<float>(2) < 0.0) continue;
            float invz = 1.0 / p3Dc1.at<float>(2);
            float x = p3Dc1.at<float>(0) * invz;
            float y = p3Dc1.at<float>(1) * invz;
            float u = fx * x + cx;
            float v = fy * y + cy;
            if (!pKF1->IsInImage(u, v)) continue;
            float maxDistance = pMP->GetMaxDistanceInvariance();
            float minDistance = pMP->GetMinDistanceInvariance();
            float dist3D = cv::norm(p3Dc1);
            if (dist3D < minDistance || dist3D > maxDistance) continue;
            P21.set(i2, u, v, b200::predictLevel(sf1, dist3D / minDistance, nMaxLevel1), pMP);
        }
        P12.trim(N1); P21.trim(N2);
        FrameArrays K1, K2;
        b200::flatten(pKF1, K1, true);
        b200::flatten(pKF2, K2, true);
        std::vector<int32_t> m12;
        ORBmatcherArrays m(orb_default_context(), mfNNratio, mbCheckOrientation);
        const int nFound = m.SearchBySim3(K1, K2, P12.active, P12.u, P12.v, P12.level, P12.desc, P21.active, P21.u, P21.v, P21.level, P21.desc, th, m12);
        for (int i1 = 0; i1 < N1; i1++) if (m12[i1] >= 0) vpMatches12[i1] = vp2[m12[i1]];      // :1488-1493
        return nFound;
    }

    // Project MapPoints into KeyFrame and search for duplicated MapPoints, src/ORBmatcher.cc:1016-1134.  The scoring loop reads no
    // assignment, so all candidates are scored in one call; the skip tests and the Replace / AddObservation bookkeeping then run
    // point by point in the reference's order (they see each other's effects, like :1036-1040 and :1116-1128 do).
    int Fuse(KeyFrame* pKF, std::vector<MapPoint*>& vpMapPoints, float th = 2.5)
    {
        cv::Mat Rcw = pKF->GetRotation();
        cv::Mat tcw = pKF->GetTranslation();
        const float &fx = pKF->fx, &fy = pKF->fy, &cx = pKF->cx, &cy = pKF->cy;
        const int nMaxLevel = pKF->GetScaleLevels() - 1;
        const std::vector<float> vfScaleFactors = pKF->GetScaleFactors();
        cv::Mat Ow = pKF->GetCameraCenter();
        const size_t n = vpMapPoints.size();
        if (n == 0) return 0;
        b200::Projected P(n);
        for (size_t i = 0; i < n; i++) {
            MapPoint* pMP = vpMapPoints[i];
            if (!pMP) continue;
            cv::Mat p3Dw = pMP->GetWorldPos();
            cv::Mat p3Dc = Rcw * p3Dw + tcw;
            if (p3Dc.at<float>(2) < 0.0f) continue;
            const float invz = 1 / p3Dc.at<float>(2);
            const float x = p3Dc.at<float>(0) * invz;
            const float y = p3Dc.at<float>(1) * invz;
            const float u = fx * x + cx;
            const float v = fy * y + cy;
            if (!pKF->IsInImage(u, v)) continue;
            const float maxDistance = pMP->GetMaxDistanceInvariance();
            const float minDistance = pMP->GetMinDistanceInvariance();
            cv::Mat PO = p3Dw - Ow;
            const float dist3D = cv::norm(PO);
            if (dist3D < minDistance || dist3D > maxDistance) continue;
            cv::Mat Pn = pMP->GetNormal();
            if (PO.dot(Pn) < 0.5 * dist3D) continue;
            P.set(i, u, v, b200::predictLevel(vfScaleFactors, dist3D / minDistance, nMaxLevel), pMP);
        }
        P.trim(n);
        FrameArrays K;
        b200::flatten(pKF, K, true);
        std::vector<int32_t> fuseIdx;
        ORBmatcherArrays m(orb_default_context(), mfNNratio, mbCheckOrientation);
        m.FuseCandidates(K, P.active, P.u, P.v, P.level, P.desc, th, fuseIdx);
        int nFused = 0;
        for (size_t i = 0; i < n; i++) {
            MapPoint* pMP = vpMapPoints[i];
            if (!pMP || pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;                         // :1033-1040, evaluated NOW
            if (!P.active[i] || fuseIdx[i] < 0) continue;
            const int bestIdx = fuseIdx[i];
            MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx);                                        // :1116-1128
            if (pMPinKF) { if (!pMPinKF->isBad()) pMP->Replace(pMPinKF); }
            else { pMP->AddObservation(pKF, bestIdx); pKF->AddMapPoint(pMP, bestIdx); }
            nFused++;
        }
        return nFused;
    }

    // Project MapPoints into KeyFrame using a given Sim3 and search for duplicated MapPoints, src/ORBmatcher.cc:1136-1265
    int Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th = 2.5)
    {
        const float &fx = pKF->fx, &fy = pKF->fy, &cx = pKF->cx, &cy = pKF->cy;
        cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);                                         // :1144-1149, as written
        const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
        cv::Mat Rcw = sRcw / scw;
        cv::Mat tcw = Scw.rowRange(0, 3).col(3) / scw;
        cv::Mat Ow = -Rcw.t() * tcw;
        const std::set<MapPoint*> spAlreadyFound = pKF->GetMapPoints();
        const int nMaxLevel = pKF->GetScaleLevels() - 1;
        const std::vector<float> vfScaleFactors = pKF->GetScaleFactors();
        const size_t n = vpPoints.size();
        if (n == 0) return 0;
        b200::Projected P(n);
        for (size_t i = 0; i < n; i++) {
            MapPoint* pMP = vpPoints[i];
            if (spAlreadyFound.count(pMP)) continue;
            cv::Mat p3Dw = pMP->GetWorldPos();
            cv::Mat p3Dc = Rcw * p3Dw + tcw;
            if (p3Dc.at<float>(2) < 0.0f) continue;
            const float invz = 1.0 / p3Dc.at<float>(2);
            const float x = p3Dc.at<float>(0) * invz;
            const float y = p3Dc.at<float>(1) * invz;
            const float u = fx * x + cx;
            const float v = fy * y + cy;
            if (!pKF->IsInImage(u, v)) continue;
            const float maxDistance = pMP->GetMaxDistanceInvariance();
            const float minDistance = pMP->GetMinDistanceInvariance();
            cv::Mat PO = p3Dw - Ow;
            const float dist3D = cv::norm(PO);
            if (dist3D < minDistance || dist3D > maxDistance) continue;
            cv::Mat Pn = pMP->GetNormal();
            if (PO.dot(Pn) < 0.5 * dist3D) continue;
            P.set(i, u, v, b200::predictLevel(vfScaleFactors, dist3D / minDistance, nMaxLevel), pMP);
        }
        P.trim(n);
        FrameArrays K;
        b200::flatten(pKF, K, true);
        std::vector<int32_t> fuseIdx;
        ORBmatcherArrays m(orb_default_context(), mfNNratio, mbCheckOrientation);
        m.FuseCandidates(K, P.active, P.u, P.v, P.level, P.desc, th, fuseIdx);
        int nFused = 0;
        for (size_t i = 0; i < n; i++) {
            MapPoint* pMP = vpPoints[i];
            if (pMP->isBad()) continue;                                                           // :1165, evaluated NOW (Replace marks points bad)
            if (!P.active[i] || fuseIdx[i] < 0) continue;
            const int bestIdx = fuseIdx[i];
            MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx);                                        // :1247-1259
            if (pMPinKF) { if (!pMPinKF->isBad()) pMPinKF->Replace(pMP); }
            else { pMP->AddObservation(pKF, bestIdx); pKF->AddMapPoint(pMP, bestIdx); }
            nFused++;
        }
        return nFused;
    }

public:
    static constexpr int TH_LOW = 50;          // src/ORBmatcher.cc:40-42
    static constexpr int TH_HIGH = 100;
    static constexpr int HISTO_LENGTH = 30;

protected:
    float mfNNratio;
    bool mbCheckOrientation;
};

} // namespace ORB_SLAM

#endif
