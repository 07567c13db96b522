/*
 * ORBmatcher.h — header-only C++ shim with the reference's class name (caomw/ORBSLAM_jpMiniPC
 * include/ORBmatcher.h:37-107) for the matcher methods on the accelerated path:
 *   DescriptorDistance (src/ORBmatcher.cc:1794-1810)
 *   SearchByProjection(Frame&, const Frame&, th) (:1507-1620)
 *   SearchByBoW(KeyFrame*, Frame&, ...) candidate scoring (:155-284), SearchByBoW(KeyFrame*, KeyFrame*, ...) (:715-850)
 *   SearchByProjection(Frame&, vector<MapPoint*>, th) (:49-125), WindowSearch (:409-516),
 *   SearchByProjection(F1, F2, windowSize, ...) (:519-594), SearchByProjection(Frame&, KeyFrame*, ...) (:1622-1746),
 *   SearchForInitialization (:598-713), SearchByProjection(KeyFrame*, Scw, ...) (:286-407), the scoring loops of Fuse
 *   (:1016-1265) and SearchBySim3 (:1267-1505), SearchForTriangulation (:852-1014)
 * plus the brute-force best/second-best + ratio test used for relocalisation-sized searches.
 * Frame / KeyFrame / MapPoint are the reference's own graph classes and stay on the host.  Two forms:
 *   - ORBmatcherArrays (below): the plain arrays those methods read, usable without any reference header;
 *   - with -DORB_B200_WITH_REFERENCE_TYPES and the reference's include/ directory behind this one on the include path:
 *     ORB_SLAM::ORBmatcher with the reference's exact constructor and method signatures (orb_b200_reftypes.h), so that the
 *     reference's call sites compile unchanged (see INTEGRATION.md and oracle/Makefile target _ref/libref_dropin.so).
 */
#ifndef ORBMATCHER_H
#define ORBMATCHER_H

#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "orb_b200.h"
#if defined(ORB_B200_WITH_REFERENCE_TYPES) && !defined(ORB_B200_WITH_OPENCV)
#define ORB_B200_WITH_OPENCV
#endif

namespace ORB_SLAM
{

// the slice of ORB_SLAM::Frame the matcher reads; grid is built on the GPU (src/Frame.cc:109-123)
struct FrameArrays {
    std::vector<orb_keypoint> mvKeysUn;
    std::vector<unsigned char> mDescriptors;          // N x 32
    float fx, fy, cx, cy;
    int mnMinX, mnMaxX, mnMinY, mnMaxY;
    int mnScaleLevels; float mfScaleFactor;
    std::vector<int32_t> cellStart, cellItems;        // CSR of mGrid[64][48]

    void AssignFeaturesToGrid(orb_ctx* ctx)
    {
        cellStart.assign(ORB_GRID_COLS * ORB_GRID_ROWS + 1, 0);
        cellItems.assign(mvKeysUn.empty() ? 1 : mvKeysUn.size(), 0);
        int rc = orb_frame_grid_build(ctx, mvKeysUn.data(), (int)mvKeysUn.size(), mnMinX, mnMaxX, mnMinY, mnMaxY,
                                      cellStart.data(), cellItems.data());
        if (rc != ORB_OK) throw std::runtime_error(std::string("Frame grid: ") + orb_error_string(rc));
    }
    orb_frame_view view() const
    {
        orb_frame_view v;
        v.n = (int32_t)mvKeysUn.size(); v.kps = mvKeysUn.data(); v.desc = mDescriptors.data();
        v.fx = fx; v.fy = fy; v.cx = cx; v.cy = cy;
        v.min_x = mnMinX; v.max_x = mnMaxX; v.min_y = mnMinY; v.max_y = mnMaxY;
        v.nlevels = mnScaleLevels; v.scale_factor = mfScaleFactor;
        v.cell_start = cellStart.data(); v.cell_items = cellItems.data();
        return v;
    }
};

// The matcher on plain arrays (always available).  Without ORB_B200_WITH_REFERENCE_TYPES it is also ORB_SLAM::ORBmatcher; with it,
// ORB_SLAM::ORBmatcher is the reference-signature class of orb_b200_reftypes.h, which flattens Frame / KeyFrame / MapPoint and calls this.
class ORBmatcherArrays
{
public:
    ORBmatcherArrays(orb_ctx* ctx, float nnratio = 0.6, bool checkOri = true) : ctx(ctx), mfNNratio(nnratio), mbCheckOrientation(checkOri)
    {
        if (!ctx) throw std::runtime_error(std::string("ORBmatcher: no context (") + orb_last_cuda_error() + ")");
    }

    // Computes the Hamming distance between two ORB descriptors (32-byte rows)
    static int DescriptorDistance(const unsigned char* a, const unsigned char* b) { return orb_descriptor_distance(a, b); }

    // Project MapPoints tracked in last frame into the current frame and search matches.
    // matchCur[i2] = index of the last-frame feature whose map point now belongs to current keypoint i2, or -1
    // (the caller maps it back: CurrentFrame.mvpMapPoints[i2] = LastFrame.mvpMapPoints[matchCur[i2]]).
    int SearchByProjection(const FrameArrays& CurrentFrame, const FrameArrays& LastFrame,
                           const std::vector<unsigned char>& lastHasMapPoint, const std::vector<unsigned char>& lastOutlier,
                           const std::vector<float>& lastWorldPos /*3 per feature*/, const float* Tcw /*4x4 row-major*/,
                           float th, std::vector<int32_t>& matchCur)
    {
        if (matchCur.size() != CurrentFrame.mvKeysUn.size()) matchCur.assign(CurrentFrame.mvKeysUn.size(), -1);
        orb_frame_view cur = CurrentFrame.view(), last = LastFrame.view();
        int n = 0;
        check(orb_search_by_projection(ctx, &cur, &last, lastHasMapPoint.data(), lastOutlier.data(), lastWorldPos.data(), Tcw, th,
                                       mbCheckOrientation ? 1 : 0, matchCur.data(), &n));
        return n;
    }

    // Brute force constrained to ORB that belong to the same vocabulary node.  FeatureVectors as CSR.
    int SearchByBoW(const orb_featvec_view& kfFeatVec, const unsigned char* kfDesc, const orb_keypoint* kfKeysUn,
                    const unsigned char* kfMapPointValid, int nKF,
                    const orb_featvec_view& fFeatVec, const unsigned char* fDesc, const orb_keypoint* fKeys, int nF,
                    std::vector<int32_t>& matchF)
    {
        matchF.assign(nF, -1);
        int n = 0;
        check(orb_search_by_bow(ctx, &kfFeatVec, kfDesc, kfKeysUn, kfMapPointValid, nKF, &fFeatVec, fDesc, fKeys, nF,
                                mfNNratio, mbCheckOrientation ? 1 : 0, matchF.data(), &n));
        return n;
    }

    // ---- the other windowed searches (all map to orb_search_window, see orb_b200.h) ----
    // Search matches between Frame keypoints and projected MapPoints (src/ORBmatcher.cc:49-125).  Per map point:
    // mbTrackInView && !isBad, mTrackProjX/Y, mnTrackScaleLevel, mTrackViewCos, GetDescriptor().
    int SearchByProjection(const FrameArrays& F, const std::vector<unsigned char>& inView, const std::vector<float>& projX,
                           const std::vector<float>& projY, const std::vector<int32_t>& level, const std::vector<float>& viewCos,
                           const std::vector<unsigned char>& mpDesc, float th, std::vector<int32_t>& matchF)
    {
        const size_t n = inView.size();
        if (matchF.size() != F.mvKeysUn.size()) matchF.assign(F.mvKeysUn.size(), -1);
        std::vector<float> sf(F.mnScaleLevels, 1.0f), radius(n);
        for (int i = 1; i < F.mnScaleLevels; i++) sf[i] = sf[i - 1] * F.mfScaleFactor;
        std::vector<int32_t> lmin(n), lmax(n);
        const bool bFactor = th != 1.0;
        for (size_t i = 0; i < n; i++) {
            float r = viewCos[i] > 0.998 ? 2.5f : 4.0f;               // RadiusByViewingCos, :127-133
            if (bFactor) r *= th;
            radius[i] = r * sf[level[i]];
            lmin[i] = level[i] - 1; lmax[i] = level[i];
        }
        orb_window_query_set q = { (int32_t)n, inView.data(), mpDesc.data(), projX.data(), projY.data(), nullptr, nullptr, 0,
                                   radius.data(), 0.f, lmin.data(), lmax.data(), nullptr };
        return window(F, q, ORB_ACCEPT_LEVEL_RATIO, TH_HIGH, false, matchF);
    }

    // WindowSearch (src/ORBmatcher.cc:409-516): matches21[i2] = i1 or -1
    int WindowSearch(const FrameArrays& F1, const FrameArrays& F2, int windowSize, const std::vector<unsigned char>& f1HasMapPoint,
                     std::vector<int32_t>& matches21, int minScaleLevel = -1, int maxScaleLevel = 2147483647)
    {
        const size_t n = F1.mvKeysUn.size();
        std::vector<unsigned char> active(f1HasMapPoint);
        std::vector<float> u(n), v(n), ang(n);
        std::vector<int32_t> lv(n);
        for (size_t i = 0; i < n; i++) {
            const orb_keypoint& k = F1.mvKeysUn[i];
            u[i] = k.x; v[i] = k.y; ang[i] = k.angle; lv[i] = k.octave;
            if (minScaleLevel > 0 && k.octave < minScaleLevel) active[i] = 0;
            if (maxScaleLevel < 2147483647 && k.octave > maxScaleLevel) active[i] = 0;
        }
        matches21.assign(F2.mvKeysUn.size(), -1);
        orb_window_query_set q = { (int32_t)n, active.data(), F1.mDescriptors.data(), u.data(), v.data(), nullptr, nullptr, 0,
                                   nullptr, (float)windowSize, lv.data(), lv.data(), ang.data() };
        return window(F2, q, ORB_ACCEPT_RATIO, TH_HIGH, mbCheckOrientation, matches21);
    }

    // SearchForInitialization (src/ORBmatcher.cc:598-713): vbPrevMatched holds x,y per F1 keypoint (in/out)
    int SearchForInitialization(const FrameArrays& F1, const FrameArrays& F2, std::vector<float>& vbPrevMatched,
                                std::vector<int32_t>& vnMatches12, int windowSize = 10)
    {
        if (vbPrevMatched.size() != 2 * F1.mvKeysUn.size()) throw std::invalid_argument("vbPrevMatched: one point per F1 keypoint");
        vnMatches12.assign(F1.mvKeysUn.size(), -1);
        orb_frame_view v1 = F1.view(), v2 = F2.view();
        int n = 0;
        check(orb_search_for_initialization(ctx, &v1, &v2, vbPrevMatched.data(), windowSize, mfNNratio, mbCheckOrientation ? 1 : 0,
                                            vnMatches12.data(), &n));
        return n;
    }

    // Refined matching with a pose guess for F2 (src/ORBmatcher.cc:519-594): matches2 starts as F2's own map points
    int SearchByProjection(const FrameArrays& F1, const FrameArrays& F2, int windowSize, const std::vector<unsigned char>& f1Active,
                           const std::vector<float>& f1WorldPos, const float* Tc2w, std::vector<int32_t>& matches2)
    {
        const size_t n = F1.mvKeysUn.size();
        std::vector<int32_t> lv(n);
        for (size_t i = 0; i < n; i++) lv[i] = F1.mvKeysUn[i].octave;
        orb_window_query_set q = { (int32_t)n, f1Active.data(), F1.mDescriptors.data(), nullptr, nullptr, f1WorldPos.data(), Tc2w, 0,
                                   nullptr, (float)windowSize, lv.data(), lv.data(), nullptr };
        return window(F2, q, ORB_ACCEPT_RATIO, TH_HIGH, false, matches2);
    }

    // Project MapPoints seen in a KeyFrame into the Frame (relocalisation, src/ORBmatcher.cc:1622-1746).  predLevel is the
    // level the reference derives from dist3D / minDistance (:1663-1669); kfAngle = pKF->GetKeyPointUn(i).angle.
    int SearchByProjection(const FrameArrays& CurrentFrame, const std::vector<unsigned char>& active, const std::vector<float>& worldPos,
                           const float* Tcw, const std::vector<int32_t>& predLevel, const std::vector<unsigned char>& mpDesc,
                           const std::vector<float>& kfAngle, float th, int ORBdist, std::vector<int32_t>& matchCur)
    {
        const size_t n = active.size();
        if (matchCur.size() != CurrentFrame.mvKeysUn.size()) matchCur.assign(CurrentFrame.mvKeysUn.size(), -1);
        std::vector<float> sf(CurrentFrame.mnScaleLevels, 1.0f), radius(n);
        for (int i = 1; i < CurrentFrame.mnScaleLevels; i++) sf[i] = sf[i - 1] * CurrentFrame.mfScaleFactor;
        std::vector<int32_t> lmin(n), lmax(n);
        for (size_t i = 0; i < n; i++) { radius[i] = th * sf[predLevel[i]]; lmin[i] = predLevel[i] - 1; lmax[i] = predLevel[i] + 1; }
        orb_window_query_set q = { (int32_t)n, active.data(), mpDesc.data(), nullptr, nullptr, worldPos.data(), Tcw, 1,
                                   radius.data(), 0.f, lmin.data(), lmax.data(), kfAngle.data() };
        return window(CurrentFrame, q, ORB_ACCEPT_BEST, ORBdist, mbCheckOrientation, matchCur);
    }

    // SearchByBoW between two keyframes (loop closing, src/ORBmatcher.cc:715-850): matches12[idx1] = idx2 or -1
    int SearchByBoW(const orb_featvec_view& fv1, const unsigned char* desc1, const orb_keypoint* kps1, const unsigned char* valid1, int n1,
                    const orb_featvec_view& fv2, const unsigned char* desc2, const orb_keypoint* kps2, const unsigned char* valid2, int n2,
                    std::vector<int32_t>& matches12)
    {
        matches12.assign(n1, -1);
        int n = 0;
        check(orb_search_by_bow_kf(ctx, &fv1, desc1, kps1, valid1, n1, &fv2, desc2, kps2, valid2, n2, mfNNratio,
                                   mbCheckOrientation ? 1 : 0, matches12.data(), &n));
        return n;
    }

    // SearchForTriangulation (src/ORBmatcher.cc:852-1014): vMatchedPairs = (index in KF1, index in KF2); F12 = 3x3 CV_32F row major,
    // levelSigma2 = pKF2's mvLevelSigma2
    int SearchForTriangulation(const orb_featvec_view& fv1, const FrameArrays& KF1, const std::vector<unsigned char>& hasMapPoint1,
                               const orb_featvec_view& fv2, const FrameArrays& KF2, const std::vector<unsigned char>& hasMapPoint2,
                               const float* F12, const std::vector<float>& levelSigma2,
                               std::vector<std::pair<size_t, size_t> >& vMatchedPairs)
    {
        std::vector<int32_t> m12(KF1.mvKeysUn.size(), -1);
        int n = 0;
        check(orb_search_for_triangulation(ctx, &fv1, KF1.mDescriptors.data(), KF1.mvKeysUn.data(), hasMapPoint1.data(), (int)KF1.mvKeysUn.size(),
                                           &fv2, KF2.mDescriptors.data(), KF2.mvKeysUn.data(), hasMapPoint2.data(), (int)KF2.mvKeysUn.size(),
                                           F12, levelSigma2.data(), (int)levelSigma2.size(), mbCheckOrientation ? 1 : 0, m12.data(), &n));
        vMatchedPairs.clear();
        for (size_t i = 0; i < m12.size(); i++) if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, (size_t)m12[i]));
        return n;
    }

    // ---- back-end searches on map points the caller has already projected (Sim3 / pose projection and level prediction,
    //      src/ORBmatcher.cc:300-349, :1033-1078, :1301-1354, stay in the adapter)
    // SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) (:286-407): matchedKF[idx] >= 0 where vpMatched[idx] is set
    int SearchByProjection(const FrameArrays& KF, const std::vector<unsigned char>& active, const std::vector<float>& u,
                           const std::vector<float>& v, const std::vector<int32_t>& predLevel, const std::vector<unsigned char>& mpDesc,
                           int th, std::vector<int32_t>& matchedKF)
    {
        const size_t n = active.size();
        std::vector<float> radius(n);
        std::vector<int32_t> lmin(n);
        const std::vector<float> sf = scaleFactors(KF);
        for (size_t i = 0; i < n; i++) { radius[i] = (float)th * sf[clampLevel(KF, predLevel[i])]; lmin[i] = predLevel[i] - 1; }
        orb_window_query_set q = { (int32_t)n, active.data(), mpDesc.data(), u.data(), v.data(), nullptr, nullptr, 0,
                                   radius.data(), 0.f, lmin.data(), predLevel.data(), nullptr };
        return window(KF, q, ORB_ACCEPT_BEST, TH_LOW, false, matchedKF);
    }

    // scoring loop of Fuse (:1016-1134, :1136-1265): fuseIdx[i] = KF keypoint map point i fuses with (bestDist <= TH_LOW) or -1
    void FuseCandidates(const FrameArrays& KF, const std::vector<unsigned char>& active, const std::vector<float>& u,
                        const std::vector<float>& v, const std::vector<int32_t>& predLevel, const std::vector<unsigned char>& mpDesc,
                        float th, std::vector<int32_t>& fuseIdx)
    {
        std::vector<int32_t> dist;
        best(KF, active, u, v, predLevel, mpDesc, th, fuseIdx, dist);
        for (size_t i = 0; i < fuseIdx.size(); i++) if (dist[i] > TH_LOW) fuseIdx[i] = -1;
    }

    // SearchBySim3 (:1267-1505): points of KF1 projected into KF2 (u12, v12, level12) and the other way round; matches12[i1] = idx2 or -1
    int SearchBySim3(const FrameArrays& KF1, const FrameArrays& KF2,
                     const std::vector<unsigned char>& active1, const std::vector<float>& u12, const std::vector<float>& v12,
                     const std::vector<int32_t>& level12, const std::vector<unsigned char>& desc1,
                     const std::vector<unsigned char>& active2, const std::vector<float>& u21, const std::vector<float>& v21,
                     const std::vector<int32_t>& level21, const std::vector<unsigned char>& desc2, float th, std::vector<int32_t>& matches12)
    {
        std::vector<int32_t> m1, d1, m2, d2;
        best(KF2, active1, u12, v12, level12, desc1, th, m1, d1);
        best(KF1, active2, u21, v21, level21, desc2, th, m2, d2);
        matches12.assign(m1.size(), -1);
        int nFound = 0;
        for (size_t i1 = 0; i1 < m1.size(); i1++) {                    // agreement, :1478-1493
            const int idx2 = d1[i1] <= TH_HIGH ? m1[i1] : -1;
            if (idx2 >= 0 && d2[idx2] <= TH_HIGH && m2[idx2] == (int32_t)i1) { matches12[i1] = idx2; nFound++; }
        }
        return nFound;
    }

    // best / second-best over all DB rows + the acceptance test of :224-226; returns the number of matches
    int MatchBruteForce(const unsigned char* q, int nq, const unsigned char* db, long long ndb, int th, std::vector<int32_t>& match)
    {
        std::vector<int32_t> idx(nq), d1(nq), d2(nq);
        match.assign(nq, -1);
        check(orb_hamming_knn2(ctx, q, nq, db, ndb, idx.data(), d1.data(), d2.data()));
        int n = 0;
        check(orb_match_ratio(ctx, idx.data(), d1.data(), d2.data(), nq, mfNNratio, th, match.data(), &n));
        return n;
    }

    static const int TH_LOW = 50;
    static const int TH_HIGH = 100;
    static const int HISTO_LENGTH = 30;

protected:
    int window(const FrameArrays& target, const orb_window_query_set& q, int accept, int thDist, bool hist, std::vector<int32_t>& match)
    {
        orb_frame_view tv = target.view();
        int n = 0;
        check(orb_search_window(ctx, &tv, &q, accept, mfNNratio, thDist, hist ? 1 : 0, match.data(), &n));
        return n;
    }
    static std::vector<float> scaleFactors(const FrameArrays& F)
    {
        std::vector<float> sf(F.mnScaleLevels > 0 ? F.mnScaleLevels : 1, 1.0f);
        for (size_t i = 1; i < sf.size(); i++) sf[i] = sf[i - 1] * F.mfScaleFactor;
        return sf;
    }
    static int clampLevel(const FrameArrays& F, int l) { return l < 0 ? 0 : (l >= F.mnScaleLevels ? F.mnScaleLevels - 1 : l); }
    void best(const FrameArrays& F, const std::vector<unsigned char>& active, const std::vector<float>& u, const std::vector<float>& v,
              const std::vector<int32_t>& predLevel, const std::vector<unsigned char>& desc, float th, std::vector<int32_t>& idx,
              std::vector<int32_t>& dist)
    {
        const size_t n = active.size();
        std::vector<float> radius(n);
        std::vector<int32_t> lmin(n);
        const std::vector<float> sf = scaleFactors(F);
        for (size_t i = 0; i < n; i++) { radius[i] = th * sf[clampLevel(F, predLevel[i])]; lmin[i] = predLevel[i] - 1; }
        orb_window_query_set q = { (int32_t)n, active.data(), desc.data(), u.data(), v.data(), nullptr, nullptr, 0,
                                   radius.data(), 0.f, lmin.data(), predLevel.data(), nullptr };
        idx.assign(n, -1); dist.assign(n, 0);
        orb_frame_view tv = F.view();
        check(orb_search_window_best(ctx, &tv, &q, idx.data(), dist.data()));
    }
    void check(int status)
    {
        if (status != ORB_OK)
            throw std::runtime_error(std::string("ORBmatcher: ") + orb_error_string(status) +
                                     (status == ORB_ERR_CUDA ? std::string(" [") + orb_last_cuda_error() + "]" : std::string()));
    }
    orb_ctx* ctx;
    float mfNNratio;
    bool mbCheckOrientation;
};

#ifndef ORB_B200_WITH_REFERENCE_TYPES
class ORBmatcher : public ORBmatcherArrays
{
public:
    ORBmatcher(orb_ctx* ctx, float nnratio = 0.6, bool checkOri = true) : ORBmatcherArrays(ctx, nnratio, checkOri) {}
    // the reference's constructor (include/ORBmatcher.h:41): runs on the process-wide default context
    explicit ORBmatcher(float nnratio = 0.6, bool checkOri = true) : ORBmatcherArrays(orb_default_context(), nnratio, checkOri) {}
};
#endif

} // namespace ORB_SLAM

#ifdef ORB_B200_WITH_REFERENCE_TYPES
#include "orb_b200_reftypes.h"
#endif

#endif
