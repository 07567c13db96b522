/*
 * ORBmatcher.h — header-only C++ shim with the reference's class name (caomw/ORBSLAM_jpMiniPC
 * include/ORBmatcher.h:37-107) for the matcher methods on the accelerated path:
 *   DescriptorDistance (src/ORBmatcher.cc:1794-1810)
 *   SearchByProjection(Frame&, const Frame&, th) (:1507-1620)
 *   SearchByBoW(KeyFrame*, Frame&, ...) candidate scoring (:155-284)
 * plus the brute-force best/second-best + ratio test used for relocalisation-sized searches.
 * Frame / KeyFrame / MapPoint are the reference's own graph classes and stay on the host: the shim
 * takes the plain arrays those methods read (see INTEGRATION.md for the adapter code).
 */
#ifndef ORBMATCHER_H
#define ORBMATCHER_H

#include <stdexcept>
#include <string>
#include <vector>
#include "orb_b200.h"

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

class ORBmatcher
{
public:
    ORBmatcher(orb_ctx* ctx, float nnratio = 0.6, bool checkOri = true) : ctx(ctx), mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

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

} // namespace ORB_SLAM

#endif
