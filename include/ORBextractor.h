/*
 * ORBextractor.h — header-only C++ shim with the reference's class name, constructor and call
 * signature (caomw/ORBSLAM_jpMiniPC include/ORBextractor.h:32-77) on top of the C ABI of
 * liborb_b200.so (orb_b200.h).  Swap it for the reference header and link -lorb_b200: Frame::Frame
 * (src/Frame.cc:60) keeps calling (*mpORBextractor)(im, cv::Mat(), mvKeys, mDescriptors).
 *
 * With OpenCV headers available (define ORB_B200_WITH_OPENCV before including) the operator takes
 * cv::InputArray / std::vector<cv::KeyPoint> / cv::OutputArray exactly like the reference; without
 * them it takes raw 8-bit buffers and orb_keypoint records (bit-compatible with cv::KeyPoint).
 * Errors: the reference asserts / throws cv::Exception; the shim throws std::runtime_error.
 */
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <stdexcept>
#include <string>
#include <vector>
#include "orb_b200.h"
#if defined(ORB_B200_WITH_REFERENCE_TYPES) && !defined(ORB_B200_WITH_OPENCV)
#define ORB_B200_WITH_OPENCV
#endif
#ifdef ORB_B200_WITH_OPENCV
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#endif

namespace ORB_SLAM
{

class ORBextractor
{
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures = 1000, float scaleFactor = 1.2f, int nlevels = 8, int scoreType = FAST_SCORE, int fastTh = 20,
                 int device = 0, int maxWidth = 1920, int maxHeight = 1200, int maxBatch = 16)
        : nfeatures(nfeatures), scaleFactor(scaleFactor), nlevels(nlevels), scoreType(scoreType), fastTh(fastTh)
    {
        ctx = orb_create(device, nfeatures, scaleFactor, nlevels, scoreType, fastTh, maxWidth, maxHeight, maxBatch);
        if (!ctx) throw std::runtime_error(std::string("ORBextractor: ") + orb_last_cuda_error());
        capacity = orb_keypoint_capacity(ctx);
    }
    ~ORBextractor() { orb_destroy(ctx); }
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image (raw-buffer form)
    void operator()(const unsigned char* image, int width, int height, int stride,
                    std::vector<orb_keypoint>& keypoints, std::vector<unsigned char>& descriptors)
    {
        if (!image || width <= 0 || height <= 0) return;            // reference: silent return on an empty image
        keypoints.resize(capacity);
        descriptors.resize((size_t)capacity * 32);
        int n = 0;
        check(orb_extract(ctx, image, width, height, stride, keypoints.data(), descriptors.data(), capacity, &n));
        keypoints.resize(n);
        descriptors.resize((size_t)n * 32);
    }

#ifdef ORB_B200_WITH_OPENCV
    // Same signature as the reference (the mask is never consumed there either, src/ORBextractor.cc:601-607)
    void operator()(cv::InputArray _image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& _keypoints, cv::OutputArray _descriptors)
    {
        if (_image.empty()) return;
        cv::Mat image = _image.getMat();
        CV_Assert(image.type() == CV_8UC1);
        static_assert(sizeof(cv::KeyPoint) == sizeof(orb_keypoint), "cv::KeyPoint layout");
        _keypoints.resize(capacity);
        cv::Mat desc(capacity, 32, CV_8U);
        int n = 0;
        check(orb_extract(ctx, image.data, image.cols, image.rows, (int)image.step,
                          reinterpret_cast<orb_keypoint*>(_keypoints.data()), desc.data, capacity, &n));
        _keypoints.resize(n);
        if (n == 0) _descriptors.release();
        else desc.rowRange(0, n).copyTo(_descriptors);
    }
#endif

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return (float)scaleFactor; }
    orb_ctx* context() { return ctx; }
    /* not in the reference: pick which reference BUILD the descriptors reproduce (orb_set_descriptor_fma, orb_b200.h) */
    void SetDescriptorFMA(bool on) { check(orb_set_descriptor_fma(ctx, on ? 1 : 0)); }

protected:
    void check(int status)
    {
        if (status != ORB_OK)
            throw std::runtime_error(std::string("ORBextractor: ") + orb_error_string(status) +
                                     (status == ORB_ERR_CUDA ? std::string(" [") + orb_last_cuda_error() + "]" : std::string()));
    }
    int nfeatures;
    double scaleFactor;
    int nlevels;
    int scoreType;
    int fastTh;
    int capacity;
    orb_ctx* ctx;
};

} // namespace ORB_SLAM

#endif
