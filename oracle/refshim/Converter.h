/* reference-build shim: shadows include/Converter.h (which needs Eigen and g2o) for the one helper the front end calls,
 * Converter::toDescriptorVector (src/Frame.cc:284, src/KeyFrame.cc:60); defined in oracle/ref_glue.cpp */
#ifndef CONVERTER_H
#define CONVERTER_H
#include <vector>
#include <opencv2/core/core.hpp>
namespace ORB_SLAM {
class Converter {
public:
    static std::vector<cv::Mat> toDescriptorVector(const cv::Mat& Descriptors);
};
}
#endif
