// Minimal stand-in for <opencv2/opencv.hpp>, used ONLY to compile the reference's
// CHaarFeature.cpp / CIntImage_to_Featurevec.cpp in place for the parity oracle
// (oracle/_ref).  The reference only touches cv::Mat in the dead function
// CHaarFeature::calcFval (reference src/CHaarFeature.cpp:82-102), so a layout-only
// declaration is enough.  TEST INFRASTRUCTURE - never linked into the product.
#ifndef HAF_ORACLE_STUB_OPENCV_HPP
#define HAF_ORACLE_STUB_OPENCV_HPP
#include <cstddef>
namespace cv {
struct Mat {
    unsigned char* data;
    size_t step;
    unsigned char* ptr() const { return data; }
};
}  // namespace cv
#endif
