// Stand-in for <opencv2/core.hpp> used ONLY by tests/test_signature_dropin.py to compile the cv::Mat overloads of the host
// mirrors in an image without OpenCV: the members of cv::Mat those overloads touch, nothing else.
#pragma once
#include <cstddef>
#include <cstdint>
#define CV_8UC1 0
typedef unsigned char uchar;
namespace cv {
struct MatStep {
    size_t v = 0;
    operator size_t() const { return v; }
};
class Mat {
  public:
    Mat() {}
    Mat(int r, int c, int /*type*/, void *d = nullptr, size_t s = 0) : data(static_cast<uchar *>(d)), rows(r), cols(c) { step.v = s ? s : (size_t)c; }
    uchar *data = nullptr;
    int rows = 0, cols = 0;
    MatStep step;
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
};
}  // namespace cv
