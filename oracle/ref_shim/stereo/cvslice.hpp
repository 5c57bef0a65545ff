// TEST INFRASTRUCTURE ONLY -- the slice of cv::Mat / cv::KeyPoint that the reference's Frame::ComputeStereoMatches
// (Frame.cc:810-984) and MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:257-322) use, modelled on what OpenCV does
// for exactly those calls: row/col views share storage, convertTo(CV_32F) widens uchar, Mat - float*ones is a per-element
// float subtraction, norm(NORM_L1) of two CV_32F matrices sums |a-b| in double, clone() copies.
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <utility>
#include <vector>

#include "../../match_oracle.h"

using namespace std;

#define CV_8U 0
#define CV_32F 5

namespace cv {

struct Point2f { float x, y; };
struct KeyPoint { Point2f pt; float size, angle, response; int octave, class_id; };
enum { NORM_L1 = 2 };

class Mat {
public:
    int rows = 0, cols = 0, type_ = CV_8U;
    size_t step = 0;                       // bytes per row
    unsigned char* data = nullptr;
    std::shared_ptr<std::vector<unsigned char>> own;   // null for views onto caller memory

    Mat() {}
    Mat(int r, int c, int type) { *this = alloc(r, c, type); }          // zero-filled (OpenCV leaves it uninitialised)
    void create(int r, int c, int type) { *this = alloc(r, c, type); }
    Mat(int r, int c, int type, void* p, size_t stepBytes) : rows(r), cols(c), type_(type), step(stepBytes), data((unsigned char*)p) {}
    static Mat alloc(int r, int c, int type) {
        Mat m;
        m.rows = r; m.cols = c; m.type_ = type;
        m.step = (size_t)c * (type == CV_32F ? 4 : 1);
        m.own = std::make_shared<std::vector<unsigned char>>(m.step * (size_t)r);
        m.data = m.own->data();
        return m;
    }
    static Mat ones(int r, int c, int type) {
        Mat m = alloc(r, c, type);
        for (int y = 0; y < r; ++y)
            for (int x = 0; x < c; ++x) m.at<float>(y, x) = 1.0f;
        return m;
    }
    size_t elem() const { return type_ == CV_32F ? 4 : 1; }
    Mat rowRange(int a, int b) const {
        if (!(0 <= a && a <= b && b <= rows)) abort();      // CV_Assert in OpenCV
        Mat m = *this;
        m.data = data + (size_t)a * step;
        m.rows = b - a;
        return m;
    }
    Mat colRange(int a, int b) const {
        if (!(0 <= a && a <= b && b <= cols)) abort();
        Mat m = *this;
        m.data = data + (size_t)a * elem();
        m.cols = b - a;
        return m;
    }
    Mat row(int y) const { return rowRange(y, y + 1); }
    Mat clone() const {
        Mat out = alloc(rows, cols, type_);
        for (int y = 0; y < rows; ++y) std::memcpy(out.data + (size_t)y * out.step, data + (size_t)y * step, (size_t)cols * elem());
        return out;
    }
    template <class T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    template <class T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    template <class T> const T* ptr() const { return (const T*)data; }
    template <class T> T* ptr() { return (T*)data; }
    void convertTo(Mat& dst, int type) const {              // only CV_8U -> CV_32F is used (Frame.cc:909, :926)
        if (type_ != CV_8U || type != CV_32F) abort();
        Mat out = alloc(rows, cols, CV_32F);
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) out.at<float>(y, x) = (float)at<unsigned char>(y, x);
        dst = out;                                          // dst may alias *this: assigned last
    }
};

inline Mat operator*(float s, const Mat& m) {
    Mat out = Mat::alloc(m.rows, m.cols, CV_32F);
    for (int y = 0; y < m.rows; ++y)
        for (int x = 0; x < m.cols; ++x) out.at<float>(y, x) = s * m.at<float>(y, x);
    return out;
}
inline Mat operator-(const Mat& a, const Mat& b) {
    Mat out = Mat::alloc(a.rows, a.cols, CV_32F);
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < a.cols; ++x) out.at<float>(y, x) = a.at<float>(y, x) - b.at<float>(y, x);
    return out;
}
inline double norm(const Mat& a, const Mat& b, int) {       // NORM_L1 of CV_32F: double accumulator
    double s = 0;
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < a.cols; ++x) s += (double)std::abs(a.at<float>(y, x) - b.at<float>(y, x));
    return s;
}

}  // namespace cv


namespace ORB_SLAM2 {
class ORBmatcher {
public:
    static const int TH_LOW = 50;                           // ORBmatcher.cc:36-37 (values pinned by tests/test_matcher_ref.py)
    static const int TH_HIGH = 100;
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {   // pinned against ORBmatcher.cc:1675-1691 there
        return orbo::descriptor_distance(a.ptr<uint8_t>(), b.ptr<uint8_t>());
    }
};
}  // namespace ORB_SLAM2
