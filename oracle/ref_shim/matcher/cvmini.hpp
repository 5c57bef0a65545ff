// TEST INFRASTRUCTURE ONLY.
// A minimal stand-in for the slice of the OpenCV C++ API that /root/reference/src/ORBmatcher.cc uses, so that the
// reference's own file can be compiled IN PLACE (never copied) into oracle/_ref/libmatch_ref.so.  cv::Mat here is a
// small dense matrix of either bytes (descriptor tables: row(), ptr<>) or floats (poses and points: * + - t() dot norm).
// A * b (+ c) on 3x3 . 3x1 (+ 3x1) CV_32F operands -- every matrix product ORBmatcher.cc forms -- follows OpenCV's
// small-matrix gemm: the expression template fuses the addend, the products are summed in float in source order and the
// addend joins in double with one rounding (oracle/cvprims.cpp::gemm3_f32, pinned against cv2.gemm on 4000 poses in
// tests/test_cvprims.py::test_gemm_3x3_projection).  Other shapes accumulate in double (not used by the pinned cases).
#pragma once
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

typedef unsigned char uchar;
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

namespace cv {

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
};
typedef Point_<float> Point2f;

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};

class Mat {
public:
    int rows, cols;
    Mat() : rows(0), cols(0), type_(CV_32F), step_(0), off_(0) {}
    Mat(int r, int c, int type) : rows(r), cols(c), type_(type), step_((size_t)c * esz(type)), off_(0) {
        buf_ = std::make_shared<std::vector<uchar> >((size_t)r * step_, (uchar)0);
    }
    static Mat eye(int r, int c, int type) {
        Mat m(r, c, type);
        for (int i = 0; i < r && i < c; ++i) m.at<float>(i, i) = 1.f;
        return m;
    }
    bool empty() const { return rows == 0 || cols == 0; }
    int type() const { return type_; }
    template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(buf_->data() + off_ + (size_t)r * step_); }
    template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(buf_->data() + off_ + (size_t)r * step_); }
    uchar* ptr(int r = 0) { return ptr<uchar>(r); }
    const uchar* ptr(int r = 0) const { return ptr<uchar>(r); }
    template <class T> T& at(int r, int c) { return ptr<T>(r)[c]; }
    template <class T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
    template <class T> T& at(int i) { return rows == 1 ? ptr<T>(0)[i] : ptr<T>(i)[0]; }
    template <class T> const T& at(int i) const { return rows == 1 ? ptr<T>(0)[i] : ptr<T>(i)[0]; }
    // views share the buffer, like OpenCV's
    Mat row(int r) const { return view(r, r + 1, 0, cols); }
    Mat col(int c) const { return view(0, rows, c, c + 1); }
    Mat rowRange(int a, int b) const { return view(a, b, 0, cols); }
    Mat colRange(int a, int b) const { return view(0, rows, a, b); }
    Mat clone() const {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; ++r) std::memcpy(m.ptr<uchar>(r), ptr<uchar>(r), (size_t)cols * esz(type_));
        return m;
    }
    Mat t() const {
        Mat m(cols, rows, CV_32F);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) m.at<float>(c, r) = at<float>(r, c);
        return m;
    }
    double dot(const Mat& o) const {
        double s = 0;
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) s += (double)at<float>(r, c) * (double)o.at<float>(r, c);
        return s;
    }

private:
    static size_t esz(int type) { return type == CV_32F ? 4 : 1; }
    Mat view(int r0, int r1, int c0, int c1) const {
        Mat m;
        m.rows = r1 - r0; m.cols = c1 - c0; m.type_ = type_; m.step_ = step_; m.buf_ = buf_;
        m.off_ = off_ + (size_t)r0 * step_ + (size_t)c0 * esz(type_);
        return m;
    }
    int type_;
    size_t step_, off_;
    std::shared_ptr<std::vector<uchar> > buf_;
};

inline Mat gemm_(const Mat& a, const Mat& b, const Mat* c) {
    assert(a.cols == b.rows);
    Mat m(a.rows, b.cols, CV_32F);
    const bool small = a.rows == 3 && a.cols == 3 && b.cols == 1;
    for (int r = 0; r < a.rows; ++r)
        for (int col = 0; col < b.cols; ++col) {
            if (small) {
                const float t = (a.at<float>(r, 0) * b.at<float>(0, 0) + a.at<float>(r, 1) * b.at<float>(1, 0)) + a.at<float>(r, 2) * b.at<float>(2, 0);
                m.at<float>(r, col) = c ? (float)((double)t + (double)c->at<float>(r, 0)) : t;
            } else {
                double s = 0;
                for (int k = 0; k < a.cols; ++k) s += (double)a.at<float>(r, k) * (double)b.at<float>(k, col);
                if (c) s += (double)c->at<float>(r, col);
                m.at<float>(r, col) = (float)s;
            }
        }
    return m;
}
// cv::MatExpr for the one fusion that matters: (A * b) + c is a single gemm
struct MatMul {
    Mat a, b;
    operator Mat() const { return gemm_(a, b, nullptr); }
    template <class T> T at(int i) const { return Mat(*this).at<T>(i); }
};
inline MatMul operator*(const Mat& a, const Mat& b) { return MatMul{a, b}; }
inline MatMul operator*(const MatMul& m, const Mat& b) { return MatMul{Mat(m), b}; }
inline Mat operator+(const MatMul& m, const Mat& c) { return gemm_(m.a, m.b, &c); }
template <class F> inline Mat map2(const Mat& a, const Mat& b, F f) {
    Mat m(a.rows, a.cols, CV_32F);
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) m.at<float>(r, c) = f(a.at<float>(r, c), b.at<float>(r, c));
    return m;
}
template <class F> inline Mat map1(const Mat& a, F f) {
    Mat m(a.rows, a.cols, CV_32F);
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) m.at<float>(r, c) = f(a.at<float>(r, c));
    return m;
}
inline Mat operator+(const Mat& a, const Mat& b) { return map2(a, b, [](float x, float y) { return x + y; }); }
inline Mat operator-(const Mat& a, const Mat& b) { return map2(a, b, [](float x, float y) { return x - y; }); }
inline Mat operator-(const Mat& a) { return map1(a, [](float x) { return -x; }); }
inline Mat operator*(double s, const Mat& a) { return map1(a, [s](float x) { return (float)(s * x); }); }
inline Mat operator*(const Mat& a, double s) { return s * a; }
inline Mat operator/(const Mat& a, double s) { return map1(a, [s](float x) { return (float)(x / s); }); }
inline double norm(const Mat& a) { return std::sqrt(a.dot(a)); }

}  // namespace cv
