// TEST INFRASTRUCTURE ONLY.
// A minimal stand-in for the slice of the OpenCV C++ API that /root/reference/src/ORBextractor.cc uses, so that the
// reference's own file can be compiled IN PLACE (it is never copied into this repo) and run as oracle/_ref.
// OpenCV has no C++ SDK in this image; the five image primitives are backed by oracle/cvprims.cpp, whose models are
// pinned bit-for-bit against cv2 4.13.0 by tests/test_cvprims.py.  Everything else here is plain container glue.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "../../cvprims.h"

typedef unsigned char uchar;

#define CV_8U 0
#define CV_8UC1 0
#define CV_Assert(expr) assert(expr)
#define CV_PI 3.1415926535897932384626433832795

inline int cvRound(double v) { return orbo::cv_round_d(v); }
inline int cvRound(float v) { return orbo::cv_round(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {

enum { INTER_LINEAR = 1 };
enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
template <class T> inline Point_<T>& operator*=(Point_<T>& a, float b) {
    a.x = (T)(a.x * b);
    a.y = (T)(a.y * b);
    return a;
}

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
    int x, y, width, height;
    Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {}
};

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};
static_assert(sizeof(KeyPoint) == 28, "layout shared with orbo::KeyPoint");

struct ZerosExpr { int rows, cols, type; };

struct MatStep {
    size_t v;
    operator size_t() const { return v; }
};

class Mat {
public:
    int rows = 0, cols = 0;
    uchar* data = nullptr;
    MatStep step{0};

    Mat() {}
    Mat(Size sz, int /*type*/) { alloc(sz.height, sz.width); }
    Mat(int r, int c, int /*type*/) { alloc(r, c); }
    static ZerosExpr zeros(int r, int c, int t) { return ZerosExpr{r, c, t}; }
    Mat& operator=(const ZerosExpr& z) {  // MatExpr assignment: keeps the buffer if size and type already match
        if (!data || rows != z.rows || cols != z.cols) alloc(z.rows, z.cols);
        for (int y = 0; y < rows; ++y) std::memset(data + (size_t)y * step.v, 0, cols);
        return *this;
    }
    int type() const { return CV_8UC1; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t step1() const { return step.v; }
    Mat operator()(const Rect& r) const { return view(r.y, r.y + r.height, r.x, r.x + r.width); }
    Mat rowRange(int a, int b) const { return view(a, b, 0, cols); }
    Mat colRange(int a, int b) const { return view(0, rows, a, b); }
    Mat clone() const {
        Mat m;
        m.alloc(rows, cols);
        for (int y = 0; y < rows; ++y) std::memcpy(m.data + (size_t)y * m.step.v, data + (size_t)y * step.v, cols);
        return m;
    }
    template <class T> T& at(int r, int c) { return *(T*)(data + (size_t)r * step.v + c * sizeof(T)); }
    template <class T> const T& at(int r, int c) const { return *(const T*)(data + (size_t)r * step.v + c * sizeof(T)); }
    uchar* ptr(int r = 0) { return data + (size_t)r * step.v; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step.v; }
    void create(int r, int c, int /*type*/) {
        if (data && rows == r && cols == c) return;
        alloc(r, c);
    }
    void release() { owner.reset(); data = nullptr; rows = cols = 0; step.v = 0; }

private:
    std::shared_ptr<std::vector<uchar>> owner;
    void alloc(int r, int c) {
        owner = std::make_shared<std::vector<uchar>>((size_t)r * c);
        rows = r; cols = c; step.v = (size_t)c; data = owner->data();
    }
    Mat view(int r0, int r1, int c0, int c1) const {
        Mat m;
        m.owner = owner; m.rows = r1 - r0; m.cols = c1 - c0; m.step = step;
        m.data = data + (size_t)r0 * step.v + c0;
        return m;
    }
};

// InputArray / OutputArray: the reference only ever passes cv::Mat (Frame.cc:591-597)
class _InputArray {
public:
    _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
    bool empty() const { return m_->empty(); }
    Mat getMat() const { return *m_; }
protected:
    Mat* m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray(Mat& m) : _InputArray(m) {}
    void create(int r, int c, int t) const { m_->create(r, c, t); }
    void release() const { m_->release(); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

inline float fastAtan2(float y, float x) { return orbo::fast_atan2(y, x); }

inline void resize(InputArray _src, OutputArray _dst, Size dsize, double = 0, double = 0, int = INTER_LINEAR) {
    Mat src = _src.getMat();
    _dst.create(dsize.height, dsize.width, CV_8UC1);
    Mat dst = _dst.getMat();
    orbo::resize_linear_u8(src.data, src.cols, src.rows, (int)src.step.v, dst.data, dst.cols, dst.rows, (int)dst.step.v);
}

// src may be a view into dst (ORBextractor.cc:1143): stage through a compact copy first
inline void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right, int /*type*/) {
    Mat src = _src.getMat().clone();
    assert(top == bottom && left == right && top == left);
    _dst.create(src.rows + top + bottom, src.cols + left + right, CV_8UC1);
    Mat dst = _dst.getMat();
    orbo::copy_make_border_reflect101(src.data, src.cols, src.rows, (int)src.step.v, dst.data, (int)dst.step.v, top);
}

inline void GaussianBlur(InputArray _src, OutputArray _dst, Size k, double sx, double sy, int /*border*/) {
    assert(k.width == 7 && k.height == 7 && sx == 2 && sy == 2);
    (void)k; (void)sx; (void)sy;
    Mat src = _src.getMat().clone();
    _dst.create(src.rows, src.cols, CV_8UC1);
    Mat dst = _dst.getMat();
    orbo::gaussian_blur_7x7_s2(src.data, src.cols, src.rows, (int)src.step.v, dst.data, (int)dst.step.v);
}

inline void FAST(InputArray _img, std::vector<KeyPoint>& kps, int threshold, bool nonmax = true) {
    Mat img = _img.getMat();
    std::vector<orbo::KeyPoint> v;
    orbo::fast9_16(img.data, img.cols, img.rows, (int)img.step.v, threshold, nonmax, v);
    kps.resize(v.size());
    if (!v.empty()) std::memcpy((void*)kps.data(), v.data(), v.size() * sizeof(KeyPoint));
}

// only referenced from the dead ComputeKeyPointsOld (ORBextractor.cc:857-1034); must link, is never run
struct KeyPointsFilter {
    static void retainBest(std::vector<KeyPoint>& k, int n) {
        if (n >= 0 && (int)k.size() > n) k.resize(n);
    }
};

}  // namespace cv
