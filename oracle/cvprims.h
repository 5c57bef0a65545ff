// TEST INFRASTRUCTURE ONLY -- CPU oracle, never linked into or called from the product path.
//
// Bit-exact scalar models of the five OpenCV primitives that the reference's ORB front-end
// calls (OpenCV itself is not vendored in /root/reference and has no C++ SDK in this image).
// Each model is pinned against cv2 4.13.0 (opencv-python-headless) by tests/test_cvprims.py.
//
// Reference call sites (all in /root/reference/src/ORBextractor.cc):
//   cv::resize(INTER_LINEAR)            :1141
//   cv::copyMakeBorder(REFLECT_101)     :1143-1149
//   cv::FAST(roi, th, nonmax=true)      :811-817
//   cv::GaussianBlur(7x7, sigma 2)      :1104
//   cv::fastAtan2                       :105
//   cvRound                             :83, 117, 121-122, 444, 462, 1133
// and, for the Frame post-processing either side of the path (/root/reference/src/Frame.cc):
//   cv::undistortPoints                 :767, :793
// and for the projections the matcher's callers do (/root/reference/src/ORBmatcher.cc):
//   cv::Mat operator* / operator+ (gemm) :1377 (3x3 * 3x1 + 3x1, CV_32F)
#pragma once
#include <cstdint>
#include <vector>

namespace orbo {

// cv::KeyPoint memory layout (28 bytes): pt.x, pt.y, size, angle, response, octave, class_id
struct KeyPoint {
    float x, y, size, angle, response;
    int32_t octave, class_id;
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

int cv_round(float v);    // round-half-to-even (SSE cvtss2si semantics)
int cv_round_d(double v);
int cv_floor(float v);
int cv_ceil(float v);

// cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR) for CV_8UC1
void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride,
                      uint8_t* dst, int dw, int dh, int dstride);

// The per-axis tables of the resize (source index, 11-bit coefficient pair); exported so the
// tests can compare the device-side table builder with the oracle.
void resize_axis_table(int ssize, int dsize, int* ofs, int16_t* coef /* [dsize][2] */);

// cv::copyMakeBorder(src, dst, b, b, b, b, BORDER_REFLECT_101): dst is (w+2b)x(h+2b)
void copy_make_border_reflect101(const uint8_t* src, int w, int h, int sstride,
                                 uint8_t* dst, int dstride, int border);

// cv::FAST(img, kps, threshold, nonmaxSuppression) TYPE_9_16, row-major emission
void fast9_16(const uint8_t* img, int w, int h, int stride, int threshold, bool nonmax,
              std::vector<KeyPoint>& out);

// Threshold-free FAST score (cornerScore<16>): S >= t  <=>  corner at threshold t
int fast_score(const uint8_t* p, int stride);

// cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) for CV_8UC1
void gaussian_blur_7x7_s2(const uint8_t* src, int w, int h, int sstride,
                          uint8_t* dst, int dstride);

// cv::fastAtan2(y, x): degrees in [0, 360)
float fast_atan2(float y, float x);

// cv::undistortPoints(pts, pts, K, D, noArray(), K) for CV_32FC2 points and CV_32F K / D (Frame.cc:748-808);
// dist = k1 k2 p1 p2 [k3], nDist = 4 or 5
void undistort_points(const float* xy, int n, float fx, float fy, float cx, float cy, const float* dist, int nDist,
                      float* out);

// cv::Mat expression Rcw * x3Dw + tcw for a 3x3 by 3x1 CV_32F product (ORBmatcher.cc:1377 and every other projection of
// the matcher): OpenCV evaluates it as ONE gemm(A, b, 1, c, 1), and its small-matrix path (inner dimension 2..4) works in
// float, not in double: t = (a0*b0 + a1*b1) + a2*b2 with every operation rounded to float, d = (float)((double)t + c).
void gemm3_f32(const float* A /* 3x3 row-major */, const float* b, const float* c /* may be null: plain product */,
               float* d);

}  // namespace orbo
