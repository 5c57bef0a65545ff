// TEST INFRASTRUCTURE ONLY -- CPU oracle (see cvprims.h). Build: -O2 -ffp-contract=off, no -march=native.
#include "cvprims.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace orbo {

int cv_round(float v) { return (int)lrintf(v); }
int cv_round_d(double v) { return (int)lrint(v); }
int cv_floor(float v) { return (int)std::floor(v); }
int cv_ceil(float v) { return (int)std::ceil(v); }

// ---------------------------------------------------------------------------------------------
// resize, INTER_LINEAR, 8UC1. Fixed point: 11-bit coefficients, horizontal pass keeps 8+11 bits,
// vertical pass drops 4 bits of the row sums and 16 of each product, rounds once (+2 >> 2).
// ---------------------------------------------------------------------------------------------
void resize_axis_table(int ssize, int dsize, int* ofs, int16_t* coef) {
    const double scale = 1.0 / ((double)dsize / (double)ssize);
    for (int d = 0; d < dsize; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cv_floor(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
        ofs[d] = s;
        coef[2 * d + 0] = (int16_t)cv_round((1.f - f) * 2048.f);
        coef[2 * d + 1] = (int16_t)cv_round(f * 2048.f);
    }
}

void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride,
                      uint8_t* dst, int dw, int dh, int dstride) {
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<int16_t> xa(2 * dw), yb(2 * dh);
    resize_axis_table(sw, dw, xofs.data(), xa.data());
    resize_axis_table(sh, dh, yofs.data(), yb.data());
    std::vector<int> row0(dw), row1(dw);
    for (int dy = 0; dy < dh; ++dy) {
        const int sy0 = yofs[dy];
        const int sy1 = std::min(sy0 + 1, sh - 1);
        const uint8_t* s0 = src + (size_t)sy0 * sstride;
        const uint8_t* s1 = src + (size_t)sy1 * sstride;
        for (int dx = 0; dx < dw; ++dx) {
            const int x0 = xofs[dx];
            const int x1 = std::min(x0 + 1, sw - 1);
            const int a0 = xa[2 * dx], a1 = xa[2 * dx + 1];
            row0[dx] = s0[x0] * a0 + s0[x1] * a1;
            row1[dx] = s1[x0] * a0 + s1[x1] * a1;
        }
        const int b0 = yb[2 * dy], b1 = yb[2 * dy + 1];
        uint8_t* d = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; ++dx)
            d[dx] = (uint8_t)((((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2);
    }
}

// ---------------------------------------------------------------------------------------------
// copyMakeBorder, BORDER_REFLECT_101:  gfedcb|abcdefgh|gfedcba
// ---------------------------------------------------------------------------------------------
static inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * n - 2 - i;
    }
    return i;
}

void copy_make_border_reflect101(const uint8_t* src, int w, int h, int sstride,
                                 uint8_t* dst, int dstride, int border) {
    for (int y = 0; y < h + 2 * border; ++y) {
        const uint8_t* s = src + (size_t)reflect101(y - border, h) * sstride;
        uint8_t* d = dst + (size_t)y * dstride;
        for (int x = 0; x < w + 2 * border; ++x) d[x] = s[reflect101(x - border, w)];
    }
}

// ---------------------------------------------------------------------------------------------
// FAST 9/16
// ---------------------------------------------------------------------------------------------
static const int kCircle[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},  {3, 0},  {3, -1}, {2, -2}, {1, -3},
                                   {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

int fast_score(const uint8_t* p, int stride) {
    const int v = p[0];
    int d[25];
    for (int k = 0; k < 16; ++k) d[k] = (int)p[kCircle[k][1] * stride + kCircle[k][0]] - v;
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int best = -256;
    for (int k = 0; k < 16; ++k) {
        int mn = 256, mx = -256;
        for (int j = 0; j < 9; ++j) {
            mn = std::min(mn, d[k + j]);
            mx = std::max(mx, d[k + j]);
        }
        best = std::max(best, std::max(mn, -mx));  // brighter arc: min(p-v); darker arc: min(v-p) = -max(p-v)
    }
    return best - 1;
}

// Necessary condition for a 9-arc at threshold t: of every opposing pair (k, k+8) at least one pixel lies in the
// arc, so all 8 pairs must hold a brighter (or all 8 a darker) pixel.  Only an accelerator: fast_score decides.
static inline bool maybe_corner(const uint8_t* p, int stride, int t) {
    const int v = p[0];
    int flags = 3;
    for (int k = 0; k < 8 && flags; ++k) {
        const int a = (int)p[kCircle[k][1] * stride + kCircle[k][0]] - v;
        const int b = (int)p[kCircle[k + 8][1] * stride + kCircle[k + 8][0]] - v;
        flags &= ((a > t || b > t) ? 1 : 0) | ((a < -t || b < -t) ? 2 : 0);
    }
    return flags != 0;
}

void fast9_16(const uint8_t* img, int w, int h, int stride, int threshold, bool nonmax,
              std::vector<KeyPoint>& out) {
    out.clear();
    threshold = std::min(std::max(threshold, 0), 255);
    if (w < 7 || h < 7) return;
    // score map over the 3-px inset interior, 0 elsewhere and for non-corners
    std::vector<int> sc((size_t)w * h, 0);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            if (!maybe_corner(img + (size_t)y * stride + x, stride, threshold)) continue;
            const int s = fast_score(img + (size_t)y * stride + x, stride);
            if (s >= threshold) sc[(size_t)y * w + x] = nonmax ? s : 1;
        }
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const int s = sc[(size_t)y * w + x];
            if (s == 0) continue;
            bool keep = true;
            if (nonmax) {
                for (int dy = -1; dy <= 1 && keep; ++dy)
                    for (int dx = -1; dx <= 1; ++dx) {
                        if (dx == 0 && dy == 0) continue;
                        if (s <= sc[(size_t)(y + dy) * w + (x + dx)]) { keep = false; break; }
                    }
            }
            if (keep) {
                KeyPoint kp;
                kp.x = (float)x; kp.y = (float)y; kp.size = 7.f; kp.angle = -1.f;
                kp.response = nonmax ? (float)s : 0.f;
                kp.octave = 0; kp.class_id = -1;
                out.push_back(kp);
            }
        }
}

// ---------------------------------------------------------------------------------------------
// GaussianBlur 7x7 sigma 2, 8U bit-exact fixed point: kernel [18 34 48 56 48 34 18]/256 per axis,
// exact integer accumulation, one rounding (v + 2^15) >> 16, reflect-101 borders.
// ---------------------------------------------------------------------------------------------
void gaussian_blur_7x7_s2(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
    // two separable passes over contiguous rows (plain loops the compiler can vectorise); same arithmetic as above
    const int pw = w + 6;
    std::vector<uint8_t> padded(pw);
    std::vector<uint16_t> tmp((size_t)w * (h + 6));
    auto hrow = [&](int sy, uint16_t* out) {
        const uint8_t* s = src + (size_t)sy * sstride;
        for (int x = 0; x < 3; ++x) padded[x] = s[reflect101(x - 3, w)];
        std::memcpy(&padded[3], s, w);
        for (int x = 0; x < 3; ++x) padded[w + 3 + x] = s[reflect101(w + x, w)];
        const uint8_t* p = padded.data();
        for (int x = 0; x < w; ++x)
            out[x] = (uint16_t)(18 * (p[x] + p[x + 6]) + 34 * (p[x + 1] + p[x + 5]) + 48 * (p[x + 2] + p[x + 4]) + 56 * p[x + 3]);
    };
    for (int y = -3; y < h + 3; ++y) hrow(reflect101(y, h), &tmp[(size_t)(y + 3) * w]);
    for (int y = 0; y < h; ++y) {
        const uint16_t *r0 = &tmp[(size_t)y * w], *r1 = r0 + w, *r2 = r1 + w, *r3 = r2 + w, *r4 = r3 + w, *r5 = r4 + w, *r6 = r5 + w;
        uint8_t* d = dst + (size_t)y * dstride;
        for (int x = 0; x < w; ++x) {
            const uint32_t acc = 18u * ((uint32_t)r0[x] + r6[x]) + 34u * ((uint32_t)r1[x] + r5[x]) + 48u * ((uint32_t)r2[x] + r4[x]) + 56u * r3[x];
            d[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fastAtan2: float32 odd polynomial on the octant, no FMA contraction (build with -ffp-contract=off)
// ---------------------------------------------------------------------------------------------
float fast_atan2(float y, float x) {
    const float P1 = 57.283626556396484f, P3 = -18.66744613647461f, P5 = 8.914000511169434f,
                P7 = -2.539724588394165f;
    const float eps = (float)DBL_EPSILON;
    const float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps);
        c2 = c * c;
        a = (((P7 * c2 + P5) * c2 + P3) * c2 + P1) * c;
    } else {
        c = ax / (ay + eps);
        c2 = c * c;
        a = 90.f - (((P7 * c2 + P5) * c2 + P3) * c2 + P1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---------------------------------------------------------------------------------------------
// cv::undistortPoints(src, dst, K, distCoeffs, noArray(), K) for CV_32FC2 points, as Frame::UndistortKeyPoints and
// Frame::ComputeImageBounds call it (Frame.cc:748-808): K and the coefficients are CV_32F and are widened to double,
// five fixed-point iterations (TermCriteria(MAX_ITER, 5, 0.01)), rational model with k4..k6 = 0, tangential p1/p2,
// no thin-prism / tilt terms, then P = K applied in double and one rounding to float.  The zero-valued terms that
// OpenCV's general formula still evaluates are kept where they could change a rounding (they cannot: x + 0 is exact).
// ---------------------------------------------------------------------------------------------
void undistort_points(const float* xy, int n, float fxF, float fyF, float cxF, float cyF, const float* dist, int nDist,
                      float* out) {
    double k[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // k1 k2 p1 p2 k3 k4 k5 k6
    for (int i = 0; i < nDist && i < 8; ++i) k[i] = (double)dist[i];
    const double fx = fxF, fy = fyF, cx = cxF, cy = cyF;
    const double ifx = 1. / fx, ify = 1. / fy;
    for (int i = 0; i < n; ++i) {
        const double u = xy[2 * i], v = xy[2 * i + 1];
        double x = (u - cx) * ifx, y = (v - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; ++j) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            if (icdist < 0) {   // OpenCV's regression_14583 guard
                x = (u - cx) * ifx;
                y = (v - cy) * ify;
                break;
            }
            const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x);
            const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        out[2 * i] = (float)(fx * x + cx);
        out[2 * i + 1] = (float)(fy * y + cy);
    }
}

// ---------------------------------------------------------------------------------------------
// gemm for the 3x3 * 3x1 (+ 3x1) CV_32F case: OpenCV's small-matrix path, float arithmetic in source order
// ---------------------------------------------------------------------------------------------
void gemm3_f32(const float* A, const float* b, const float* c, float* d) {
    for (int r = 0; r < 3; ++r) {
        const float t = (A[3 * r] * b[0] + A[3 * r + 1] * b[1]) + A[3 * r + 2] * b[2];
        d[r] = c ? (float)((double)t + (double)c[r]) : t;
    }
}

}  // namespace orbo
