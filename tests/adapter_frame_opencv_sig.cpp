// adapter/Frame.h with -DORBB200_WITH_OPENCV against the OpenCV stand-in of oracle/ref_shim/include: the overloads on the
// reference's member types (cv::Mat mK / mDistCoef, std::vector<cv::KeyPoint>, cv::Mat descriptors) must compile and
// link, and the parts that need no device must behave: MakeCamera reads the reference's matrix layout, a null matcher
// is reported as an exception with the library's message.
#include <cstdio>
#include <cstring>
#include <vector>

#include <opencv2/core/core.hpp>

#define ORBB200_WITH_OPENCV
#include "Frame.h"

// the stand-in cv::Mat is byte-typed (all the extractor needs): a CV_32F r x c matrix is an r x 4c byte matrix, which gives
// at<float>(r, c) the row stride of the real thing
static cv::Mat f32(int r, int c) { return cv::Mat(r, c * 4, CV_8UC1); }

int main() {
    namespace fo = ORB_SLAM2::frame_ops;
    cv::Mat K = f32(3, 3), D = f32(5, 1), D4 = f32(4, 1);
    std::memset(K.data, 0, 9 * sizeof(float));
    K.at<float>(0, 0) = 458.654f; K.at<float>(1, 1) = 457.296f; K.at<float>(0, 2) = 367.215f; K.at<float>(1, 2) = 248.375f;
    K.at<float>(2, 2) = 1.f;
    const float d[5] = {-0.28340811f, 0.07395907f, 0.00019359f, 1.76187114e-05f, 0.25f};
    for (int i = 0; i < 5; ++i) D.at<float>(i, 0) = d[i];
    for (int i = 0; i < 4; ++i) D4.at<float>(i, 0) = d[i];
    const orb_camera c5 = fo::MakeCamera(K, D), c4 = fo::MakeCamera(K, D4);
    if (c5.fx != 458.654f || c5.fy != 457.296f || c5.cx != 367.215f || c5.cy != 248.375f) return 3;
    if (c5.k1 != d[0] || c5.k2 != d[1] || c5.p1 != d[2] || c5.p2 != d[3] || c5.k3 != d[4] || c4.k3 != 0.f) return 4;
    std::vector<cv::KeyPoint> keys(3), un;
    bool threw = false;
    try {
        fo::UndistortKeyPoints(nullptr, K, D, keys, un);
    } catch (const std::exception& e) {
        threw = std::strstr(e.what(), "null") != nullptr;
    }
    if (!threw) return 5;
    float minX = -1, maxX = -1, minY = -1, maxY = -1;
    cv::Mat im(480, 752, CV_8UC1), D0 = f32(4, 1);
    std::memset(D0.data, 0, 4 * sizeof(float));
    fo::ComputeImageBounds(nullptr, K, D0, im, minX, maxX, minY, maxY);      // k1 == 0: plain arithmetic, Frame.cc:803-808
    if (minX != 0.f || minY != 0.f || maxX != 752.f || maxY != 480.f) return 6;
    // ComputeStereoMatches is only instantiated here (it needs a device to run)
    if (keys.empty()) {
        ORB_SLAM2::ORBextractor l(1000, 1.2f, 8, 20, 7), r(1000, 1.2f, 8, 20, 7);
        std::vector<float> u, z;
        cv::Mat dl, dr;
        fo::ComputeStereoMatches(l, r, keys, dl, keys, dr, 0.11f, 47.9f, u, z);
    }
    printf("ok\n");
    return 0;
}
