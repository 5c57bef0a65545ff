// The adapter's reference-signature operator()(cv::InputArray, cv::InputArray, std::vector<cv::KeyPoint>&, cv::OutputArray)
// compiled against the OpenCV stand-in of oracle/ref_shim/include (the one the reference's own ORBextractor.cc is built
// against for oracle/_ref/orb_ref): must give the same keypoints and descriptors as the raw-pointer overload.
// usage: adapter_opencv_sig image.raw width height   -> prints "n checksum levels" and exits 0 on agreement
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <opencv2/core/core.hpp>

#define ORBB200_WITH_OPENCV
#include "ORBextractor.h"

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const int w = atoi(argv[2]), h = atoi(argv[3]);
    cv::Mat img(h, w, CV_8UC1);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(img.data, 1, (size_t)w * h, f) != (size_t)w * h) return 3;
    fclose(f);
    ORB_SLAM2::ORBextractor ex(1000, 1.2f, 8, 20, 7);
    std::vector<cv::KeyPoint> kps;
    cv::Mat desc;
    ex(img, cv::Mat(), kps, desc);                       // Frame.cc:591-597
    std::vector<orb_keypoint> k2;
    std::vector<unsigned char> d2;
    ex(img.data, w, h, (int)img.step, k2, d2);
    if (kps.size() != k2.size() || desc.rows != (int)k2.size() || desc.cols != 32) return 4;
    unsigned long long s = 0;
    for (size_t i = 0; i < k2.size(); ++i) {
        if (kps[i].pt.x != k2[i].x || kps[i].pt.y != k2[i].y || kps[i].octave != k2[i].octave || kps[i].angle != k2[i].angle) return 5;
        for (int b = 0; b < 32; ++b) {
            if (desc.ptr((int)i)[b] != d2[i * 32 + b]) return 6;
            s = s * 131 + d2[i * 32 + b];
        }
    }
    ex.SetSyncPyramid(true);
    ex(img, cv::Mat(), kps, desc);
    if ((int)ex.mvImagePyramid.size() != 8 || ex.mvImagePyramid[0].cols != w || ex.mvImagePyramid[0].rows != h) return 7;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            if (ex.mvImagePyramid[0].ptr(y)[x] != img.ptr(y)[x]) return 8;
    printf("%zu %llu %d\n", k2.size(), s, ex.GetLevels());
    return 0;
}
