// Host-side mirror of ORB_SLAM2::ORBextractor (reference: include/ORBextractor.h:44-131) over the orbb200 C ABI.
//
// Same class name, constructor, operator() and getters as the reference, so Frame.cc:591-597
//     (*mpORBextractorLeft)(im, cv::Mat(), mvKeys, mDescriptors);
// compiles unchanged when this header replaces the reference's.  All pixel and keypoint work runs in the CUDA library;
// this file only converts containers.  Two surfaces:
//   * always: operator() on a raw 8-bit image + std::vector outputs (what the tests in this repo drive, no OpenCV needed)
//   * with -DORBB200_WITH_OPENCV (OpenCV headers present): the reference's exact cv::InputArray / cv::OutputArray
//     signature and the public mvImagePyramid member (filled on demand, see SyncPyramid()).
// Error behaviour: the reference returns void and never reports; here a failed call throws std::runtime_error with
// orb_last_error() (a CUDA failure must not look like "no keypoints").
#ifndef ORBB200_ADAPTER_ORBEXTRACTOR_H
#define ORBB200_ADAPTER_ORBEXTRACTOR_H

#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/orbb200.h"

#ifdef ORBB200_WITH_OPENCV
#include <opencv2/core/core.hpp>
#endif

namespace ORB_SLAM2 {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    // max_width/max_height/device are additions with defaults: the arena is sized on the first image otherwise
    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int max_width = 0,
                 int max_height = 0, int device = 0)
        : nfeatures(nfeatures), scaleFactor(scaleFactor), nlevels(nlevels), iniThFAST(iniThFAST), minThFAST(minThFAST) {
        check(orbx_create(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_width, max_height, 1, device, &h_));
        mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
        mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
        mnFeaturesPerLevel.resize(nlevels); umax.resize(16);
        check(orbx_get_scale_tables(h_, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                                    mvInvLevelSigma2.data(), mnFeaturesPerLevel.data(), umax.data()));
    }
    ~ORBextractor() { orbx_destroy(h_); }
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image (mask is ignored, as in the reference: ORBextractor.h:76).
    void operator()(const unsigned char* image, int width, int height, int stride, std::vector<orb_keypoint>& keypoints,
                    std::vector<unsigned char>& descriptors) {
        if (!image || width <= 0 || height <= 0) return;   // ORBextractor.cc:1048: outputs untouched
        int cap = 0;
        check(orbx_keypoint_capacity(h_, &cap));
        keypoints.resize(cap);
        descriptors.resize((size_t)cap * 32);
        int n = 0;
        int st = orbx_extract(h_, image, width, height, stride, keypoints.data(), descriptors.data(), cap, &n);
        if (st == ORB_ERR_CAPACITY) {   // the bound depends on the image size: retry once with the refreshed capacity
            check(orbx_keypoint_capacity(h_, &cap));
            keypoints.resize(cap);
            descriptors.resize((size_t)cap * 32);
            st = orbx_extract(h_, image, width, height, stride, keypoints.data(), descriptors.data(), cap, &n);
        }
        check(st);
        keypoints.resize(n);
        descriptors.resize((size_t)n * 32);
    }

#ifdef ORBB200_WITH_OPENCV
    void operator()(cv::InputArray _image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& _keypoints,
                    cv::OutputArray _descriptors) {
        if (_image.empty()) return;
        cv::Mat image = _image.getMat();
        CV_Assert(image.type() == CV_8UC1);
        static_assert(sizeof(cv::KeyPoint) == sizeof(orb_keypoint), "cv::KeyPoint layout");
        std::vector<orb_keypoint> k;
        std::vector<unsigned char> d;
        (*this)(image.data, image.cols, image.rows, (int)image.step, k, d);
        _keypoints.resize(k.size());
        if (k.empty()) { _descriptors.release(); return; }   // ORBextractor.cc:1080-1081
        std::memcpy((void*)_keypoints.data(), k.data(), k.size() * sizeof(orb_keypoint));
        _descriptors.create((int)k.size(), 32, CV_8U);
        cv::Mat out = _descriptors.getMat();
        for (size_t i = 0; i < k.size(); ++i) std::memcpy(out.ptr((int)i), &d[i * 32], 32);
        if (syncPyramid_) SyncPyramid();
    }
    // mvImagePyramid is a public member of the reference (ORBextractor.h:103) read by Frame::ComputeStereoMatches.
    // Copying 1.3 MB back per frame is wasted for monocular use, so it is opt-in.
    std::vector<cv::Mat> mvImagePyramid;
    void SetSyncPyramid(bool on) { syncPyramid_ = on; }
    void SyncPyramid() {
        mvImagePyramid.resize(nlevels);
        for (int l = 0; l < nlevels; ++l) {
            int w = 0, h = 0;
            check(orbx_get_level(h_, 0, l, nullptr, &w, &h));
            cv::Mat padded(h + 38, w + 38, CV_8UC1);
            check(orbx_get_level(h_, 0, l, padded.data, &w, &h));
            mvImagePyramid[l] = padded(cv::Rect(19, 19, w, h));   // ROI inside the framed buffer, as ORBextractor.cc:1135-1136
        }
    }
#endif

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return (float)scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // VI-ORB-SLAM additions (ORBextractor.h:51-53): milliseconds of the last call, measured with CUDA events
    double GetTimeOfComputePyramid(void) { return stage(0); }
    double GetTimeOfComputeKeyPointsOctTree(void) { return stage(1); }
    double GetTImeOfComputeDescriptor(void) { return stage(2); }

    orbx_handle handle() { return h_; }

protected:
    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<int> umax;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;

private:
    orbx_handle h_ = nullptr;
    bool syncPyramid_ = false;
    double stage(int i) {
        double ms[3] = {0, 0, 0};
        orbx_stage_times(h_, ms);
        return ms[i];
    }
    static void check(int st) {
        if (st != ORB_OK) throw std::runtime_error(std::string("orbb200: ") + orb_last_error());
    }
};

}  // namespace ORB_SLAM2

#endif
