// Host-side mirror of the data-parallel member functions of ORB_SLAM2::Frame / MapPoint that sit either side of the
// extractor and the matcher (reference: src/Frame.cc, src/MapPoint.cc), over the orbb200 C ABI:
//
//   Frame::UndistortKeyPoints()            Frame.cc:748-778   mvKeys -> mvKeysUn  (cv::undistortPoints, bit for bit)
//   Frame::ComputeImageBounds(imLeft)      Frame.cc:780-808   mnMinX, mnMaxX, mnMinY, mnMaxY
//   Frame::ComputeStereoMatches()          Frame.cc:810-984   mvuRight, mvDepth
//   MapPoint::ComputeDistinctiveDescriptors()  MapPoint.cc:257-322  (batched over map points)
//   Frame::ComputeBoW()                    Frame.cc:736-745   mBowVec, mFeatVec (DBoW2 vocabulary-tree descent)
//
// The reference keeps these as member functions working on the object's own fields; here they are free functions named
// alike that take exactly the fields the reference reads and write the fields it writes, so the bodies in Frame.cc /
// MapPoint.cc become one call each.  Errors throw std::runtime_error with orb_last_error().
#ifndef ORBB200_ADAPTER_FRAME_H
#define ORBB200_ADAPTER_FRAME_H

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/orbb200.h"
#include "ORBextractor.h"

namespace ORB_SLAM2 {
namespace frame_ops {

inline void check(int st) {
    if (st != ORB_OK) throw std::runtime_error(std::string("orbb200: ") + orb_last_error());
}

// mK = [fx 0 cx; 0 fy cy; 0 0 1] and mDistCoef = (k1 k2 p1 p2 [k3]) as Tracking.cc reads them from the settings file
inline orb_camera MakeCamera(float fx, float fy, float cx, float cy, float k1, float k2, float p1, float p2, float k3 = 0.f) {
    orb_camera c = {fx, fy, cx, cy, k1, k2, p1, p2, k3};
    return c;
}

// Frame.cc:748-778.  mDistCoef.at<float>(0) == 0 copies mvKeys (Frame.cc:750-754).
inline void UndistortKeyPoints(orbm_handle matcher, const orb_camera& cam, const std::vector<orb_keypoint>& mvKeys,
                               std::vector<orb_keypoint>& mvKeysUn) {
    const int N = (int)mvKeys.size();
    std::vector<float> xy(2 * (size_t)N), un(2 * (size_t)N);
    for (int i = 0; i < N; ++i) { xy[2 * i] = mvKeys[i].x; xy[2 * i + 1] = mvKeys[i].y; }
    check(orbm_undistort_points(matcher, &cam, xy.data(), N, un.data()));
    mvKeysUn = mvKeys;
    for (int i = 0; i < N; ++i) { mvKeysUn[i].x = un[2 * i]; mvKeysUn[i].y = un[2 * i + 1]; }
}

// Frame.cc:780-808
inline void ComputeImageBounds(orbm_handle matcher, const orb_camera& cam, int cols, int rows, float& mnMinX, float& mnMaxX,
                               float& mnMinY, float& mnMaxY) {
    float b[4];
    check(orbm_image_bounds(matcher, &cam, cols, rows, b));
    mnMinX = b[0]; mnMinY = b[1]; mnMaxX = b[2]; mnMaxY = b[3];
}

// Frame.cc:810-984: both extractors hold the pyramids of the stereo pair they just processed (Frame.cc:422-425)
inline int ComputeStereoMatches(ORBextractor& left, ORBextractor& right, const std::vector<orb_keypoint>& mvKeys,
                                const std::vector<unsigned char>& mDescriptors, const std::vector<orb_keypoint>& mvKeysRight,
                                const std::vector<unsigned char>& mDescriptorsRight, float mb, float mbf,
                                std::vector<float>& mvuRight, std::vector<float>& mvDepth) {
    const int N = (int)mvKeys.size();
    mvuRight.assign(N, -1.0f);
    mvDepth.assign(N, -1.0f);
    int kept = 0;
    check(orbx_compute_stereo_matches(left.handle(), 0, right.handle(), 0, mvKeys.data(), mDescriptors.data(), N,
                                      mvKeysRight.data(), mDescriptorsRight.data(), (int)mvKeysRight.size(), mb, mbf,
                                      mvuRight.data(), mvDepth.data(), &kept));
    return kept;
}

// MapPoint.cc:257-322 for many map points at once: descriptors of point p = rows start[p] .. start[p+1] of `descriptors`
// (mObservations order, bad keyframes dropped).  best[p] = row inside the run to clone into mDescriptor, -1 = keep.
inline void ComputeDistinctiveDescriptors(orbm_handle matcher, const std::vector<unsigned char>& descriptors,
                                          const std::vector<int>& start, std::vector<int>& best) {
    const int nPoints = start.empty() ? 0 : (int)start.size() - 1;
    best.assign(nPoints, -1);
    if (nPoints == 0) return;
    check(orbm_distinctive_descriptors(matcher, descriptors.data(), start.data(), nPoints, best.data(), nullptr));
}

// Frame::ComputeBoW / KeyFrame::ComputeBoW (Frame.cc:736-745): mBowVec as (ascending word id, L1-normalised value) pairs
// and mFeatVec as node-sorted CSR (node id, run start, feature indices) -- the form orbm_search_by_bow and
// orbm_search_for_triangulation take.  `vocabulary` comes from orbm_vocabulary_create (the host project's loader fills it
// from ORBvoc); levelsup = 4 in the reference.
struct BowResult {
    std::vector<int> bowWord;
    std::vector<double> bowValue;
    std::vector<int> fvNode, fvStart, fvIdx;
};
inline void ComputeBoW(orbm_handle matcher, orbm_vocabulary vocabulary, const std::vector<unsigned char>& mDescriptors,
                       BowResult& out, int levelsup = 4) {
    const int n = (int)(mDescriptors.size() / 32);
    out.bowWord.assign(n + 1, 0); out.bowValue.assign(n + 1, 0.0);
    out.fvNode.assign(n + 1, 0); out.fvStart.assign(n + 2, 0); out.fvIdx.assign(n + 1, 0);
    int nWords = 0, nNodes = 0;
    check(orbm_bow_transform(matcher, vocabulary, mDescriptors.data(), n, levelsup, nullptr, nullptr, nullptr,
                             out.bowWord.data(), out.bowValue.data(), &nWords, out.fvNode.data(), out.fvStart.data(),
                             out.fvIdx.data(), &nNodes));
    out.bowWord.resize(nWords); out.bowValue.resize(nWords);
    out.fvNode.resize(nNodes); out.fvStart.resize(nNodes + 1);
    out.fvIdx.resize(out.fvStart[nNodes]);
}

#ifdef ORBB200_WITH_OPENCV
// The same with the reference's own member types (cv::Mat mK 3x3 CV_32F, cv::Mat mDistCoef 4x1 or 5x1 CV_32F,
// std::vector<cv::KeyPoint>, cv::Mat descriptors N x 32 CV_8U), so that the bodies in Frame.cc shrink to one line each.
inline orb_camera MakeCamera(const cv::Mat& mK, const cv::Mat& mDistCoef) {
    return MakeCamera(mK.at<float>(0, 0), mK.at<float>(1, 1), mK.at<float>(0, 2), mK.at<float>(1, 2),
                      mDistCoef.at<float>(0, 0), mDistCoef.at<float>(1, 0), mDistCoef.at<float>(2, 0), mDistCoef.at<float>(3, 0),
                      mDistCoef.rows > 4 ? mDistCoef.at<float>(4, 0) : 0.f);   // k3 is optional (Tracking.cc:776-781)
}

inline void UndistortKeyPoints(orbm_handle matcher, const cv::Mat& mK, const cv::Mat& mDistCoef,
                               const std::vector<cv::KeyPoint>& mvKeys, std::vector<cv::KeyPoint>& mvKeysUn) {
    static_assert(sizeof(cv::KeyPoint) == sizeof(orb_keypoint), "cv::KeyPoint layout");
    const int N = (int)mvKeys.size();
    std::vector<float> xy(2 * (size_t)N), un(2 * (size_t)N);
    for (int i = 0; i < N; ++i) { xy[2 * i] = mvKeys[i].pt.x; xy[2 * i + 1] = mvKeys[i].pt.y; }
    const orb_camera cam = MakeCamera(mK, mDistCoef);
    check(orbm_undistort_points(matcher, &cam, xy.data(), N, un.data()));
    mvKeysUn = mvKeys;
    for (int i = 0; i < N; ++i) { mvKeysUn[i].pt.x = un[2 * i]; mvKeysUn[i].pt.y = un[2 * i + 1]; }
}

inline void ComputeImageBounds(orbm_handle matcher, const cv::Mat& mK, const cv::Mat& mDistCoef, const cv::Mat& imLeft,
                               float& mnMinX, float& mnMaxX, float& mnMinY, float& mnMaxY) {
    ComputeImageBounds(matcher, MakeCamera(mK, mDistCoef), imLeft.cols, imLeft.rows, mnMinX, mnMaxX, mnMinY, mnMaxY);
}

inline int ComputeStereoMatches(ORBextractor& left, ORBextractor& right, const std::vector<cv::KeyPoint>& mvKeys,
                                const cv::Mat& mDescriptors, const std::vector<cv::KeyPoint>& mvKeysRight,
                                const cv::Mat& mDescriptorsRight, float mb, float mbf, std::vector<float>& mvuRight,
                                std::vector<float>& mvDepth) {
    const int N = (int)mvKeys.size(), Nr = (int)mvKeysRight.size();
    mvuRight.assign(N, -1.0f);
    mvDepth.assign(N, -1.0f);
    // rows of a cv::Mat may be padded: hand the C ABI a dense N x 32 copy
    std::vector<unsigned char> dl((size_t)N * 32), dr((size_t)Nr * 32);
    for (int i = 0; i < N; ++i) std::memcpy(&dl[(size_t)i * 32], mDescriptors.ptr(i), 32);
    for (int i = 0; i < Nr; ++i) std::memcpy(&dr[(size_t)i * 32], mDescriptorsRight.ptr(i), 32);
    int kept = 0;
    check(orbx_compute_stereo_matches(left.handle(), 0, right.handle(), 0, reinterpret_cast<const orb_keypoint*>(mvKeys.data()),
                                      dl.data(), N, reinterpret_cast<const orb_keypoint*>(mvKeysRight.data()), dr.data(), Nr,
                                      mb, mbf, mvuRight.data(), mvDepth.data(), &kept));
    return kept;
}
#endif  // ORBB200_WITH_OPENCV

}  // namespace frame_ops
}  // namespace ORB_SLAM2

#endif
