// TEST INFRASTRUCTURE ONLY.
// Stand-ins for ORB_SLAM2::MapPoint / KeyFrame / Frame with exactly the members /root/reference/src/ORBmatcher.cc reads,
// so that the reference's search loops can be compiled in place and run on flat arrays.  The reference's own headers
// need Eigen, DBoW2's vocabulary, g2o and the IMU classes; they are kept out by pre-defining their include guards.
// GetFeaturesInArea / the 64x48 grid live in Frame.cc / KeyFrame.cc, which cannot be compiled whole here.  With
// ORB_REF_GRID (libmatch_ref.so) the stand-ins carry the reference's own member names and the reference's own text of
// Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea and KeyFrame::GetFeaturesInArea is streamed into the build
// (Makefile: grid_ref.o), so the reference's searches run on the reference's grid.  Without it (libmatch_adapter.so, whose
// ORBmatcher never asks the stand-ins for the grid) the members call the oracle's restatement.
#pragma once
#define MAPPOINT_H
#define KEYFRAME_H
#define FRAME_H

#include <map>
#include <set>
#include <vector>

#include "../../match_oracle.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"   // the reference's own (a std::map), found through -I$(REF)
#include "cvmini.hpp"

namespace ORB_SLAM2 {

class KeyFrame;
class Frame;

class MapPoint {
public:
    // what the searches read
    bool bad = false;
    int nObs = 0;
    cv::Mat descriptor, worldPos, normal;
    float minDist = 0.f, maxDist = 1e30f;
    int predictedLevel = 0;
    std::map<const void*, int> indexIn;          // keyframe -> keypoint index
    // tracking fields (SearchByProjection(Frame, vector<MapPoint*>))
    bool mbTrackInView = false;
    float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0, mTrackViewCos = 1;
    int mnTrackScaleLevel = 0;
    // what Fuse does to the map, recorded
    MapPoint* replacedBy = nullptr;
    std::vector<std::pair<KeyFrame*, size_t> > added;

    bool isBad() { return bad; }
    int Observations() { return nObs; }
    cv::Mat GetDescriptor() { return descriptor; }
    cv::Mat GetWorldPos() { return worldPos; }
    cv::Mat GetNormal() { return normal; }
    float GetMinDistanceInvariance() { return minDist; }
    float GetMaxDistanceInvariance() { return maxDist; }
    int PredictScale(const float&, KeyFrame*) { return predictedLevel; }
    int PredictScale(const float&, Frame*) { return predictedLevel; }
    bool IsInKeyFrame(KeyFrame* kf) { return indexIn.count(kf) != 0; }
    int GetIndexInKeyFrame(KeyFrame* kf) { return indexIn.count(kf) ? indexIn[kf] : -1; }
    void Replace(MapPoint* p) { replacedBy = p; }
    void AddObservation(KeyFrame* kf, size_t idx) { added.push_back(std::make_pair(kf, idx)); }
};

// arrays + grid shared by the two frame stand-ins
struct FrameData {
    int N = 0;
    std::vector<cv::KeyPoint> mvKeysUn, mvKeys;
    cv::Mat mDescriptors;
    std::vector<float> mvuRight, mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    DBoW2::FeatureVector mFeatVec;
    float fx = 1, fy = 1, cx = 0, cy = 0, mbf = 0, mb = 0;
    float mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
    // Frame.h:190 / KeyFrame.h: every object gets the next id at construction, copies keep it
    static unsigned long& nNextId() { static unsigned long n = 0; return n; }
    unsigned long mnId = nNextId()++;
    cv::Mat mDistCoef = cv::Mat(4, 1, CV_32F);   // zero-initialised: no distortion
    orbo::FrameArrays fa;

    virtual ~FrameData() {}
    virtual void afterSet() {}          // ORB_REF_GRID: the reference's AssignFeaturesToGrid on the reference's members
    void set(const cv::KeyPoint* keys, const uint8_t* desc, int n, float minX, float minY, float maxX, float maxY) {
        N = n;
        mvKeysUn.assign(keys, keys + n);
        mvKeys = mvKeysUn;
        mDescriptors = cv::Mat(n > 0 ? n : 1, 32, CV_8U);
        if (n) std::memcpy(mDescriptors.ptr<uchar>(0), desc, (size_t)n * 32);
        mvuRight.assign(n, -1.f);
        mnMinX = minX; mnMinY = minY; mnMaxX = maxX; mnMaxY = maxY;
        static_assert(sizeof(cv::KeyPoint) == sizeof(orbo::KeyPoint), "keypoint layouts must agree");
        fa.n = n;
        fa.keysUn = reinterpret_cast<const orbo::KeyPoint*>(mvKeysUn.data());
        fa.desc = mDescriptors.ptr<uchar>(0);
        fa.minX = minX; fa.minY = minY; fa.maxX = maxX; fa.maxY = maxY;
        fa.invW = (float)orbo::GRID_COLS / (maxX - minX);    // Frame.cc:104-105 / KeyFrame ctor
        fa.invH = (float)orbo::GRID_ROWS / (maxY - minY);
        fa.buildGrid();
        afterSet();
    }
    std::vector<size_t> area(float x, float y, float r, int minLevel, int maxLevel) const {
        std::vector<int> v;
        fa.featuresInArea(x, y, r, minLevel, maxLevel, v);
        return std::vector<size_t>(v.begin(), v.end());
    }
    bool inImage(float x, float y) const { return x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY; }
};

#ifdef ORB_REF_GRID
#define FRAME_GRID_ROWS 48    // Frame.h:41-42
#define FRAME_GRID_COLS 64
#endif

class Frame : public FrameData {
public:
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    cv::Mat mTcw = cv::Mat::eye(4, 4, CV_32F);
#ifdef ORB_REF_GRID
    // the reference's members (Frame.h:233-235); the three functions are the reference's own text (grid_ref.o)
    static float mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    void AssignFeaturesToGrid();
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1,
                                          const int maxLevel = -1) const;
    void afterSet() override {
        mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (mnMaxX - mnMinX);     // Frame.cc:445-446
        mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (mnMaxY - mnMinY);
        for (int i = 0; i < FRAME_GRID_COLS; ++i)
            for (int j = 0; j < FRAME_GRID_ROWS; ++j) mGrid[i][j].clear();
        AssignFeaturesToGrid();
    }
#else
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1,
                                          const int maxLevel = -1) const {
        return area(x, y, r, minLevel, maxLevel);
    }
#endif
};

class KeyFrame : public FrameData {
public:
    std::vector<MapPoint*> mvpMapPoints;
    cv::Mat Rcw = cv::Mat::eye(3, 3, CV_32F), tcw = cv::Mat(3, 1, CV_32F), Ow = cv::Mat(3, 1, CV_32F);
    std::vector<std::pair<MapPoint*, size_t> > added;

    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    std::set<MapPoint*> GetMapPoints() {
        std::set<MapPoint*> s;
        for (size_t i = 0; i < mvpMapPoints.size(); ++i)
            if (mvpMapPoints[i] && !mvpMapPoints[i]->isBad()) s.insert(mvpMapPoints[i]);
        return s;
    }
    MapPoint* GetMapPoint(const size_t& i) { return mvpMapPoints[i]; }
    cv::Mat GetRotation() { return Rcw.clone(); }
    cv::Mat GetTranslation() { return tcw.clone(); }
    cv::Mat GetCameraCenter() { return Ow.clone(); }
    bool IsInImage(const float& x, const float& y) const { return inImage(x, y); }
#ifdef ORB_REF_GRID
    // the reference's members (KeyFrame.h:226-229, :306), filled as the KeyFrame constructor does from a Frame
    // (KeyFrame.cc:55-56, :84-90); GetFeaturesInArea is the reference's own text (grid_ref.o)
    int mnGridCols = FRAME_GRID_COLS, mnGridRows = FRAME_GRID_ROWS;
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    std::vector<std::vector<std::vector<size_t> > > mGrid;
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const;
    void afterSet() override {
        Frame F;
        F.set(mvKeysUn.data(), mDescriptors.ptr<uchar>(0), N, mnMinX, mnMinY, mnMaxX, mnMaxY);
        mfGridElementWidthInv = F.mfGridElementWidthInv;
        mfGridElementHeightInv = F.mfGridElementHeightInv;
        mGrid.resize(mnGridCols);
        for (int i = 0; i < mnGridCols; i++) {
            mGrid[i].resize(mnGridRows);
            for (int j = 0; j < mnGridRows; j++) mGrid[i][j] = F.mGrid[i][j];
        }
    }
#else
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const { return area(x, y, r, -1, -1); }
#endif
    void AddMapPoint(MapPoint* p, const size_t& idx) { added.push_back(std::make_pair(p, idx)); }
};

}  // namespace ORB_SLAM2
