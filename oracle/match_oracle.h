// TEST INFRASTRUCTURE ONLY -- CPU oracle of the reference's ORBmatcher search loops on plain arrays.
//
// Restates (all in /root/reference/src):
//   ORBmatcher::DescriptorDistance                       ORBmatcher.cc:1675-1691
//   ORBmatcher::ComputeThreeMaxima                       ORBmatcher.cc:1629-1670
//   Frame::AssignFeaturesToGrid / PosInGrid              Frame.cc:574-589, 726-736
//   Frame::GetFeaturesInArea                             Frame.cc:671-724  (KeyFrame.cc:1138-1177)
//   ORBmatcher::SearchForInitialization                  ORBmatcher.cc:405-520
//   ORBmatcher::SearchByProjection(Frame&,Frame&,th,mono) ORBmatcher.cc:1341-1498
//   ORBmatcher::SearchByProjection(Frame&,vector<MP*>,th) ORBmatcher.cc:45-129
//   ORBmatcher::SearchForTriangulation                   ORBmatcher.cc:657-823 (+140-157)
//   ORBmatcher::SearchByBoW(KF,KF) inner loop            ORBmatcher.cc:566-618 (brute-force / all-pairs)
// The object graphs (Frame, KeyFrame, MapPoint) are flattened to the arrays the loops actually read.
// Pinned: tests/test_matcher_ref.py runs every search here against the reference's own ORBmatcher.cc compiled in place
// (oracle/_ref/libmatch_ref.so) on identical inputs; the grid lookup is pinned by known-answer tests only.
#pragma once
#include <cstdint>
#include <vector>

#include "cvprims.h"

namespace orbo {

constexpr int TH_HIGH = 100;
constexpr int TH_LOW = 50;
constexpr int HISTO_LENGTH = 30;
constexpr int GRID_COLS = 64;   // Frame.h:42
constexpr int GRID_ROWS = 48;   // Frame.h:41

int descriptor_distance(const uint8_t* a, const uint8_t* b);
void three_maxima(const int* histoSizes, int L, int& i1, int& i2, int& i3);
int rotation_bin(float a1, float a2);  // bin of (a1 - a2) as every search computes it

// What the search loops read of a Frame / KeyFrame.
struct FrameArrays {
    int n = 0;
    const KeyPoint* keysUn = nullptr;   // pt, angle, octave
    const uint8_t* desc = nullptr;      // n x 32
    float minX = 0, minY = 0, maxX = 0, maxY = 0;   // mnMinX ... (static Frame members)
    float invW = 0, invH = 0;           // mfGridElementWidthInv / HeightInv
    // CSR form of mGrid[64][48]: cell id = ix*48+iy, members in push_back order
    std::vector<int> cellStart, cellIdx;
    void buildGrid();                   // AssignFeaturesToGrid
    // GetFeaturesInArea; minLevel=-1,maxLevel=-1 reproduces the defaults (and the KeyFrame overload)
    void featuresInArea(float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& out) const;
};

int search_for_initialization(const FrameArrays& F1, const FrameArrays& F2, float* prevMatchedXY /* n1 x 2, in/out */,
                              int* matches12 /* n1 */, int windowSize, float nnratio, bool checkOri);

struct ProjQuery {          // one LastFrame keypoint with a map point (ORBmatcher.cc:1365-1410)
    float u, v;             // projection into the current frame (host computes it)
    float invz;             // 1/z in the current camera, for the stereo check
    int32_t octave;         // LastFrame.mvKeys[i].octave
    int32_t valid;          // pMP != NULL && !outlier && invz >= 0
    int32_t obsPositive;    // pMP->Observations() > 0 (what a later query's skip test reads)
    float angle;            // LastFrame.mvKeysUn[i].angle
};
// mode: 0 = [oct-1, oct+1] (mono / no motion), 1 = forward (>= oct), 2 = backward (0..oct)
// curOccupied[i2] != 0 : CurrentFrame.mvpMapPoints[i2] already holds a point with observations
// curMatch[i2] (out)   : index of the query assigned to keypoint i2, or -1
int search_by_projection_frame(const FrameArrays& cur, const float* scaleFactors, const float* uRight /* may be null */,
                               float mbf, const ProjQuery* q, const uint8_t* qdesc, int nq, float th, int mode,
                               const uint8_t* curOccupied, int* curMatch, bool checkOri, int maxDist = TH_HIGH);

struct MapPointQuery {      // ORBmatcher.cc:45-129
    float projX, projY, projXR;
    float viewCos;
    int32_t level;          // mnTrackScaleLevel
    int32_t inView;         // mbTrackInView && !isBad()
    int32_t obsPositive;
};
int search_by_projection_points(const FrameArrays& F, const float* scaleFactors, const float* uRight,
                                const MapPointQuery* q, const uint8_t* qdesc, int nq, float th, float nnratio,
                                const uint8_t* occupied, int* match /* n, out */);

// SearchForTriangulation. Feature vectors are given as node-sorted CSR: nodeId[k], start[k]..start[k+1] into idx[].
struct FeatVec { int nNodes; const int* nodeId; const int* start; const int* idx; };
struct EpiParams { float F12[9]; float ex, ey; const float* scaleFactors2; const float* levelSigma2_2; };
int search_for_triangulation(const FrameArrays& K1, const FrameArrays& K2, const FeatVec& fv1, const FeatVec& fv2,
                             const uint8_t* hasMapPoint1, const uint8_t* hasMapPoint2, const float* uRight1,
                             const float* uRight2, const EpiParams& ep, bool onlyStereo, bool checkOri,
                             int* matches12 /* n1, out */);

// SearchByBoW, both overloads (ORBmatcher.cc:159-288 KeyFrame->Frame, 522-655 KeyFrame->KeyFrame): brute force inside
// shared vocabulary nodes, one-to-one through the "already matched" flags of the second frame.
//   strictLow = false : accept best <= TH_LOW (KF->Frame, :227);  true : best < TH_LOW (KF->KF, :598)
//   valid1 / valid2   : keypoint has a usable map point (KF->Frame has no test on the frame side: pass null)
// matches12[i1] = i2 or -1, matches21[i2] = i1 or -1 (the Frame overload reports per frame keypoint, the KF one per idx1).
int search_by_bow(const FrameArrays& K1, const FrameArrays& K2, const FeatVec& fv1, const FeatVec& fv2,
                  const uint8_t* valid1, const uint8_t* valid2, float nnratio, bool checkOri, bool strictLow,
                  int* matches12, int* matches21);

// The independent projected search shared by Fuse (ORBmatcher.cc:892-944), Fuse with Sim3 (:1051-1075) and both directions
// of SearchBySim3 (:1191-1215, :1271-1295): KeyFrame::GetFeaturesInArea(u, v, radius), keep candidates whose octave is in
// [level-1, level], optionally Fuse's reprojection chi-square test, strict '<' on the Hamming distance (first wins).
struct BestQuery { float u, v, radius, ur; int32_t level, valid; };
void search_projected_best(const FrameArrays& K, const BestQuery* q, const uint8_t* qdesc, int nq, bool chi2Filter,
                           const float* uRight, const float* invLevelSigma2, int* bestIdx, int* bestDist);

// Brute force, every query against every train descriptor (no window, no one-to-one constraint):
// best / second / index with strict '<' (first wins), accept best<=TH_LOW && best < (float)second*ratio,
// then the rotation-histogram pruning.  (SearchForInitialization's inner loop + accept rule, ORBmatcher.cc:432-461)
int bruteforce_match(const uint8_t* q, const float* qAngle, int nq, const uint8_t* t, const float* tAngle, int nt,
                     float nnratio, bool checkOri, int* bestDist, int* secondDist, int* bestIdx, int* matches12);

// All-pairs keyframe matching count with SearchByBoW(KF,KF) semantics minus the BoW gating
// (ORBmatcher.cc:566-618 + 634-652): one-to-one through vbMatched2, best<TH_LOW, ratio on floats.
int kf_pair_match_count(const uint8_t* d1, const float* a1, int n1, const uint8_t* d2, const float* a2, int n2,
                        float nnratio, bool checkOri, int* matches12 /* may be null */);

// The projection at the head of SearchByProjection(Frame&, const Frame&, th, mono) (ORBmatcher.cc:1376-1393): x3Dc =
// Rcw * x3Dw + tcw (gemm3_f32), invzc = (float)(1.0 / zc), u = fx*xc*invzc + cx, v = fy*yc*invzc + cy in float, source
// order.  Fills u, v, invz per point and valid = invzc >= 0 && u, v inside [minX, maxX] x [minY, maxY] -- the fields of
// ProjQuery that the caller computes today.
void project_points(const float* Rcw /* 3x3 row-major */, const float* tcw, float fx, float fy, float cx, float cy,
                    const float* bounds /* minX, minY, maxX, maxY */, const float* xyzWorld, int n, float* u, float* v,
                    float* invz, int32_t* valid);

// MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:257-322): among the n descriptors observing one map point (in
// mObservations order, bad keyframes removed) the one whose MEDIAN Hamming distance to all of them (its own 0 included,
// median = sorted[(size_t)(0.5*(n-1))]) is smallest; the first such row wins.  Returns its index (-1 for n == 0) and
// the median through *bestMedian.
int distinctive_descriptor(const uint8_t* desc, int n, int* bestMedian);

// Frame::ComputeStereoMatches (Frame.cc:810-984): for every left keypoint the closest right descriptor among the right
// keypoints whose row band covers the left keypoint's row (band = +-2*scale[octave]), octave within +-1, uR in
// [uL - mbf/mb, uL]; accepted below (TH_HIGH+TH_LOW)/2; refined by an 11x11 centre-normalised SAD over 11 horizontal
// shifts on the pyramid level of the left keypoint and a parabola through the three SADs around the best shift; finally
// matches whose SAD is >= 1.5*1.4*median are withdrawn.  Levels are given as the padded buffers of mvImagePyramid
// ((w+38) x (h+38), the level's pixel (0,0) at (19,19)), so that the shifted windows stay addressable.
// Keys are mvKeys / mvKeysRight (NOT undistorted: stereo input is rectified).  Writes mvuRight / mvDepth (-1 = none) and
// the SAD of each accepted match (-1 otherwise); returns the number of stereo matches kept.
// Two deliberate definitions where the reference has undefined behaviour: a right keypoint's row band is clipped to
// the image rows, and with no accepted match at all the median step is skipped (the reference indexes an empty vector).
struct StereoLevel { const uint8_t* padded; int stride; int cols, rows; };
int compute_stereo_matches(const KeyPoint* keysL, const uint8_t* descL, int nL, const KeyPoint* keysR, const uint8_t* descR,
                           int nR, const StereoLevel* levelsL, const StereoLevel* levelsR, int nLevels,
                           const float* scaleFactors, const float* invScaleFactors, float mb, float mbf,
                           float* uRight, float* depth, int* sad);

}  // namespace orbo
