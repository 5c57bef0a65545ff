// TEST INFRASTRUCTURE ONLY -- CPU oracle of the reference's ORB extractor, never part of the product path.
//
// Restates /root/reference/src/ORBextractor.cc on plain arrays:
//   ctor tables            :412-472      ComputePyramid           :1128-1153
//   IC_Angle               :79-106       ComputeKeyPointsOctTree  :767-855
//   computeOrbDescriptor   :110-149      DivideNode               :483-539
//   operator()             :1045-1126    DistributeOctTree        :541-765
// OpenCV primitives come from cvprims.{h,cpp} (pinned against cv2 4.13.0).
//
// One rule is OURS, not the reference's: std::sort at :686 orders pair<int, ExtractorNode*>, i.e. ties on
// node size are broken by heap address, which the reference leaves allocator-dependent.  The oracle
// defines the address order as creation order (what a monotonic allocator yields): a node created later
// compares greater.  oracle/_ref (the reference's own file compiled against a shim) enforces the same rule
// with a bump allocator, see oracle/ref_shim/.
#pragma once
#include <cstdint>
#include <vector>

#include "cvprims.h"

namespace orbo {

constexpr int kPatchSize = 31;
constexpr int kHalfPatch = 15;
constexpr int kEdge = 19;

struct PyramidLevel {
    int w = 0, h = 0;       // level size (without the 19-px frame)
    int stride = 0;         // = w + 2*kEdge
    std::vector<uint8_t> padded;  // (w+38) x (h+38), reflect-101 frame
    const uint8_t* roi() const { return padded.data() + (size_t)kEdge * stride + kEdge; }
    uint8_t* roi() { return padded.data() + (size_t)kEdge * stride + kEdge; }
};

// DistributeOctTree on its own (ORBextractor.cc:541-765). Keys are in detection-window coordinates.
std::vector<KeyPoint> distribute_octree(const std::vector<KeyPoint>& keys, int minX, int maxX, int minY,
                                        int maxY, int N);

class Extractor {
public:
    Extractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);

    // ORBextractor::operator(): returns keypoints (level-0 coordinates) and nkp x 32 descriptor bytes.
    // Returns false when the reference would hit undefined behaviour (a level too small for the cell grid).
    bool extract(const uint8_t* img, int w, int h, int stride, std::vector<KeyPoint>& kps,
                 std::vector<uint8_t>& desc);

    // tables (ORBextractor.h:81-101 getters)
    int nfeatures, nlevels, iniTh, minTh;
    double scaleFactor;
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> featuresPerLevel, umax;

    // intermediates of the last extract(), kept for stage-by-stage parity tests
    std::vector<PyramidLevel> pyramid;
    std::vector<std::vector<uint8_t>> blurred;           // per level, w*h, empty if level had no keypoints
    std::vector<std::vector<KeyPoint>> candidates;        // per level, before the quadtree (window coords)
    std::vector<std::vector<KeyPoint>> selected;          // per level, level coords, with angle
    double msPyramid = 0, msKeypoints = 0, msDescriptors = 0;  // ORBextractor.h:51-53

    void computePyramid(const uint8_t* img, int w, int h, int stride);
    bool computeKeyPoints();
};

float ic_angle(const uint8_t* center, int stride, const std::vector<int>& umax);
void orb_descriptor(float angleDeg, const uint8_t* center, int stride, uint8_t* desc32);
const int8_t* brief_pattern();  // 1024 values

}  // namespace orbo
