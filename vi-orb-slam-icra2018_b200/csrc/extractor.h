// Internal layout of the extractor: per-handle geometry, the device arena and the kernel launchers.
//
// HBM layout per handle (all sized once at orbx_create for max_width x max_height x max_batch):
//   pyr    [frame][level]  padded pyramid level, uint8: row pitch = round16(32 + w + 19), 19 + h + 19 rows.
//                          The level's pixel (0,0) sits at row 19, column 32, so interior rows are 16-byte aligned;
//                          columns 13..31 / 32+w..32+w+18 and rows 0..18 / 19+h.. hold the reflect-101 frame
//                          (mvImagePyramid, ORBextractor.cc:1128-1153).
//   blur   [frame][level]  GaussianBlur of the level, uint8, pitch = round16(w)           (ORBextractor.cc:1103-1104)
//   slots  [frame][cell][cap]  FAST keypoints of one 30x30-ish cell, packed x:12|y:12|score:8 in (y,x) order
//   ccount [frame][cell]   keypoints per cell                                               (ORBextractor.cc:790-829)
//   sel    [frame][level][cap]  quadtree survivors: x, y, response, angle                  (ORBextractor.cc:833-854)
//   selcnt [frame][level]
//   keyws  [frame][level]  spill space for candidate lists that do not fit shared memory
#pragma once
#include <vector>

#include "common.cuh"

namespace orbb {

constexpr int kMaxLevels = 16;
constexpr int kPadLeft = 32;   // column of level pixel x=0 inside the padded row
constexpr int kCellMax = 60;   // a FAST cell interior is < 60 px on a side (wCell = ceil(w/floor(w/30)))

struct LevelGeom {
    int w, h;            // level size
    int pitch;           // padded row pitch (bytes)
    int bpitch;          // blurred row pitch
    long long pyrOff;    // byte offset of the padded buffer inside one frame's pyramid block
    long long blurOff;   // byte offset inside one frame's blur block
    int xTab, yTab;      // offsets into the resize tables (entries)
    int pyCol, pyRow;    // offsets of the level's column / row records of the resize kernel (uint4 entries)
    int pyFast;          // 1: the level is resized by pyramid_resize2_kernel from those records
    int pyBand;          // offset of the level's band records (int4 entries) of the staged kernel
    int pyBulk;          // 1: batches are resized by pyramid_resize3_kernel (source rows staged by bulk-async copies)
    int pyBufBytes;      // bytes of one staging buffer
    int pyBulkCtas;      // resident CTAs of that kernel on the device
    int blCtas;          // resident CTAs of the staged blur kernel for this level (0: the level does not fit it)
    int cellBase, nCells;    // this level's cells inside the frame's cell table
    int slotCap;         // entries per cell slot
    long long slotBase;  // entry offset of the level's first slot inside one frame's slot block
    int nFeatures;       // mnFeaturesPerLevel
    int nIni;            // quadtree root count
    float hX;            // root strip width
    int winW, winH;      // maxBorderX-minBorderX, maxBorderY-minBorderY
    int selBase, selCap; // survivors: offset/capacity inside one frame's sel block
    long long keyWsOff;  // entry offset inside one frame's key workspace
    int keyWsCap;
    float scale;         // mvScaleFactor[level]
    float patchSize;     // (float)(int)(31*scale)
};

struct alignas(16) Cell {    // one FAST cell with a non-empty interior; 16 bytes = one 128-bit load
    short level;
    short x0, y0;        // interior origin in level coordinates (first pixel that can be a keypoint)
    short cw, ch;        // interior size
    short pad;
    int slot;            // entry offset of its slot inside one frame's slot block
};
static_assert(sizeof(Cell) == 16, "Cell is read as one uint4");

// Plan of the warp-per-cell FAST kernel (fast_warp.cu), host-computed per image size.  Every warp owns one region of
// the CTA's dynamic shared memory: the tile's mbarrier, the TMA tile buffer (byte tile of bw x bh), the score map with
// its zero ring, the candidate queue (compacted in place into the corner list) and the NMS survivor bitmap.
struct FastWarpPlan {
    int bw, bh;              // TMA box: bytes per tile row (80 / 96) and rows (largest cell + 6)
    int tileBytes;           // bw * bh = bytes one TMA load delivers
    int scorePitch;          // bytes per score-map row (interior + 1-px zero ring)
    int scoreVec;            // uint4 count of the score map
    int lutBytes;            // bit -> pixel tables at the start of the CTA's dynamic shared memory (256 bytes per level)
    int offTile, tileStride, offScore, offQueue, offBitmap, warpBytes;   // inside a warp's region
    int queueCap;            // entries (16 bit each) of the shared-memory queue
    int smemBytes;
    int frameBase;           // arena slot of this launch's frame 0 (the tensor maps are anchored at the arena base)
    const void* maps;        // device array of kMaxLevels CUtensorMap (one per pyramid level)
    unsigned int* counters;  // {next work item, finished warps}: dynamic cell scheduling; the last warp out zeroes both
    unsigned short* scratch; // global-memory queues for cells with more candidates than queueCap (one per resident warp)
    int scratchCap;          // entries per warp: 2 per pixel of the largest cell
    int maxWarps;            // warps the scratch was sized for
    unsigned int one;        // = 1, opaque to the compiler: adds written as IMAD x, one, y go to the FMA pipe instead of the
                             // (binding) ALU pipe, which takes one warp instruction every other clock (profiles/r02_microbench_pipes.txt)
    // pre-test schedule of a level's cells: a step is one pixel row of 4 adjacent 4-px groups per 8 rows (lane = row r
    // of 8, group q of 4); T steps cover the width, chunks of `chunkSteps` (a multiple of T, <= 8) fill one 32-bit mask
    struct Level { unsigned char T, chunkSteps, steps, pad; } lv[kMaxLevels];
};

struct SelKey {          // quadtree survivor in level coordinates
    float x, y, response, angle;
};

struct ExtractParams {
    int nLevels;
    int nFrames;
    int iniTh, minTh;
    int nCellsTotal;
    int selPerFrame;         // entries
    int outCapacity;         // caller's per-frame output capacity
    int maxCellW, maxCellH;  // largest FAST cell interior of this image size (sizes the FAST kernel's shared memory)
    FastWarpPlan fw;
    long long pyrFrameBytes, blurFrameBytes, slotFrameEntries, keyWsFrameEntries;
    unsigned char* pyr;
    unsigned char* blur;
    unsigned int* slots;
    int* cellCount;
    SelKey* sel;
    int* selCount;
    unsigned int* keyWs;
    const Cell* cells;
    const int* tabOfs;       // resize tables: source index per destination index
    const short2* tabCoef;   // 11-bit coefficient pairs
    const uint4* pyColTab;   // pyramid_resize2_kernel: per 4-byte destination group {base, shift, sel01, sel23}, {cf[0..3]}
    const uint4* pyRowTab;   //                         per padded destination row {off(sy0), off(sy0+1), b0 << 12, b1 << 12}
    const void* brMaps;      // brief_staged_kernel: device array of kMaxLevels CUtensorMap over the blurred levels (or null)
    const int4* pyBandTab;   // pyramid_resize3_kernel: per band of 16 destination rows {source offset, bytes, offset(first sy0), -}
    int pyBulkMinFrames;     // batches of at least this many frames use the staged kernel
    unsigned int* pyBarrier; // kMaxLevels counters of the fused small-batch pyramid kernel's grid barrier
    LevelGeom lv[kMaxLevels];
};

struct StereoParams {    // Frame::ComputeStereoMatches: one frame's pyramid block of the left and of the right extractor
    const unsigned char* pyrL;
    const unsigned char* pyrR;
    int nLevels, nRows;          // nRows = mvImagePyramid[0].rows
    float maxD, mbf;             // maxD = mbf / mb
    long long lvOff[kMaxLevels];
    int pitch[kMaxLevels], cols[kMaxLevels], rows[kMaxLevels];
    float scale[kMaxLevels], invScale[kMaxLevels];
};

struct BlurTile { int level, cta; };   // one CTA of the blur kernel: level and CTA index inside the level

// launchers (each returns orb_status and bumps *launches)
int launch_pyramid(const ExtractParams& P, const unsigned char* dImages, int width, int height, int stride,
                   size_t frameStride, cudaStream_t st, int* launches);
bool pyramid_level_plan(const LevelGeom& S, const LevelGeom& D, const int* xofs, const short2* xcoef, const int* yofs,
                        const short2* ycoef, std::vector<uint4>& col, std::vector<uint4>& row);
int pyramid_band_plan(const LevelGeom& S, const LevelGeom& D, const int* yofs, std::vector<int4>& bands);
int pyramid_bulk_ctas(int groups, int bufBytes);
int fast_warp_plan(int nLevels, const int* cellW, const int* cellH, int slotCapMax, FastWarpPlan* plan);
int fast_warp_max_warps(const FastWarpPlan& plan, int* warps);   // resident warps of one launch on the current device
int launch_fast_warp(const ExtractParams& P, cudaStream_t st, int* launches);
// one CUtensorMap (128 bytes, host copy) per level over [frames][rows][pitch] of the padded pyramid arena
int fast_warp_encode_maps(const ExtractParams& P, int arenaFrames, void* hostMaps128xLevels);
int launch_octree(const ExtractParams& P, int smemBytes, int keyCapSmem, int nodeCap, int cellCap, cudaStream_t st,
                  int* launches);
int launch_blur(const ExtractParams& P, const BlurTile* dTiles, int nTiles, cudaStream_t st, int* launches);
int launch_brief(const ExtractParams& P, int maxKeypoints, orb_keypoint* dKps, unsigned char* dDesc, int* dCount,
                 cudaStream_t st, int* launches);
int launch_stereo(const StereoParams& P, const orb_keypoint* dKeysL, const unsigned char* dDescL, int nLmax, const int* dNL,
                  const orb_keypoint* dKeysR, const unsigned char* dDescR, int nRmax, const int* dNR, float* dURight,
                  float* dDepth, int* dSad, int* dKept, cudaStream_t st, int* launches);
int octree_smem_plan(int nodeCap, int cellCap, int* smemBytes, int* keyCapSmem);
int blur_cta_count(int w, int h);
int blur_staged_ctas(const LevelGeom& L);
int upload_brief_pattern();
int brief_encode_maps(const ExtractParams& P, int arenaFrames, void* hostMaps128xLevels);
int upload_orientation_table(const int* umax);

}  // namespace orbb
