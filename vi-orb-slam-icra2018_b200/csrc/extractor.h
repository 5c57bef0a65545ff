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

struct Cell {            // one FAST cell with a non-empty interior
    short level;
    short x0, y0;        // interior origin in level coordinates (first pixel that can be a keypoint)
    short cw, ch;        // interior size
    short pad;
    int slot;            // entry offset of its slot inside one frame's slot block
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
    int umax[16];
    LevelGeom lv[kMaxLevels];
};

struct BlurTile { short level, tx, ty, pad; };

// launchers (each returns orb_status and bumps *launches)
int launch_pyramid(const ExtractParams& P, const unsigned char* dImages, int width, int height, int stride,
                   size_t frameStride, cudaStream_t st, int* launches);
int launch_fast(const ExtractParams& P, cudaStream_t st, int* launches);
int launch_octree(const ExtractParams& P, int smemBytes, int keyCapSmem, int nodeCap, int cellCap, cudaStream_t st,
                  int* launches);
int launch_blur(const ExtractParams& P, const BlurTile* dTiles, int nTiles, cudaStream_t st, int* launches);
int launch_brief(const ExtractParams& P, int maxKeypoints, orb_keypoint* dKps, unsigned char* dDesc, int* dCount,
                 cudaStream_t st, int* launches);
int octree_smem_plan(int nodeCap, int cellCap, int* smemBytes, int* keyCapSmem);
int blur_tile_dims(int* tw, int* th);
int upload_brief_pattern();

}  // namespace orbb
