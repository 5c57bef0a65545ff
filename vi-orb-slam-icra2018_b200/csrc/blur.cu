// 7x7 sigma-2 Gaussian blur of every pyramid level (north-star kernel 5): ORBextractor.cc:1103-1104,
// cv::GaussianBlur(level, Size(7,7), 2, 2, BORDER_REFLECT_101) on CV_8UC1.
//
// OpenCV's 8-bit path is fixed point and separable: kernel [18 34 48 56 48 34 18]/256 per axis, horizontal sums exact
// (<= 255*256, 16 bits), vertical sums exact in 32 bits, one rounding (v + 2^15) >> 16.  The reference blurs a clone
// of the level, so the border is the level's own reflection -- which is exactly what the pyramid's 19-px frame already
// holds, so the tile is loaded with its 3-px halo straight from the padded level: no border logic here.
// Tile: 128 x 16 outputs per CTA; input staged in shared memory with aligned 32-bit loads, horizontal pass into a
// 16-bit shared buffer, vertical pass from it, 4 output bytes packed per 32-bit store.  All levels in one launch.
#include "extractor.h"

namespace orbb {

constexpr int BL_TW = 128, BL_TH = 16, BL_THREADS = 256;
constexpr int BL_IN_W = BL_TW + 8;           // x0-4 .. x0+TW+3, word aligned
constexpr int BL_IN_ROWS = BL_TH + 6;

__global__ void __launch_bounds__(BL_THREADS) blur_kernel(const __grid_constant__ ExtractParams P, const BlurTile* __restrict__ tiles) {
    __shared__ unsigned int in[BL_IN_ROWS * BL_IN_W / 4];
    __shared__ unsigned short hsum[BL_IN_ROWS * BL_TW];
    const BlurTile t = tiles[blockIdx.x];
    const LevelGeom& L = P.lv[t.level];
    const int frame = blockIdx.y, tid = threadIdx.x;
    const int x0 = t.tx * BL_TW, y0 = t.ty * BL_TH;
    const unsigned char* padded = P.pyr + (size_t)frame * P.pyrFrameBytes + L.pyrOff;
    const int pitchW = L.pitch >> 2, rows = L.h + 2 * kEdge;

    // stage rows y0-3 .. y0+TH+2, columns x0-4 .. x0+TW+3 (buffer coordinates: +19 rows, +32 columns)
    const int gx0w = (x0 - 4 + kPadLeft) >> 2;
    for (int i = tid; i < BL_IN_ROWS * (BL_IN_W / 4); i += BL_THREADS) {
        const int r = i / (BL_IN_W / 4), wI = i - r * (BL_IN_W / 4);
        const int gr = y0 - 3 + r + kEdge, gw = gx0w + wI;
        unsigned int v = 0;
        if (gr < rows && gw < pitchW) v = __ldg(reinterpret_cast<const unsigned int*>(padded + (size_t)gr * L.pitch) + gw);
        in[i] = v;
    }
    __syncthreads();
    const unsigned char* inb = reinterpret_cast<const unsigned char*>(in);
    for (int i = tid; i < BL_IN_ROWS * BL_TW; i += BL_THREADS) {
        const int r = i / BL_TW, x = i - r * BL_TW;
        const unsigned char* p = inb + r * BL_IN_W + x + 1;   // column x-3 of the tile row
        hsum[i] = (unsigned short)(18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3]);
    }
    __syncthreads();
    unsigned char* out = P.blur + (size_t)frame * P.blurFrameBytes + L.blurOff;
    for (int i = tid; i < BL_TH * (BL_TW / 4); i += BL_THREADS) {
        const int r = i / (BL_TW / 4), xq = (i - r * (BL_TW / 4)) * 4;
        const int gy = y0 + r, gx = x0 + xq;
        if (gy >= L.h || gx >= L.w) continue;
        unsigned int word = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned short* h = hsum + r * BL_TW + xq + k;
            const unsigned int v = 18u * (h[0] + h[6 * BL_TW]) + 34u * (h[BL_TW] + h[5 * BL_TW]) +
                                   48u * (h[2 * BL_TW] + h[4 * BL_TW]) + 56u * h[3 * BL_TW];
            word |= ((v + 32768u) >> 16) << (8 * k);
        }
        *reinterpret_cast<unsigned int*>(out + (size_t)gy * L.bpitch + gx) = word;
    }
}

int launch_blur(const ExtractParams& P, const BlurTile* dTiles, int nTiles, cudaStream_t st, int* launches) {
    if (nTiles == 0) return ORB_OK;
    dim3 grid(nTiles, P.nFrames);
    blur_kernel<<<grid, BL_THREADS, 0, st>>>(P, dTiles);
    ++*launches;
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

int blur_tile_dims(int* tw, int* th) { *tw = BL_TW; *th = BL_TH; return ORB_OK; }

}  // namespace orbb
