// 7x7 sigma-2 Gaussian blur of every pyramid level (north-star kernel 5): ORBextractor.cc:1103-1104,
// cv::GaussianBlur(level, Size(7,7), 2, 2, BORDER_REFLECT_101) on CV_8UC1.
//
// OpenCV's 8-bit path is fixed point and separable: kernel [18 34 48 56 48 34 18]/256 per axis, horizontal sums exact
// (<= 255*256, 16 bits), vertical sums exact in 32 bits, one rounding (v + 2^15) >> 16.  The reference blurs a clone
// of the level, so the border is the level's own reflection -- which is exactly what the pyramid's 19-px frame already
// holds, so the halo is read straight from the padded level: no border logic here.
//
// The kernel is instruction-bound, not DRAM-bound (first version: ~75 instructions per pixel, ncu), so it is built
// around registers, not shared memory:
//   * a thread owns a 4-pixel-wide column (one 32-bit word per row) and walks down 16 output rows; a warp is a
//     128-pixel-wide band.  Per input row each lane loads ONE aligned word; the words to its left and right come from
//     the neighbouring lanes by shuffle (the two edge lanes load theirs);
//   * horizontal pass on packed pairs: two adjacent pixels sit in the two 16-bit halves of a register, so one IMAD
//     advances two horizontal sums (each sum <= 65280 fits its half; symmetric taps are added first);
//   * vertical pass from a 7-row register ring (the row loop is fully unrolled, ring indices are compile-time);
//     accumulators start at 2^15 so the rounding is free, and the four result bytes are picked with PRMT.
// No shared memory, no barriers; all levels in one launch.
#include "extractor.h"

namespace orbb {

constexpr int BL_ROWS = 16;                 // output rows per warp
constexpr int BL_WARPS = 4;                 // warps per CTA, stacked vertically
constexpr int BL_TW = 128, BL_TH = BL_ROWS * BL_WARPS;

// pair (b[j], b[j+1]) of the 12 bytes w0|w1|w2 (j counted from the first byte of w1), zero-extended to 16x2
__device__ __forceinline__ unsigned int pair_at(unsigned int w0, unsigned int w1, unsigned int w2, int j) {
    // compile-time j in [-3, 5]
    switch (j) {
        case -3: return __byte_perm(w0, 0, 0x4241);
        case -2: return __byte_perm(w0, 0, 0x4342);
        case -1: return __byte_perm(w0, w1, 0x7473) & 0x00ff00ffu;   // (w0.b3, w1.b0)
        case 0: return __byte_perm(w1, 0, 0x4140);
        case 1: return __byte_perm(w1, 0, 0x4241);
        case 2: return __byte_perm(w1, 0, 0x4342);
        case 3: return __byte_perm(w1, w2, 0x7473) & 0x00ff00ffu;    // (w1.b3, w2.b0)
        case 4: return __byte_perm(w2, 0, 0x4140);
        default: return __byte_perm(w2, 0, 0x4241);
    }
}

__global__ void __launch_bounds__(BL_WARPS * 32) blur_kernel(const __grid_constant__ ExtractParams P, const BlurTile* __restrict__ tiles) {
    const BlurTile t = tiles[blockIdx.x];
    const LevelGeom& L = P.lv[t.level];
    const int frame = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = t.tx * BL_TW + lane * 4;
    const int y0 = t.ty * BL_TH + warp * BL_ROWS;
    if (y0 >= L.h) return;
    const bool loads = x0 < L.w + 4;          // this lane's word is needed (by itself or by its left neighbour)
    const bool active = x0 < L.w;
    const int lastRow = L.h + 2 * kEdge - 1;
    const unsigned char* col = P.pyr + (size_t)frame * P.pyrFrameBytes + L.pyrOff + kPadLeft + x0;
    unsigned char* out = P.blur + (size_t)frame * P.blurFrameBytes + L.blurOff + x0;

    unsigned int ring[7][4];                  // horizontal sums of the last 7 input rows, one 32-bit value per pixel
#pragma unroll
    for (int r = 0; r < BL_ROWS + 6; ++r) {
        const int gr = min(y0 - 3 + r + kEdge, lastRow);        // buffer row (frame rows included)
        const unsigned char* p = col + (size_t)gr * L.pitch;
        unsigned int w1 = 0;
        if (loads) w1 = __ldg(reinterpret_cast<const unsigned int*>(p));
        unsigned int w0 = __shfl_up_sync(0xffffffffu, w1, 1);
        unsigned int w2 = __shfl_down_sync(0xffffffffu, w1, 1);
        if (lane == 0 && loads) w0 = __ldg(reinterpret_cast<const unsigned int*>(p - 4));
        if (lane == 31 && active) w2 = __ldg(reinterpret_cast<const unsigned int*>(p + 4));
        // horizontal: pixels (x0, x0+1) use pairs j = -3..3, pixels (x0+2, x0+3) use j = -1..5
        const unsigned int pm3 = pair_at(w0, w1, w2, -3), pm2 = pair_at(w0, w1, w2, -2), pm1 = pair_at(w0, w1, w2, -1),
                           p0 = pair_at(w0, w1, w2, 0), p1 = pair_at(w0, w1, w2, 1), p2 = pair_at(w0, w1, w2, 2),
                           p3 = pair_at(w0, w1, w2, 3), p4 = pair_at(w0, w1, w2, 4), p5 = pair_at(w0, w1, w2, 5);
        const unsigned int hA = 18u * (pm3 + p3) + 34u * (pm2 + p2) + 48u * (pm1 + p1) + 56u * p0;
        const unsigned int hB = 18u * (pm1 + p5) + 34u * (p0 + p4) + 48u * (p1 + p3) + 56u * p2;
        unsigned int* slot = ring[r % 7];
        slot[0] = hA & 0xffffu; slot[1] = hA >> 16; slot[2] = hB & 0xffffu; slot[3] = hB >> 16;
        if (r >= 6) {
            const int gy = y0 + r - 6;
            unsigned int v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                v[k] = 32768u + 18u * (ring[(r - 6) % 7][k] + ring[r % 7][k]) + 34u * (ring[(r - 5) % 7][k] + ring[(r - 1) % 7][k]) +
                       48u * (ring[(r - 4) % 7][k] + ring[(r - 2) % 7][k]) + 56u * ring[(r - 3) % 7][k];
            // byte 2 of each sum is (v >> 16) & 0xff (sums stay below 2^24)
            const unsigned int word = __byte_perm(__byte_perm(v[0], v[1], 0x0062), __byte_perm(v[2], v[3], 0x0062), 0x5410);
            if (active && gy < L.h) *reinterpret_cast<unsigned int*>(out + (size_t)gy * L.bpitch) = word;
        }
    }
}

int launch_blur(const ExtractParams& P, const BlurTile* dTiles, int nTiles, cudaStream_t st, int* launches) {
    if (nTiles == 0) return ORB_OK;
    dim3 grid(nTiles, P.nFrames);
    blur_kernel<<<grid, BL_WARPS * 32, 0, st>>>(P, dTiles);
    ++*launches;
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

int blur_tile_dims(int* tw, int* th) { *tw = BL_TW; *th = BL_TH; return ORB_OK; }

}  // namespace orbb
