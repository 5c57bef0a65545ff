// Per-cell FAST-9/16 detection (north-star kernel 2): the cell loop of ORBextractor::ComputeKeyPointsOctTree,
// ORBextractor.cc:767-831, with cv::FAST(roi, th, nonmaxSuppression=true) inside.
//
// One CTA per FAST cell per frame; all pyramid levels go in one launch (the cell table carries the level).
// Reference semantics reproduced exactly:
//   * keypoints can only lie in the cell's interior = the 3-px-inset of the (wCell+6)x(hCell+6) ROI; interiors of
//     neighbouring cells tile the detection window without overlap;
//   * score S = cornerScore<16> = (max over the 16 contiguous 9-arcs of min |p_k - v| on one side) - 1; a pixel is a
//     corner at threshold t  <=>  S >= t, so ONE score map serves both thresholds;
//   * non-max suppression is per cell: keep iff S > all 8 neighbours, where neighbours outside the interior and
//     non-corners count as 0;
//   * threshold fallback: the cell is re-run at minThFAST only if NOTHING survives NMS at iniThFAST (:811-818);
//   * emission order inside a cell is row-major (y, x); cells are consumed in (row, col) order by the quadtree kernel.
// Stages: tile (+3 halo) -> shared memory with aligned 32-bit loads; cheap necessary test on all pixels (of each
// opposing circle pair one must be brighter / darker) with survivors compacted into a queue; exact packed-16-bit
// sliding-window score on the queue only; NMS + ordered compaction with a block scan.
#include "extractor.h"

namespace orbb {

constexpr int FAST_THREADS = 128;
constexpr int TILE_PITCH = 72;                    // bytes; 3 (alignment shift) + 66 + slack, multiple of 4
constexpr int TILE_ROWS = kCellMax + 6;
constexpr int SC_PITCH = 64;                      // score map pitch, interior + 1-px zero ring (<= 62)
constexpr int SC_ROWS = kCellMax + 2;

// offsets of the 16 circle pixels in the shared tile, OpenCV order (dx,dy) = (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)
// (0,-3)(-1,-3)(-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)
#define CIRCLE_OFFSET(k)                                                                                           \
    ((k) == 0 ? 3 * TILE_PITCH : (k) == 1 ? 3 * TILE_PITCH + 1 : (k) == 2 ? 2 * TILE_PITCH + 2 : (k) == 3 ? TILE_PITCH + 3 \
     : (k) == 4 ? 3 : (k) == 5 ? -TILE_PITCH + 3 : (k) == 6 ? -2 * TILE_PITCH + 2 : (k) == 7 ? -3 * TILE_PITCH + 1       \
     : (k) == 8 ? -3 * TILE_PITCH : (k) == 9 ? -3 * TILE_PITCH - 1 : (k) == 10 ? -2 * TILE_PITCH - 2                   \
     : (k) == 11 ? -TILE_PITCH - 3 : (k) == 12 ? -3 : (k) == 13 ? TILE_PITCH - 3 : (k) == 14 ? 2 * TILE_PITCH - 2       \
                                                                                           : 3 * TILE_PITCH - 1)

// Necessary condition for S >= t: every opposing pair (k, k+8) holds a pixel of the arc, so all 8 pairs need a
// brighter (> v+t) member, or all 8 a darker (< v-t) one.
__device__ __forceinline__ bool maybe_corner(const unsigned char* c, int t) {
    const int v = c[0];
    const int hi = v + t, lo = v - t;
    int bright = 1, dark = 1;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int a = c[CIRCLE_OFFSET(k)], b = c[CIRCLE_OFFSET(k + 8)];
        bright &= (a > hi) | (b > hi);
        dark &= (a < lo) | (b < lo);
        if (k == 1 && !(bright | dark)) return false;   // two pairs read: most flat pixels leave here
    }
    return (bright | dark) != 0;
}

// Exact threshold-free score. Both polarities ride in one register: low half p_k - v, high half v - p_k.
__device__ __forceinline__ int fast_score(const unsigned char* c) {
    const int v = c[0];
    unsigned int d[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int diff = (int)c[CIRCLE_OFFSET(k)] - v;
        d[k] = ((unsigned int)diff & 0xffffu) | ((unsigned int)(-diff) << 16);
    }
    unsigned int m3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m3[k] = __vimin3_s16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
    unsigned int best = 0x80008000u;   // (-32768, -32768)
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const unsigned int a = __vimin3_s16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
        const unsigned int b = __vimin3_s16x2(m3[k + 1], m3[(k + 4) & 15], m3[(k + 7) & 15]);
        best = __vimax3_s16x2(best, a, b);
    }
    const int lo = (int)(short)(best & 0xffffu), hi = (int)(short)(best >> 16);
    return max(lo, hi) - 1;
}

struct FastShared {
    unsigned int tile[TILE_ROWS * TILE_PITCH / 4];          // pixels; reused as per-pixel NMS flags afterwards
    unsigned int score[SC_ROWS * SC_PITCH / 4];             // uint8 scores with a zero ring
    unsigned short queue[kCellMax * kCellMax];              // pixels that passed the necessary test
    int warpSums[FAST_THREADS / 32];
    int queueLen;
};

__global__ void __launch_bounds__(FAST_THREADS) fast_cells_kernel(const __grid_constant__ ExtractParams P) {
    __shared__ FastShared S;
    const Cell cell = P.cells[blockIdx.x];
    const LevelGeom& L = P.lv[cell.level];
    const int frame = blockIdx.y, tid = threadIdx.x;
    const int cw = cell.cw, ch = cell.ch, npix = cw * ch;
    const unsigned char* level0 = P.pyr + (size_t)frame * P.pyrFrameBytes + L.pyrOff + (size_t)kEdge * L.pitch + kPadLeft;

    // ---- stage the (cw+6) x (ch+6) tile with aligned 32-bit loads
    const int tx0 = cell.x0 - 3, ty0 = cell.y0 - 3;
    const int shift = (tx0 + kPadLeft) & 3;
    const int wordsPerRow = (shift + cw + 6 + 3) >> 2;
    const unsigned char* src = level0 + (long long)ty0 * L.pitch + (tx0 - shift);
    for (int i = tid; i < (ch + 6) * wordsPerRow; i += FAST_THREADS) {
        const int r = i / wordsPerRow, wI = i - r * wordsPerRow;
        S.tile[r * (TILE_PITCH / 4) + wI] = __ldg(reinterpret_cast<const unsigned int*>(src + (size_t)r * L.pitch) + wI);
    }
    for (int i = tid; i < SC_ROWS * SC_PITCH / 4; i += FAST_THREADS) S.score[i] = 0;
    if (tid == 0) S.queueLen = 0;
    __syncthreads();

    const unsigned char* tile = reinterpret_cast<const unsigned char*>(S.tile);
    unsigned char* score = reinterpret_cast<unsigned char*>(S.score);

    // ---- necessary test on every interior pixel, survivors queued
    for (int p = tid; p < npix; p += FAST_THREADS) {
        const int y = p / cw, x = p - y * cw;
        if (maybe_corner(tile + (y + 3) * TILE_PITCH + shift + x + 3, P.minTh)) S.queue[atomicAdd(&S.queueLen, 1)] = (unsigned short)p;
    }
    __syncthreads();
    // ---- exact score for the queue
    const int qn = S.queueLen;
    for (int q = tid; q < qn; q += FAST_THREADS) {
        const int p = S.queue[q];
        const int y = p / cw, x = p - y * cw;
        const int s = fast_score(tile + (y + 3) * TILE_PITCH + shift + x + 3);
        if (s >= P.minTh) score[(y + 1) * SC_PITCH + x + 1] = (unsigned char)s;
    }
    __syncthreads();

    // ---- per-cell NMS; each thread owns a contiguous run of the row-major pixel order
    unsigned char* flags = reinterpret_cast<unsigned char*>(S.tile);
    const int per = (npix + FAST_THREADS - 1) / FAST_THREADS;
    const int p0 = min(tid * per, npix), p1 = min(p0 + per, npix);
    int nIni = 0, nMin = 0;
    for (int p = p0; p < p1; ++p) {
        const int y = p / cw, x = p - y * cw;
        const unsigned char* sc = score + (y + 1) * SC_PITCH + x + 1;
        const int s = sc[0];
        unsigned char f = 0;
        if (s > 0) {
            const int m = max(max(max(sc[-SC_PITCH - 1], sc[-SC_PITCH]), max(sc[-SC_PITCH + 1], sc[-1])),
                              max(max(sc[1], sc[SC_PITCH - 1]), max(sc[SC_PITCH], sc[SC_PITCH + 1])));
            if (s > m) {
                f = s >= P.iniTh ? 2 : 1;
                ++nMin;
                nIni += f == 2;
            }
        }
        flags[p] = f;
    }
    const int anyIni = __syncthreads_or(nIni > 0);   // also orders the flag writes
    const int need = anyIni ? 2 : 1;
    int mine = anyIni ? nIni : nMin;

    // ---- ordered compaction: exclusive block scan of the per-thread counts
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += n;
    }
    if ((tid & 31) == 31) S.warpSums[tid >> 5] = incl;
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int wI = 0; wI < FAST_THREADS / 32; ++wI) {
        const int s = S.warpSums[wI];
        if (wI < (tid >> 5)) base += s;
        total += s;
    }
    int pos = base + incl - mine;
    unsigned int* slot = P.slots + (size_t)frame * P.slotFrameEntries + cell.slot;
    for (int p = p0; p < p1; ++p) {
        if (flags[p] >= need) {
            const int y = p / cw, x = p - y * cw;
            const unsigned int s = score[(y + 1) * SC_PITCH + x + 1];
            slot[pos++] = ((unsigned int)(cell.x0 + x - 16) << 20) | ((unsigned int)(cell.y0 + y - 16) << 8) | s;
        }
    }
    if (tid == 0) P.cellCount[(size_t)frame * P.nCellsTotal + blockIdx.x] = total;
}

int launch_fast(const ExtractParams& P, cudaStream_t st, int* launches) {
    if (P.nCellsTotal == 0) return ORB_OK;
    dim3 grid(P.nCellsTotal, P.nFrames);
    fast_cells_kernel<<<grid, FAST_THREADS, 0, st>>>(P);
    ++*launches;
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orbb
