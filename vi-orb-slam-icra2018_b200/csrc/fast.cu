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
//
// Instruction budget is what bounds this kernel (ncu: issue-bound, DRAM < 2%), so the stages are shaped for dense
// lanes and packed arithmetic:
//   0. tile (+halo) -> shared memory, widened to 16 bits per pixel so that two horizontally adjacent pixels are one
//      ready-made 16x2 operand;
//   A. necessary test on EVERY pixel, 4 pixels per thread, no divergence: of each opposing circle pair (k, k+8) one
//      pixel lies on any 9-arc, so min_j max(p_j, p_j+8) > v+t (or max_j min(..) < v-t) must hold.  All in DPX 16x2
//      min/max (VIMNMX); neighbours at odd offsets come from 16-bit funnel shifts.  Survivors are queued;
//   B. exact score on the queue only (dense lanes again): both polarities in one 16x2 register, sliding-window minimum
//      with 3-input min/max;
//   C. NMS on the queue entries, survivors set bits in two row-major bitmaps (>= iniTh, >= minTh);
//   D. one warp prefix-sums the popcounts of the chosen bitmap; E. survivors store themselves at their rank.
#include "extractor.h"

namespace orbb {

constexpr int FAST_THREADS = 128;
constexpr int TPX = 72;                           // tile pitch in pixels (16-bit each); interior x=0 sits at column 4
constexpr int TILE_ROWS = kCellMax + 6;
constexpr int SC_PITCH = 64;                      // score map pitch (bytes), interior + 1-px zero ring (<= 62)
constexpr int SC_ROWS = kCellMax + 2;
constexpr int BM_WORDS = (kCellMax * kCellMax + 31) / 32;   // 113

// circle offsets in tile pixels, OpenCV order (dx,dy) = (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)
// (-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)
#define CIRCLE_OFFSET(k)                                                                                   \
    ((k) == 0 ? 3 * TPX : (k) == 1 ? 3 * TPX + 1 : (k) == 2 ? 2 * TPX + 2 : (k) == 3 ? TPX + 3             \
     : (k) == 4 ? 3 : (k) == 5 ? -TPX + 3 : (k) == 6 ? -2 * TPX + 2 : (k) == 7 ? -3 * TPX + 1              \
     : (k) == 8 ? -3 * TPX : (k) == 9 ? -3 * TPX - 1 : (k) == 10 ? -2 * TPX - 2 : (k) == 11 ? -TPX - 3     \
     : (k) == 12 ? -3 : (k) == 13 ? TPX - 3 : (k) == 14 ? 2 * TPX - 2 : 3 * TPX - 1)

// Dynamic shared memory, sized by the largest cell of the current image size (typical 36x34 cells: 13 KB, so the SM holds
// enough CTAs to hide the barriers between the phases; the 60x60 worst case needs 29 KB).
struct FastLayout {
    int tile, score, queue, alive, bmMin, bmIni, total, bmWords;
};
__host__ __device__ inline FastLayout fast_layout(int maxW, int maxH) {
    FastLayout f;
    const int npix = maxW * maxH;
    int p = 0;
    auto take = [&](int bytes) { int r = p; p += (bytes + 15) & ~15; return r; };
    f.tile = take((maxH + 6) * TPX * 2);          // 16-bit pixels; reused as the survivor list after phase B
    f.score = take((maxH + 2) * SC_PITCH);        // uint8 scores with a zero ring
    f.queue = take(npix * 2);                     // pixels that pass the necessary test: x | y<<6
    f.alive = take(npix * 2);                     // corners (S >= minTh): x | y<<6 | ini<<14
    f.bmWords = ((npix + 31) >> 5) + 1;
    f.bmMin = take(f.bmWords * 4);                // survivors at minTh / iniTh, bit = y*cw + x
    f.bmIni = take(f.bmWords * 4);
    f.total = p;
    return f;
}

// Exact threshold-free score. Both polarities ride in one register: low half p_k - v, high half v - p_k.
__device__ __forceinline__ int fast_score(const unsigned short* c) {
    const int v = c[0];
    unsigned int d[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int diff = (int)c[CIRCLE_OFFSET(k)] - v;
        d[k] = __byte_perm((unsigned int)diff, (unsigned int)(-diff), 0x5410);
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

// per-half test of the necessary condition: bright possible (mm >= v+t+1) or dark possible (v >= nn+t+1); 2 result bits
__device__ __forceinline__ unsigned int pass_bits(unsigned int mm, unsigned int nn, unsigned int c, unsigned int T1) {
    bool bh, bl, dh, dl;
    __vibmax_u16x2(mm, c + T1, &bh, &bl);
    __vibmax_u16x2(c, nn + T1, &dh, &dl);
    return ((bl | dl) ? 1u : 0u) | ((bh | dh) ? 2u : 0u);
}

__global__ void __launch_bounds__(FAST_THREADS, 10) fast_cells_kernel(const __grid_constant__ ExtractParams P) {
    extern __shared__ __align__(16) unsigned char fsm[];
    __shared__ int sQueueLen, sAliveLen, sSurvLen, sIniLen;
    const FastLayout lay = fast_layout(P.maxCellW, P.maxCellH);
    unsigned int* sTile = reinterpret_cast<unsigned int*>(fsm + lay.tile);
    unsigned int* sScore = reinterpret_cast<unsigned int*>(fsm + lay.score);
    unsigned short* sQueue = reinterpret_cast<unsigned short*>(fsm + lay.queue);
    unsigned short* sAlive = reinterpret_cast<unsigned short*>(fsm + lay.alive);
    unsigned int* sBmMin = reinterpret_cast<unsigned int*>(fsm + lay.bmMin);
    unsigned int* sBmIni = reinterpret_cast<unsigned int*>(fsm + lay.bmIni);
    const Cell cell = P.cells[blockIdx.x];
    const LevelGeom& L = P.lv[cell.level];
    const int frame = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const int cw = cell.cw, ch = cell.ch;
    const unsigned char* level0 = P.pyr + (size_t)frame * P.pyrFrameBytes + L.pyrOff + (size_t)kEdge * L.pitch + kPadLeft;

    // ---- 0. stage level pixels [x0-4, x0+cw+4) x [y0-3, y0+ch+3) as 16-bit values; aligned 32-bit global loads,
    //         funnel shift to the tile's alignment, widen, one 64-bit shared store per 4 pixels
    {
        const int tx0 = cell.x0 - 4, ty0 = cell.y0 - 3;
        const int shift8 = ((tx0 + kPadLeft) & 3) * 8;
        const int quads = (cw + 8 + 3) >> 2;                 // 4-pixel groups per tile row (<= 17)
        const unsigned int rq = (65536u + quads - 1) / quads;   // i / quads == (i * rq) >> 16 for i * quads < 65536
        const unsigned char* src = level0 + (long long)ty0 * L.pitch + (tx0 - (shift8 >> 3));
        uint2* tile64 = reinterpret_cast<uint2*>(sTile);
        for (int i = tid; i < (ch + 6) * quads; i += FAST_THREADS) {
            const int r = (int)(((unsigned int)i * rq) >> 16), q = i - r * quads;
            const unsigned int* g = reinterpret_cast<const unsigned int*>(src + (size_t)r * L.pitch) + q;
            const unsigned int w0 = __ldg(g), w1 = __ldg(g + 1);
            const unsigned int px = __funnelshift_r(w0, w1, shift8);
            tile64[r * (TPX / 4) + q] = make_uint2(__byte_perm(px, 0, 0x4140), __byte_perm(px, 0, 0x4342));
        }
        uint4* sc4 = reinterpret_cast<uint4*>(sScore);
        for (int i = tid; i < (ch + 2) * (SC_PITCH / 16); i += FAST_THREADS) sc4[i] = make_uint4(0, 0, 0, 0);
        for (int i = tid; i < ((cw * ch + 31) >> 5) + 1; i += FAST_THREADS) { sBmMin[i] = 0; sBmIni[i] = 0; }
        if (tid == 0) { sQueueLen = 0; sAliveLen = 0; sSurvLen = 0; sIniLen = 0; }
    }
    __syncthreads();

    // ---- A. necessary test, 4 pixels (two 16x2 pairs) per thread; warp-uniform loop so that a warp whose 128 pixels
    //         all fail after the three middle-row pairs skips the other five
    {
        const int groups = (cw + 3) >> 2, total = groups * ch;
        const unsigned int rg = (65536u + groups - 1) / groups;
        const unsigned int T1 = (unsigned int)(P.minTh + 1) * 0x00010001u;
        const uint2* tile64 = reinterpret_cast<const uint2*>(sTile);
        for (int i0 = tid - lane; i0 < total; i0 += FAST_THREADS) {
            const int i = min(i0 + lane, total - 1);
            const bool valid = i0 + lane < total;
            const int y = (int)(((unsigned int)i * rg) >> 16), g = i - y * groups;
            const uint2* row = tile64 + (y + 3) * (TPX / 4) + g;   // row[0] = pixels x0-4..x0-1, row[1] = x0..x0+3, row[2] = x0+4..
            // A = pixels (x0, x0+1), B = (x0+2, x0+3). M = pair maxima, m = pair minima of the opposing circle points
            unsigned int MA[8], MB[8], mA[8], mB[8], cA, cB;
#define PAIR(j, aA, bA, aB, bB)                                                            \
    MA[j] = __vmaxu2(aA, bA); mA[j] = __vminu2(aA, bA); MB[j] = __vmaxu2(aB, bB); mB[j] = __vminu2(aB, bB);
            {   // rows +-1 and 0: dx = +-3 -> circle points 3/11, 5/13, 4/12
                const uint2* up = row - (TPX / 4);
                const uint2* dn = row + (TPX / 4);
                const uint2 ul = up[0], uc = up[1], ur = up[2], dl = dn[0], dc = dn[1], dr = dn[2];
                const uint2 ml = row[0], mc = row[1], mr = row[2];
                cA = mc.x; cB = mc.y;
                // +3: A' = (x0+3, x0+4), B' = (x0+5, x0+6);  -3: A' = (x0-3, x0-2), B' = (x0-1, x0)
                const unsigned int uPA = __funnelshift_r(uc.y, ur.x, 16), uPB = __funnelshift_r(ur.x, ur.y, 16);
                const unsigned int uMA = __funnelshift_r(ul.x, ul.y, 16), uMB = __funnelshift_r(ul.y, uc.x, 16);
                const unsigned int dPA = __funnelshift_r(dc.y, dr.x, 16), dPB = __funnelshift_r(dr.x, dr.y, 16);
                const unsigned int dMA = __funnelshift_r(dl.x, dl.y, 16), dMB = __funnelshift_r(dl.y, dc.x, 16);
                const unsigned int mPA = __funnelshift_r(mc.y, mr.x, 16), mPB = __funnelshift_r(mr.x, mr.y, 16);
                const unsigned int mMA = __funnelshift_r(ml.x, ml.y, 16), mMB = __funnelshift_r(ml.y, mc.x, 16);
                PAIR(0, dPA, uMA, dPB, uMB)                                        // k=3 (3,1) with k=11 (-3,-1)
                PAIR(1, uPA, dMA, uPB, dMB)                                        // k=5 (3,-1) with k=13 (-3,1)
                PAIR(2, mPA, mMA, mPB, mMB)                                        // k=4 (3,0) with k=12 (-3,0)
            }
            unsigned int mmA = __vimin3_u16x2(MA[0], MA[1], MA[2]), mmB = __vimin3_u16x2(MB[0], MB[1], MB[2]);
            unsigned int nnA = __vimax3_u16x2(mA[0], mA[1], mA[2]), nnB = __vimax3_u16x2(mB[0], mB[1], mB[2]);
            unsigned int flags = pass_bits(mmA, nnA, cA, T1) | (pass_bits(mmB, nnB, cB, T1) << 2);
            if (!__any_sync(0xffffffffu, valid && flags)) continue;
            {   // rows +-3: dx = 0, +1, -1  -> circle points 0/8, 1/9, 15/7
                const uint2* up = row - 3 * (TPX / 4);
                const uint2* dn = row + 3 * (TPX / 4);
                const unsigned int u1 = up[0].y, u4 = up[2].x, d1 = dn[0].y, d4 = dn[2].x;
                const uint2 uc = up[1], dc = dn[1];
                const unsigned int uf12 = __funnelshift_r(u1, uc.x, 16), uf23 = __funnelshift_r(uc.x, uc.y, 16),
                                   uf34 = __funnelshift_r(uc.y, u4, 16);
                const unsigned int df12 = __funnelshift_r(d1, dc.x, 16), df23 = __funnelshift_r(dc.x, dc.y, 16),
                                   df34 = __funnelshift_r(dc.y, d4, 16);
                PAIR(3, dc.x, uc.x, dc.y, uc.y)                                    // k=0 (0,3) with k=8 (0,-3)
                PAIR(4, df23, uf12, df34, uf23)                                    // k=1 (1,3) with k=9 (-1,-3)
                PAIR(5, uf23, df12, uf34, df23)                                    // k=7 (1,-3) with k=15 (-1,3)
            }
            {   // rows +-2: dx = +-2 -> circle points 2/10, 6/14; no shifts
                const uint2* up = row - 2 * (TPX / 4);
                const uint2* dn = row + 2 * (TPX / 4);
                const unsigned int u1 = up[0].y, u4 = up[2].x, d1 = dn[0].y, d4 = dn[2].x;
                const uint2 uc = up[1], dc = dn[1];
                PAIR(6, dc.y, u1, d4, uc.x)                                        // k=2 (2,2) with k=10 (-2,-2)
                PAIR(7, uc.y, d1, u4, dc.x)                                        // k=6 (2,-2) with k=14 (-2,2)
            }
#undef PAIR
            mmA = __vimin3_u16x2(mmA, __vimin3_u16x2(MA[3], MA[4], MA[5]), __vminu2(MA[6], MA[7]));
            mmB = __vimin3_u16x2(mmB, __vimin3_u16x2(MB[3], MB[4], MB[5]), __vminu2(MB[6], MB[7]));
            nnA = __vimax3_u16x2(nnA, __vimax3_u16x2(mA[3], mA[4], mA[5]), __vmaxu2(mA[6], mA[7]));
            nnB = __vimax3_u16x2(nnB, __vimax3_u16x2(mB[3], mB[4], mB[5]), __vmaxu2(mB[6], mB[7]));
            const int x0 = 4 * g;
            flags = pass_bits(mmA, nnA, cA, T1) | (pass_bits(mmB, nnB, cB, T1) << 2);
            flags &= (1u << min(4, cw - x0)) - 1u;
            if (valid && flags) {
                int pos = atomicAdd(&sQueueLen, __popc(flags));
                const unsigned int base = (unsigned int)x0 | ((unsigned int)y << 6);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (flags & (1u << k)) sQueue[pos++] = (unsigned short)(base + k);
            }
        }
    }
    __syncthreads();

    // ---- B. exact score for the queue; corners go to the dense `alive` list
    const int qn = sQueueLen;
    const unsigned short* tile16 = reinterpret_cast<const unsigned short*>(sTile);
    unsigned char* score = reinterpret_cast<unsigned char*>(sScore);
    for (int q = tid; q < qn; q += FAST_THREADS) {
        const unsigned int e = sQueue[q];
        const int x = e & 63, y = (e >> 6) & 63;
        const int s = fast_score(tile16 + (y + 3) * TPX + x + 4);
        if (s >= P.minTh) {
            score[(y + 1) * SC_PITCH + x + 1] = (unsigned char)s;
            sAlive[atomicAdd(&sAliveLen, 1)] = (unsigned short)(e | (s >= P.iniTh ? 0x4000u : 0u));
        }
    }
    __syncthreads();

    // ---- C. per-cell NMS on the corners; survivors are listed (in the tile, no longer needed) and set bitmap bits
    const int an = sAliveLen;
    unsigned short* surv = reinterpret_cast<unsigned short*>(sTile);
    for (int q = tid; q < an; q += FAST_THREADS) {
        const unsigned int e = sAlive[q];
        const int x = e & 63, y = (e >> 6) & 63;
        const unsigned char* sc = score + (y + 1) * SC_PITCH + x + 1;
        const int s = sc[0];
        const int m = max(max(max(sc[-SC_PITCH - 1], sc[-SC_PITCH]), max(sc[-SC_PITCH + 1], sc[-1])),
                          max(max(sc[1], sc[SC_PITCH - 1]), max(sc[SC_PITCH], sc[SC_PITCH + 1])));
        if (s > m) {
            const int bit = y * cw + x;
            atomicOr(&sBmMin[bit >> 5], 1u << (bit & 31));
            if (e & 0x4000u) {
                atomicOr(&sBmIni[bit >> 5], 1u << (bit & 31));
                atomicAdd(&sIniLen, 1);
            }
            surv[atomicAdd(&sSurvLen, 1)] = (unsigned short)e;
        }
    }
    __syncthreads();

    // ---- D. survivors of the chosen threshold store themselves at their row-major rank = number of set bits below
    //         their own in the bitmap (a few dozen POPCs each; there are only ~15 survivors per cell)
    const int nIni = sIniLen, sn = sSurvLen;
    const bool useIni = nIni > 0;
    const unsigned int* bm = useIni ? sBmIni : sBmMin;
    if (tid == 0) P.cellCount[(size_t)frame * P.nCellsTotal + blockIdx.x] = useIni ? nIni : sn;
    unsigned int* slot = P.slots + (size_t)frame * P.slotFrameEntries + cell.slot;
    for (int q = tid; q < sn; q += FAST_THREADS) {
        const unsigned int e = surv[q];
        if (useIni && !(e & 0x4000u)) continue;
        const int x = e & 63, y = (e >> 6) & 63;
        const int bit = y * cw + x;
        int rank = __popc(bm[bit >> 5] & ((1u << (bit & 31)) - 1u));
        for (int w = 0; w < (bit >> 5); ++w) rank += __popc(bm[w]);
        const unsigned int s = score[(y + 1) * SC_PITCH + x + 1];
        slot[rank] = ((unsigned int)(cell.x0 + x - 16) << 20) | ((unsigned int)(cell.y0 + y - 16) << 8) | s;
    }
}

int launch_fast(const ExtractParams& P, cudaStream_t st, int* launches) {
    if (P.nCellsTotal == 0) return ORB_OK;
    dim3 grid(P.nCellsTotal, P.nFrames);
    const int smem = fast_layout(P.maxCellW, P.maxCellH).total;
    fast_cells_kernel<<<grid, FAST_THREADS, smem, st>>>(P);
    ++*launches;
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orbb
