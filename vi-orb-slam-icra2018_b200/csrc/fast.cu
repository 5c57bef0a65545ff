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
// lanes and packed arithmetic, and everything that is the same for all frames (index divisions, the shared-memory
// plan, the tile's address) is precomputed on the host into the cell table:
//   0. tile (+halo) -> shared memory, widened to 16 bits per pixel so that two horizontally adjacent pixels are one
//      ready-made 16x2 operand; a thread owns one 4-pixel column group and walks down the rows;
//   A. necessary test on EVERY pixel, 4 pixels per thread, no divergence: of each opposing circle pair (k, k+8) one
//      pixel lies on any 9-arc, so min_j max(p_j, p_j+8) > v+t (or max_j min(..) < v-t) must hold.  All in DPX 16x2
//      min/max (VIMNMX); neighbours at odd offsets come from 16-bit funnel shifts.  The same minima/maxima are tested
//      against BOTH thresholds: pixels that can be corners at iniThFAST are queued from the front, those that can
//      only be corners at minThFAST from the back;
//   B. exact score on the front queue only (dense lanes again): both polarities in one 16x2 register, sliding-window
//      minimum with 3-input min/max;
//   C. NMS on the corners with S >= iniThFAST; survivors are listed;
//   B', C'. only if nothing survived: the back queue is scored too and NMS runs at minThFAST;
//   D. survivors store themselves at their row-major rank (count of smaller keys in the list).
#include "extractor.h"

namespace orbb {

constexpr int FAST_THREADS = 128;
constexpr int TPX = 68;                           // tile pitch in pixels (16-bit each); interior x=0 sits at column 4.
                                                  // 17 eight-byte words per row: with items two rows apart, 2 * 17 = 2 (mod
                                                  // 16), so 8 row pairs x 2 adjacent groups cover all 16 bank pairs
constexpr int SC_PITCH = 68;                      // score map pitch (bytes), interior + 1-px zero ring (<= 62); 17 words:
                                                  // odd, so vertically adjacent corners fall into different banks

// circle offsets in tile pixels, OpenCV order (dx,dy) = (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)
// (-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)
#define CIRCLE_OFFSET(k)                                                                                   \
    ((k) == 0 ? 3 * TPX : (k) == 1 ? 3 * TPX + 1 : (k) == 2 ? 2 * TPX + 2 : (k) == 3 ? TPX + 3             \
     : (k) == 4 ? 3 : (k) == 5 ? -TPX + 3 : (k) == 6 ? -2 * TPX + 2 : (k) == 7 ? -3 * TPX + 1              \
     : (k) == 8 ? -3 * TPX : (k) == 9 ? -3 * TPX - 1 : (k) == 10 ? -2 * TPX - 2 : (k) == 11 ? -TPX - 3     \
     : (k) == 12 ? -3 : (k) == 13 ? TPX - 3 : (k) == 14 ? 2 * TPX - 2 : 3 * TPX - 1)

// Dynamic shared memory, sized by the largest cell of the current image size (typical 36x34 cells: 12 KB, so the SM holds
// enough CTAs to hide the barriers between the phases; the 60x60 worst case needs 28 KB).
FastLayout fast_layout(int maxW, int maxH) {
    FastLayout f;
    const int npix = maxW * maxH;
    int p = 0;
    auto take = [&](int bytes) { int r = p; p += (bytes + 15) & ~15; return r; };
    f.tile = take((maxH + 6) * TPX * 2);          // 16-bit pixels
    f.score = take((maxH + 2) * SC_PITCH);        // uint8 scores with a zero ring
    f.queue = take(npix * 2);                     // x | y<<6: front = may be a corner at iniTh, back = only at minTh
    f.alive = take(((maxW + 1) / 2) * ((maxH + 1) / 2) * 2);   // NMS survivors (no two are 8-adjacent)
    f.total = p;
    f.zeroVec = (f.queue - f.score) / 16;
    f.qCap = npix;
    return f;
}

// Per-cell constants of the kernel's two thread mappings (ORBextractor.cc:775-789 gives the cell; this only adds
// where its tile starts in the padded level and how 128 threads are laid over it).
void fast_cell_setup(Cell& c, long long levelPyrOff, int pitch) {
    const int tx0 = c.x0 - 4, ty0 = c.y0 - 3;
    const int mis = (tx0 + kPadLeft) & 3;
    c.pitch = (unsigned short)pitch;
    c.tileOff = (int)(levelPyrOff + (long long)(kEdge + ty0) * pitch + kPadLeft + tx0 - mis);
    c.shift8 = (unsigned char)(mis * 8);
    c.quads = (unsigned char)((c.cw + 8 + 3) >> 2);
    c.groups = (unsigned char)((c.cw + 3) >> 2);
    c.rowsStage = (unsigned char)(FAST_THREADS / c.quads);
    c.rq = (unsigned short)((32768 + c.quads - 1) / c.quads);
    { const unsigned int colItems = 2u * ((c.ch + 1) >> 1); c.rci = ((1u << 20) + colItems - 1) / colItems; }
    c.pad = 0;
    c.pad1 = 0;
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

// Necessary condition per 16-bit half, branch- and predicate-free: with K = 0x0200 - (t+1) in both halves,
// bit 9 of (mm - c + K) is set iff mm >= c + t + 1 (bright arc possible) and bit 9 of (c - nn + K) iff c >= nn + t + 1
// (dark arc possible). All halves hold 8-bit values, so every half stays in [1, 0x2fe]: no borrow crosses the halves and
// each expression is a single three-input add.
constexpr unsigned int PASS_MASK = 0x02000200u;
__device__ __forceinline__ unsigned int pass_word(unsigned int mm, unsigned int nn, unsigned int c, unsigned int K) {
    return ((mm - c + K) | (c - nn + K)) & PASS_MASK;
}

struct FastShared {
    unsigned int* tile;
    unsigned char* score;
    unsigned short* queue;
    unsigned short* surv;
};

// B: exact score of queue entries first, first + step, ... (step = +1 walks the front, -1 the back); corners (S >= minTh)
//    enter the score map
__device__ __forceinline__ void score_queue(const FastShared& S, int first, int step, int n, int tid, int minTh) {
    const unsigned short* tile16 = reinterpret_cast<const unsigned short*>(S.tile);
#pragma unroll 1
    for (int q = tid; q < n; q += FAST_THREADS) {
        const unsigned int e = S.queue[first + q * step];
        const int x = e & 63, y = e >> 6;
        const int s = fast_score(tile16 + (y + 3) * TPX + x + 4);
        if (s >= minTh) S.score[(y + 1) * SC_PITCH + x + 1] = (unsigned char)s;
    }
}

// C: per-cell NMS over the queue entries whose score lies in [lo, hi); survivors are listed
__device__ __forceinline__ void nms_queue(const FastShared& S, int first, int step, int n, int tid, int cw, int lo, int hi,
                                          int* survLen) {
#pragma unroll 1
    for (int q = tid; q < n; q += FAST_THREADS) {
        const unsigned int e = S.queue[first + q * step];
        const int x = e & 63, y = e >> 6;
        const unsigned char* sc = S.score + (y + 1) * SC_PITCH + x + 1;
        const int s = sc[0];
        if (s < lo || s >= hi) continue;
        const int m = max(max(max(sc[-SC_PITCH - 1], sc[-SC_PITCH]), max(sc[-SC_PITCH + 1], sc[-1])),
                          max(max(sc[1], sc[SC_PITCH - 1]), max(sc[SC_PITCH], sc[SC_PITCH + 1])));
        if (s > m) S.surv[atomicAdd(survLen, 1)] = (unsigned short)e;
    }
}

__global__ void __launch_bounds__(FAST_THREADS, 12) fast_cells_kernel(const __grid_constant__ ExtractParams P) {
    extern __shared__ __align__(16) unsigned char fsm[];
    __shared__ unsigned int sQueueLens;   // front length | back length << 16
    __shared__ int sSurvLen;
    const int frame = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const uint4* cellWords = reinterpret_cast<const uint4*>(P.cells + blockIdx.x);
    const uint4 cw0 = __ldg(cellWords), cw1 = __ldg(cellWords + 1);
    FastShared S;
    S.tile = reinterpret_cast<unsigned int*>(fsm + P.fast.tile);
    S.score = fsm + P.fast.score;
    S.queue = reinterpret_cast<unsigned short*>(fsm + P.fast.queue);
    S.surv = reinterpret_cast<unsigned short*>(fsm + P.fast.alive);

    // the score map starts as zero (independent of the cell: overlaps the cell-table load)
    {
        uint4* z = reinterpret_cast<uint4*>(S.score);
#pragma unroll 1
        for (int i = tid; i < P.fast.zeroVec; i += FAST_THREADS) z[i] = make_uint4(0, 0, 0, 0);
        if (tid == 0) { sQueueLens = 0; sSurvLen = 0; }
    }
    // unpack the cell record (layout of struct Cell)
    const int cellX0 = (int)(short)(cw0.x >> 16), cellY0 = (int)(short)(cw0.y & 0xffffu);
    const int cw = (int)(cw0.y >> 16), ch = (int)(cw0.z & 0xffffu), pitch = (int)(cw0.z >> 16);
    const int cellSlot = (int)cw0.w, tileOff = (int)cw1.x;
    const int shift8 = (int)(cw1.y & 0xffu), quads = (int)((cw1.y >> 8) & 0xffu), rowsStage = (int)((cw1.y >> 16) & 0xffu),
              groups = (int)(cw1.y >> 24);
    const unsigned int rq = cw1.z >> 16, rci = cw1.w;

    // ---- 0. stage level pixels [x0-4, x0+cw+4) x [y0-3, y0+ch+3) as 16-bit values; aligned 32-bit global loads,
    //         funnel shift to the tile's alignment, widen, one 64-bit shared store per 4 pixels
    {
        const int r0 = (int)(((unsigned int)tid * rq) >> 15), q = tid - r0 * quads;
        if (r0 < rowsStage) {
            const unsigned int* g =
                reinterpret_cast<const unsigned int*>(P.pyr + (size_t)frame * P.pyrFrameBytes + tileOff + (size_t)r0 * pitch) + q;
            uint2* t = reinterpret_cast<uint2*>(S.tile) + r0 * (TPX / 4) + q;
            const int gStep = rowsStage * (pitch >> 2), tStep = rowsStage * (TPX / 4), rows = ch + 6;
            // four rows per trip, all eight loads issued before the first use (one DRAM/L2 round trip per trip)
#pragma unroll 1
            for (int r = r0; r < rows; r += 4 * rowsStage, g += 4 * gStep, t += 4 * tStep) {
                unsigned int w0[4], w1[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (r + j * rowsStage < rows) { w0[j] = __ldg(g + j * gStep); w1[j] = __ldg(g + j * gStep + 1); }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (r + j * rowsStage < rows) {
                        const unsigned int px = __funnelshift_r(w0[j], w1[j], shift8);
                        t[j * tStep] = make_uint2(__byte_perm(px, 0, 0x4140), __byte_perm(px, 0, 0x4342));
                    }
            }
        }
    }
    __syncthreads();

    const int qLastAll = P.fast.qCap - 1;
    // ---- A. necessary test on every pixel. A work item is a 4-pixel group on TWO vertically adjacent rows (y, y+1): the
    //         two rows share six of the eight tile rows they read, which is what matters here -- the kernel is bound by
    //         shared-memory wavefronts. Items are laid out so that 16 consecutive lanes (8 row pairs x 2 adjacent groups)
    //         hit 16 different 8-byte bank pairs (see TPX). A warp whose items all fail after the middle rows skips the rest.
    {
        const unsigned int Kmin = PASS_MASK - (unsigned int)(P.minTh + 1) * 0x00010001u,
                           Kini = PASS_MASK - (unsigned int)(P.iniTh + 1) * 0x00010001u;
        constexpr int PW = TPX / 4;                       // tile pitch in uint2
        const int colItems = 2 * ((ch + 1) >> 1);         // items per pair of group columns: (row pair, left/right group)
        const int total = colItems * ((groups + 1) >> 1);
        const int qLast = P.fast.qCap - 1;
        // pair maxima / minima of opposing circle points folded into running registers
        auto acc2 = [](unsigned int& mm, unsigned int& nn, unsigned int a1, unsigned int b1, unsigned int a2, unsigned int b2) {
            mm = __vimin3_u16x2(mm, __vmaxu2(a1, b1), __vmaxu2(a2, b2));
            nn = __vimax3_u16x2(nn, __vminu2(a1, b1), __vminu2(a2, b2));
        };
        auto acc1 = [](unsigned int& mm, unsigned int& nn, unsigned int a, unsigned int b) {
            mm = __vminu2(mm, __vmaxu2(a, b));
            nn = __vmaxu2(nn, __vminu2(a, b));
        };
#pragma unroll 1
        for (int i0 = tid - lane; i0 < total; i0 += FAST_THREADS) {
            const int i = min(i0 + lane, total - 1);
            const int cp = (int)(((unsigned int)i * rci) >> 20), rem = i - cp * colItems;   // i * rci < 512 * 2^20
            const int gRaw = 2 * cp + (rem & 1), y = rem & ~1;
            const bool valid = (i0 + lane < total) && gRaw < groups;
            const int g = min(gRaw, groups - 1), x0 = 4 * g;
            // pixels x0, x0+1 answer in bits 9 / 25 of the A word, x0+2, x0+3 in those of the B word
            const unsigned int colA = (x0 < cw ? 0x200u : 0u) | (x0 + 1 < cw ? 0x02000000u : 0u),
                               colB = (x0 + 2 < cw ? 0x200u : 0u) | (x0 + 3 < cw ? 0x02000000u : 0u);
            const uint2* rowC = reinterpret_cast<const uint2*>(S.tile) + (y + 3) * PW + g;   // tile row of pixel row y
            // rowC[0] = pixels x0-4..x0-1, rowC[1] = x0..x0+3, rowC[2] = x0+4..   (A = pixels x0, x0+1; B = x0+2, x0+3)
            // Row r relative to y is "up/mid/down k" for pixel row 0 (= y) and/or pixel row 1 (= y+1).
            unsigned int mm0A, nn0A, mm0B, nn0B, mm1A, nn1A, mm1B, nn1B, c0A, c0B, c1A, c1B;
            unsigned int m1w1, m1cx, m1cy, m1w4;          // row -1 raw words (pixel row 1's "up 2")
            unsigned int p2w1, p2cx, p2cy, p2w4;          // row +2 raw words (pixel row 0's "down 2")
            {
                // rows -1..+2: the +-3 shifted pairs. +3: A' = (x0+3, x0+4), B' = (x0+5, x0+6); -3: A' = (x0-3, x0-2), B' = (x0-1, x0)
                uint2 l = rowC[-PW], c = rowC[-PW + 1], r = rowC[-PW + 2];
                const unsigned int am1PA = __funnelshift_r(c.y, r.x, 16), am1PB = __funnelshift_r(r.x, r.y, 16),
                                   am1MA = __funnelshift_r(l.x, l.y, 16), am1MB = __funnelshift_r(l.y, c.x, 16);
                m1w1 = l.y; m1cx = c.x; m1cy = c.y; m1w4 = r.x;
                l = rowC[0]; c = rowC[1]; r = rowC[2];
                const unsigned int a0PA = __funnelshift_r(c.y, r.x, 16), a0PB = __funnelshift_r(r.x, r.y, 16),
                                   a0MA = __funnelshift_r(l.x, l.y, 16), a0MB = __funnelshift_r(l.y, c.x, 16);
                c0A = c.x; c0B = c.y;
                // pixel row 0: k=4 (3,0) with k=12 (-3,0)
                mm0A = __vmaxu2(a0PA, a0MA); nn0A = __vminu2(a0PA, a0MA);
                mm0B = __vmaxu2(a0PB, a0MB); nn0B = __vminu2(a0PB, a0MB);
                l = rowC[PW]; c = rowC[PW + 1]; r = rowC[PW + 2];
                const unsigned int a1PA = __funnelshift_r(c.y, r.x, 16), a1PB = __funnelshift_r(r.x, r.y, 16),
                                   a1MA = __funnelshift_r(l.x, l.y, 16), a1MB = __funnelshift_r(l.y, c.x, 16);
                c1A = c.x; c1B = c.y;
                // pixel row 0: k=3 (3,1) with k=11 (-3,-1); k=5 (3,-1) with k=13 (-3,1)
                acc2(mm0A, nn0A, a1PA, am1MA, am1PA, a1MA);
                acc2(mm0B, nn0B, a1PB, am1MB, am1PB, a1MB);
                // pixel row 1: k=4 with k=12
                mm1A = __vmaxu2(a1PA, a1MA); nn1A = __vminu2(a1PA, a1MA);
                mm1B = __vmaxu2(a1PB, a1MB); nn1B = __vminu2(a1PB, a1MB);
                l = rowC[2 * PW]; c = rowC[2 * PW + 1]; r = rowC[2 * PW + 2];
                const unsigned int a2PA = __funnelshift_r(c.y, r.x, 16), a2PB = __funnelshift_r(r.x, r.y, 16),
                                   a2MA = __funnelshift_r(l.x, l.y, 16), a2MB = __funnelshift_r(l.y, c.x, 16);
                p2w1 = l.y; p2cx = c.x; p2cy = c.y; p2w4 = r.x;
                // pixel row 1: k=3 with k=11; k=5 with k=13
                acc2(mm1A, nn1A, a2PA, a0MA, a0PA, a2MA);
                acc2(mm1B, nn1B, a2PB, a0MB, a0PB, a2MB);
            }
            if (!__any_sync(0xffffffffu, valid && (pass_word(mm0A, nn0A, c0A, Kmin) | pass_word(mm0B, nn0B, c0B, Kmin) |
                                                   pass_word(mm1A, nn1A, c1A, Kmin) | pass_word(mm1B, nn1B, c1B, Kmin))))
                continue;
            unsigned int u3f12, u3f23, u3f34, u3cx, u3cy;   // row -2 as pixel row 1's "up 3"
            {
                // row -2: pixel row 0's "up 2" (dx = +-2, no shifts) against row +2
                const unsigned int w1 = rowC[-2 * PW].y, w4 = rowC[-2 * PW + 2].x;
                const uint2 c = rowC[-2 * PW + 1];
                acc2(mm0A, nn0A, p2cy, w1, c.y, p2w1);     // k=2 (2,2) with k=10 (-2,-2); k=6 (2,-2) with k=14 (-2,2)
                acc2(mm0B, nn0B, p2w4, c.x, w4, p2cx);
                u3f12 = __funnelshift_r(w1, c.x, 16); u3f23 = __funnelshift_r(c.x, c.y, 16); u3f34 = __funnelshift_r(c.y, w4, 16);
                u3cx = c.x; u3cy = c.y;
            }
            {
                // row +3: pixel row 1's "down 2" against row -1; pixel row 0's "down 3" against row -3
                const unsigned int w1 = rowC[3 * PW].y, w4 = rowC[3 * PW + 2].x;
                const uint2 c = rowC[3 * PW + 1];
                acc2(mm1A, nn1A, c.y, m1w1, m1cy, w1);
                acc2(mm1B, nn1B, w4, m1cx, m1w4, c.x);
                const unsigned int df12 = __funnelshift_r(w1, c.x, 16), df23 = __funnelshift_r(c.x, c.y, 16),
                                   df34 = __funnelshift_r(c.y, w4, 16);
                const unsigned int v1 = rowC[-3 * PW].y, v4 = rowC[-3 * PW + 2].x;
                const uint2 uc = rowC[-3 * PW + 1];
                const unsigned int uf12 = __funnelshift_r(v1, uc.x, 16), uf23 = __funnelshift_r(uc.x, uc.y, 16),
                                   uf34 = __funnelshift_r(uc.y, v4, 16);
                // k=0 (0,3) with k=8 (0,-3); k=1 (1,3) with k=9 (-1,-3); k=7 (1,-3) with k=15 (-1,3)
                acc2(mm0A, nn0A, c.x, uc.x, df23, uf12);
                acc2(mm0B, nn0B, c.y, uc.y, df34, uf23);
                acc1(mm0A, nn0A, uf23, df12);
                acc1(mm0B, nn0B, uf34, df23);
            }
            {
                // row +4: pixel row 1's "down 3" against row -2
                const unsigned int w1 = rowC[4 * PW].y, w4 = rowC[4 * PW + 2].x;
                const uint2 c = rowC[4 * PW + 1];
                const unsigned int df12 = __funnelshift_r(w1, c.x, 16), df23 = __funnelshift_r(c.x, c.y, 16),
                                   df34 = __funnelshift_r(c.y, w4, 16);
                acc2(mm1A, nn1A, c.x, u3cx, df23, u3f12);
                acc2(mm1B, nn1B, c.y, u3cy, df34, u3f23);
                acc1(mm1A, nn1A, u3f23, df12);
                acc1(mm1B, nn1B, u3f34, df23);
            }
            const bool valid1 = valid && y + 1 < ch;
            const unsigned int pass0A = valid ? pass_word(mm0A, nn0A, c0A, Kmin) & colA : 0u,
                               pass0B = valid ? pass_word(mm0B, nn0B, c0B, Kmin) & colB : 0u,
                               pass1A = valid1 ? pass_word(mm1A, nn1A, c1A, Kmin) & colA : 0u,
                               pass1B = valid1 ? pass_word(mm1B, nn1B, c1B, Kmin) & colB : 0u;
            const unsigned int front0A = pass_word(mm0A, nn0A, c0A, Kini) & pass0A, front0B = pass_word(mm0B, nn0B, c0B, Kini) & pass0B,
                               front1A = pass_word(mm1A, nn1A, c1A, Kini) & pass1A, front1B = pass_word(mm1B, nn1B, c1B, Kini) & pass1B;
            const unsigned int back0A = pass0A ^ front0A, back0B = pass0B ^ front0B, back1A = pass1A ^ front1A, back1B = pass1B ^ front1B;
            // queue slots: warp scan of (front count | back count << 16), one shared atomic per warp
            const unsigned int mine = (unsigned int)(__popc(front0A | (front0B << 1)) + __popc(front1A | (front1B << 1))) |
                                      ((unsigned int)(__popc(back0A | (back0B << 1)) + __popc(back1A | (back1B << 1))) << 16);
            unsigned int incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned int up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += up;
            }
            unsigned int start = 0;
            if (lane == 31 && incl) start = atomicAdd(&sQueueLens, incl);
            start = __shfl_sync(0xffffffffu, start, 31) + incl - mine;
            const unsigned int e0 = (unsigned int)x0 | ((unsigned int)y << 6), e1 = e0 + 64u;
            int pf = (int)(start & 0xffffu), pb = qLast - (int)(start >> 16);
            if (front0A & 0x200u) S.queue[pf++] = (unsigned short)e0;
            if (front0A & 0x02000000u) S.queue[pf++] = (unsigned short)(e0 + 1);
            if (front0B & 0x200u) S.queue[pf++] = (unsigned short)(e0 + 2);
            if (front0B & 0x02000000u) S.queue[pf++] = (unsigned short)(e0 + 3);
            if (front1A & 0x200u) S.queue[pf++] = (unsigned short)e1;
            if (front1A & 0x02000000u) S.queue[pf++] = (unsigned short)(e1 + 1);
            if (front1B & 0x200u) S.queue[pf++] = (unsigned short)(e1 + 2);
            if (front1B & 0x02000000u) S.queue[pf] = (unsigned short)(e1 + 3);
            if (back0A & 0x200u) S.queue[pb--] = (unsigned short)e0;
            if (back0A & 0x02000000u) S.queue[pb--] = (unsigned short)(e0 + 1);
            if (back0B & 0x200u) S.queue[pb--] = (unsigned short)(e0 + 2);
            if (back0B & 0x02000000u) S.queue[pb--] = (unsigned short)(e0 + 3);
            if (back1A & 0x200u) S.queue[pb--] = (unsigned short)e1;
            if (back1A & 0x02000000u) S.queue[pb--] = (unsigned short)(e1 + 1);
            if (back1B & 0x200u) S.queue[pb--] = (unsigned short)(e1 + 2);
            if (back1B & 0x02000000u) S.queue[pb] = (unsigned short)(e1 + 3);
        }
    }
    __syncthreads();

    // ---- B, C at iniThFAST
    const int nFront = (int)(sQueueLens & 0xffffu), nBack = (int)(sQueueLens >> 16);
    score_queue(S, 0, 1, nFront, tid, P.minTh);
    __syncthreads();
    nms_queue(S, 0, 1, nFront, tid, cw, P.iniTh, 256, &sSurvLen);
    __syncthreads();
    int sn = sSurvLen;
    if (sn == 0) {
        // ---- B', C': the reference's second cv::FAST call at minThFAST (:811-818). Corners found so far keep their
        //      scores; those >= iniThFAST have just lost their NMS and would lose it again.
        score_queue(S, qLastAll, -1, nBack, tid, P.minTh);
        __syncthreads();
        nms_queue(S, 0, 1, nFront, tid, cw, P.minTh, P.iniTh, &sSurvLen);
        nms_queue(S, qLastAll, -1, nBack, tid, cw, P.minTh, P.iniTh, &sSurvLen);
        __syncthreads();
        sn = sSurvLen;
    }

    // ---- D. survivors store themselves at their row-major rank = number of survivors with a smaller (y, x) key
    //         (there are only ~15 survivors per cell; the list reads are warp-wide broadcasts)
    if (tid == 0) P.cellCount[(size_t)frame * P.nCellsTotal + blockIdx.x] = sn;
    unsigned int* slot = P.slots + (size_t)frame * P.slotFrameEntries + cellSlot;
    for (int q = tid; q < sn; q += FAST_THREADS) {
        const unsigned int e = S.surv[q];
        const int x = e & 63, y = e >> 6;
        int rank = 0;                                   // entries are x | y << 6: numeric order == (y, x) order
        for (int j = 0; j < sn; ++j) rank += S.surv[j] < e;
        const unsigned int s = S.score[(y + 1) * SC_PITCH + x + 1];
        slot[rank] = ((unsigned int)(cellX0 + x - 16) << 20) | ((unsigned int)(cellY0 + y - 16) << 8) | s;
    }
}

int launch_fast(const ExtractParams& P, cudaStream_t st, int* launches) {
    if (P.nCellsTotal == 0) return ORB_OK;
    dim3 grid(P.nCellsTotal, P.nFrames);
    fast_cells_kernel<<<grid, FAST_THREADS, P.fast.total, st>>>(P);
    ++*launches;
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orbb
