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
//   C. NMS on the corners with S >= iniThFAST; survivors set bits in a row-major bitmap;
//   B', C'. only if nothing survived: the back queue is scored too and NMS runs at minThFAST;
//   D. survivors store themselves at their row-major rank (bitmap popcounts).
#include "extractor.h"

namespace orbb {

constexpr int FAST_THREADS = 128;
constexpr int TPX = 76;                           // tile pitch in pixels (16-bit each); interior x=0 sits at column 4.
                                                  // 19 eight-byte words per row: ODD, so the 16 lanes of a half warp that
                                                  // walk down a column of 4-pixel groups hit 16 different bank pairs
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
    f.bitmap = take((((npix + 31) >> 5) + 1) * 4);    // NMS survivors, bit = y*cw + x
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
    c.rch = ((1u << 20) + c.ch - 1) / c.ch;
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
    unsigned int* bitmap;
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

// C: per-cell NMS over the queue entries whose score lies in [lo, hi); survivors are listed and set their bitmap bit
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
        if (s > m) {
            const int bit = y * cw + x;
            atomicOr(&S.bitmap[bit >> 5], 1u << (bit & 31));
            S.surv[atomicAdd(survLen, 1)] = (unsigned short)e;
        }
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
    S.bitmap = reinterpret_cast<unsigned int*>(fsm + P.fast.bitmap);
    S.queue = reinterpret_cast<unsigned short*>(fsm + P.fast.queue);
    S.surv = reinterpret_cast<unsigned short*>(fsm + P.fast.alive);

    // score map and bitmap start as zero (independent of the cell: overlaps the cell-table load)
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
    const unsigned int rq = cw1.z >> 16, rch = cw1.w;

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
    // ---- A. necessary test, 4 pixels (two 16x2 pairs) per thread; warp-uniform loop so that a warp whose 128 pixels
    //         all fail after the three middle-row pairs skips the other five
    {
        const unsigned int Kmin = PASS_MASK - (unsigned int)(P.minTh + 1) * 0x00010001u,
                           Kini = PASS_MASK - (unsigned int)(P.iniTh + 1) * 0x00010001u;
        // work item i = g * ch + y: the lanes of a warp walk DOWN a column of 4-pixel groups (conflict-free shared loads,
        // see TPX); i / ch by multiply-shift (exact for i < 2^20 / ch)
        const int total = groups * ch;
        const int qLast = P.fast.qCap - 1;
#pragma unroll 1
        for (int i0 = tid - lane; i0 < total; i0 += FAST_THREADS) {
            const int i = min(i0 + lane, total - 1);
            const bool valid = i0 + lane < total;
            const int g = (int)(((unsigned int)i * rch) >> 20), y = i - g * ch;   // i * rch < 900 * 2^20
            const int x0 = 4 * g;
            // pixels x0, x0+1 answer in bits 9 / 25 of the A word, x0+2, x0+3 in those of the B word
            const unsigned int colA = (x0 < cw ? 0x200u : 0u) | (x0 + 1 < cw ? 0x02000000u : 0u),
                               colB = (x0 + 2 < cw ? 0x200u : 0u) | (x0 + 3 < cw ? 0x02000000u : 0u);
            const uint2* rowC = reinterpret_cast<const uint2*>(S.tile) + (y + 3) * (TPX / 4) + g;
            // rowC[0] = pixels x0-4..x0-1, rowC[1] = x0..x0+3, rowC[2] = x0+4..
            // A = pixels (x0, x0+1), B = (x0+2, x0+3). mm = running minimum of the pair maxima, nn = running maximum of
            // the pair minima of opposing circle points (folded in as they are formed: few live registers)
            unsigned int mmA, mmB, nnA, nnB, cA, cB;
            {   // rows +-1 and 0: dx = +-3 -> circle points 3/11, 5/13, 4/12
                const uint2* up = rowC - (TPX / 4);
                const uint2* dn = rowC + (TPX / 4);
                const uint2 ul = up[0], uc = up[1], ur = up[2], dl = dn[0], dc = dn[1], dr = dn[2];
                const uint2 ml = rowC[0], mc = rowC[1], mr = rowC[2];
                cA = mc.x; cB = mc.y;
                // +3: A' = (x0+3, x0+4), B' = (x0+5, x0+6);  -3: A' = (x0-3, x0-2), B' = (x0-1, x0)
                const unsigned int uPA = __funnelshift_r(uc.y, ur.x, 16), uPB = __funnelshift_r(ur.x, ur.y, 16);
                const unsigned int uMA = __funnelshift_r(ul.x, ul.y, 16), uMB = __funnelshift_r(ul.y, uc.x, 16);
                const unsigned int dPA = __funnelshift_r(dc.y, dr.x, 16), dPB = __funnelshift_r(dr.x, dr.y, 16);
                const unsigned int dMA = __funnelshift_r(dl.x, dl.y, 16), dMB = __funnelshift_r(dl.y, dc.x, 16);
                const unsigned int mPA = __funnelshift_r(mc.y, mr.x, 16), mPB = __funnelshift_r(mr.x, mr.y, 16);
                const unsigned int mMA = __funnelshift_r(ml.x, ml.y, 16), mMB = __funnelshift_r(ml.y, mc.x, 16);
                // k=3 (3,1) with k=11 (-3,-1); k=5 (3,-1) with k=13 (-3,1); k=4 (3,0) with k=12 (-3,0)
                mmA = __vimin3_u16x2(__vmaxu2(dPA, uMA), __vmaxu2(uPA, dMA), __vmaxu2(mPA, mMA));
                nnA = __vimax3_u16x2(__vminu2(dPA, uMA), __vminu2(uPA, dMA), __vminu2(mPA, mMA));
                mmB = __vimin3_u16x2(__vmaxu2(dPB, uMB), __vmaxu2(uPB, dMB), __vmaxu2(mPB, mMB));
                nnB = __vimax3_u16x2(__vminu2(dPB, uMB), __vminu2(uPB, dMB), __vminu2(mPB, mMB));
            }
            if (!__any_sync(0xffffffffu, valid && (pass_word(mmA, nnA, cA, Kmin) | pass_word(mmB, nnB, cB, Kmin)))) continue;
            {   // rows +-3: dx = 0, +1, -1  -> circle points 0/8, 1/9, 15/7
                const uint2* up = rowC - 3 * (TPX / 4);
                const uint2* dn = rowC + 3 * (TPX / 4);
                const unsigned int u1 = up[0].y, u4 = up[2].x, d1 = dn[0].y, d4 = dn[2].x;
                const uint2 uc = up[1], dc = dn[1];
                const unsigned int uf12 = __funnelshift_r(u1, uc.x, 16), uf23 = __funnelshift_r(uc.x, uc.y, 16),
                                   uf34 = __funnelshift_r(uc.y, u4, 16);
                const unsigned int df12 = __funnelshift_r(d1, dc.x, 16), df23 = __funnelshift_r(dc.x, dc.y, 16),
                                   df34 = __funnelshift_r(dc.y, d4, 16);
                // k=0 (0,3) with k=8 (0,-3); k=1 (1,3) with k=9 (-1,-3)
                mmA = __vimin3_u16x2(mmA, __vmaxu2(dc.x, uc.x), __vmaxu2(df23, uf12));
                nnA = __vimax3_u16x2(nnA, __vminu2(dc.x, uc.x), __vminu2(df23, uf12));
                mmB = __vimin3_u16x2(mmB, __vmaxu2(dc.y, uc.y), __vmaxu2(df34, uf23));
                nnB = __vimax3_u16x2(nnB, __vminu2(dc.y, uc.y), __vminu2(df34, uf23));
                // k=7 (1,-3) with k=15 (-1,3) is folded in below with k=2/10
                const unsigned int pA7 = __vmaxu2(uf23, df12), qA7 = __vminu2(uf23, df12);
                const unsigned int pB7 = __vmaxu2(uf34, df23), qB7 = __vminu2(uf34, df23);
                // rows +-2: dx = +-2 -> circle points 2/10, 6/14; no shifts
                const uint2* up2 = rowC - 2 * (TPX / 4);
                const uint2* dn2 = rowC + 2 * (TPX / 4);
                const unsigned int v1 = up2[0].y, v4 = up2[2].x, e1 = dn2[0].y, e4 = dn2[2].x;
                const uint2 vc = up2[1], ec = dn2[1];
                // k=2 (2,2) with k=10 (-2,-2)
                mmA = __vimin3_u16x2(mmA, pA7, __vmaxu2(ec.y, v1));
                nnA = __vimax3_u16x2(nnA, qA7, __vminu2(ec.y, v1));
                mmB = __vimin3_u16x2(mmB, pB7, __vmaxu2(e4, vc.x));
                nnB = __vimax3_u16x2(nnB, qB7, __vminu2(e4, vc.x));
                // k=6 (2,-2) with k=14 (-2,2)
                mmA = __vminu2(mmA, __vmaxu2(vc.y, e1));
                nnA = __vmaxu2(nnA, __vminu2(vc.y, e1));
                mmB = __vminu2(mmB, __vmaxu2(v4, ec.x));
                nnB = __vmaxu2(nnB, __vminu2(v4, ec.x));
            }
            const unsigned int passA = valid ? pass_word(mmA, nnA, cA, Kmin) & colA : 0u,
                               passB = valid ? pass_word(mmB, nnB, cB, Kmin) & colB : 0u;
            const unsigned int frontA = pass_word(mmA, nnA, cA, Kini) & passA, frontB = pass_word(mmB, nnB, cB, Kini) & passB;
            const unsigned int backA = passA ^ frontA, backB = passB ^ frontB;
            // queue slots: warp scan of (front count | back count << 16), one shared atomic per warp
            const unsigned int mine = (unsigned int)(__popc(frontA) + __popc(frontB)) | ((unsigned int)(__popc(backA) + __popc(backB)) << 16);
            unsigned int incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned int up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += up;
            }
            unsigned int start = 0;
            if (lane == 31 && incl) start = atomicAdd(&sQueueLens, incl);
            start = __shfl_sync(0xffffffffu, start, 31) + incl - mine;
            const unsigned int e = (unsigned int)x0 | ((unsigned int)y << 6);
            int pf = (int)(start & 0xffffu), pb = qLast - (int)(start >> 16);
            if (frontA & 0x200u) S.queue[pf++] = (unsigned short)e;
            if (frontA & 0x02000000u) S.queue[pf++] = (unsigned short)(e + 1);
            if (frontB & 0x200u) S.queue[pf++] = (unsigned short)(e + 2);
            if (frontB & 0x02000000u) S.queue[pf] = (unsigned short)(e + 3);
            if (backA & 0x200u) S.queue[pb--] = (unsigned short)e;
            if (backA & 0x02000000u) S.queue[pb--] = (unsigned short)(e + 1);
            if (backB & 0x200u) S.queue[pb--] = (unsigned short)(e + 2);
            if (backB & 0x02000000u) S.queue[pb] = (unsigned short)(e + 3);
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

    // ---- D. survivors store themselves at their row-major rank = number of set bits below their own in the bitmap
    //         (a few dozen POPCs each; there are only ~15 survivors per cell)
    if (tid == 0) P.cellCount[(size_t)frame * P.nCellsTotal + blockIdx.x] = sn;
    unsigned int* slot = P.slots + (size_t)frame * P.slotFrameEntries + cellSlot;
    for (int q = tid; q < sn; q += FAST_THREADS) {
        const unsigned int e = S.surv[q];
        const int x = e & 63, y = e >> 6;
        const int bit = y * cw + x;
        int rank = __popc(S.bitmap[bit >> 5] & ((1u << (bit & 31)) - 1u));
        for (int w = 0; w < (bit >> 5); ++w) rank += __popc(S.bitmap[w]);
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
