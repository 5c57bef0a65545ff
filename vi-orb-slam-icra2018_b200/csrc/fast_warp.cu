// Per-cell FAST-9/16 detection (north-star kernel 2): the cell loop of ORBextractor::ComputeKeyPointsOctTree,
// ORBextractor.cc:767-831, with cv::FAST(roi, th, nonmaxSuppression=true) inside.
//
// Reference semantics (unchanged from round 1, see the parity tests):
//   * keypoints can only lie in the cell's interior = the 3-px inset of the (wCell+6)x(hCell+6) ROI;
//   * score S = cornerScore<16> = (max over the 16 contiguous 9-arcs of min |p_k - v| on one side) - 1, and a pixel is
//     a corner at threshold t  <=>  S >= t;
//   * non-max suppression is per cell: keep iff S > all 8 neighbours, non-corners and pixels outside the interior
//     count as 0;
//   * the cell is re-run at minThFAST only if NOTHING survives at iniThFAST (:811-818);
//   * emission order inside a cell is row-major (y, x).
//
// Mapping (round 2): ONE WARP PER CELL, persistent warps, no block-wide barrier anywhere.
//   The round-1 kernel gave every cell a 128-thread CTA: ~7 pixels per thread, so the per-thread prologue/epilogue
//   (cell record, staging loop set-up, five __syncthreads phases) cost more instructions than the pixels did, and the
//   barrier was the top stall.  Here a warp walks a strided list of (frame, cell) items; every phase is warp-synchronous:
//     0. the cell's tile (interior + 3-px ring) is fetched by ONE TMA tensor load (cp.async.bulk.tensor.3d, box
//        bw x bh x 1 out of [frame][row][pitch]) into the warp's own shared-memory buffer; the load of the NEXT cell is
//        issued before the current one is processed (two buffers, two mbarriers), so global latency is hidden without
//        holding registers.  A TMA box must start on a 16-byte boundary of global memory (measured: an unaligned
//        innermost coordinate raises "illegal instruction", tools/microbench/tma_probe.cu), so the box starts at the
//        aligned column below x0-4 and the cell carries its misalignment `mis` (0..15); the tile pitch is 80 or 96
//        bytes (20 / 24 words), for which rows two apart fall into disjoint bank octets;
//     A. pre-test on every pixel, 4 pixels x 1 row per step in packed 16x2 arithmetic: a 9-arc contains one pixel of
//        each opposing circle pair, so min(max(N,S), max(E,W)) > v+t  (bright) or  max(min(N,S), min(E,W)) < v-t (dark)
//        is necessary.  8 aligned word loads, 5 PRMT that undo the misalignment (selectors uniform per cell), 10 PRMT
//        that widen, 12 VIMNMX, 8 threshold ops per 4 pixels; the results of 8 rows x 4 pixels accumulate in one
//        32-bit mask per polarity (no per-row bookkeeping);
//     B. the set bits become queue entries (x | y<<6 | polarity<<12), one ballot per round so that consecutive entries
//        come from different lanes (different banks);
//     C. exact score of the queued pixels, TWO per lane: the pre-test told the polarity, so entry A rides in the low
//        and entry B in the high 16-bit half of every operand (d = 256 +- (p - v) by one IMAD each), and the 40
//        three-input min/max of the 9-arc scan serve both.  Corners (S >= t) go to the score map and are compacted in
//        place to the front of the queue;
//     D. NMS over the corner list, survivors listed, ranked by (y, x), packed x:12|y:12|score:8 into the cell's slot.
//   A cell with no survivor at iniThFAST simply runs A-D again at minThFAST (the reference's second cv::FAST call).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "extractor.h"

namespace orbb {

constexpr int FW_WARPS = 4;                       // warps per CTA; they share nothing but the 128-byte bit->pixel table
constexpr unsigned int FW_FULL = 0xffffffffu;
constexpr unsigned int FW_PASS = 0x02000200u;     // bit 9 of each 16-bit half

// ---- mbarrier / TMA (PTX ISA 8.x, sm_90+; SASS: SYNCS.*, UTMALDG) ----------------------------------------------------
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned int bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned int bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned int bar, unsigned int parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "FW_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra FW_DONE_%=;\n"
        "bra FW_WAIT_%=;\n"
        "FW_DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_tile(unsigned int dst, const void* map, int c0, int c1, int c2, unsigned int bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}

// ---- A: one pixel row of one 4-pixel group --------------------------------------------------------------------------
// r points at the aligned tile word that holds pixel (x0-4, y); the pixel itself is byte s = mis & 3 of it.  The five
// extractions (pixels x0-3.., x0.., x0+3.. of row y, x0.. of rows y-3 and y+3) are PRMTs over two adjacent words with
// per-cell selectors: selC = 0x3210 + 0x1111*s, selL = selC + 0x1111, and x0+3.. lies in words (1,2) for s < 2 and in
// words (2,3) otherwise (hiR).  bright / dark get bits 9, 25 (pixels x0, x0+1) and 10, 26 (x0+2, x0+3).
// K = 0x0200 - (t+1) in both halves: bit 9 of (m - c + K) is set iff m >= c + t + 1; all halves stay in [1, 0x2fe], so
// no borrow crosses the halves and each expression is one three-input add.
struct FwAlign {
    unsigned int selC, selL, selR;
    bool hiR;
};
template <int BW>
__device__ __forceinline__ void pretest_row(const unsigned char* r, const FwAlign& al, unsigned int K, unsigned int& bright,
                                            unsigned int& dark) {
    const unsigned int* w = reinterpret_cast<const unsigned int*>(r);
    const unsigned int* wu = reinterpret_cast<const unsigned int*>(r - 3 * BW);
    const unsigned int* wd = reinterpret_cast<const unsigned int*>(r + 3 * BW);
    const unsigned int A0 = w[0], A1 = w[1], A2 = w[2], A3 = w[3];
    const unsigned int U1 = wu[1], U2 = wu[2], D1 = wd[1], D2 = wd[2];
    const unsigned int W1 = __byte_perm(A1, A2, al.selC);                                  // pixels x0 .. x0+3
    const unsigned int FL = __byte_perm(A0, A1, al.selL);                                  // pixels x0-3 .. x0
    const unsigned int FR = __byte_perm(al.hiR ? A2 : A1, al.hiR ? A3 : A2, al.selR);      // pixels x0+3 .. x0+6
    const unsigned int U = __byte_perm(U1, U2, al.selC), D = __byte_perm(D1, D2, al.selC); // rows y-3, y+3
    const unsigned int cA = __byte_perm(W1, 0, 0x4140), cB = __byte_perm(W1, 0, 0x4342);
    const unsigned int uA = __byte_perm(U, 0, 0x4140), uB = __byte_perm(U, 0, 0x4342);
    const unsigned int dA = __byte_perm(D, 0, 0x4140), dB = __byte_perm(D, 0, 0x4342);
    const unsigned int lA = __byte_perm(FL, 0, 0x4140), lB = __byte_perm(FL, 0, 0x4342);
    const unsigned int rA = __byte_perm(FR, 0, 0x4140), rB = __byte_perm(FR, 0, 0x4342);
    const unsigned int mmA = __vminu2(__vmaxu2(uA, dA), __vmaxu2(lA, rA)), mmB = __vminu2(__vmaxu2(uB, dB), __vmaxu2(lB, rB));
    const unsigned int nnA = __vmaxu2(__vminu2(uA, dA), __vminu2(lA, rA)), nnB = __vmaxu2(__vminu2(uB, dB), __vminu2(lB, rB));
    const unsigned int bA = (mmA - cA + K) & FW_PASS, bB = (mmB - cB + K) & FW_PASS;
    const unsigned int kA = (cA - nnA + K) & FW_PASS, kB = (cB - nnB + K) & FW_PASS;
    bright = bA + 2u * bB;
    dark = kA + 2u * kB;
}

// mask bits of the first n (<= 8) rows of an item: row j owns bits 9+2j, 10+2j, 25+2j, 26+2j (mod 32)
__device__ __forceinline__ unsigned int rows_mask(int n) {
    const unsigned int h = (1u << (2 * n)) - 1u;
    return __funnelshift_l(h * 0x00010001u, h * 0x00010001u, 9);
}

// B: append the set bits of m to the queue; one ballot per round, so a round's entries come from distinct lanes
__device__ __forceinline__ int enqueue_bits(unsigned int m, unsigned int baseEntry, const unsigned short* lut, unsigned short* queue,
                                            int n, unsigned int ltMask) {
    unsigned int bal;
    while ((bal = __ballot_sync(FW_FULL, m != 0u)) != 0u) {
        if (m) {
            const int k = 31 - __clz(m);
            m ^= 1u << k;
            queue[n + __popc(bal & ltMask)] = (unsigned short)(baseEntry + lut[k]);
        }
        n += __popc(bal);
    }
    return n;
}

// circle offsets in tile bytes, OpenCV order (dx,dy) = (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)
// (-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)
#define FW_CIRCLE(k, P)                                                                                    \
    ((k) == 0 ? 3 * (P) : (k) == 1 ? 3 * (P) + 1 : (k) == 2 ? 2 * (P) + 2 : (k) == 3 ? (P) + 3             \
     : (k) == 4 ? 3 : (k) == 5 ? -(P) + 3 : (k) == 6 ? -2 * (P) + 2 : (k) == 7 ? -3 * (P) + 1              \
     : (k) == 8 ? -3 * (P) : (k) == 9 ? -3 * (P)-1 : (k) == 10 ? -2 * (P)-2 : (k) == 11 ? -(P)-3           \
     : (k) == 12 ? -3 : (k) == 13 ? (P)-3 : (k) == 14 ? 2 * (P)-2 : 3 * (P)-1)

struct FwWarp {            // the warp's shared-memory arrays
    const unsigned char* tile;
    unsigned char* score;
    unsigned short* queue;
    unsigned short* surv;
};

// A-D for one cell at threshold th; returns the number of NMS survivors (listed in W.surv)
template <int BW>
__device__ __forceinline__ int fast_cell(const FwWarp& W, const FastWarpPlan::Level& F, int scorePitch, const unsigned short* lut,
                                         int cw, int ch, int mis, int th, int lane, unsigned int ltMask) {
    // ---- A + B
    const unsigned int K = FW_PASS - (unsigned int)(th + 1) * 0x00010001u;
    FwAlign al;
    {
        const unsigned int sh = (unsigned int)(mis & 3);
        al.selC = 0x3210u + 0x1111u * sh;
        al.selL = al.selC + 0x1111u;
        al.hiR = sh >= 2u;
        al.selR = 0x3210u + 0x1111u * (al.hiR ? sh - 1u : sh + 3u);
    }
    const unsigned char* tileAligned = W.tile + (mis & ~3);      // word that holds pixel (-4, -3)
    const unsigned char* tilePix = W.tile + mis + 3 * BW + 4;    // pixel (0, 0)
    const int items = F.groups * F.bands;
    const int pairStep = 2 * F.bands * BW;
    int nq = 0;
#pragma unroll 1
    for (int i0 = 0; i0 < items; i0 += 32) {
        const int i = min(i0 + lane, items - 1);
        const int b = (int)(((unsigned int)i * F.rcpGroups) >> 16), g = i - b * F.groups;
        // valid rows are a prefix of the item's rows y(j) = 2b + 2*bands*(j>>1) + (j&1), valid pixels a prefix of its 4
        const int rem = ch - 2 * b;
        int nRows = 0;
        if (rem >= 2) nRows = 2 * ((int)(((unsigned int)(rem - 2) * F.rcpBandStep) >> 16) + 1);
        if (rem >= 1) {
            const int q = (int)(((unsigned int)(rem - 1) * F.rcpBandStep) >> 16);
            if (q * 2 * F.bands == rem - 1) nRows += 1;
        }
        nRows = min(nRows, 2 * F.halfRows);
        const int nCols = cw - 4 * g;
        unsigned int colMask = nCols >= 4 ? 0xffffffffu : nCols == 3 ? 0xabfffeaau : nCols == 2 ? 0xaaaaaaaau : nCols == 1 ? 0x00aaaa00u : 0u;
        if (i0 + lane >= items) colMask = 0u;
        const unsigned char* r = tileAligned + (2 * b + 3) * BW + 4 * g;
        unsigned int mb0 = 0, md0 = 0, mb1 = 0, md1 = 0;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            if (jj < F.halfRows) {
                unsigned int tb, td;
                pretest_row<BW>(r + jj * pairStep, al, K, tb, td);
                mb0 |= __funnelshift_l(tb, tb, 4 * jj);
                md0 |= __funnelshift_l(td, td, 4 * jj);
                pretest_row<BW>(r + jj * pairStep + BW, al, K, tb, td);
                mb0 |= __funnelshift_l(tb, tb, 4 * jj + 2);
                md0 |= __funnelshift_l(td, td, 4 * jj + 2);
            }
        }
#pragma unroll 1
        for (int jj = 4; jj < F.halfRows; ++jj) {
            unsigned int tb, td;
            pretest_row<BW>(r + jj * pairStep, al, K, tb, td);
            mb1 |= __funnelshift_l(tb, tb, 4 * (jj - 4));
            md1 |= __funnelshift_l(td, td, 4 * (jj - 4));
            pretest_row<BW>(r + jj * pairStep + BW, al, K, tb, td);
            mb1 |= __funnelshift_l(tb, tb, 4 * (jj - 4) + 2);
            md1 |= __funnelshift_l(td, td, 4 * (jj - 4) + 2);
        }
        const unsigned int v0 = colMask & rows_mask(min(nRows, 8)), v1 = colMask & rows_mask(max(nRows - 8, 0));
        const unsigned int baseEntry = (unsigned int)(4 * g) | ((unsigned int)(2 * b) << 6);
        nq = enqueue_bits(mb0 & v0, baseEntry, lut, W.queue, nq, ltMask);
        nq = enqueue_bits(md0 & v0, baseEntry + 0x1000u, lut, W.queue, nq, ltMask);
        if (F.halfRows > 4) {
            nq = enqueue_bits(mb1 & v1, baseEntry, lut + 32, W.queue, nq, ltMask);
            nq = enqueue_bits(md1 & v1, baseEntry + 0x1000u, lut + 32, W.queue, nq, ltMask);
        }
    }
    __syncwarp();

    // ---- C: exact score, two queue entries per lane (A in the low halves, B in the high halves)
    int nc = 0;
#pragma unroll 1
    for (int q0 = 0; q0 < nq; q0 += 64) {
        const int iA = q0 + lane, iB = iA + 32;
        const bool vA = iA < nq, vB = iB < nq;
        const unsigned int eA = vA ? W.queue[iA] : 0u, eB = vB ? W.queue[iB] : 0u;
        const int xA = eA & 63, yA = (eA >> 6) & 63, xB = eB & 63, yB = (eB >> 6) & 63;
        const unsigned char* pA = tilePix + yA * BW + xA;
        const unsigned char* pB = tilePix + yB * BW + xB;
        // d = 256 + (p - v) for a bright candidate, 256 + (v - p) for a dark one: every half stays in [1, 511]
        const int mulA = (eA & 0x1000u) ? -1 : 1, mulB = (eB & 0x1000u) ? -65536 : 65536;
        const unsigned int C = 0x01000100u - (unsigned int)mulA * pA[0] - (unsigned int)mulB * pB[0];
        unsigned int d[16];
#pragma unroll
        for (int k = 0; k < 16; ++k)
            d[k] = (unsigned int)pA[FW_CIRCLE(k, BW)] * (unsigned int)mulA + ((unsigned int)pB[FW_CIRCLE(k, BW)] * (unsigned int)mulB + C);
        unsigned int m3[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) m3[k] = __vimin3_u16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
        unsigned int best = 0u;
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
            const unsigned int a = __vimin3_u16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
            const unsigned int b = __vimin3_u16x2(m3[k + 1], m3[(k + 4) & 15], m3[(k + 7) & 15]);
            best = __vimax3_u16x2(best, a, b);
        }
        const int sA = (int)(best & 0xffffu) - 257, sB = (int)(best >> 16) - 257;
        const bool cA = vA && sA >= th, cB = vB && sB >= th;
        if (cA) W.score[(yA + 1) * scorePitch + xA + 1] = (unsigned char)sA;
        if (cB) W.score[(yB + 1) * scorePitch + xB + 1] = (unsigned char)sB;
        // corners move to the front of the queue (never past the entries already read: nc <= q0 + 64)
        const unsigned int balA = __ballot_sync(FW_FULL, cA), balB = __ballot_sync(FW_FULL, cB);
        if (cA) W.queue[nc + __popc(balA & ltMask)] = (unsigned short)(eA & 0xfffu);
        nc += __popc(balA);
        if (cB) W.queue[nc + __popc(balB & ltMask)] = (unsigned short)(eB & 0xfffu);
        nc += __popc(balB);
        __syncwarp();
    }

    // ---- D: non-max suppression over the corners
    int sn = 0;
#pragma unroll 1
    for (int q0 = 0; q0 < nc; q0 += 32) {
        const bool v = q0 + lane < nc;
        const unsigned int e = v ? W.queue[q0 + lane] : 0u;
        const int x = e & 63, y = e >> 6;
        const unsigned char* sc = W.score + (y + 1) * scorePitch + x + 1;
        const int SP = scorePitch;
        const int s = sc[0];
        const int m = max(max(max(sc[-SP - 1], sc[-SP]), max(sc[-SP + 1], sc[-1])), max(max(sc[1], sc[SP - 1]), max(sc[SP], sc[SP + 1])));
        const bool keep = v && s > m;
        const unsigned int bal = __ballot_sync(FW_FULL, keep);
        if (keep) W.surv[sn + __popc(bal & ltMask)] = (unsigned short)e;
        sn += __popc(bal);
    }
    __syncwarp();
    return sn;
}

template <int BW>
__global__ void __launch_bounds__(FW_WARPS * 32, 4) fast_warp_kernel(const __grid_constant__ ExtractParams P) {
    extern __shared__ __align__(128) unsigned char fsm[];
    __shared__ unsigned short sLut[kMaxLevels][64];   // per level: mask bit -> px | row offset << 6 (second half: rows 8..15)
    const FastWarpPlan& F = P.fw;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int ltMask = (1u << lane) - 1u;
    unsigned char* wb = fsm + (size_t)warp * F.warpBytes;
    const unsigned int bar = smem_u32(wb);   // two mbarriers
    FwWarp W;
    W.score = wb + F.offScore;
    W.queue = reinterpret_cast<unsigned short*>(wb + F.offQueue);
    W.surv = reinterpret_cast<unsigned short*>(wb + F.offSurv);

    for (int t = threadIdx.x; t < 64 * P.nLevels; t += FW_WARPS * 32) {
        const int k = t & 31, kk = (k - 9) & 31;
        const int j = ((kk >> 1) & 7) + 8 * ((t >> 5) & 1), px = ((kk & 1) << 1) | (kk >> 4);
        sLut[t >> 6][t & 63] = (unsigned short)(px | ((2 * F.lv[t >> 6].bands * (j >> 1) + (j & 1)) << 6));
    }
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const unsigned int nCells = (unsigned int)P.nCellsTotal;
    const unsigned int total = (unsigned int)P.nFrames * nCells;
    const unsigned int stride = gridDim.x * FW_WARPS;
    unsigned int it = blockIdx.x * FW_WARPS + warp;
    if (it >= total) return;
    const unsigned int strideFrames = stride / nCells, strideCells = stride - strideFrames * nCells;
    unsigned int frame = it / nCells, cell = it - frame * nCells;
    const unsigned char* maps = static_cast<const unsigned char*>(F.maps);

    uint4 c0 = __ldg(reinterpret_cast<const uint4*>(P.cells + cell));
    if (lane == 0) {
        mbar_expect_tx(bar, (unsigned int)F.tileBytes);
        tma_load_tile(smem_u32(wb + F.offTile), maps + 128 * (int)(short)(c0.x & 0xffffu),
                      (kPadLeft + (int)(short)(c0.x >> 16) - 4) & ~15, kEdge + (int)(short)(c0.y & 0xffffu) - 3,
                      F.frameBase + (int)frame, bar);
    }
    unsigned int parity = 0;   // bit b: the phase of buffer b's barrier to wait for
    int buf = 0;
    for (;;) {
        // ---- next item: its tile goes into the other buffer (all reads of that buffer ended with the previous cell)
        unsigned int nFrame = frame + strideFrames, nCell = cell + strideCells;
        if (nCell >= nCells) { nCell -= nCells; ++nFrame; }
        const bool more = it + stride < total && it + stride > it;
        uint4 n0 = c0;
        if (more) {
            n0 = __ldg(reinterpret_cast<const uint4*>(P.cells + nCell));
            if (lane == 0) {
                const unsigned int nb = bar + 8u * (unsigned int)(buf ^ 1);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(nb, (unsigned int)F.tileBytes);
                tma_load_tile(smem_u32(wb + F.offTile + (buf ^ 1) * F.tileStride), maps + 128 * (int)(short)(n0.x & 0xffffu),
                              (kPadLeft + (int)(short)(n0.x >> 16) - 4) & ~15, kEdge + (int)(short)(n0.y & 0xffffu) - 3,
                              F.frameBase + (int)nFrame, nb);
            }
        }
        // ---- this cell
        const int cellX0 = (int)(short)(c0.x >> 16), cellY0 = (int)(short)(c0.y & 0xffffu);
        const int cw = (int)(c0.y >> 16), ch = (int)(c0.z & 0xffffu), cellSlot = (int)c0.w;
        {
            uint4* z = reinterpret_cast<uint4*>(W.score);
#pragma unroll 1
            for (int i = lane; i < F.scoreVec; i += 32) z[i] = make_uint4(0, 0, 0, 0);
        }
        __syncwarp();
        mbar_wait(bar + 8u * (unsigned int)buf, (parity >> buf) & 1u);
        parity ^= 1u << buf;
        W.tile = wb + F.offTile + buf * F.tileStride;

        const int level = (int)(short)(c0.x & 0xffffu);
        int sn, th = P.iniTh;
        for (;;) {   // the second round is the reference's second cv::FAST call at minThFAST (:811-818)
            sn = fast_cell<BW>(W, F.lv[level], F.scorePitch, sLut[level], cw, ch, (kPadLeft + cellX0 - 4) & 15, th, lane, ltMask);
            if (sn != 0 || th <= P.minTh) break;
            th = P.minTh;
        }

        // ---- survivors store themselves at their row-major rank = number of survivors with a smaller (y, x) key
        if (lane == 0) P.cellCount[(size_t)frame * P.nCellsTotal + cell] = sn;
        unsigned int* slot = P.slots + (size_t)frame * P.slotFrameEntries + cellSlot;
#pragma unroll 1
        for (int q = lane; q < sn; q += 32) {
            const unsigned int e = W.surv[q];
            const int x = e & 63, y = e >> 6;
            int rank = 0;                                   // entries are x | y << 6: numeric order == (y, x) order
            for (int j = 0; j < sn; ++j) rank += W.surv[j] < e;
            const unsigned int s = W.score[(y + 1) * F.scorePitch + x + 1];
            slot[rank] = ((unsigned int)(cellX0 + x - 16) << 20) | ((unsigned int)(cellY0 + y - 16) << 8) | s;
        }
        __syncwarp();
        if (!more) break;
        it += stride;
        frame = nFrame;
        cell = nCell;
        c0 = n0;
        buf ^= 1;
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------

int fast_warp_plan(int nLevels, const int* cellW, const int* cellH, int slotCapMax, FastWarpPlan* plan) {
    FastWarpPlan f;
    std::memset(&f, 0, sizeof f);
    int maxCellW = 1, maxCellH = 1, maxRows = 0;
    for (int l = 0; l < nLevels; ++l) {
        FastWarpPlan::Level& L = f.lv[l];
        const int w = std::max(cellW[l], 1), h = std::max(cellH[l], 1);
        maxCellW = std::max(maxCellW, w);
        maxCellH = std::max(maxCellH, h);
        L.groups = (unsigned short)((w + 3) / 4);
        // bands x rows per item: as few warp passes and masked rows as possible (31x31 cells: 8 groups x 4 bands x 8 rows)
        long long bestCost = -1;
        for (int bands = 1; bands <= 32; ++bands) {
            const int halfRows = (h + 2 * bands - 1) / (2 * bands);
            if (halfRows > 8) continue;                   // two 32-bit masks hold 16 rows
            if (2 * bands * halfRows > 64) continue;      // y < 64 in a queue entry
            const int items = L.groups * bands, passes = (items + 31) / 32;
            const long long cost = (long long)passes * (2 * halfRows * 10 + 12);
            if (bestCost < 0 || cost < bestCost) { bestCost = cost; L.bands = (unsigned short)bands; L.halfRows = (unsigned short)halfRows; }
        }
        if (bestCost < 0) return fail(ORB_ERR_INVALID, "FAST: no work-item shape for %dx%d cells", w, h);
        L.rcpGroups = (65536u + L.groups - 1) / L.groups;
        L.rcpBandStep = (65536u + 2 * L.bands - 1) / (2 * L.bands);
        maxRows = std::max(maxRows, 2 * L.bands * L.halfRows);
    }
    // box width: up to 15 bytes of misalignment + pixels -4 .. 4*groups+3 (+ the fourth word of the last group); 80 and
    // 96 bytes are the pitches whose rows two apart do not share banks
    f.bw = std::max(80, ((15 + 4 * ((maxCellW + 3) / 4) + 8) + 15) / 16 * 16);
    f.bh = maxCellH + 6;
    f.tileBytes = f.bw * f.bh;
    f.scorePitch = ((maxCellW + 2) + 3) / 4 * 4;
    if (((f.scorePitch / 4) & 1) == 0) f.scorePitch += 4;     // odd word pitch: vertical neighbours in different banks
    const int scoreBytes = ((maxCellH + 2) * f.scorePitch + 15) / 16 * 16;
    f.scoreVec = scoreBytes / 16;
    f.queueCap = 2 * maxCellW * maxCellH;                      // a pixel can pass the pre-test with both polarities
    f.survCap = slotCapMax;
    int p = 128;                                               // [0, 16): the two mbarriers
    f.offTile = p;
    f.tileStride = (f.tileBytes + 127) / 128 * 128;
    p += 2 * f.tileStride;
    f.offScore = p;
    p += scoreBytes;
    f.offQueue = p;
    p += (f.queueCap * 2 + 15) / 16 * 16;
    f.offSurv = p;
    p += (f.survCap * 2 + 15) / 16 * 16;
    // rows past a cell's last one are computed and masked: they read up to maxRows + 6 tile rows, which must stay inside
    // the warp's own region (buffer 0 overhangs into buffer 1, buffer 1 into the arrays behind it)
    const int overhang = (maxRows + 6 - f.bh) * f.bw;
    if (overhang > p - f.offScore) p = f.offScore + overhang;
    f.warpBytes = (p + 127) / 128 * 128;
    f.smemBytes = FW_WARPS * f.warpBytes;
    if (f.smemBytes > 227 * 1024)
        return fail(ORB_ERR_INVALID, "FAST: %d bytes of shared memory for %dx%d cells", f.smemBytes, maxCellW, maxCellH);
    *plan = f;
    return ORB_OK;
}

int fast_warp_encode_maps(const ExtractParams& P, int arenaFrames, void* hostMaps) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        ORB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !fn) return fail(ORB_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        encode = (EncodeFn)fn;
    }
    CUtensorMap* maps = static_cast<CUtensorMap*>(hostMaps);
    std::memset(maps, 0, sizeof(CUtensorMap) * kMaxLevels);
    for (int l = 0; l < P.nLevels; ++l) {
        const LevelGeom& L = P.lv[l];
        const cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)(L.h + 2 * kEdge), (cuuint64_t)arenaFrames};
        const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)P.pyrFrameBytes};
        const cuuint32_t box[3] = {(cuuint32_t)P.fw.bw, (cuuint32_t)P.fw.bh, 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        const CUresult r = encode(&maps[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, P.pyr + L.pyrOff, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(ORB_ERR_CUDA, "cuTensorMapEncodeTiled(level %d, %dx%d box %dx%d) -> %d", l, L.pitch, L.h + 2 * kEdge, P.fw.bw, P.fw.bh, (int)r);
    }
    return ORB_OK;
}

template <int BW>
static int launch_fast_warp_bw(const ExtractParams& P, cudaStream_t st) {
    static thread_local int ctasPerSm = 0, nSm = 0, smemSet = -1, devSet = -1;
    int dev = 0;
    ORB_CUDA(cudaGetDevice(&dev));
    if (smemSet != P.fw.smemBytes || devSet != dev) {
        ORB_CUDA(cudaFuncSetAttribute(fast_warp_kernel<BW>, cudaFuncAttributeMaxDynamicSharedMemorySize, P.fw.smemBytes));
        ORB_CUDA(cudaFuncSetAttribute(fast_warp_kernel<BW>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        ORB_CUDA(cudaDeviceGetAttribute(&nSm, cudaDevAttrMultiProcessorCount, dev));
        ORB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctasPerSm, fast_warp_kernel<BW>, FW_WARPS * 32, P.fw.smemBytes));
        if (ctasPerSm < 1) return fail(ORB_ERR_CUDA, "FAST kernel does not fit an SM (%d bytes of shared memory)", P.fw.smemBytes);
        smemSet = P.fw.smemBytes;
        devSet = dev;
    }
    const long long total = (long long)P.nFrames * P.nCellsTotal;
    const long long want = (total + FW_WARPS - 1) / FW_WARPS;
    const int grid = (int)std::min<long long>(want, (long long)nSm * ctasPerSm);
    fast_warp_kernel<BW><<<grid, FW_WARPS * 32, P.fw.smemBytes, st>>>(P);
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

int launch_fast_warp(const ExtractParams& P, cudaStream_t st, int* launches) {
    if (P.nCellsTotal == 0) return ORB_OK;
    if ((long long)P.nFrames * P.nCellsTotal >= (1ll << 31)) return fail(ORB_ERR_INVALID, "FAST: more than 2^31 cells in one launch");
    ++*launches;
    switch (P.fw.bw) {
        case 80: return launch_fast_warp_bw<80>(P, st);
        case 96: return launch_fast_warp_bw<96>(P, st);
    }
    return fail(ORB_ERR_INVALID, "FAST: unsupported tile pitch %d", P.fw.bw);
}

}  // namespace orbb
