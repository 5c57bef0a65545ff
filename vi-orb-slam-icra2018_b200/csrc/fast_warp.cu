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
//        issued as soon as the current one has read its tile for the last time (end of its last scoring pass), so
//        the copy overlaps the current cell's NMS / emission / clean-up without holding registers.  A TMA box must start on a 16-byte boundary of global memory (measured: an unaligned
//        innermost coordinate raises "illegal instruction", tools/microbench/tma_probe.cu), so the box starts at the
//        aligned column below x0-4 and the cell carries its misalignment `mis` (0..15); the tile pitch is 80 or 96
//        bytes (20 / 24 words), for which rows two apart fall into disjoint bank octets;
//     A. pre-test on every pixel, 4 pixels x 1 row per step in packed 16x2 arithmetic: a 9-arc contains one pixel of
//        each opposing circle pair, so min(max(N,S), max(E,W)) > v+t  (bright) or  max(min(N,S), min(E,W)) < v-t (dark)
//        is necessary.  8 aligned word loads, 5 PRMT that undo the misalignment (selectors uniform per cell), 10 PRMT
//        that widen, 12 VIMNMX, 8 threshold ops per 4 pixels; the results of 8 rows x 4 pixels accumulate in one
//        32-bit mask per polarity (no per-row bookkeeping);
//     B. the set bits become queue entries (x | y<<6 | polarity<<15).  Lane (r, q) of a step owns row 8m + r and the
//        4-px group 4t + q, so a step's 32 lanes read 32 different banks; the per-lane counts of a chunk's four step
//        pairs are prefix-summed over the warp (two packed shuffle scans) and every lane writes its own entries, which
//        makes the queue block-major, lane-major: 32 consecutive entries = 4-5 pixel rows = distinct bank octets, so
//        the byte gathers of phase C see ~1.1 wavefronts per load instead of 2.5;
//     C. exact score of the queued pixels, TWO per lane: the pre-test told the polarity, so entry A rides in the low
//        and entry B in the high 16-bit half of every operand (d = 256 +- (p - v) by one IMAD each), and the 40
//        three-input min/max of the 9-arc scan serve both.  Corners (S >= t) go to the score map and are compacted in
//        place to the front of the queue;
//     D. NMS over the corner list, survivors listed, ranked by (y, x), packed x:12|y:12|score:8 into the cell's slot;
//        the score map is zeroed again through the corner list.
//   The shared-memory queue holds a typical cell (~300 candidates) with room to spare; a cell with more candidates
//   than that (up to 2 per pixel) is redone with its queue in a per-warp global scratch area.
//   A cell with no survivor at iniThFAST simply runs A-D again at minThFAST (the reference's second cv::FAST call).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "extractor.h"

namespace orbb {

#ifdef ORBB_FW_STATS
__device__ unsigned long long gFwStats[8];   // cells, rounds, second rounds, overflows, queue entries, corners, survivors
#define FW_STAT(i, v) do { if (lane == 0) atomicAdd(&gFwStats[i], (unsigned long long)(v)); } while (0)
#else
#define FW_STAT(i, v) do { } while (0)
#endif

constexpr int FW_WARPS = 8;                       // warps per CTA; they share nothing but the bit -> pixel tables
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
// no borrow crosses the halves and each expression is one three-input add.  vA / vB are FW_PASS restricted to the
// pixels that exist (cell edge), so validity costs nothing here.
struct FwAlign {
    unsigned int selC, selL, selR;
    bool hiR;
    unsigned int one, neg;   // 1 and -1 the compiler cannot see through (adds as IMAD on the FMA pipe)
};
// a * m + b on the FMA pipe (IMAD); m is a runtime +-1 or power of two
__device__ __forceinline__ unsigned int fma_u32(unsigned int a, unsigned int m, unsigned int b) {
    unsigned int r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(m), "r"(b));
    return r;
}
template <int BW>
__device__ __forceinline__ void pretest_row(const unsigned char* r, const FwAlign& al, unsigned int K, unsigned int vA, unsigned int vB,
                                            unsigned int& bright, unsigned int& dark) {
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
    // bit 9 / 25 of (mm - c + K) and of (c - nn + K); pair B's bits move one up; the adds run as IMADs (FMA pipe)
    const unsigned int kmA = fma_u32(cA, al.neg, K), kmB = fma_u32(cB, al.neg, K);     // K - c
    const unsigned int kpA = fma_u32(cA, al.one, K), kpB = fma_u32(cB, al.one, K);     // K + c
    const unsigned int bA = fma_u32(mmA, al.one, kmA), bB = fma_u32(mmB, al.one, kmB);
    const unsigned int kA = fma_u32(nnA, al.neg, kpA), kB = fma_u32(nnB, al.neg, kpB);
    bright = fma_u32(bB & vB, al.one + al.one, bA & vA);
    dark = fma_u32(kB & vB, al.one + al.one, kA & vA);
}

// FW_PASS restricted to the first n (of 4) pixels of a group: pair A = pixels 0, 1 (bits 9, 25), pair B = pixels 2, 3
__device__ __forceinline__ void valid_pairs(int n, unsigned int& vA, unsigned int& vB) {
    vA = (n >= 1 ? 0x00000200u : 0u) | (n >= 2 ? 0x02000000u : 0u);
    vB = (n >= 3 ? 0x00000200u : 0u) | (n >= 4 ? 0x02000000u : 0u);
}

// circle offsets in tile bytes, OpenCV order (dx,dy) = (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)
// (-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)
#define FW_CIRCLE(k, P)                                                                                    \
    ((k) == 0 ? 3 * (P) : (k) == 1 ? 3 * (P) + 1 : (k) == 2 ? 2 * (P) + 2 : (k) == 3 ? (P) + 3             \
     : (k) == 4 ? 3 : (k) == 5 ? -(P) + 3 : (k) == 6 ? -2 * (P) + 2 : (k) == 7 ? -3 * (P) + 1              \
     : (k) == 8 ? -3 * (P) : (k) == 9 ? -3 * (P)-1 : (k) == 10 ? -2 * (P)-2 : (k) == 11 ? -(P)-3           \
     : (k) == 12 ? -3 : (k) == 13 ? (P)-3 : (k) == 14 ? 2 * (P)-2 : 3 * (P)-1)

struct FwWarp {            // the warp's arrays
    const unsigned char* tile;     // shared memory: current tile buffer
    unsigned char* scorePix;       // shared memory: score of pixel (x, y) at scorePix[y * scorePitch + x], zero ring around it
    unsigned short* queue;         // shared memory, or the warp's global scratch for a cell that overflowed it
    unsigned int* bitmap;          // shared memory: NMS survivors, one 64-bit row per pixel row
    int queueCap;
    unsigned short* globalQueue;
    int globalCap;
};

struct FwSurvivors {       // a lane's view of the survivor bitmap: pixel rows `lane` and `lane + 32`
    uint2 rowLo, rowHi;
    int rankLo, rankHi;    // survivors in the rows before them
};

// B: the set bits of one chunk's masks become queue entries (x | y << 6 | polarity << 15).  A "unit" is a group of four
// steps (two 8-row blocks when T = 2); the two units' per-lane counts are prefix-summed over the warp in one packed
// shuffle scan.  Returns the new queue length, or -1 (nothing written) if the chunk does not fit.
// SMEM: the queue is the warp's shared-memory one (the common case): 32-bit addresses and st.shared instead of generic stores.
template <bool SMEM>
__device__ __forceinline__ int enqueue_chunk(unsigned int mb, unsigned int md, unsigned int laneEntry, const unsigned short (*lut)[32],
                                             unsigned short* queue, int nq, int cap, int lane) {
    constexpr unsigned int U0 = 0xFE01FE01u, U1 = 0x01FE01FEu;   // steps 0-3: bits 9..16, 25..31, 0; steps 4-7: the rest
    const unsigned int c0 = __popc(mb & U0) + __popc(md & U0), c1 = __popc(mb & U1) + __popc(md & U1);
    unsigned int s01 = c0 | (c1 << 16);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int a = __shfl_up_sync(FW_FULL, s01, d);
        if (lane >= d) s01 += a;
    }
    const unsigned int t01 = __shfl_sync(FW_FULL, s01, 31);
    const int base1 = nq + (int)(t01 & 0xffffu), nqNew = base1 + (int)(t01 >> 16);
    if (nqNew > cap) return -1;
    const int pos[2] = {nq + (int)(s01 & 0xffffu) - (int)c0, base1 + (int)(s01 >> 16) - (int)c1};
    const unsigned int um[2] = {U0, U1};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        // bright bits stay, dark bits move 8 up (into the other unit's positions, which are masked out here)
        unsigned int c = (mb & um[u]) | __funnelshift_l(md & um[u], md & um[u], 8);
        c = __funnelshift_r(c, c, 8 * u);
        if (SMEM) {
            unsigned int w = smem_u32(queue) + 2u * (unsigned int)pos[u];
            const unsigned int l = smem_u32(lut[u]);
            while (c) {
                const unsigned int k = 31u - (unsigned int)__clz(c);
                c ^= 1u << k;
                unsigned short v;
                asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(l + 2u * k));
                v = (unsigned short)(laneEntry + v);
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(w), "h"(v) : "memory");
                w += 2;
            }
        } else {
            unsigned short* w = queue + pos[u];
            const unsigned short* l = lut[u];
            while (c) {
                const int k = 31 - __clz(c);
                c ^= 1u << k;
                *w++ = (unsigned short)(laneEntry + l[k]);
            }
        }
    }
    return nqNew;
}

// A-C for one cell at threshold th: everything that reads the tile.  Returns the number of corners (W.queue[0, nc),
// scores in the score map).  A cell with more candidates than the shared-memory queue holds moves its queue to the
// warp's global scratch.
// T == 2: the mask bits of this lane's pixels (4-px groups q and q + 4 of every row) that lie inside a cell of width cw
__device__ __forceinline__ unsigned int column_mask2(int cw, int q) {
    unsigned int vA0, vB0, vA1, vB1;
    valid_pairs(cw - 4 * q, vA0, vB0);
    valid_pairs(cw - 16 - 4 * q, vA1, vB1);
    const unsigned int m0 = vA0 | (vB0 << 1), m1 = vA1 | (vB1 << 1);   // step 0 and step 1; steps 2i, 2i + 1 are these rotated by 4i
    const unsigned int m = m0 | __funnelshift_l(m1, m1, 2);
    const unsigned int m2 = m | __funnelshift_l(m, m, 4);
    return m2 | __funnelshift_l(m2, m2, 8);
}

template <int BW>
__device__ __forceinline__ int fast_front(FwWarp& W, const FastWarpPlan::Level& F, const unsigned short (*lut)[32], int SP, int cw, int ch,
                                          int mis, int th, int lane, unsigned int ltMask, unsigned int one, unsigned int colMask2) {
    // ---- A + B
    const unsigned int K = FW_PASS - (unsigned int)(th + 1) * 0x00010001u;
    FwAlign al;
    {
        const unsigned int sh = (unsigned int)(mis & 3);
        al.selC = 0x3210u + 0x1111u * sh;
        al.selL = al.selC + 0x1111u;
        al.hiR = sh >= 2u;
        al.selR = 0x3210u + 0x1111u * (al.hiR ? sh - 1u : sh + 3u);
        al.one = one;
        al.neg = 0u - one;
    }
    const int r = lane >> 2, q = lane & 3;
    const unsigned char* tilePix = W.tile + mis + 3 * BW + 4;              // pixel (0, 0)
    const unsigned char* rowPtr = W.tile + (mis & ~3) + (r + 3) * BW + 4 * q;   // aligned word of pixel (4q - 4, r)
    const int T = F.T, chunkSteps = F.chunkSteps, steps = F.steps;
    const int chunkRows = 8 * (chunkSteps / T);
    int nq = 0;
#pragma unroll 1
    for (int j0 = 0, y0 = r; j0 < steps; j0 += chunkSteps, y0 += chunkRows, rowPtr += chunkRows * BW) {
        const int nSteps = min(chunkSteps, steps - j0);
        unsigned int mb = 0, md = 0;
        if (T == 2) {
            // 8 groups per row: step i = (block i>>1, half i&1).  Pixels past the cell's right edge and rows past its last
            // one are computed like the others and struck from the finished masks: step i's bits sit at 9, 10, 25, 26 rotated
            // by 2i, so the valid rows of the chunk are two runs of 4 bits per 8-row block, and the columns' mask is the
            // same for every chunk of the cell
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < nSteps) {
                    unsigned int tb, td;
                    pretest_row<BW>(rowPtr + (i >> 1) * 8 * BW + 16 * (i & 1), al, K, FW_PASS, FW_PASS, tb, td);
                    mb |= __funnelshift_l(tb, tb, 2 * i);
                    md |= __funnelshift_l(td, td, 2 * i);
                }
            }
            const int nb = (ch - y0 + 7) >> 3;                      // 8-row blocks of this chunk with a row inside the cell (for this lane)
            const unsigned int run = nb >= 8 ? 0xffffffffu : nb <= 0 ? 0u : (1u << (4 * nb)) - 1u;
            const unsigned int ok = colMask2 & (__funnelshift_l(run, run, 9) | __funnelshift_l(run, run, 25));
            mb &= ok;
            md &= ok;
        } else {
            int t = 0, y = y0;
            const unsigned char* rp = rowPtr;
#pragma unroll 1
            for (int i = 0; i < nSteps; ++i) {
                unsigned int vA, vB, tb, td;
                valid_pairs(y < ch ? cw - 16 * t - 4 * q : 0, vA, vB);
                pretest_row<BW>(rp + 16 * t, al, K, vA, vB, tb, td);
                mb |= __funnelshift_l(tb, tb, 2 * i);
                md |= __funnelshift_l(td, td, 2 * i);
                if (++t == T) { t = 0; y += 8; rp += 8 * BW; }
            }
        }
        int n2 = W.queue != W.globalQueue ? enqueue_chunk<true>(mb, md, (unsigned int)((y0 << 6) | (4 * q)), lut, W.queue, nq, W.queueCap, lane)
                                           : enqueue_chunk<false>(mb, md, (unsigned int)((y0 << 6) | (4 * q)), lut, W.queue, nq, W.queueCap, lane);
        if (n2 < 0) {
            __syncwarp();
            for (int i = lane; i < nq; i += 32) W.globalQueue[i] = W.queue[i];
            W.queue = W.globalQueue;
            W.queueCap = W.globalCap;
            __syncwarp();
            n2 = enqueue_chunk<false>(mb, md, (unsigned int)((y0 << 6) | (4 * q)), lut, W.queue, nq, W.queueCap, lane);
        }
        nq = n2;
    }
    __syncwarp();

    // ---- C: exact score, two queue entries per lane (A in the low halves, B in the high halves)
    int nc = 0;
#pragma unroll 1
    for (int q0 = 0; q0 < nq; q0 += 64) {
        const int iA = q0 + lane, iB = iA + 32;
        const bool vA = iA < nq, vB = iB < nq;
        const unsigned int eA = vA ? W.queue[iA] : 0u, eB = vB ? W.queue[iB] : 0u;
        __syncwarp();   // every lane has read its entries before any lane overwrites them with corners below
        const int xA = eA & 63, yA = (eA >> 6) & 63, xB = eB & 63, yB = (eB >> 6) & 63;
        const unsigned char* pA = tilePix + yA * BW + xA;
        const unsigned char* pB = tilePix + yB * BW + xB;
        // d = 256 + (p - v) for a bright candidate, 256 + (v - p) for a dark one: every half stays in [1, 511]
        // the multipliers go through an opaque mad (fma_u32): seeing the +-1, the compiler rewrote every d[k] as
        // (pA - vA) -> predicated negate -> (pB - vB) * mulB + .. + constant, five instructions where two do
        const unsigned int mulA = (eA & 0x8000u) ? 0xffffffffu : 1u, mulB = (eB & 0x8000u) ? 0xffff0000u : 0x00010000u;
        const unsigned int C = 0x01000100u - mulA * pA[0] - mulB * pB[0];
        unsigned int d[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) d[k] = fma_u32(pA[FW_CIRCLE(k, BW)], mulA, fma_u32(pB[FW_CIRCLE(k, BW)], mulB, C));
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
        if (cA) W.scorePix[yA * SP + xA] = (unsigned char)sA;
        if (cB) W.scorePix[yB * SP + xB] = (unsigned char)sB;
        // corners move to the front of the queue (never past the entries already read: nc <= q0 + 64)
        const unsigned int balA = __ballot_sync(FW_FULL, cA), balB = __ballot_sync(FW_FULL, cB);
        if (cA) W.queue[nc + __popc(balA & ltMask)] = (unsigned short)(eA & 0xfffu);
        nc += __popc(balA);
        if (cB) W.queue[nc + __popc(balB & ltMask)] = (unsigned short)(eB & 0xfffu);
        nc += __popc(balB);
        __syncwarp();
    }
    FW_STAT(4, nq);
    return nc;
}

// D for one cell: non-max suppression over the corners; survivors become bits of the row bitmap.  Returns their number;
// S = this lane's rows of the bitmap and the ranks of the rows' first survivors.
__device__ __forceinline__ int fast_back(const FwWarp& W, int SP, int nc, int lane, FwSurvivors& S) {
    reinterpret_cast<uint4*>(W.bitmap)[lane] = make_uint4(0, 0, 0, 0);
    __syncwarp();
#pragma unroll 1
    for (int q0 = 0; q0 < nc; q0 += 32) {
        const bool v = q0 + lane < nc;
        const unsigned int e = v ? W.queue[q0 + lane] : 0u;
        const int x = e & 63, y = e >> 6;
        const unsigned char* sc = W.scorePix + y * SP + x;
        const unsigned char* up = sc - SP;
        const unsigned char* dn = sc + SP;
        const int s = sc[0];
        const int m = max(max(max(up[-1], up[0]), max(up[1], sc[-1])), max(max(sc[1], dn[-1]), max(dn[0], dn[1])));
        if (v && s > m) atomicOr(W.bitmap + 2 * y + (x >> 5), 1u << (x & 31));
    }
    __syncwarp();
    // survivors per row -> ranks of the rows' first survivors (row-major emission order), one packed warp scan
    S.rowLo = reinterpret_cast<const uint2*>(W.bitmap)[lane];
    S.rowHi = reinterpret_cast<const uint2*>(W.bitmap)[lane + 32];
    const unsigned int cl = __popc(S.rowLo.x) + __popc(S.rowLo.y), chh = __popc(S.rowHi.x) + __popc(S.rowHi.y);
    unsigned int incl = cl | (chh << 16);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int a = __shfl_up_sync(FW_FULL, incl, d);
        if (lane >= d) incl += a;
    }
    const unsigned int tot = __shfl_sync(FW_FULL, incl, 31);
    S.rankLo = (int)(incl & 0xffffu) - (int)cl;
    S.rankHi = (int)(tot & 0xffffu) + (int)(incl >> 16) - (int)chh;
    return (int)(tot & 0xffffu) + (int)(tot >> 16);
}

template <int BW>
__global__ void __launch_bounds__(FW_WARPS * 32, 3) fast_warp_kernel(const __grid_constant__ ExtractParams P) {
    extern __shared__ __align__(128) unsigned char fsm[];
    const FastWarpPlan& F = P.fw;
    // per level and unit: normalised mask bit -> x | y << 6 | polarity << 15 relative to the lane's first pixel of the chunk
    unsigned short (*sLut)[2][32] = reinterpret_cast<unsigned short (*)[2][32]>(fsm);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int ltMask = (1u << lane) - 1u;
    unsigned char* wb = fsm + F.lutBytes + (size_t)warp * F.warpBytes;
    const unsigned int bar = smem_u32(wb);              // the tile's mbarrier
    const unsigned int tileAddr = smem_u32(wb + F.offTile);
    const int SP = F.scorePitch;
    FwWarp W;
    W.tile = wb + F.offTile;
    W.scorePix = wb + F.offScore + SP + 16;
    W.bitmap = reinterpret_cast<unsigned int*>(wb + F.offBitmap);
    unsigned short* const smemQueue = reinterpret_cast<unsigned short*>(wb + F.offQueue);

    for (int idx = threadIdx.x; idx < P.nLevels * 64; idx += FW_WARPS * 32) {
        // after the rotation by 8u a unit's bits sit at 9 + kk: kk = 2 * step + (0: pixel 0, 1: pixel 2), + 16 for pixels 1 / 3,
        // + 8 for the dark polarity
        const int level = idx >> 6, u = (idx >> 5) & 1, k = idx & 31, kk = (k - 9) & 31;
        const int pol = (kk >> 3) & 1, st = (kk & 7) >> 1, px = ((kk & 1) << 1) | (kk >> 4);
        const int i = 4 * u + st, T = F.lv[level].T, dm = i / T, t = i - dm * T;
        sLut[level][u][k] = (unsigned short)((16 * t + px) | ((8 * dm) << 6) | (pol << 15));
    }
    {
        uint4* z = reinterpret_cast<uint4*>(wb + F.offScore);
        for (int i = lane; i < F.scoreVec; i += 32) z[i] = make_uint4(0, 0, 0, 0);
    }
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const unsigned int nCells = (unsigned int)P.nCellsTotal;
    const unsigned int total = (unsigned int)P.nFrames * nCells;
    const unsigned int warpId = blockIdx.x * FW_WARPS + warp;
    W.globalQueue = F.scratch + (size_t)warpId * F.scratchCap;
    W.globalCap = F.scratchCap;
    const unsigned char* maps = static_cast<const unsigned char*>(F.maps);
    // one tile buffer per warp: the load of a cell's tile is issued by lane 0 as soon as the previous cell has read its
    // tile for the last time (end of its last scoring pass), so it overlaps that cell's NMS / emission / clean-up
    auto request_tile = [&](const uint4& c, unsigned int frame) {
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar, (unsigned int)F.tileBytes);
            tma_load_tile(tileAddr, maps + 128 * (int)(short)(c.x & 0xffffu), (kPadLeft + (int)(short)(c.x >> 16) - 4) & ~15,
                          kEdge + (int)(short)(c.y & 0xffffu) - 3, F.frameBase + (int)frame, bar);
        }
    };
    // cells are handed out one at a time (their cost varies 10x between a flat and a busy cell)
    unsigned int it = 0;
    if (lane == 0) it = atomicAdd(F.counters, 1u);
    it = __shfl_sync(FW_FULL, it, 0);
    uint4 c0 = make_uint4(0, 0, 0, 0);
    unsigned int frame = 0, cell = 0;
    if (it < total) {
        frame = it / nCells;
        cell = it - frame * nCells;
        c0 = __ldg(reinterpret_cast<const uint4*>(P.cells + cell));
        request_tile(c0, frame);
    }
    unsigned int parity = 0;
    int maskCw = -1;
    unsigned int colMask2 = 0;
    while (it < total) {
        // ---- ticket of the next cell: asked for now, looked at after this cell's first scoring pass
        unsigned int nIt = 0;
        if (lane == 0) nIt = atomicAdd(F.counters, 1u);
        // ---- this cell
        const int cellX0 = (int)(short)(c0.x >> 16), cellY0 = (int)(short)(c0.y & 0xffffu);
        const int cw = (int)(c0.y >> 16), ch = (int)(c0.z & 0xffffu), cellSlot = (int)c0.w;
        const int level = (int)(short)(c0.x & 0xffffu), mis = (kPadLeft + cellX0 - 4) & 15;
        if (cw != maskCw) {   // the column mask of the pre-test changes only at a cell of another width (the last column of a level)
            maskCw = cw;
            colMask2 = column_mask2(cw, lane & 3);
        }
        mbar_wait(bar, parity);
        parity ^= 1u;
        W.queue = smemQueue;
        W.queueCap = F.queueCap;

        FwSurvivors S;
        unsigned int nFrame = 0, nCell = 0;
        uint4 n0 = c0;
        bool nextKnown = false, nextRequested = false;
        int sn, nc, th = P.iniTh;
        for (;;) {   // a second round at minThFAST is the reference's second cv::FAST call (:811-818)
            nc = fast_front<BW>(W, F.lv[level], sLut[level], SP, cw, ch, mis, th, lane, ltMask, F.one, colMask2);
            FW_STAT(1, 1);
            if (!nextKnown) {
                nIt = __shfl_sync(FW_FULL, nIt, 0);
                if (nIt < total) {
                    nFrame = nIt / nCells;
                    nCell = nIt - nFrame * nCells;
                    n0 = __ldg(reinterpret_cast<const uint4*>(P.cells + nCell));
                }
                nextKnown = true;
            }
            const bool lastRound = th <= P.minTh;
            // the tile is dead unless a second round follows, which needs nc == 0 or (rarely) every corner losing its NMS
            if (nIt < total && !nextRequested && (lastRound || nc > 0)) {
                __syncwarp();
                request_tile(n0, nFrame);
                nextRequested = true;
            }
            sn = fast_back(W, SP, nc, lane, S);
            if (sn != 0 || lastRound) break;
            th = P.minTh;
            FW_STAT(2, 1);
            if (nextRequested) {   // the next cell's tile is on its way into the buffer: let it land, then fetch this cell's again
                mbar_wait(bar, parity);
                parity ^= 1u;
                __syncwarp();
                request_tile(c0, frame);
                mbar_wait(bar, parity);
                parity ^= 1u;
                nextRequested = false;
            }
        }
        FW_STAT(0, 1);
        FW_STAT(5, nc);
        FW_STAT(6, sn);

        // ---- every lane emits the survivors of its two pixel rows in x order, starting at the rows' ranks
        if (lane == 0) P.cellCount[(size_t)frame * P.nCellsTotal + cell] = sn;
        unsigned int* slot = P.slots + (size_t)frame * P.slotFrameEntries + cellSlot;
        {
            unsigned long long bits = (unsigned long long)S.rowLo.x | ((unsigned long long)S.rowLo.y << 32);
            const unsigned int yKey = (unsigned int)(cellY0 + lane - 16) << 8;
            const unsigned char* srow = W.scorePix + lane * SP;
            unsigned int* o = slot + S.rankLo;
            while (bits) {
                const int x = __ffsll((long long)bits) - 1;
                bits &= bits - 1;
                *o++ = ((unsigned int)(cellX0 + x - 16) << 20) | yKey | srow[x];
            }
        }
        if (ch > 32) {
            unsigned long long bits = (unsigned long long)S.rowHi.x | ((unsigned long long)S.rowHi.y << 32);
            const unsigned int yKey = (unsigned int)(cellY0 + lane + 32 - 16) << 8;
            const unsigned char* srow = W.scorePix + (lane + 32) * SP;
            unsigned int* o = slot + S.rankHi;
            while (bits) {
                const int x = __ffsll((long long)bits) - 1;
                bits &= bits - 1;
                *o++ = ((unsigned int)(cellX0 + x - 16) << 20) | yKey | srow[x];
            }
        }
        __syncwarp();
        // the score map goes back to zero by the corner list (a second round's corners include the first round's)
#pragma unroll 1
        for (int i = lane; i < nc; i += 32) {
            const unsigned int e = W.queue[i];
            W.scorePix[(int)(e >> 6) * SP + (int)(e & 63)] = 0;
        }
        __syncwarp();
        it = nIt;
        frame = nFrame;
        cell = nCell;
        c0 = n0;
    }
    // the last warp out re-arms the counters for the next launch
    if (lane == 0) {
        __threadfence();
        if (atomicAdd(F.counters + 1, 1u) == gridDim.x * FW_WARPS - 1) {
            F.counters[0] = 0;
            F.counters[1] = 0;
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------

int fast_warp_plan(int nLevels, const int* cellW, const int* cellH, int slotCapMax, FastWarpPlan* plan) {
    FastWarpPlan f;
    std::memset(&f, 0, sizeof f);
    int maxCellW = 1, maxCellH = 1, maxRows = 0, maxGroups = 1;
    for (int l = 0; l < nLevels; ++l) {
        FastWarpPlan::Level& L = f.lv[l];
        const int w = std::max(cellW[l], 1), h = std::max(cellH[l], 1);
        maxCellW = std::max(maxCellW, w);
        maxCellH = std::max(maxCellH, h);
        const int groups = (w + 3) / 4, T = std::max(2, (groups + 3) / 4), blocks = (h + 7) / 8;
        if (T > 4 || blocks > 8) return fail(ORB_ERR_INVALID, "FAST: cell %dx%d is larger than 64x64", w, h);
        L.T = (unsigned char)T;
        L.chunkSteps = (unsigned char)(8 / T * T);
        L.steps = (unsigned char)(T * blocks);
        maxRows = std::max(maxRows, 8 * blocks);
        maxGroups = std::max(maxGroups, 4 * T);
    }
    // box width: up to 15 bytes of misalignment + pixels -4 .. 4*groups+3 (+ the fourth word of the last group); 80 and
    // 96 bytes are the pitches whose rows fall into distinct bank octets (20 y mod 32 takes 8 values, 24 y takes 4)
    f.bw = std::max(80, ((15 + 4 * maxGroups + 12) + 15) / 16 * 16);
    if (f.bw > 96) return fail(ORB_ERR_INVALID, "FAST: tile pitch %d", f.bw);
    f.bh = maxCellH + 6;
    f.tileBytes = f.bw * f.bh;
    f.scorePitch = ((maxCellW + 2) + 3) / 4 * 4;
    if ((f.scorePitch / 4) % 2 == 0) f.scorePitch += 4;        // odd word pitch: the rows of a 3x3 neighbourhood in different banks
    const int scoreBytes = ((maxCellH + 2) * f.scorePitch + 32 + 15) / 16 * 16;
    f.scoreVec = scoreBytes / 16;
    f.lutBytes = nLevels * 128;
    f.one = 1u;
    f.scratchCap = 2 * maxCellW * maxCellH + 64;               // a pixel can pass the pre-test with both polarities
    f.queueCap = std::min(f.scratchCap, 640);                  // a typical 31x31 cell queues ~300
    (void)slotCapMax;
    int p = 128;                                               // [0, 8): the tile's mbarrier
    f.offTile = p;
    f.tileStride = (f.tileBytes + 127) / 128 * 128;
    p += f.tileStride;
    f.offScore = p;
    p += scoreBytes;
    f.offQueue = p;
    p += (f.queueCap * 2 + 15) / 16 * 16;
    f.offBitmap = p;
    p += 512;                                                  // 64 pixel rows x 64 bits
    // rows past a cell's last one are computed and masked: they read up to maxRows + 6 tile rows, which must stay inside
    // the warp's own region (the tile buffer overhangs into the arrays behind it)
    const int overhang = (maxRows + 6 - f.bh) * f.bw + 64;
    if (overhang > p - f.offScore) p = f.offScore + overhang;
    f.warpBytes = (p + 127) / 128 * 128;
    f.smemBytes = f.lutBytes + FW_WARPS * f.warpBytes;
    if (f.smemBytes > 220 * 1024)
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
static int fast_warp_occupancy(const FastWarpPlan& f, int* nSm, int* ctasPerSm) {
    static thread_local int cSm = 0, cCtas = 0, cSmem = -1, cDev = -1;
    int dev = 0;
    ORB_CUDA(cudaGetDevice(&dev));
    if (cSmem != f.smemBytes || cDev != dev) {
        ORB_CUDA(cudaFuncSetAttribute(fast_warp_kernel<BW>, cudaFuncAttributeMaxDynamicSharedMemorySize, f.smemBytes));
        ORB_CUDA(cudaFuncSetAttribute(fast_warp_kernel<BW>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        ORB_CUDA(cudaDeviceGetAttribute(&cSm, cudaDevAttrMultiProcessorCount, dev));
        ORB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cCtas, fast_warp_kernel<BW>, FW_WARPS * 32, f.smemBytes));
        if (cCtas < 1) return fail(ORB_ERR_CUDA, "FAST kernel does not fit an SM (%d bytes of shared memory)", f.smemBytes);
        cSmem = f.smemBytes;
        cDev = dev;
    }
    *nSm = cSm;
    *ctasPerSm = cCtas;
    return ORB_OK;
}

int fast_warp_max_warps(const FastWarpPlan& f, int* warps) {
    int nSm = 0, ctas = 0;
    if (f.bw == 80) ORB_CHECK(fast_warp_occupancy<80>(f, &nSm, &ctas));
    else ORB_CHECK(fast_warp_occupancy<96>(f, &nSm, &ctas));
    *warps = nSm * ctas * FW_WARPS;
    return ORB_OK;
}

template <int BW>
static int launch_fast_warp_bw(const ExtractParams& P, cudaStream_t st) {
    int nSm = 0, ctasPerSm = 0;
    ORB_CHECK(fast_warp_occupancy<BW>(P.fw, &nSm, &ctasPerSm));
    const long long total = (long long)P.nFrames * P.nCellsTotal;
    const long long want = (total + FW_WARPS - 1) / FW_WARPS;
    const int grid = (int)std::min<long long>(want, (long long)nSm * ctasPerSm);
    if (grid * FW_WARPS > P.fw.maxWarps) return fail(ORB_ERR_CUDA, "FAST: scratch sized for %d warps, launch has %d", P.fw.maxWarps, grid * FW_WARPS);
    fast_warp_kernel<BW><<<grid, FW_WARPS * 32, P.fw.smemBytes, st>>>(P);
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

#ifdef ORBB_FW_STATS
extern "C" int orbx_debug_fast_stats(unsigned long long* out8, int reset) {
    ORB_CUDA(cudaDeviceSynchronize());
    ORB_CUDA(cudaMemcpyFromSymbol(out8, gFwStats, sizeof gFwStats));
    if (reset) {
        unsigned long long z[8] = {0};
        ORB_CUDA(cudaMemcpyToSymbol(gFwStats, z, sizeof z));
    }
    return ORB_OK;
}
#endif

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
