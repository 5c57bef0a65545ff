// Image pyramid (north-star kernel 1): ORBextractor::ComputePyramid, ORBextractor.cc:1128-1153.
//
//   level 0      = copyMakeBorder(image, 19, BORDER_REFLECT_101)
//   level l >= 1 = resize(level l-1, INTER_LINEAR) then copyMakeBorder(.., 19, BORDER_REFLECT_101 | ISOLATED)
//
// OpenCV's 8-bit INTER_LINEAR is fixed point: per axis a source index and an 11-bit coefficient pair (tables built on
// the host with the exact float recipe, see extractor.cu), horizontal sums kept at 19 bits, vertical pass
// (((b0*(T0>>4))>>16) + ((b1*(T1>>4))>>16) + 2) >> 2.  The kernels write the 19-px reflect-101 frame in the same pass
// by evaluating the reflected interior coordinate, so the border costs no extra launch and no read-after-write.
//
// Shape (instruction-bound kernel; every lane does the same thing):
//   * work item = 4 horizontally adjacent output bytes (one aligned 32-bit store) x 16 output rows; items are numbered
//     row-band-major and laid over the threads linearly, so warps are full whatever the level's width;
//   * a work item resolves its four destination columns ONCE: reflected column -> source offset + coefficient pair.
//     Interior, mirrored (left/right frame) and straddling groups all read their source bytes from one 8-byte window
//     that starts at the group's smallest source offset, so they share a single code path: per source row three aligned
//     32-bit loads, two funnel shifts that bring the window to byte 0, two byte permutes with per-item selectors and
//     four IDP.2A (16-bit coefficient pair x two bytes) give the four horizontal sums;
//   * consecutive output rows usually step one source row (scale 1.2: five times out of six), so the lower row's sums
//     are kept for the next output row;
//   * the vertical pass (b0 * (T0 >> 4) >> 16) + (b1 * (T1 >> 4) >> 16) + 2 uses 32-bit products (T >> 4 < 2^15, b <= 2^11):
//     two IMAD, one PRMT that picks the two high halves, one IDP.2A that adds them and the rounding term.  (Until late in
//     round 2 this was two multiply-high per pixel; IMAD.HI issues at about half the rate of IMAD and made the
//     FMA-heavy pipe the kernel's busiest one: 3.95 -> 3.49 ms per 4096 frames.)
// Reading S[x+1] / row y+1 one past the level is harmless: the table's coefficient there is 0 and the source level
// has its own frame.
//
// Kernels in this file (round 2), all with the arithmetic above:
//   pyramid_level0_kernel / _wide_kernel  level 0 = copy + frame (wide: 128-bit interior copies for 16-byte aligned rows)
//   pyramid_level0_bulk_kernel            level 0 of a batch with 16-byte aligned rows: bulk copies in, frame filled in shared
//                                         memory, one bulk store per band of 16 padded rows
//   pyramid_resize3_kernel                batches: a band of 16 destination rows per CTA, its contiguous source rows staged
//                                         in shared memory by one cp.async.bulk (mbarrier, persistent CTAs)
//   pyramid_fused_kernel                  small calls: all levels in ONE launch, grid-wide barrier between levels
//   pyramid_resize2_kernel                record-driven, straight from global memory (fallback)
//   pyramid_resize_kernel                 table-driven, any scale factor (fallback)
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "extractor.h"

namespace orbb {

constexpr int PY_ROWS = 16;       // output rows per work item

// one output pixel of the vertical pass from the two source rows' horizontal sums (already >> 4) and coefficients
__device__ __forceinline__ unsigned int py_vsum(unsigned int t0, unsigned int b0, unsigned int t1, unsigned int b1) {
    return __dp2a_lo(__byte_perm(t0 * b0, t1 * b1, 0x7632), 0x0101u, 2u);
}
constexpr int PY_THREADS = 128;

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

// destination column of output byte bx + k: reflect-101 inside the 19-px frame; the unused alignment bytes beyond the
// frame get a clamped (deterministic, never used) column
__device__ __forceinline__ int frame_column(int lx, int w) { return min(max(reflect101(lx, w), 0), w - 1); }

__global__ void __launch_bounds__(PY_THREADS)
pyramid_level0_kernel(const unsigned char* __restrict__ images, int w, int h, int stride, size_t frameStride,
                      unsigned char* __restrict__ pyr, long long pyrFrameBytes, long long pyrOff, int pitch, int groups,
                      int nItems, int wordLoads) {
    const int item = blockIdx.x * PY_THREADS + threadIdx.x;
    if (item >= nItems) return;
    const int band = item / groups, g = item - band * groups;
    const int bx = 4 * g, lx0 = bx - kPadLeft;
    const unsigned char* img = images + (size_t)blockIdx.y * frameStride;
    unsigned char* dst = pyr + (size_t)blockIdx.y * pyrFrameBytes + pyrOff + bx;
    int col[4], cmin = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < 4; ++k) { col[k] = frame_column(lx0 + k, w); cmin = min(cmin, col[k]); }
    const int base = cmin & ~3;
    unsigned int sel = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) sel |= (unsigned int)(col[k] - base) << (4 * k);   // the columns span <= 4 bytes: index <= 6
    const int by0 = band * PY_ROWS, rowsTotal = h + 2 * kEdge;
    if (wordLoads) {
        // image rows are 4-byte aligned: two aligned words hold the four (possibly mirrored) source bytes
        const bool second = base + 4 < ((w + 3) & ~3);   // stay inside the row's last word
#pragma unroll 4
        for (int r = 0; r < PY_ROWS; ++r) {
            const int by = by0 + r;
            if (by >= rowsTotal) break;
            const unsigned int* src = reinterpret_cast<const unsigned int*>(img + (size_t)reflect101(by - kEdge, h) * stride + base);
            const unsigned int w0 = __ldg(src), w1 = second ? __ldg(src + 1) : 0u;
            *reinterpret_cast<unsigned int*>(dst + (size_t)by * pitch) = __byte_perm(w0, w1, sel);
        }
        return;
    }
#pragma unroll 1
    for (int r = 0; r < PY_ROWS; ++r) {
        const int by = by0 + r;
        if (by >= rowsTotal) break;
        const unsigned char* src = img + (size_t)reflect101(by - kEdge, h) * stride;
        unsigned int word = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) word |= (unsigned int)__ldg(src + col[k]) << (8 * k);
        *reinterpret_cast<unsigned int*>(dst + (size_t)by * pitch) = word;
    }
}

// Level 0 for images whose rows are 16-byte aligned and a multiple of 16 wide (EuRoC: 752): the interior of a padded row is
// a straight copy, so an interior work item is 16 bytes x 16 rows (one 128-bit load and store per row); the frame left and
// right of it is dealt out in 4-byte groups that take pyramid_level0_kernel's two-word path with the reflected columns.
// Items of a band: nInt = w / 16 interior groups, then (pitch - w) / 4 frame groups (8 on the left, the rest on the right).
__global__ void __launch_bounds__(PY_THREADS)
pyramid_level0_wide_kernel(const unsigned char* __restrict__ images, int w, int h, int stride, size_t frameStride,
                           unsigned char* __restrict__ pyr, long long pyrFrameBytes, long long pyrOff, int pitch, int groups,
                           int nItems) {
    const int item = blockIdx.x * PY_THREADS + threadIdx.x;
    if (item >= nItems) return;
    const int band = item / groups, g = item - band * groups;
    const int nInt = w >> 4;
    const unsigned char* img = images + (size_t)blockIdx.y * frameStride;
    unsigned char* dstFrame = pyr + (size_t)blockIdx.y * pyrFrameBytes + pyrOff;
    const int by0 = band * PY_ROWS, rowsTotal = h + 2 * kEdge;
    if (g < nInt) {
        const int lx0 = 16 * g;
        unsigned char* dst = dstFrame + kPadLeft + lx0;
#pragma unroll 4
        for (int r = 0; r < PY_ROWS; ++r) {
            const int by = by0 + r;
            if (by >= rowsTotal) break;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(img + (size_t)reflect101(by - kEdge, h) * stride + lx0));
            *reinterpret_cast<uint4*>(dst + (size_t)by * pitch) = v;
        }
        return;
    }
    const int j = g - nInt;
    const int bx = j < kPadLeft / 4 ? 4 * j : w + 4 * j;      // left: columns 0 .. 31; right: from column 32 + w on
    const int lx0 = bx - kPadLeft;
    int col[4], cmin = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < 4; ++k) { col[k] = frame_column(lx0 + k, w); cmin = min(cmin, col[k]); }
    const int base = cmin & ~3;
    unsigned int sel = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) sel |= (unsigned int)(col[k] - base) << (4 * k);
    const bool second = base + 4 < w;
    unsigned char* dst = dstFrame + bx;
#pragma unroll 4
    for (int r = 0; r < PY_ROWS; ++r) {
        const int by = by0 + r;
        if (by >= rowsTotal) break;
        const unsigned int* src = reinterpret_cast<const unsigned int*>(img + (size_t)reflect101(by - kEdge, h) * stride + base);
        const unsigned int w0 = __ldg(src), w1 = second ? __ldg(src + 1) : 0u;
        *reinterpret_cast<unsigned int*>(dst + (size_t)by * pitch) = __byte_perm(w0, w1, sel);
    }
}

__global__ void __launch_bounds__(PY_THREADS)
pyramid_resize_kernel(unsigned char* __restrict__ pyr, long long pyrFrameBytes, long long srcOff, int srcPitch,
                      long long dstOff, int dstPitch, int dw, int dh, int groups, int nItems, const int* __restrict__ xofs,
                      const unsigned int* __restrict__ xcoef, const int* __restrict__ yofs, const short2* __restrict__ ycoef) {
    const int item = blockIdx.x * PY_THREADS + threadIdx.x;
    if (item >= nItems) return;
    const int band = item / groups, g = item - band * groups;
    const int bx = 4 * g, lx0 = bx - kPadLeft;
    unsigned char* frame = pyr + (size_t)blockIdx.y * pyrFrameBytes;
    const unsigned char* src0 = frame + srcOff + (size_t)kEdge * srcPitch + kPadLeft;   // source level pixel (0,0)
    unsigned char* dst = frame + dstOff + bx;

    // per-item column invariants
    unsigned int cf[4];
    int ofs[4], omin = 0x7fffffff, omax = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int dx = frame_column(lx0 + k, dw);
        ofs[k] = __ldg(xofs + dx);
        cf[k] = __ldg(xcoef + dx);                    // a0 | a1 << 16
        omin = min(omin, ofs[k]);
        omax = max(omax, ofs[k]);
    }
    const int by0 = band * PY_ROWS, rowsTotal = dh + 2 * kEdge;
    if (omax - omin <= 6) {
        // the 8 bytes from omin hold S[x0], S[x0+1] of all four pixels (any scale factor up to 2)
        const int base = omin & ~3, shift = (omin & 3) * 8;
        unsigned int sel01, sel23;
        {
            const unsigned int i0 = ofs[0] - omin, i1 = ofs[1] - omin, i2 = ofs[2] - omin, i3 = ofs[3] - omin;
            sel01 = i0 | ((i0 + 1) << 4) | (i1 << 8) | ((i1 + 1) << 12);
            sel23 = i2 | ((i2 + 1) << 4) | (i3 << 8) | ((i3 + 1) << 12);
        }
        // horizontal sums of one source row, low 4 bits dropped (the vertical pass uses T >> 4 only)
        auto hsum = [&](const unsigned char* row, unsigned int (&t)[4]) {
            const unsigned int* p = reinterpret_cast<const unsigned int*>(row + base);
            const unsigned int w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
            const unsigned int X = __funnelshift_r(w0, w1, shift), Y = __funnelshift_r(w1, w2, shift);
            const unsigned int v01 = __byte_perm(X, Y, sel01), v23 = __byte_perm(X, Y, sel23);
            t[0] = __dp2a_lo(cf[0], v01, 0u) >> 4;          // S[x0]*a0 + S[x0+1]*a1
            t[1] = __dp2a_hi(cf[1], v01, 0u) >> 4;
            t[2] = __dp2a_lo(cf[2], v23, 0u) >> 4;
            t[3] = __dp2a_hi(cf[3], v23, 0u) >> 4;
        };
        int keptRow = -0x40000000;
        unsigned int kept[4] = {0, 0, 0, 0};
#pragma unroll 2
        for (int r = 0; r < PY_ROWS; ++r) {
            const int by = by0 + r;
            if (by >= rowsTotal) break;
            const int dy = reflect101(by - kEdge, dh);
            const int sy0 = __ldg(yofs + dy);
            const short2 b = __ldg(ycoef + dy);
            const unsigned int B0 = (unsigned int)b.x, B1 = (unsigned int)b.y;
            const unsigned char* r0 = src0 + (size_t)sy0 * srcPitch;
            unsigned int t0[4], t1[4];
            if (sy0 == keptRow) {
#pragma unroll
                for (int k = 0; k < 4; ++k) t0[k] = kept[k];
            } else {
                hsum(r0, t0);
            }
            hsum(r0 + srcPitch, t1);
            keptRow = sy0 + 1;
            unsigned int s[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                kept[k] = t1[k];
                s[k] = py_vsum(t0[k], B0, t1[k], B1);    // <= 1023
            }
            const unsigned int q01 = __byte_perm(s[0], s[1], 0x5410) >> 2, q23 = __byte_perm(s[2], s[3], 0x5410) >> 2;
            *reinterpret_cast<unsigned int*>(dst + (size_t)by * dstPitch) = __byte_perm(q01, q23, 0x6420);
        }
        return;
    }
    // unusual scale factors (> 2): per-pixel path
#pragma unroll 1
    for (int r = 0; r < PY_ROWS; ++r) {
        const int by = by0 + r;
        if (by >= rowsTotal) break;
        const int dy = reflect101(by - kEdge, dh);
        const int sy0 = __ldg(yofs + dy);
        const short2 b = __ldg(ycoef + dy);
        const unsigned char* r0 = src0 + (size_t)sy0 * srcPitch;
        const unsigned char* r1 = r0 + srcPitch;      // row sh is the source's own frame when sy0 == sh-1 (b.y == 0 there)
        unsigned int word = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ax = (int)(cf[k] & 0xffffu), ay = (int)(cf[k] >> 16);
            const int t0 = (int)r0[ofs[k]] * ax + (int)r0[ofs[k] + 1] * ay;
            const int t1 = (int)r1[ofs[k]] * ax + (int)r1[ofs[k] + 1] * ay;
            const unsigned int v = (unsigned int)(((((int)b.x * (t0 >> 4)) >> 16) + (((int)b.y * (t1 >> 4)) >> 16) + 2) >> 2);
            word |= (v & 0xffu) << (8 * k);
        }
        *reinterpret_cast<unsigned int*>(dst + (size_t)by * dstPitch) = word;
    }
}

// ---- the resize kernel of the usual case (every 4-byte group of the level reads an 8-byte source window: scale <= 2) ----
// Same arithmetic as pyramid_resize_kernel; what changed is everything around it (22 -> 13 thread-instructions per pixel):
// the per-group column invariants and the per-row source offsets / vertical coefficients are host-built records (one or
// two 128-bit loads instead of table look-ups, reflections and 64-bit multiplies per row), and the only 64-bit address
// arithmetic left per row is one add per source row and one for the destination.
//   colTab[2g], colTab[2g+1] : {window base (bytes from the source pixel (0,0), multiple of 4), shift, sel01, sel23}, {cf[0..3]}
//   rowTab[by]               : {byte offset of source row sy0, of row sy0+1, b0, b1}   (by = padded row)
__global__ void __launch_bounds__(PY_THREADS)
pyramid_resize2_kernel(unsigned char* __restrict__ pyr, long long pyrFrameBytes, long long srcPix0, long long dstOff, int dstPitch,
                       int rowsTotal, int groups, int nItems, const uint4* __restrict__ colTab, const uint4* __restrict__ rowTab) {
    const int item = blockIdx.x * PY_THREADS + threadIdx.x;
    if (item >= nItems) return;
    const int band = item / groups, g = item - band * groups;
    unsigned char* frame = pyr + (size_t)blockIdx.y * pyrFrameBytes;
    const uint4 ca = __ldg(colTab + 2 * g), cb = __ldg(colTab + 2 * g + 1);
    const unsigned char* srcCol = frame + srcPix0 + ca.x;
    const int by0 = band * PY_ROWS;
    unsigned char* dst = frame + dstOff + 4 * g + (size_t)by0 * dstPitch;
    const uint4* rt = rowTab + by0;
    const int rows = min(PY_ROWS, rowsTotal - by0);
    const unsigned int shift = ca.y, sel01 = ca.z, sel23 = ca.w;
    auto hsum = [&](const unsigned char* row, unsigned int (&t)[4]) {
        const unsigned int* p = reinterpret_cast<const unsigned int*>(row);
        const unsigned int w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
        const unsigned int X = __funnelshift_r(w0, w1, shift), Y = __funnelshift_r(w1, w2, shift);
        const unsigned int v01 = __byte_perm(X, Y, sel01), v23 = __byte_perm(X, Y, sel23);
        t[0] = __dp2a_lo(cb.x, v01, 0u) >> 4;
        t[1] = __dp2a_hi(cb.y, v01, 0u) >> 4;
        t[2] = __dp2a_lo(cb.z, v23, 0u) >> 4;
        t[3] = __dp2a_hi(cb.w, v23, 0u) >> 4;
    };
    unsigned int keptOff = 0xffffffffu;
    unsigned int kept[4] = {0, 0, 0, 0};
#pragma unroll 2
    for (int r = 0; r < rows; ++r) {
        const uint4 rr = __ldg(rt + r);
        unsigned int t0[4], t1[4];
        if (rr.x == keptOff) {
#pragma unroll
            for (int k = 0; k < 4; ++k) t0[k] = kept[k];
        } else {
            hsum(srcCol + rr.x, t0);
        }
        hsum(srcCol + rr.y, t1);
        keptOff = rr.y;
        unsigned int s[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            kept[k] = t1[k];
            s[k] = py_vsum(t0[k], rr.z, t1[k], rr.w);    // <= 1023
        }
        const unsigned int q01 = __byte_perm(s[0], s[1], 0x5410) >> 2, q23 = __byte_perm(s[2], s[3], 0x5410) >> 2;
        *reinterpret_cast<unsigned int*>(dst) = __byte_perm(q01, q23, 0x6420);
        dst += dstPitch;
    }
}

// host: the records of one level (dst) resized from the level above it (src); returns false if a group needs more than the
// 8-byte window (scale factor > 2), in which case the level keeps the table-driven kernel
bool pyramid_level_plan(const LevelGeom& S, const LevelGeom& D, const int* xofs, const short2* xcoef, const int* yofs,
                        const short2* ycoef, std::vector<uint4>& col, std::vector<uint4>& row) {
    auto reflect = [](int i, int n) { if (i < 0) i = -i; if (i >= n) i = 2 * n - 2 - i; return i; };
    const int groups = D.pitch / 4;
    for (int g = 0; g < groups; ++g) {
        int ofs[4], omin = 0x7fffffff, omax = 0;
        unsigned int cf[4];
        for (int k = 0; k < 4; ++k) {
            const int dx = std::min(std::max(reflect(4 * g - kPadLeft + k, D.w), 0), D.w - 1);
            ofs[k] = xofs[dx];
            cf[k] = (unsigned int)(unsigned short)xcoef[dx].x | ((unsigned int)(unsigned short)xcoef[dx].y << 16);
            omin = std::min(omin, ofs[k]);
            omax = std::max(omax, ofs[k]);
        }
        if (omax - omin > 6) return false;
        const unsigned int i0 = ofs[0] - omin, i1 = ofs[1] - omin, i2 = ofs[2] - omin, i3 = ofs[3] - omin;
        col.push_back(make_uint4((unsigned int)(omin & ~3), (unsigned int)(omin & 3) * 8u, i0 | ((i0 + 1) << 4) | (i1 << 8) | ((i1 + 1) << 12),
                                 i2 | ((i2 + 1) << 4) | (i3 << 8) | ((i3 + 1) << 12)));
        col.push_back(make_uint4(cf[0], cf[1], cf[2], cf[3]));
    }
    for (int by = 0; by < D.h + 2 * kEdge; ++by) {
        const int dy = reflect(by - kEdge, D.h);
        const unsigned int off0 = (unsigned int)yofs[dy] * (unsigned int)S.pitch;
        row.push_back(make_uint4(off0, off0 + (unsigned int)S.pitch, (unsigned int)ycoef[dy].x, (unsigned int)ycoef[dy].y));
    }
    row.push_back(make_uint4(0, 0, 0, 0));     // spare entry: the staged kernel reads one record ahead
    return true;
}

// ---- the same resize with the source rows staged through shared memory by bulk-async copies (TMA engine) ------------
// pyramid_resize2_kernel issues 3 word loads per source row per thread straight to L1/L2: a warp's windows overlap, every
// source byte is requested ~3.6 times and the kernel sits at the latency x requests-in-flight limit (measured: 40 % fewer
// instructions changed nothing, 5.16 -> 5.18 ms).  Here a CTA owns a band of 16 destination rows over the full width; the
// source rows that band needs are CONTIGUOUS in the padded source level, so one cp.async.bulk (global -> shared, completion
// on an mbarrier) fetches them, each byte once, without holding registers, and the next band's copy is in flight while
// this one is computed (two buffers, persistent CTAs striding over (frame, band)).  The arithmetic reads shared memory.
//   bandTab[b] : {byte offset of the band's first source row inside the source level buffer, bytes, offset(sy0 = first), -}
__device__ __forceinline__ unsigned int py_smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }

// Level 0 for BATCHES of images with 16-byte aligned rows (EuRoC: 752): the copy engine moves the pixels.  A CTA owns a band
// of 16 padded rows: one cp.async.bulk per row brings the (reflected) source row to its place inside the padded row in
// shared memory, the threads fill the 2 x 32 frame bytes of every row from there, and since the padded rows of a level
// are contiguous in the arena ONE bulk store writes the whole band back.  No pixel passes through a register; persistent
// CTAs stride over (frame, band), several per SM so that one CTA's store drains under the others' loads.
constexpr int PY0_THREADS = 64;
__global__ void __launch_bounds__(PY0_THREADS)
pyramid_level0_bulk_kernel(const unsigned char* __restrict__ images, int w, int h, int stride, size_t frameStride,
                           unsigned char* __restrict__ pyr, long long pyrFrameBytes, long long pyrOff, int pitch, int nBands, int nTiles) {
    extern __shared__ __align__(128) unsigned char l0sm[];
    const int tid = threadIdx.x;
    const unsigned int bar = py_smem_u32(l0sm), buf = bar + 128;
    unsigned char* rowsSm = l0sm + 128;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int rowsTotal = h + 2 * kEdge, nb = (pitch - w) >> 2;      // frame groups of 4 bytes per row: 8 left, the rest right
    unsigned int parity = 0;
    for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
        const int frame = tile / nBands, band = tile - frame * nBands;
        const int by0 = band * PY_ROWS, rows = min(PY_ROWS, rowsTotal - by0);
        if (tid == 0) {
            const unsigned char* img = images + (size_t)frame * frameStride;
            // the previous band's store has read the buffer, and every thread's accesses to it are behind the CTA barrier below
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned int)(rows * w)) : "memory");
            for (int r = 0; r < rows; ++r) {
                const unsigned char* src = img + (size_t)reflect101(by0 + r - kEdge, h) * stride;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 buf + (unsigned int)(r * pitch + kPadLeft)),
                             "l"(src), "r"((unsigned int)w), "r"(bar)
                             : "memory");
            }
        }
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "PY0_WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra PY0_DONE_%=;\n"
            "bra PY0_WAIT_%=;\n"
            "PY0_DONE_%=:\n"
            "}\n" ::"r"(bar),
            "r"(parity)
            : "memory");
        parity ^= 1u;
        for (int i = tid; i < rows * nb; i += PY0_THREADS) {
            const int r = i / nb, j = i - r * nb;
            const int bx = j < kPadLeft / 4 ? 4 * j : w + 4 * j;      // left: bytes 0 .. 31 of the padded row; right: from byte 32 + w on
            unsigned char* row = rowsSm + r * pitch;
            unsigned int word = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) word |= (unsigned int)row[kPadLeft + frame_column(bx - kPadLeft + k, w)] << (8 * k);
            *reinterpret_cast<unsigned int*>(row + bx) = word;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the frame bytes, before the copy engine reads the band
        __syncthreads();
        if (tid == 0) {
            unsigned char* dst = pyr + (size_t)frame * pyrFrameBytes + pyrOff + (size_t)by0 * pitch;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(buf), "r"((unsigned int)(rows * pitch)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(512)
pyramid_resize3_kernel(unsigned char* __restrict__ pyr, long long pyrFrameBytes, long long srcOff, long long dstOff, int dstPitch,
                       int rowsTotal, int groups, int nBands, int nTiles, int bufBytes, int nBuf, const uint4* __restrict__ colTab,
                       const uint4* __restrict__ rowTab, const int4* __restrict__ bandTab) {
    extern __shared__ __align__(128) unsigned char psm[];
    const int tid = threadIdx.x;
    const unsigned int bar0 = py_smem_u32(psm);                    // two mbarriers at bytes 0 and 8, tiles from byte 128
    const unsigned int tile0 = bar0 + 128;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const bool active = tid < groups;
    // the copies are issued by a thread that has no pixels to compute when the block has one (block size = groups rounded up
    // to a warp): looking up the band record and issuing the copy is a chain of ~600 cycles that would otherwise delay one
    // working warp, and with it the whole CTA at the barrier, on every band
    const int issuerTid = (int)blockDim.x > groups ? (int)blockDim.x - 1 : 0;
    uint4 ca = make_uint4(0, 0, 0, 0), cb = ca;
    if (active) { ca = __ldg(colTab + 2 * tid); cb = __ldg(colTab + 2 * tid + 1); }
    const unsigned int shift = ca.y, sel01 = ca.z, sel23 = ca.w;
    auto issue = [&](int tile, int buf) {
        if (tid == issuerTid) {
            const int frame = tile / nBands, band = tile - frame * nBands;
            const int4 bt = __ldg(bandTab + band);
            const unsigned char* src = pyr + (size_t)frame * pyrFrameBytes + srcOff + bt.x;
            const unsigned int bar = bar0 + 8 * buf, dstS = tile0 + buf * bufBytes;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the buffer's last generic-proxy reads are behind the CTA barrier
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bt.y) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dstS), "l"(src),
                         "r"(bt.y), "r"(bar)
                         : "memory");
        }
    };
    auto hsum = [&](unsigned int addr, unsigned int (&t)[4]) {
        unsigned int w0, w1, w2;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(addr));
        asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(addr));
        asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(w2) : "r"(addr));
        const unsigned int X = __funnelshift_r(w0, w1, shift), Y = __funnelshift_r(w1, w2, shift);
        const unsigned int v01 = __byte_perm(X, Y, sel01), v23 = __byte_perm(X, Y, sel23);
        t[0] = __dp2a_lo(cb.x, v01, 0u) >> 4;
        t[1] = __dp2a_hi(cb.y, v01, 0u) >> 4;
        t[2] = __dp2a_lo(cb.z, v23, 0u) >> 4;
        t[3] = __dp2a_hi(cb.w, v23, 0u) >> 4;
    };
    // nBuf == 2: the next band's copy is in flight while this one is computed; nBuf == 1: one buffer, twice the CTAs per SM
    int tile = blockIdx.x;
    if (nBuf == 2 && tile < nTiles) issue(tile, 0);
    for (int it = 0; tile < nTiles; tile += gridDim.x, ++it) {
        const int buf = nBuf == 2 ? (it & 1) : 0;
        if (nBuf == 2) {
            if (tile + (int)gridDim.x < nTiles) issue(tile + gridDim.x, buf ^ 1);
        } else {
            issue(tile, 0);
        }
        {
            const unsigned int bar = bar0 + 8 * buf, parity = (unsigned int)(nBuf == 2 ? it >> 1 : it) & 1u;
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "PY_WAIT_%=:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                "@p bra PY_DONE_%=;\n"
                "bra PY_WAIT_%=;\n"
                "PY_DONE_%=:\n"
                "}\n" ::"r"(bar),
                "r"(parity)
                : "memory");
        }
        if (active) {
            const int frame = tile / nBands, band = tile - frame * nBands;
            const int by0 = band * PY_ROWS;
            const int4 bt = __ldg(bandTab + band);
            // shared-memory address of the window of source row sy0 = (offset in rowTab) + this
            const unsigned int colBase = tile0 + buf * bufBytes + kPadLeft + ca.x - (unsigned int)bt.z;
            unsigned char* dst = pyr + (size_t)frame * pyrFrameBytes + dstOff + 4 * tid + (size_t)by0 * dstPitch;
            const uint4* rt = rowTab + by0;
            const int rows = min(PY_ROWS, rowsTotal - by0);
            unsigned int keptOff = 0xffffffffu;
            unsigned int kept[4] = {0, 0, 0, 0};
            uint4 rrNext = __ldg(rt);
#pragma unroll 2
            for (int r = 0; r < rows; ++r) {
                const uint4 rr = rrNext;
                rrNext = __ldg(rt + r + 1);      // the next row's record is requested before this row is computed (rowTab has a spare entry)
                unsigned int t0[4], t1[4];
                if (rr.x == keptOff) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) t0[k] = kept[k];
                } else {
                    hsum(colBase + rr.x, t0);
                }
                hsum(colBase + rr.y, t1);
                keptOff = rr.y;
                unsigned int s[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    kept[k] = t1[k];
                    s[k] = py_vsum(t0[k], rr.z, t1[k], rr.w);    // <= 1023
                }
                const unsigned int q01 = __byte_perm(s[0], s[1], 0x5410) >> 2, q23 = __byte_perm(s[2], s[3], 0x5410) >> 2;
                *reinterpret_cast<unsigned int*>(dst) = __byte_perm(q01, q23, 0x6420);
                dst += dstPitch;
            }
        }
        __syncthreads();   // every thread is done with this buffer before its next refill is issued
    }
}

// host: band records of one level; returns the largest number of source bytes a band needs
int pyramid_band_plan(const LevelGeom& S, const LevelGeom& D, const int* yofs, std::vector<int4>& bands) {
    auto reflect = [](int i, int n) { if (i < 0) i = -i; if (i >= n) i = 2 * n - 2 - i; return i; };
    const int rowsTotal = D.h + 2 * kEdge;
    int maxBytes = 0;
    for (int by0 = 0; by0 < rowsTotal; by0 += PY_ROWS) {
        int lo = 0x7fffffff, hi = 0;
        for (int by = by0; by < std::min(by0 + PY_ROWS, rowsTotal); ++by) {
            const int sy = yofs[reflect(by - kEdge, D.h)];
            lo = std::min(lo, sy);
            hi = std::max(hi, sy);
        }
        const int bytes = (hi + 2 - lo) * S.pitch;
        bands.push_back(make_int4((kEdge + lo) * S.pitch, bytes, lo * S.pitch, 0));
        maxBytes = std::max(maxBytes, bytes);
    }
    return maxBytes;
}

// ---- small batches (the live loop: one frame or a stereo pair): all levels in ONE launch -----------------------------
// Eight dependent launches of a few microseconds each cost a single frame ~100 us, most of it launch gaps and kernel
// tails.  Here one grid of co-resident CTAs walks the levels; between levels every CTA passes a grid-wide barrier (arrive
// on a zeroed counter with a release fence, spin with acquire loads).  Level l's items (the work items of
// pyramid_level0_kernel / pyramid_resize2_kernel, frame-major) are dealt over the whole grid.  Data written earlier in
// the same launch is read with ld.global.cg (never through the non-coherent path).
constexpr int PYF_ROWS = 2;      // rows per work item of the fused kernel: latency matters here, not instruction count
struct PyFusedLevel {
    long long srcPix0, dstOff;   // source pixel (0,0) / destination buffer inside a frame's pyramid block
    int dstPitch, rowsTotal, groups, nItems;
    const uint4* colTab;
    const uint4* rowTab;
};
struct PyFusedArgs {
    const unsigned char* images;
    size_t frameStride;
    int w, h, stride, wordLoads, nFrames, nLevels;
    unsigned char* pyr;
    long long pyrFrameBytes;
    unsigned int* barrier;       // nLevels counters, zeroed before the launch
    PyFusedLevel lv[kMaxLevels];
};

__global__ void __launch_bounds__(PY_THREADS)
pyramid_fused_kernel(const __grid_constant__ PyFusedArgs A) {
    const int nThreads = gridDim.x * PY_THREADS, tid0 = blockIdx.x * PY_THREADS + threadIdx.x;
    {   // level 0: copy with the reflect-101 frame
        const PyFusedLevel& L = A.lv[0];
        for (int it = tid0; it < L.nItems * A.nFrames; it += nThreads) {
            const int frame = it / L.nItems, item = it - frame * L.nItems;
            const int band = item / L.groups, g = item - band * L.groups;
            const int bx = 4 * g, lx0 = bx - kPadLeft;
            const unsigned char* img = A.images + (size_t)frame * A.frameStride;
            unsigned char* dst = A.pyr + (size_t)frame * A.pyrFrameBytes + L.dstOff + bx;
            int col[4], cmin = 0x7fffffff;
#pragma unroll
            for (int k = 0; k < 4; ++k) { col[k] = frame_column(lx0 + k, A.w); cmin = min(cmin, col[k]); }
            const int base = cmin & ~3;
            unsigned int sel = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) sel |= (unsigned int)(col[k] - base) << (4 * k);
            const int by0 = band * PYF_ROWS;
            const bool second = base + 4 < ((A.w + 3) & ~3);
            for (int r = 0; r < PYF_ROWS; ++r) {
                const int by = by0 + r;
                if (by >= L.rowsTotal) break;
                const unsigned char* src = img + (size_t)reflect101(by - kEdge, A.h) * A.stride;
                unsigned int word;
                if (A.wordLoads) {
                    const unsigned int* sw = reinterpret_cast<const unsigned int*>(src + base);
                    word = __byte_perm(__ldg(sw), second ? __ldg(sw + 1) : 0u, sel);
                } else {
                    word = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) word |= (unsigned int)__ldg(src + col[k]) << (8 * k);
                }
                *reinterpret_cast<unsigned int*>(dst + (size_t)by * L.dstPitch) = word;
            }
        }
    }
    for (int l = 1; l < A.nLevels; ++l) {
        // grid barrier: level l-1 is complete (and visible) before anyone reads it
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(A.barrier + l, 1u);
            unsigned int seen;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(A.barrier + l) : "memory");
            } while (seen < gridDim.x);
        }
        __syncthreads();
        const PyFusedLevel& L = A.lv[l];
        for (int it = tid0; it < L.nItems * A.nFrames; it += nThreads) {
            const int frame = it / L.nItems, item = it - frame * L.nItems;
            const int band = item / L.groups, g = item - band * L.groups;
            unsigned char* fb = A.pyr + (size_t)frame * A.pyrFrameBytes;
            const uint4 ca = __ldg(L.colTab + 2 * g), cb = __ldg(L.colTab + 2 * g + 1);
            const unsigned char* srcCol = fb + L.srcPix0 + ca.x;
            const int by0 = band * PYF_ROWS;
            unsigned char* dst = fb + L.dstOff + 4 * g + (size_t)by0 * L.dstPitch;
            const uint4* rt = L.rowTab + by0;
            const int rows = min(PYF_ROWS, L.rowsTotal - by0);
            const unsigned int shift = ca.y, sel01 = ca.z, sel23 = ca.w;
            auto hsum = [&](const unsigned char* row, unsigned int (&t)[4]) {
                const unsigned int* q = reinterpret_cast<const unsigned int*>(row);
                const unsigned int w0 = __ldcg(q), w1 = __ldcg(q + 1), w2 = __ldcg(q + 2);
                const unsigned int X = __funnelshift_r(w0, w1, shift), Y = __funnelshift_r(w1, w2, shift);
                const unsigned int v01 = __byte_perm(X, Y, sel01), v23 = __byte_perm(X, Y, sel23);
                t[0] = __dp2a_lo(cb.x, v01, 0u) >> 4;
                t[1] = __dp2a_hi(cb.y, v01, 0u) >> 4;
                t[2] = __dp2a_lo(cb.z, v23, 0u) >> 4;
                t[3] = __dp2a_hi(cb.w, v23, 0u) >> 4;
            };
            unsigned int keptOff = 0xffffffffu;
            unsigned int kept[4] = {0, 0, 0, 0};
#pragma unroll 2
            for (int r = 0; r < rows; ++r) {
                const uint4 rr = __ldg(rt + r);
                unsigned int t0[4], t1[4];
                if (rr.x == keptOff) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) t0[k] = kept[k];
                } else {
                    hsum(srcCol + rr.x, t0);
                }
                hsum(srcCol + rr.y, t1);
                keptOff = rr.y;
                unsigned int sv[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    kept[k] = t1[k];
                    sv[k] = py_vsum(t0[k], rr.z, t1[k], rr.w);
                }
                const unsigned int q01 = __byte_perm(sv[0], sv[1], 0x5410) >> 2, q23 = __byte_perm(sv[2], sv[3], 0x5410) >> 2;
                *reinterpret_cast<unsigned int*>(dst) = __byte_perm(q01, q23, 0x6420);
                dst += L.dstPitch;
            }
        }
    }
}

// co-resident CTAs the fused kernel may use on the current device (0: unknown)
static int pyramid_fused_capacity() {
    static thread_local int cached = -1, cachedDev = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cached < 0 || cachedDev != dev) {
        int nSm = 0, perSm = 0;
        if (cudaDeviceGetAttribute(&nSm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, pyramid_fused_kernel, PY_THREADS, 0) != cudaSuccess) return 0;
        cached = nSm * perSm;
        cachedDev = dev;
    }
    return cached;
}

// resident CTAs of the staged kernel on the current device for a level's block size and shared memory (0: does not fit)
int pyramid_bulk_buffers() {
    static const int n = getenv("ORBB_PYR_NBUF") ? std::max(1, std::min(2, atoi(getenv("ORBB_PYR_NBUF")))) : 2;   // tuning aid
    return n;
}

int pyramid_bulk_ctas(int groups, int bufBytes) {
    const int threads = (groups + 31) / 32 * 32, smem = 128 + pyramid_bulk_buffers() * bufBytes;
    if (threads > 512 || smem > 200 * 1024) return 0;
    int dev = 0, nSm = 0, perSm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaFuncSetAttribute(pyramid_resize3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&nSm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, pyramid_resize3_kernel, threads, smem) != cudaSuccess) return 0;
    return nSm * perSm;
}

int launch_pyramid(const ExtractParams& P, const unsigned char* dImages, int width, int height, int stride,
                   size_t frameStride, cudaStream_t st, int* launches) {
    static const bool noFuse = getenv("ORBB_PYR_NOFUSE") != nullptr;    // A/B aid: one launch per level for small batches too
    bool allFast = P.pyBarrier != nullptr && !noFuse && P.nFrames < P.pyBulkMinFrames;
    for (int l = 1; l < P.nLevels; ++l) allFast = allFast && P.lv[l].pyFast;
    if (allFast) {
        const int cap = pyramid_fused_capacity();
        if (cap > 0) {
            PyFusedArgs A;
            A.images = dImages;
            A.frameStride = frameStride;
            A.w = width; A.h = height; A.stride = stride;
            A.wordLoads = ((((size_t)dImages) | (size_t)stride | frameStride) & 3) == 0 ? 1 : 0;
            A.nFrames = P.nFrames;
            A.nLevels = P.nLevels;
            A.pyr = P.pyr;
            A.pyrFrameBytes = P.pyrFrameBytes;
            A.barrier = P.pyBarrier;
            int maxItems = 0;
            for (int l = 0; l < P.nLevels; ++l) {
                const LevelGeom& D = P.lv[l];
                PyFusedLevel& F = A.lv[l];
                F.dstOff = D.pyrOff;
                F.dstPitch = D.pitch;
                F.rowsTotal = D.h + 2 * kEdge;
                F.groups = D.pitch / 4;
                F.nItems = F.groups * ceil_div(F.rowsTotal, PYF_ROWS);
                F.srcPix0 = l ? P.lv[l - 1].pyrOff + (long long)kEdge * P.lv[l - 1].pitch + kPadLeft : 0;
                F.colTab = l ? P.pyColTab + D.pyCol : nullptr;
                F.rowTab = l ? P.pyRowTab + D.pyRow : nullptr;
                maxItems = std::max(maxItems, F.nItems * P.nFrames);
            }
            const int grid = std::max(1, std::min(cap, ceil_div(maxItems, PY_THREADS)));
            ORB_CUDA(cudaMemsetAsync(P.pyBarrier, 0, sizeof(unsigned int) * kMaxLevels, st));
            pyramid_fused_kernel<<<grid, PY_THREADS, 0, st>>>(A);
            ++*launches;
            ORB_CUDA(cudaGetLastError());
            return ORB_OK;
        }
    }
    {
        const LevelGeom& L = P.lv[0];
        const int groups = L.pitch / 4, nItems = groups * ceil_div(L.h + 2 * kEdge, PY_ROWS);
        const int wordLoads = ((((size_t)dImages) | (size_t)stride | frameStride) & 3) == 0 ? 1 : 0;
        static const bool noWide = getenv("ORBB_PYR_NOWIDE") != nullptr;     // A/B aid
        static const bool noBulk0 = getenv("ORBB_PYR_NOBULK0") != nullptr;   // A/B aid
        const bool aligned16 = ((((size_t)dImages) | (size_t)stride | frameStride | (size_t)width) & 15) == 0 && (L.pitch & 15) == 0;
        const int l0Smem = 128 + PY_ROWS * L.pitch;
        if (!noWide && !noBulk0 && aligned16 && P.nFrames >= P.pyBulkMinFrames && ((L.pitch - width) & 3) == 0 && kPadLeft % 16 == 0 &&
            l0Smem <= 48 * 1024) {
            static thread_local int cDev = -1, cCtas = 0;
            int dev = 0;
            ORB_CUDA(cudaGetDevice(&dev));
            if (cDev != dev) {
                int nSm = 0;
                ORB_CUDA(cudaDeviceGetAttribute(&nSm, cudaDevAttrMultiProcessorCount, dev));
                cCtas = nSm * (getenv("ORBB_PYR_L0CTAS") ? std::max(1, atoi(getenv("ORBB_PYR_L0CTAS"))) : 6);   // tuning aid (4: 3.46, 6: 3.34, 8: 3.37, 12: 3.38, 16: 3.37 ms for the whole pyramid of 4096 frames)
                cDev = dev;
            }
            const int nBands = ceil_div(L.h + 2 * kEdge, PY_ROWS);
            const long long nTiles = (long long)nBands * P.nFrames;
            const int gridB = (int)std::min<long long>(nTiles, cCtas);
            pyramid_level0_bulk_kernel<<<gridB, PY0_THREADS, l0Smem, st>>>(dImages, width, height, stride, frameStride, P.pyr, P.pyrFrameBytes,
                                                                          L.pyrOff, L.pitch, nBands, (int)nTiles);
        } else if (!noWide && aligned16) {
            const int groups16 = width / 16 + (L.pitch - width) / 4, nItems16 = groups16 * ceil_div(L.h + 2 * kEdge, PY_ROWS);
            dim3 grid16(ceil_div(nItems16, PY_THREADS), P.nFrames);
            pyramid_level0_wide_kernel<<<grid16, PY_THREADS, 0, st>>>(dImages, width, height, stride, frameStride, P.pyr,
                                                                      P.pyrFrameBytes, L.pyrOff, L.pitch, groups16, nItems16);
        } else {
            dim3 grid(ceil_div(nItems, PY_THREADS), P.nFrames);
            pyramid_level0_kernel<<<grid, PY_THREADS, 0, st>>>(dImages, width, height, stride, frameStride, P.pyr, P.pyrFrameBytes,
                                                               L.pyrOff, L.pitch, groups, nItems, wordLoads);
        }
        ++*launches;
    }
    for (int l = 1; l < P.nLevels; ++l) {
        const LevelGeom& S = P.lv[l - 1];
        const LevelGeom& D = P.lv[l];
        const int groups = D.pitch / 4, nItems = groups * ceil_div(D.h + 2 * kEdge, PY_ROWS);
        dim3 grid(ceil_div(nItems, PY_THREADS), P.nFrames);
        if (D.pyFast && D.pyBulk && P.nFrames >= P.pyBulkMinFrames) {
            const int threads = (groups + 31) / 32 * 32, nBands = ceil_div(D.h + 2 * kEdge, PY_ROWS);
            const int nBuf = pyramid_bulk_buffers(), smem = 128 + nBuf * D.pyBufBytes;
            const long long nTiles = (long long)nBands * P.nFrames;
            const int gridX = (int)std::min<long long>(nTiles, (long long)D.pyBulkCtas);
            pyramid_resize3_kernel<<<gridX, threads, smem, st>>>(P.pyr, P.pyrFrameBytes, S.pyrOff, D.pyrOff, D.pitch, D.h + 2 * kEdge, groups,
                                                                 nBands, (int)nTiles, D.pyBufBytes, nBuf, P.pyColTab + D.pyCol, P.pyRowTab + D.pyRow,
                                                                 P.pyBandTab + D.pyBand);
            ++*launches;
            continue;
        }
        if (D.pyFast) {
            pyramid_resize2_kernel<<<grid, PY_THREADS, 0, st>>>(P.pyr, P.pyrFrameBytes, S.pyrOff + (long long)kEdge * S.pitch + kPadLeft,
                                                                D.pyrOff, D.pitch, D.h + 2 * kEdge, groups, nItems, P.pyColTab + D.pyCol,
                                                                P.pyRowTab + D.pyRow);
            ++*launches;
            continue;
        }
        pyramid_resize_kernel<<<grid, PY_THREADS, 0, st>>>(P.pyr, P.pyrFrameBytes, S.pyrOff, S.pitch, D.pyrOff, D.pitch, D.w,
                                                           D.h, groups, nItems, P.tabOfs + D.xTab,
                                                           reinterpret_cast<const unsigned int*>(P.tabCoef + D.xTab),
                                                           P.tabOfs + D.yTab, P.tabCoef + D.yTab);
        ++*launches;
    }
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orbb
