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
//   * the vertical pass is two multiply-high per pixel: hi32((T & ~15) * (b << 12)) == (b * (T >> 4)) >> 16.
// Reading S[x+1] / row y+1 one past the level is harmless: the table's coefficient there is 0 and the source level
// has its own frame.
#include "extractor.h"

namespace orbb {

constexpr int PY_ROWS = 16;       // output rows per work item
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
            t[0] = __dp2a_lo(cf[0], v01, 0u) & ~15u;          // S[x0]*a0 + S[x0+1]*a1
            t[1] = __dp2a_hi(cf[1], v01, 0u) & ~15u;
            t[2] = __dp2a_lo(cf[2], v23, 0u) & ~15u;
            t[3] = __dp2a_hi(cf[3], v23, 0u) & ~15u;
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
            const unsigned int B0 = (unsigned int)b.x << 12, B1 = (unsigned int)b.y << 12;
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
                s[k] = __umulhi(t0[k], B0) + __umulhi(t1[k], B1) + 2u;    // <= 1023
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

int launch_pyramid(const ExtractParams& P, const unsigned char* dImages, int width, int height, int stride,
                   size_t frameStride, cudaStream_t st, int* launches) {
    {
        const LevelGeom& L = P.lv[0];
        const int groups = L.pitch / 4, nItems = groups * ceil_div(L.h + 2 * kEdge, PY_ROWS);
        const int wordLoads = ((((size_t)dImages) | (size_t)stride | frameStride) & 3) == 0 ? 1 : 0;
        dim3 grid(ceil_div(nItems, PY_THREADS), P.nFrames);
        pyramid_level0_kernel<<<grid, PY_THREADS, 0, st>>>(dImages, width, height, stride, frameStride, P.pyr, P.pyrFrameBytes,
                                                           L.pyrOff, L.pitch, groups, nItems, wordLoads);
        ++*launches;
    }
    for (int l = 1; l < P.nLevels; ++l) {
        const LevelGeom& S = P.lv[l - 1];
        const LevelGeom& D = P.lv[l];
        const int groups = D.pitch / 4, nItems = groups * ceil_div(D.h + 2 * kEdge, PY_ROWS);
        dim3 grid(ceil_div(nItems, PY_THREADS), P.nFrames);
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
