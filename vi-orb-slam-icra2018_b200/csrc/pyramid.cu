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
// Shape (the first version was issue-bound at ~45 instructions per pixel): a thread owns 4 horizontally adjacent output
// bytes (one aligned 32-bit store) and walks down 8 output rows, so the per-column table entries, byte offsets and
// funnel-shift amounts are loop invariants.  Per source row it loads three aligned words, funnel-shifts each pixel's
// two source bytes into place and forms the horizontal sum with ONE IDP.2A (16-bit coefficient pair x two bytes).
// Reading S[x+1] / row y+1 one past the level is harmless: the table's coefficient there is 0 and the source level
// has its own frame.  Threads that touch the left/right frame take the per-pixel reflected path.
#include "extractor.h"

namespace orbb {

constexpr int PY_ROWS = 16;   // output rows per thread
constexpr int PY_TY = 4;      // row bands per CTA

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

__global__ void __launch_bounds__(32 * PY_TY)
pyramid_level0_kernel(const unsigned char* __restrict__ images, int w, int h, int stride, size_t frameStride,
                      unsigned char* __restrict__ pyr, long long pyrFrameBytes, long long pyrOff, int pitch) {
    const int bx = (blockIdx.x * 32 + threadIdx.x) * 4;
    if (bx >= pitch) return;
    const int lx0 = bx - kPadLeft;
    const unsigned char* img = images + (size_t)blockIdx.z * frameStride;
    unsigned char* dst = pyr + (size_t)blockIdx.z * pyrFrameBytes + pyrOff + bx;
    const bool wordCopy = lx0 >= 0 && lx0 + 3 < w && (((size_t)img | (size_t)stride) & 3) == 0;
    const int by0 = (blockIdx.y * PY_TY + threadIdx.y) * PY_ROWS;
#pragma unroll
    for (int r = 0; r < PY_ROWS; ++r) {
        const int by = by0 + r;
        if (by >= h + 2 * kEdge) break;
        const unsigned char* src = img + (size_t)reflect101(by - kEdge, h) * stride;
        unsigned int word;
        if (wordCopy) {
            word = __ldg(reinterpret_cast<const unsigned int*>(src + lx0));
        } else {
            word = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int lx = lx0 + k;
                unsigned int v = 0;
                if (lx >= -kEdge && lx < w + kEdge) v = __ldg(src + reflect101(lx, w));
                word |= v << (8 * k);
            }
        }
        *reinterpret_cast<unsigned int*>(dst + (size_t)by * pitch) = word;
    }
}

__global__ void __launch_bounds__(32 * PY_TY)
pyramid_resize_kernel(unsigned char* __restrict__ pyr, long long pyrFrameBytes, long long srcOff, int srcPitch, int sw,
                      int sh, long long dstOff, int dstPitch, int dw, int dh, const int* __restrict__ xofs,
                      const unsigned int* __restrict__ xcoef, const int* __restrict__ yofs, const short2* __restrict__ ycoef) {
    const int bx = (blockIdx.x * 32 + threadIdx.x) * 4;
    if (bx >= dstPitch) return;
    const int lx0 = bx - kPadLeft;
    unsigned char* frame = pyr + (size_t)blockIdx.z * pyrFrameBytes;
    const unsigned char* src0 = frame + srcOff + (size_t)kEdge * srcPitch + kPadLeft;   // source level pixel (0,0)
    unsigned char* dst = frame + dstOff + bx;

    // per-column invariants
    unsigned int cf[4];
    int shiftBits[4], rel[4];
    bool hiWin[4];
    int base = 0;
    bool fast = lx0 >= 0 && lx0 + 3 < dw;
    if (fast) {
        const int o0 = __ldg(xofs + lx0);
        base = o0 & ~3;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            rel[k] = __ldg(xofs + lx0 + k) - base;
            cf[k] = __ldg(xcoef + lx0 + k);           // a0 | a1 << 16
            hiWin[k] = rel[k] >= 4;
            shiftBits[k] = (rel[k] & 3) * 8;
        }
        fast = rel[3] <= 7;                           // both source bytes of every pixel inside the 12-byte window
    }
    const int by0 = (blockIdx.y * PY_TY + threadIdx.y) * PY_ROWS;
    const int rowsTotal = dh + 2 * kEdge;
    if (fast) {
        // the common case; kept in its own loop so that nothing of the per-pixel border path is hoisted into it
        // Consecutive output rows usually step one source row (scale 1.2: five times out of six), so the lower source
        // row's horizontal sums become the next output row's upper ones: kept in registers, chosen by a warp-uniform test.
        int keptRow = -0x40000000, kept[4] = {0, 0, 0, 0};
        auto hsum = [&](const unsigned char* row, int (&t)[4]) {
            const unsigned int* p = reinterpret_cast<const unsigned int*>(row + base);
            const unsigned int w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned int v = __funnelshift_r(hiWin[k] ? w1 : w0, hiWin[k] ? w2 : w1, shiftBits[k]);
                t[k] = (int)__dp2a_lo(cf[k], v, 0u);              // S[x0]*a0 + S[x0+1]*a1
            }
        };
#pragma unroll 2
        for (int r = 0; r < PY_ROWS; ++r) {
            const int by = by0 + r;
            if (by >= rowsTotal) break;
            const int dy = reflect101(by - kEdge, dh);
            const int sy0 = __ldg(yofs + dy);
            const short2 b = __ldg(ycoef + dy);
            const unsigned char* r0 = src0 + (size_t)sy0 * srcPitch;
            int t0[4], t1[4];
            if (sy0 == keptRow) {
#pragma unroll
                for (int k = 0; k < 4; ++k) t0[k] = kept[k];
            } else {
                hsum(r0, t0);
            }
            hsum(r0 + srcPitch, t1);
            keptRow = sy0 + 1;
            unsigned int word = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                kept[k] = t1[k];
                const unsigned int v = (unsigned int)(((((int)b.x * (t0[k] >> 4)) >> 16) + (((int)b.y * (t1[k] >> 4)) >> 16) + 2) >> 2);
                word |= (v & 0xffu) << (8 * k);
            }
            *reinterpret_cast<unsigned int*>(dst + (size_t)by * dstPitch) = word;
        }
        return;
    }
    // threads that touch the left/right frame (or an unusual scale factor): per-pixel reflected coordinates
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
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
            const int lx = lx0 + k;
            unsigned int v = 0;
            if (lx >= -kEdge && lx < dw + kEdge) {
                const int dx = reflect101(lx, dw);
                const int x0 = __ldg(xofs + dx);
                const unsigned int a = __ldg(xcoef + dx);
                const int ax = (int)(a & 0xffffu), ay = (int)(a >> 16);
                const int t0 = (int)r0[x0] * ax + (int)r0[x0 + 1] * ay;
                const int t1 = (int)r1[x0] * ax + (int)r1[x0 + 1] * ay;
                v = (unsigned int)(((((int)b.x * (t0 >> 4)) >> 16) + (((int)b.y * (t1 >> 4)) >> 16) + 2) >> 2);
            }
            word |= (v & 0xffu) << (8 * k);
        }
        *reinterpret_cast<unsigned int*>(dst + (size_t)by * dstPitch) = word;
    }
}

int launch_pyramid(const ExtractParams& P, const unsigned char* dImages, int width, int height, int stride,
                   size_t frameStride, cudaStream_t st, int* launches) {
    const dim3 block(32, PY_TY);
    {
        const LevelGeom& L = P.lv[0];
        dim3 grid(ceil_div(L.pitch, 128), ceil_div(L.h + 2 * kEdge, PY_TY * PY_ROWS), P.nFrames);
        pyramid_level0_kernel<<<grid, block, 0, st>>>(dImages, width, height, stride, frameStride, P.pyr, P.pyrFrameBytes,
                                                      L.pyrOff, L.pitch);
        ++*launches;
    }
    for (int l = 1; l < P.nLevels; ++l) {
        const LevelGeom& S = P.lv[l - 1];
        const LevelGeom& D = P.lv[l];
        dim3 grid(ceil_div(D.pitch, 128), ceil_div(D.h + 2 * kEdge, PY_TY * PY_ROWS), P.nFrames);
        pyramid_resize_kernel<<<grid, block, 0, st>>>(P.pyr, P.pyrFrameBytes, S.pyrOff, S.pitch, S.w, S.h, D.pyrOff,
                                                      D.pitch, D.w, D.h, P.tabOfs + D.xTab,
                                                      reinterpret_cast<const unsigned int*>(P.tabCoef + D.xTab),
                                                      P.tabOfs + D.yTab, P.tabCoef + D.yTab);
        ++*launches;
    }
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orbb
