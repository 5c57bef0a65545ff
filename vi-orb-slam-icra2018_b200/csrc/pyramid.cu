// Image pyramid (north-star kernel 1): ORBextractor::ComputePyramid, ORBextractor.cc:1128-1153.
//
//   level 0      = copyMakeBorder(image, 19, BORDER_REFLECT_101)
//   level l >= 1 = resize(level l-1, INTER_LINEAR) then copyMakeBorder(.., 19, BORDER_REFLECT_101 | ISOLATED)
//
// OpenCV's 8-bit INTER_LINEAR is fixed point: per axis a source index and an 11-bit coefficient pair (tables built on
// the host with the exact float recipe, see extractor.cu), horizontal sums kept at 19 bits, vertical pass
// (((b0*(T0>>4))>>16) + ((b1*(T1>>4))>>16) + 2) >> 2.  The kernels write the 19-px reflect-101 frame in the same pass
// by evaluating the reflected interior coordinate, so the border costs no extra launch and no read-after-write.
// Each thread produces 4 horizontally adjacent bytes and stores one 32-bit word (rows are 16-byte aligned).
#include "extractor.h"

namespace orbb {

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

__global__ void __launch_bounds__(256)
pyramid_level0_kernel(const unsigned char* __restrict__ images, int w, int h, int stride, size_t frameStride,
                      unsigned char* __restrict__ pyr, long long pyrFrameBytes, long long pyrOff, int pitch) {
    const int bx = (blockIdx.x * 64 + threadIdx.x) * 4;
    const int by = blockIdx.y * 4 + threadIdx.y;
    if (bx >= pitch || by >= h + 2 * kEdge) return;
    const unsigned char* src = images + (size_t)blockIdx.z * frameStride + (size_t)reflect101(by - kEdge, h) * stride;
    unsigned int word = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int lx = bx + k - kPadLeft;
        unsigned int v = 0;
        if (lx >= -kEdge && lx < w + kEdge) v = __ldg(src + reflect101(lx, w));
        word |= v << (8 * k);
    }
    unsigned char* dst = pyr + (size_t)blockIdx.z * pyrFrameBytes + pyrOff + (size_t)by * pitch + bx;
    *reinterpret_cast<unsigned int*>(dst) = word;
}

__global__ void __launch_bounds__(256)
pyramid_resize_kernel(unsigned char* __restrict__ pyr, long long pyrFrameBytes, long long srcOff, int srcPitch, int sw,
                      int sh, long long dstOff, int dstPitch, int dw, int dh, const int* __restrict__ xofs,
                      const short2* __restrict__ xcoef, const int* __restrict__ yofs, const short2* __restrict__ ycoef) {
    const int bx = (blockIdx.x * 64 + threadIdx.x) * 4;
    const int by = blockIdx.y * 4 + threadIdx.y;
    if (bx >= dstPitch || by >= dh + 2 * kEdge) return;
    unsigned char* frame = pyr + (size_t)blockIdx.z * pyrFrameBytes;
    const int dy = reflect101(by - kEdge, dh);
    const int sy0 = __ldg(yofs + dy);
    const int sy1 = min(sy0 + 1, sh - 1);
    const short2 b = __ldg(ycoef + dy);
    const unsigned char* r0 = frame + srcOff + (size_t)(sy0 + kEdge) * srcPitch + kPadLeft;
    const unsigned char* r1 = frame + srcOff + (size_t)(sy1 + kEdge) * srcPitch + kPadLeft;
    unsigned int word = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int lx = bx + k - kPadLeft;
        unsigned int v = 0;
        if (lx >= -kEdge && lx < dw + kEdge) {
            const int dx = reflect101(lx, dw);
            const int x0 = __ldg(xofs + dx);
            const int x1 = min(x0 + 1, sw - 1);
            const short2 a = __ldg(xcoef + dx);
            const int t0 = (int)r0[x0] * a.x + (int)r0[x1] * a.y;
            const int t1 = (int)r1[x0] * a.x + (int)r1[x1] * a.y;
            v = (unsigned int)(((((int)b.x * (t0 >> 4)) >> 16) + (((int)b.y * (t1 >> 4)) >> 16) + 2) >> 2);
        }
        word |= (v & 0xffu) << (8 * k);
    }
    *reinterpret_cast<unsigned int*>(frame + dstOff + (size_t)by * dstPitch + bx) = word;
}

int launch_pyramid(const ExtractParams& P, const unsigned char* dImages, int width, int height, int stride,
                   size_t frameStride, cudaStream_t st, int* launches) {
    const dim3 block(64, 4);
    {
        const LevelGeom& L = P.lv[0];
        dim3 grid(ceil_div(L.pitch, 256), ceil_div(L.h + 2 * kEdge, 4), P.nFrames);
        pyramid_level0_kernel<<<grid, block, 0, st>>>(dImages, width, height, stride, frameStride, P.pyr, P.pyrFrameBytes,
                                                      L.pyrOff, L.pitch);
        ++*launches;
    }
    for (int l = 1; l < P.nLevels; ++l) {
        const LevelGeom& S = P.lv[l - 1];
        const LevelGeom& D = P.lv[l];
        dim3 grid(ceil_div(D.pitch, 256), ceil_div(D.h + 2 * kEdge, 4), P.nFrames);
        pyramid_resize_kernel<<<grid, block, 0, st>>>(P.pyr, P.pyrFrameBytes, S.pyrOff, S.pitch, S.w, S.h, D.pyrOff,
                                                      D.pitch, D.w, D.h, P.tabOfs + D.xTab, P.tabCoef + D.xTab,
                                                      P.tabOfs + D.yTab, P.tabCoef + D.yTab);
        ++*launches;
    }
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orbb
