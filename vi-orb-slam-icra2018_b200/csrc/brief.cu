// Rotated-BRIEF descriptors + final keypoint assembly (north-star kernel 6):
//   computeOrbDescriptor / computeDescriptors   ORBextractor.cc:110-149, 1036-1043 (pattern table :152-410)
//   coordinate rescale and concatenation         ORBextractor.cc:1112-1121
//
// One warp per keypoint: lane i produces descriptor byte i from pattern pairs 8i..8i+7.  The rotation is
// a = cosf(angle*factorPI), b = sinf(..) with glibc's cosf/sinf reproduced bit for bit (glibc_sincosf.h); sample
// coordinates are cvRound(x*b + y*a), cvRound(x*a - y*b) with every float op rounded separately (the reference's
// source-level order, no FMA) and round-half-even conversion.  Samples come from the BLURRED level.
// Keypoints of a frame are written level 0..nLevels-1, within a level in quadtree list order; pt is scaled by
// mvScaleFactor[level] only after the descriptor is taken, as the reference does.
#include "extractor.h"
#include "glibc_sincosf.h"

namespace orbb {

__constant__ char4 cPattern[256];   // (x0, y0, x1, y1) per test pair

static const signed char kPatternHost[1024] = {
#include "brief_pattern.inc"
};

int upload_brief_pattern() {
    ORB_CUDA(cudaMemcpyToSymbol(cPattern, kPatternHost, sizeof kPatternHost));
    return ORB_OK;
}

constexpr int BR_WARPS = 8;

__global__ void __launch_bounds__(BR_WARPS * 32)
brief_kernel(const __grid_constant__ ExtractParams P, orb_keypoint* __restrict__ kps, unsigned char* __restrict__ desc,
             int* __restrict__ nOut) {
    __shared__ char4 pat[256];
    const int frame = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pat[tid] = cPattern[(tid & 31) * 8 + (tid >> 5)];   // transposed: pat[j*32 + lane], conflict-free
    __syncthreads();
    const int* selCount = P.selCount + (size_t)frame * P.nLevels;
    const int g = blockIdx.x * BR_WARPS + warp;
    int level = -1, idx = 0, acc = 0;
    for (int l = 0; l < P.nLevels; ++l) {
        const int c = selCount[l];
        if (level < 0 && g < acc + c) { level = l; idx = g - acc; }
        acc += c;
    }
    if (blockIdx.x == 0 && tid == 0) nOut[frame] = acc;
    if (level < 0 || g >= P.outCapacity) return;
    const LevelGeom& L = P.lv[level];
    const SelKey k = P.sel[(size_t)frame * P.selPerFrame + L.selBase + idx];

    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    float a, b;
    sincosf_glibc(__fmul_rn(k.angle, factorPI), &b, &a);   // a = cos, b = sin
    const unsigned char* center = P.blur + (size_t)frame * P.blurFrameBytes + L.blurOff + (size_t)(int)k.y * L.bpitch + (int)k.x;
    unsigned int val = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const char4 p = pat[j * 32 + lane];
        const float x0 = (float)p.x, y0 = (float)p.y, x1 = (float)p.z, y1 = (float)p.w;
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int t0 = center[r0 * L.bpitch + c0], t1 = center[r1 * L.bpitch + c1];
        val |= (unsigned int)(t0 < t1) << j;
    }
    desc[((size_t)frame * P.outCapacity + g) * 32 + lane] = (unsigned char)val;
    if (lane == 0) {
        orb_keypoint o;
        o.x = level ? __fmul_rn(k.x, L.scale) : k.x;
        o.y = level ? __fmul_rn(k.y, L.scale) : k.y;
        o.size = L.patchSize;
        o.angle = k.angle;
        o.response = k.response;
        o.octave = level;
        o.class_id = -1;
        kps[(size_t)frame * P.outCapacity + g] = o;
    }
}

int launch_brief(const ExtractParams& P, int maxKeypoints, orb_keypoint* dKps, unsigned char* dDesc, int* dCount,
                 cudaStream_t st, int* launches) {
    dim3 grid(ceil_div(maxKeypoints > 0 ? maxKeypoints : 1, BR_WARPS), P.nFrames);
    brief_kernel<<<grid, BR_WARPS * 32, 0, st>>>(P, dKps, dDesc, dCount);
    ++*launches;
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orbb
