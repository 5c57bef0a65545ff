// Orientation + rotated-BRIEF descriptors + final keypoint assembly (north-star kernels 4 and 6):
//   IC_Angle / computeOrientation                ORBextractor.cc:79-106, 474-481
//   computeOrbDescriptor / computeDescriptors   ORBextractor.cc:110-149, 1036-1043 (pattern table :152-410)
//   coordinate rescale and concatenation         ORBextractor.cc:1112-1121
//
// One warp per keypoint.  Orientation: integer moments of the circular patch of the UNBLURRED level by dot products of
// aligned pixel words with tabulated coordinate bytes, then float32 fastAtan2 with the reference's operation order (no
// FMA).  Descriptor: lane i produces descriptor byte i from pattern pairs 8i..8i+7.  The rotation is
// a = cosf(angle*factorPI), b = sinf(..) with glibc's cosf/sinf reproduced bit for bit (glibc_sincosf.h); sample
// coordinates are cvRound(x*b + y*a), cvRound(x*a - y*b) with every float op rounded separately (the reference's
// source-level order, no FMA) and round-half-even conversion.  Samples come from the BLURRED level.
// Keypoints of a frame are written level 0..nLevels-1, within a level in quadtree list order; pt is scaled by
// mvScaleFactor[level] only after the descriptor is taken, as the reference does.
#include <cfloat>

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

// IC_Angle (ORBextractor.cc:79-106) as dot products.  The 31 rows of the circular patch are read as aligned 32-bit words
// (9 per row); for each of the four alignments of the patch's left edge and each (row, word) item the table holds the
// signed u and v coordinates of the word's four bytes (0 outside the circle), the row and the word's byte offset, so
//   m10 += dp4a(u bytes, pixels),  m01 += dp4a(v bytes, pixels).
// 279 items are dealt to the 32 lanes of a warp, 9 each (the last 9 table slots are zero).
constexpr int kOriItems = 9 * 32;
__device__ int4 gOriTable[4 * kOriItems];

int upload_orientation_table(const int* umax) {
    static int4 host[4 * kOriItems];
    for (int a = 0; a < 4; ++a)
        for (int t = 0; t < kOriItems; ++t) {
            int4 e = make_int4(0, 0, 0, 0);
            if (t < 9 * kPatch) {
                const int v = t / 9 - kHalfPatch, j = t % 9;
                unsigned int uc = 0, vc = 0;
                for (int b = 0; b < 4; ++b) {
                    const int u = 4 * j + b - a - kHalfPatch;
                    if (u < -kHalfPatch || u > kHalfPatch || (u < 0 ? -u : u) > umax[v < 0 ? -v : v]) continue;
                    uc |= (unsigned int)(u & 0xff) << (8 * b);
                    vc |= (unsigned int)(v & 0xff) << (8 * b);
                }
                e = make_int4((int)uc, (int)vc, v, 4 * j);
            }
            host[a * kOriItems + t] = e;
        }
    ORB_CUDA(cudaMemcpyToSymbol(gOriTable, host, sizeof host));
    return ORB_OK;
}

__device__ __forceinline__ int dp4a_su(int coef, unsigned int pixels, int acc) {   // signed bytes x unsigned bytes
    int d;
    asm("dp4a.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(coef), "r"(pixels), "r"(acc));
    return d;
}

// cv::fastAtan2 (degrees), float32 with the reference operation order, every op rounded (no contraction)
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float P1 = 57.283626556396484f, P3 = -18.66744613647461f, P5 = 8.914000511169434f, P7 = -2.539724588394165f;
    const float eps = (float)DBL_EPSILON;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(P7, c2), P5), c2), P3), c2), P1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(P7, c2), P5), c2), P3), c2), P1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
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

    // orientation: 279 (row, word) items of the 31x31 patch, 9 per lane; all loads are issued before the first use
    float angle;
    {
        const int x = (int)k.x, y = (int)k.y;
        const int al = (x - kHalfPatch) & 3;          // the level's pixel (0, y) is 4-byte aligned
        const unsigned char* c0 = P.pyr + (size_t)frame * P.pyrFrameBytes + L.pyrOff + (size_t)(kEdge + y) * L.pitch + kPadLeft +
                                  (x - kHalfPatch - al);
        const int4* tab = gOriTable + al * kOriItems + lane;
        int4 e[9];
        unsigned int w[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) e[i] = __ldg(tab + 32 * i);
#pragma unroll
        for (int i = 0; i < 9; ++i) w[i] = __ldg(reinterpret_cast<const unsigned int*>(c0 + e[i].z * L.pitch + e[i].w));
        int m10 = 0, m01 = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            m10 = dp4a_su(e[i].x, w[i], m10);
            m01 = dp4a_su(e[i].y, w[i], m01);
        }
        m10 = __reduce_add_sync(0xffffffffu, m10);
        m01 = __reduce_add_sync(0xffffffffu, m01);
        angle = fast_atan2_deg((float)m01, (float)m10);
    }
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    float a, b;
    sincosf_glibc(__fmul_rn(angle, factorPI), &b, &a);   // a = cos, b = sin
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
        o.angle = angle;
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
