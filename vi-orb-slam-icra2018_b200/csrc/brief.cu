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
#include <cuda.h>

#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <cstring>

#include "extractor.h"
#include "glibc_sincosf.h"

namespace orbb {

__constant__ char4 cPattern[256];   // (x0, y0, x1, y1) per test pair

static const signed char kPatternHost[1024] = {
#include "brief_pattern.inc"
};

__device__ char4 gPatternT[256];   // transposed copy for the staged kernel: gPatternT[j * 32 + lane] = pair 8 * lane + j

int upload_brief_pattern() {
    ORB_CUDA(cudaMemcpyToSymbol(cPattern, kPatternHost, sizeof kPatternHost));
    static char4 hostT[256];
    for (int lane = 0; lane < 32; ++lane)
        for (int j = 0; j < 8; ++j) {
            const signed char* q = kPatternHost + 4 * (8 * lane + j);
            hostT[j * 32 + lane] = make_char4(q[0], q[1], q[2], q[3]);
        }
    ORB_CUDA(cudaMemcpyToSymbol(gPatternT, hostT, sizeof hostT));
    return ORB_OK;
}

// IC_Angle (ORBextractor.cc:79-106) as dot products.  The 31 rows of the circular patch are read as aligned 32-bit words
// (9 per row); for each of the four alignments of the patch's left edge and each (row, word) item the table holds the
// signed u and v coordinates of the word's four bytes (0 outside the circle), the row and the word's byte offset, so
//   m10 += dp4a(u bytes, pixels),  m01 += dp4a(v bytes, pixels).
// 279 items are dealt to the 32 lanes of a warp, 9 each (the last 9 table slots are zero).
constexpr int kOriItems = 9 * 32;
__device__ int4 gOriTable[4 * kOriItems];
__device__ int2 gOriCoef[4 * kOriItems];   // staged kernel: .x = u bytes, .y = v bytes of gOriTable (row and word come from the item index)

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
    static int2 coef[4 * kOriItems];
    for (int i = 0; i < 4 * kOriItems; ++i) coef[i] = make_int2(host[i].x, host[i].y);
    ORB_CUDA(cudaMemcpyToSymbol(gOriCoef, coef, sizeof coef));
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

// ---- the same kernel with the blurred patch staged through shared memory by one TMA tensor load per keypoint --------
// brief_kernel's 16 byte gathers per lane go to 32 different rows of the blurred level: ~11 L1 wavefronts per load
// instruction, and the LSU wavefront pipe was the kernel's bound (78 %).  Here lane 0 requests the keypoint's 37 x 64-byte
// box of the blurred level ([x0a, x0a + 64) x [y - 18, y + 18], x0a = the 16-byte boundary below x - 18: TMA cannot
// realign a byte image) as soon as the keypoint is known; orientation (global loads of the unblurred level) and sin / cos
// run while it is in flight, then the samples are shared-memory byte loads at 32-bit addresses.  The level of a keypoint
// comes from one warp scan of the per-level counts instead of a serial loop, the pattern from a float table.
constexpr int BRS_TILE_W = 80, BRS_TILE_H = 37, BRS_TILE_BYTES = BRS_TILE_W * BRS_TILE_H;   // 80: rows spread over all banks
constexpr int BRS_TILE_STRIDE = (BRS_TILE_BYTES + 127) / 128 * 128;
constexpr int BRS_PER_WARP = 8;      // consecutive keypoints of a frame per warp: pattern, level scan and tables are loaded once


__global__ void __launch_bounds__(BR_WARPS * 32, 4)
brief_staged_kernel(const __grid_constant__ ExtractParams P, orb_keypoint* __restrict__ kps, unsigned char* __restrict__ desc,
                    int* __restrict__ nOut, int perWarp) {
    __shared__ __align__(128) unsigned char tiles[BR_WARPS * BRS_TILE_STRIDE];
    __shared__ __align__(8) unsigned long long bars[BR_WARPS];
    const int frame = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int bar = (unsigned int)__cvta_generic_to_shared(&bars[warp]);
    const unsigned int tileAddr = (unsigned int)__cvta_generic_to_shared(tiles + warp * BRS_TILE_STRIDE);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // levels: lane l holds the count of level l; one inclusive scan
    const int cnt = lane < P.nLevels ? P.selCount[(size_t)frame * P.nLevels + lane] : 0;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < kMaxLevels; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = min(__shfl_sync(0xffffffffu, incl, kMaxLevels - 1), P.outCapacity);
    if (blockIdx.x == 0 && tid == 0) nOut[frame] = __shfl_sync(1u, incl, 0) * 0 + total;
    const int g0 = (blockIdx.x * BR_WARPS + warp) * perWarp;
    if (g0 >= total) return;
    char4 pat[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) pat[j] = gPatternT[j * 32 + lane];
    // orientation items of this lane: t = lane + 32 i -> patch row v = t / 9 - 15, word t % 9; 32 items further on is
    // 3 rows and 5 words further on
    const int oriV0 = lane / 9 - kHalfPatch, oriW0 = lane % 9;
    const unsigned char* pyrFrame = P.pyr + (size_t)frame * P.pyrFrameBytes;
    const SelKey* selFrame = P.sel + (size_t)frame * P.selPerFrame;
    const int gEnd = min(g0 + perWarp, total);
    unsigned int parity = 0;
    // the record of keypoint g + 1 is requested at the end of keypoint g, before its descriptor byte and keypoint go out
    // (3.04 -> 2.92 ms per 4096 frames; requested earlier -- at the top of g: 3.19 ms with 60 bytes of spills, before the
    // tile wait: 2.98 ms)
    auto locate = [&](int g, int& level) -> const SelKey* {
        const unsigned int below = __ballot_sync(0xffffffffu, incl <= g);     // levels wholly before keypoint g
        level = __popc(below & ((1u << kMaxLevels) - 1u));
        const int idx = g - (__shfl_sync(0xffffffffu, incl, level) - __shfl_sync(0xffffffffu, cnt, level));
        return selFrame + P.lv[level].selBase + idx;
    };
    int level, levelNext;
    SelKey k, kNext = *locate(g0, levelNext);
    for (int g = g0; g < gEnd; ++g) {
        k = kNext;
        level = levelNext;
        const LevelGeom& L = P.lv[level];
        const int x = (int)k.x, y = (int)k.y;
        const int x0a = (x - 18) & ~15;
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the previous keypoint's reads of the tile are done (__syncwarp below)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(BRS_TILE_BYTES) : "memory");
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(tileAddr),
                "l"(static_cast<const unsigned char*>(P.brMaps) + 128 * level), "r"(x0a), "r"(y - 18), "r"(P.fw.frameBase + frame), "r"(bar)
                : "memory");
        }
        float angle;
        {
            const int al = (x - kHalfPatch) & 3;          // the level's pixel (0, y) is 4-byte aligned
            const int pitch = L.pitch;
            const unsigned char* c0 = pyrFrame + L.pyrOff + (size_t)(kEdge + y) * pitch + kPadLeft + (x - kHalfPatch - al);
            const int2* tab = gOriCoef + al * kOriItems + lane;
            int2 e[9];
            unsigned int w[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) e[i] = __ldg(tab + 32 * i);
            {
                int off = oriV0 * pitch + 4 * oriW0, wd = oriW0;
#pragma unroll
                for (int i = 0; i < 9; ++i) {
                    w[i] = __ldg(reinterpret_cast<const unsigned int*>(c0 + off));
                    const bool wrap = wd >= 4;                     // wd + 5 >= 9: one more row, nine words back
                    off += wrap ? 4 * pitch - 16 : 3 * pitch + 20;
                    wd += wrap ? -4 : 5;
                }
            }
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
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "BRS_WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra BRS_DONE_%=;\n"
            "bra BRS_WAIT_%=;\n"
            "BRS_DONE_%=:\n"
            "}\n" ::"r"(bar),
            "r"(parity)
            : "memory");
        parity ^= 1u;
#ifdef BRS_F2I
        const unsigned int center = tileAddr + 18 * BRS_TILE_W + (unsigned int)(x - x0a);
#else
        // cvRound without the conversion pipe (F2I issues on the XU pipe, the kernel's busiest): for |v| < 2^22 the float
        // v + 1.5 * 2^23 has its integer value, rounded to nearest-even like lrintf, in its low mantissa bits; the bias
        // 0x4B400000 of both coordinates is taken out of the tile address once (32-bit wrap-around arithmetic)
        const float kRound = 12582912.f;
        const unsigned int center = tileAddr + 18 * BRS_TILE_W + (unsigned int)(x - x0a) - 0x4B400000u * (unsigned int)(BRS_TILE_W + 1);
#endif
        unsigned int val = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float px = (float)pat[j].x, py = (float)pat[j].y, pz = (float)pat[j].z, pw = (float)pat[j].w;
#ifdef BRS_F2I
            const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)));
            const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)));
            const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(pz, b), __fmul_rn(pw, a)));
            const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(pz, a), __fmul_rn(pw, b)));
#else
            const unsigned int r0 = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)), kRound));
            const unsigned int c0 = __float_as_uint(__fadd_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)), kRound));
            const unsigned int r1 = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(pz, b), __fmul_rn(pw, a)), kRound));
            const unsigned int c1 = __float_as_uint(__fadd_rn(__fsub_rn(__fmul_rn(pz, a), __fmul_rn(pw, b)), kRound));
#endif
            unsigned int t0, t1;
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t0) : "r"(center + (unsigned int)(r0 * BRS_TILE_W + c0)));
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t1) : "r"(center + (unsigned int)(r1 * BRS_TILE_W + c1)));
            val |= (unsigned int)(t0 < t1) << j;
        }
        __syncwarp();      // every lane has read its samples: the tile may be refilled
        if (g + 1 < gEnd) kNext = *locate(g + 1, levelNext);
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
}

// one CUtensorMap per level over [arena frames][h][bpitch] of the blurred arena, box = 64 x 37 x 1 bytes
int brief_encode_maps(const ExtractParams& P, int arenaFrames, void* hostMaps) {
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
        const cuuint64_t dims[3] = {(cuuint64_t)L.bpitch, (cuuint64_t)L.h, (cuuint64_t)arenaFrames};
        const cuuint64_t strides[2] = {(cuuint64_t)L.bpitch, (cuuint64_t)P.blurFrameBytes};
        const cuuint32_t box[3] = {(cuuint32_t)BRS_TILE_W, (cuuint32_t)BRS_TILE_H, 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        const CUresult r = encode(&maps[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, P.blur + L.blurOff, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(ORB_ERR_CUDA, "cuTensorMapEncodeTiled(blurred level %d, %dx%d) -> %d", l, L.bpitch, L.h, (int)r);
    }
    return ORB_OK;
}

int launch_brief(const ExtractParams& P, int maxKeypoints, orb_keypoint* dKps, unsigned char* dDesc, int* dCount,
                 cudaStream_t st, int* launches) {
    dim3 grid(ceil_div(maxKeypoints > 0 ? maxKeypoints : 1, BR_WARPS), P.nFrames);
    // batches: 8 consecutive keypoints per warp (pattern, level scan and tables loaded once); small calls: one per warp, so
    // that a single frame's ~1000 keypoints spread over the whole GPU instead of running eight deep
    static const int batchPerWarp = getenv("ORBB_BRIEF_PERWARP") ? std::max(1, atoi(getenv("ORBB_BRIEF_PERWARP"))) : BRS_PER_WARP;   // tuning aid
    const int perWarp = P.nFrames >= P.pyBulkMinFrames ? batchPerWarp : 1;
    const dim3 gridStaged(ceil_div(maxKeypoints > 0 ? maxKeypoints : 1, BR_WARPS * perWarp), P.nFrames);
    static const bool noStage = getenv("ORBB_BRIEF_DIRECT") != nullptr;      // A/B aid: the direct-gather kernel
    if (P.brMaps && !noStage && P.nLevels <= kMaxLevels) {
        brief_staged_kernel<<<gridStaged, BR_WARPS * 32, 0, st>>>(P, dKps, dDesc, dCount, perWarp);
        ++*launches;
        ORB_CUDA(cudaGetLastError());
        return ORB_OK;
    }
    brief_kernel<<<grid, BR_WARPS * 32, 0, st>>>(P, dKps, dDesc, dCount);
    ++*launches;
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orbb
