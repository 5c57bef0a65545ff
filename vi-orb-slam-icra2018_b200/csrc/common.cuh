// Shared host/device helpers of liborbb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/orbb200.h"

namespace orbb {

// ---- error plumbing: every C-ABI entry point returns orb_status, message kept per host thread ----
std::string& last_error();
int fail(orb_status st, const char* fmt, ...);

#define ORB_CUDA(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return ::orbb::fail(ORB_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    } while (0)

#define ORB_CHECK(expr)                                  \
    do {                                                 \
        int st__ = (expr);                               \
        if (st__ != ORB_OK) return st__;                 \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// growable device buffer (never shrinks); used for matcher workspaces
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return ORB_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        ORB_CUDA(cudaMalloc(&p, want));
        cap = want;
        return ORB_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return (T*)p; }
};

// growable page-locked host staging (batched host entry points: one memcpy into it, one full-rate H2D out of it)
struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return ORB_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        ORB_CUDA(cudaHostAlloc(&p, want, cudaHostAllocDefault));
        cap = want;
        return ORB_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return (T*)p; }
};

constexpr int kEdge = 19;         // EDGE_THRESHOLD, ORBextractor.cc:76
constexpr int kHalfPatch = 15;    // HALF_PATCH_SIZE
constexpr int kPatch = 31;        // PATCH_SIZE
constexpr int kGridCols = 64;     // FRAME_GRID_COLS, Frame.h:42
constexpr int kGridRows = 48;     // FRAME_GRID_ROWS, Frame.h:41
constexpr int kThHigh = 100;      // ORBmatcher.cc:36
constexpr int kThLow = 50;        // ORBmatcher.cc:37
constexpr int kHistoLength = 30;  // ORBmatcher.cc:38

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
// 256-bit Hamming distance: 8 x (LOP3 xor, POPC), summed with IADD3  (ORBmatcher.cc:1675-1691 computes the same
// number with a SWAR bit count)
__device__ __forceinline__ int hamming256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
    return (__popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z)) +
           (__popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y)) +
           (__popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w));
}

// rot = a1 - a2; if (rot < 0) rot += 360; bin = round(rot * (1.0f/HISTO_LENGTH)); bin==30 -> 0   (e.g. ORBmatcher.cc:475-481)
__device__ __forceinline__ int rotation_bin(float a1, float a2) {
    float rot = __fsub_rn(a1, a2);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, 1.0f / kHistoLength));
    if (bin == kHistoLength) bin = 0;
    return bin;
}

// ComputeThreeMaxima (ORBmatcher.cc:1629-1670), run by one thread over 30 bin sizes
__device__ __forceinline__ void three_maxima(const int* sz, int& i1, int& i2, int& i3) {
    int m1 = 0, m2 = 0, m3 = 0;
    i1 = i2 = i3 = -1;
    for (int i = 0; i < kHistoLength; ++i) {
        const int s = sz[i];
        if (s > m1) { m3 = m2; m2 = m1; m1 = s; i3 = i2; i2 = i1; i1 = i; }
        else if (s > m2) { m3 = m2; m2 = s; i3 = i2; i2 = i; }
        else if (s > m3) { m3 = s; i3 = i; }
    }
    if ((float)m2 < __fmul_rn(0.1f, (float)m1)) { i2 = -1; i3 = -1; }
    else if ((float)m3 < __fmul_rn(0.1f, (float)m1)) { i3 = -1; }
}
#endif

}  // namespace orbb
