// sinf / cosf with glibc 2.39's results, bit for bit, for arguments in [0, 2*pi] (plus a little slack).
//
// Why: the reference computes the descriptor rotation as (float)cos(angle), (float)sin(angle) with a float argument
// under `using namespace std` (ORBextractor.cc:114-115), i.e. libm cosf/sinf.  glibc's implementation is not
// correctly rounded (it differs from the rounded double result on ~1.3% of inputs) and CUDA's cosf/sinf differ again,
// so bit-exact descriptors need the same algorithm: argument widened to double, reduction by pi/2 with a 2^24-scaled
// 2/pi, degree-7 / degree-8 minimax polynomials in double, one final rounding to float (the ARM optimized-routines
// algorithm that glibc adopted in 2.28).  tests/test_sincosf.py compiles this header for the host and compares it
// with libm for EVERY float in [0, 6.2832] (1.09e9 values, 0 mismatches, with or without FMA contraction).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define ORB_HD __host__ __device__ __forceinline__
#define ORB_FMA(a, b, c) __fma_rn((a), (b), (c))
#define ORB_MUL(a, b) __dmul_rn((a), (b))
#elif defined(__CUDACC__)
#define ORB_HD __host__ __device__ __forceinline__
#define ORB_FMA(a, b, c) ((a) * (b) + (c))
#define ORB_MUL(a, b) ((a) * (b))
#else
#define ORB_HD static inline
#define ORB_FMA(a, b, c) ((a) * (b) + (c))
#define ORB_MUL(a, b) ((a) * (b))
#endif

namespace orbb {

// even n: sine polynomial on x; odd n: cosine polynomial with coefficient sign `cs`
ORB_HD float sincosf_poly(double x, double x2, int n, double cs) {
    const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10,
                 C4 = 0x1.99343027bf8c3p-16;
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0) {
        const double x3 = ORB_MUL(x, x2);
        const double s1 = ORB_FMA(x2, S3, S2);
        const double x7 = ORB_MUL(x3, x2);
        const double s = ORB_FMA(x3, S1, x);
        return (float)ORB_FMA(x7, s1, s);
    }
    const double x4 = ORB_MUL(x2, x2);
    const double c2 = ORB_FMA(x2, cs * C4, cs * C3);
    const double c1 = ORB_FMA(x2, cs * C2, cs * C1);
    const double x6 = ORB_MUL(x4, x2);
    const double c = ORB_FMA(x2, c1, cs * C0);
    return (float)ORB_FMA(x6, c2, c);
}

// valid for 0 <= y < 120 (the extractor only passes [0, 2*pi])
ORB_HD void sincosf_glibc(float y, float* sinOut, float* cosOut) {
    const double hpiInv = 0x1.45F306DC9C883p+23;   // 2/pi * 2^24
    const double hpi = 0x1.921FB54442D18p0;
    double x = (double)y;
    if (y < 0.75f) {                               // glibc compares the top 12 bits with those of pi/4: that is y < 0.75
        if (y < 0x1p-12f) {
            *sinOut = y;
            *cosOut = 1.0f;
            return;
        }
        const double x2 = ORB_MUL(x, x);
        *sinOut = sincosf_poly(x, x2, 0, 1.0);
        *cosOut = sincosf_poly(x, x2, 1, 1.0);
        return;
    }
    const double r = ORB_MUL(x, hpiInv);
    const int n = (int)(((int32_t)r + 0x800000) >> 24);
    x = ORB_FMA(-(double)n, hpi, x);
    const double s = ((n + 1) & 2) ? -1.0 : 1.0;   // sign[n & 3] = {1, -1, -1, 1}
    const double cs = (n & 2) ? -1.0 : 1.0;
    const double xs = ORB_MUL(x, s), x2 = ORB_MUL(x, x);
    *sinOut = sincosf_poly(xs, x2, n, cs);
    *cosOut = sincosf_poly(xs, x2, n ^ 1, cs);
}

}  // namespace orbb
