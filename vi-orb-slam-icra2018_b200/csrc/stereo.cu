// Frame::ComputeStereoMatches (Frame.cc:810-984) on the two extractors' device-resident pyramids.
//
// The reference builds a row table of the right keypoints (each covers the rows kpY -+ 2*scale[octave]), then for every
// left keypoint walks the candidates of its row: octave within +-1, uR in [uL - mbf/mb, uL], closest descriptor with a
// strict '<' (so the lowest right index wins a tie), accepted below (TH_HIGH + TH_LOW) / 2.  The match is refined by an
// 11x11 SAD of centre-normalised patches over the shifts -5..+5 on mvImagePyramid[kpL.octave] of both images and a
// parabola through the three SADs around the best shift; finally matches with SAD >= 1.5*1.4*median are withdrawn.
//
// Here every left keypoint is independent until the median step, so: one warp per left keypoint; lanes stride over ALL
// right keypoints and apply the row-band test directly (nR is one or two thousand: the row table would cost more than
// the 12 bytes per test it saves); (distance, index) keys are min-reduced by shuffle; the same warp stages the two
// patches in shared memory, the 11 shifts x 11 patch rows are summed by all lanes in integers (the reference's float
// differences and its double accumulation are exact on these integer-valued operands), lane 0 fits the parabola with
// separately rounded float operations.  A second one-block kernel finds the median by a two-pass radix select and
// withdraws the outliers.
#include "extractor.h"

namespace orbb {

namespace {

constexpr int kWarpsPerCta = 4;
constexpr int kW = 5, kL = 5;                      // Frame.cc:907, :915
constexpr int kPatchL = (2 * kW + 1) * (2 * kW + 1);                 // 121
constexpr int kRowR = 2 * (kW + kL) + 1;                              // 21 columns cover every shifted window
constexpr int kPatchR = (2 * kW + 1) * kRowR;                         // 231

__device__ __forceinline__ int band_lo(float y, float r) { return (int)floorf(__fsub_rn(y, r)); }   // :833
__device__ __forceinline__ int band_hi(float y, float r) { return (int)ceilf(__fadd_rn(y, r)); }    // :832

__global__ void __launch_bounds__(kWarpsPerCta * 32)
stereo_match_kernel(StereoParams P, const orb_keypoint* __restrict__ keysL, const uint4* __restrict__ descL, int nLmax,
                    const int* __restrict__ dNL, const orb_keypoint* __restrict__ keysR, const uint4* __restrict__ descR,
                    int nRmax, const int* __restrict__ dNR, float* __restrict__ uRight, float* __restrict__ depth,
                    int* __restrict__ sad) {
    __shared__ short sL[kWarpsPerCta][kPatchL + 7];
    __shared__ short sR[kWarpsPerCta][kPatchR + 1];
    __shared__ int sSad[kWarpsPerCta][2 * kL + 2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int iL = blockIdx.x * kWarpsPerCta + warp;
    const int nL = dNL ? min(*dNL, nLmax) : nLmax;
    const int nR = dNR ? min(*dNR, nRmax) : nRmax;
    if (iL >= nL) return;
    if (lane == 0) { uRight[iL] = -1.0f; depth[iL] = -1.0f; sad[iL] = -1; }     // :812-813
    const orb_keypoint kpL = keysL[iL];
    const int levelL = kpL.octave;
    const float uL = kpL.x, vL = kpL.y;
    const int row = (int)vL;                                                    // vRowIndices[vL], :856
    if (row < 0 || row >= P.nRows) return;
    const float minU = __fsub_rn(uL, P.maxD), maxU = uL;                        // :861-862 (minD = 0)
    if (maxU < 0) return;
    const uint4 a0 = descL[2 * iL], a1 = descL[2 * iL + 1];
    unsigned best = (unsigned)kThHigh << 16;                                    // bestDist = TH_HIGH, strict '<'
    // Four right keypoints per lane and round: their twelve field loads are issued before any test, so a round costs one
    // memory latency instead of four (the kernel is bound by exactly that latency: ncu long_scoreboard 76 % of samples).
    for (int base = 0; base < nR; base += 128) {
        float ry[4], rx[4];
        int ro[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int iR = base + j * 32 + lane;
            ro[j] = -1;
            if (iR < nR) {
                ry[j] = keysR[iR].y;
                rx[j] = keysR[iR].x;
                ro[j] = keysR[iR].octave;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int iR = base + j * 32 + lane;
            if ((unsigned)ro[j] >= (unsigned)P.nLevels) continue;
            const float r = __fmul_rn(2.0f, P.scale[ro[j]]);                    // :831
            const int lo = max(band_lo(ry[j], r), 0), hi = min(band_hi(ry[j], r), P.nRows - 1);
            if (row < lo || row > hi) continue;
            if (ro[j] < levelL - 1 || ro[j] > levelL + 1) continue;             // :878
            if (!(rx[j] >= minU && rx[j] <= maxU)) continue;                    // :883
            const unsigned key = ((unsigned)hamming256(a0, a1, descR[2 * iR], descR[2 * iR + 1]) << 16) | (unsigned)iR;
            best = min(best, key);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    const int bestDist = (int)(best >> 16);
    if (!(bestDist < (kThHigh + kThLow) / 2)) return;                           // thOrbDist, :815, :898
    const int bestIdxR = (int)(best & 0xffffu);
    if ((unsigned)levelL >= (unsigned)P.nLevels) return;

    const float uR0 = keysR[bestIdxR].x;
    const float isf = P.invScale[levelL];
    const int su = (int)roundf(__fmul_rn(uL, isf));                             // :903-905
    const int sv = (int)roundf(__fmul_rn(vL, isf));
    const int sr0 = (int)roundf(__fmul_rn(uR0, isf));
    const int cols = P.cols[levelL], rows = P.rows[levelL];
    // cv::Mat::rowRange / colRange assert that a window stays inside the level (:908, :925): no match out there
    if (su - kW < 0 || su + kW + 1 > cols || sv - kW < 0 || sv + kW + 1 > rows) return;
    if (sr0 < 0 || sr0 + kL + kW + 1 >= cols) return;                           // iniu < 0 || endu >= cols, :920
    if (sr0 - kL - kW < 0) return;
    const int pitch = P.pitch[levelL];
    const unsigned char* baseL = P.pyrL + P.lvOff[levelL] + (size_t)(sv - kW + kEdge) * pitch + kPadLeft + (su - kW);
    const unsigned char* baseR = P.pyrR + P.lvOff[levelL] + (size_t)(sv - kW + kEdge) * pitch + kPadLeft + (sr0 - kL - kW);
    for (int i = lane; i < kPatchL; i += 32) sL[warp][i] = baseL[(i / 11) * pitch + (i % 11)];
    for (int i = lane; i < kPatchR; i += 32) sR[warp][i] = baseR[(i / kRowR) * pitch + (i % kRowR)];
    __syncwarp();
    // 11 shifts x 11 patch rows = 121 row sums spread over the 32 lanes, added up per shift in shared memory
    if (lane < 2 * kL + 1) sSad[warp][lane] = 0;
    __syncwarp();
    {
        const int cL = sL[warp][kW * 11 + kW];
        for (int t = lane; t < (2 * kL + 1) * 11; t += 32) {
            const int inc = t / 11, y = t - inc * 11;                           // shift incR = inc - L, :923-937
            const int cR = sR[warp][kW * kRowR + inc + kW];
            int acc = 0;
#pragma unroll
            for (int x = 0; x < 11; ++x)
                acc += abs((sL[warp][y * 11 + x] - cL) - (sR[warp][y * kRowR + inc + x] - cR));
            atomicAdd(&sSad[warp][inc], acc);
        }
    }
    __syncwarp();
    int mySad = 0x7fffffff >> 4;
    if (lane < 2 * kL + 1) mySad = sSad[warp][lane];
    unsigned sk = ((unsigned)mySad << 4) | (unsigned)(lane & 15);               // first minimum wins
    if (lane >= 2 * kL + 1) sk = 0xffffffffu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sk = min(sk, __shfl_xor_sync(0xffffffffu, sk, o));
    const int bi = (int)(sk & 15u);
    const int bestSad = (int)(sk >> 4);
    const int i1 = max(bi - 1, 0), i3 = min(bi + 1, 2 * kL);
    const float dist1 = (float)__shfl_sync(0xffffffffu, mySad, i1);
    const float dist2 = (float)bestSad;
    const float dist3 = (float)__shfl_sync(0xffffffffu, mySad, i3);
    if (bi == 0 || bi == 2 * kL) return;                                        // :939
    if (lane != 0) return;
    const float den = __fmul_rn(2.0f, __fsub_rn(__fadd_rn(dist1, dist3), __fmul_rn(2.0f, dist2)));
    const float deltaR = __fdiv_rn(__fsub_rn(dist1, dist3), den);               // :947
    if (deltaR < -1 || deltaR > 1) return;
    float bestuR = __fmul_rn(P.scale[levelL], __fadd_rn(__fadd_rn((float)sr0, (float)(bi - kL)), deltaR));   // :953
    float disparity = __fsub_rn(uL, bestuR);
    if (disparity >= 0.0f && disparity < P.maxD) {                              // :957
        if (disparity <= 0) {
            disparity = 0.01f;
            bestuR = (float)__dsub_rn((double)uL, 0.01);                        // uL - 0.01 is double arithmetic, :962
        }
        depth[iL] = __fdiv_rn(P.mbf, disparity);
        uRight[iL] = bestuR;
        sad[iL] = bestSad;
    }
}

// sort(vDistIdx); median = vDistIdx[size/2].first; thDist = 1.5f*1.4f*median; withdraw every match with SAD >= thDist
// (Frame.cc:971-984).  Only the VALUE of the (size/2)-th smallest SAD matters, and a SAD is at most 121 * 510 < 2^16:
// two-pass radix select over 256-bin shared histograms (high byte, then low byte inside the selected bin).
__device__ __forceinline__ int select_bin(const int* hist, int rank, int* before) {   // warp 0: first bin with cum > rank
    const int lane = threadIdx.x & 31;
    int c[8], sum = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { c[k] = hist[lane * 8 + k]; sum += c[k]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    int run = incl - sum, bin = 0x7fff, bef = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (bin == 0x7fff && run + c[k] > rank) { bin = lane * 8 + k; bef = run; }
        run += c[k];
    }
    const int key = __reduce_min_sync(0xffffffffu, bin);
    const int src = __ffs(__ballot_sync(0xffffffffu, bin == key)) - 1;
    *before = __shfl_sync(0xffffffffu, bef, src);
    return key;
}

__global__ void __launch_bounds__(1024)
stereo_median_kernel(int nLmax, const int* __restrict__ dNL, float* __restrict__ uRight, float* __restrict__ depth,
                     int* __restrict__ sad, int* __restrict__ kept) {
    __shared__ int hist[256];
    __shared__ int sCount, sHigh, sRank, sMedian, sKept;
    const int nL = dNL ? min(*dNL, nLmax) : nLmax;
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) { sCount = 0; sKept = 0; }
    __syncthreads();
    int c = 0;
    for (int i = threadIdx.x; i < nL; i += blockDim.x) {
        const int s = sad[i];
        if (s >= 0) { ++c; atomicAdd(&hist[min(s >> 8, 255)], 1); }
    }
    if (c) atomicAdd(&sCount, c);
    __syncthreads();
    const int count = sCount;
    if (count == 0) {           // the reference indexes an empty vector here; defined as "nothing to withdraw"
        if (threadIdx.x == 0) *kept = 0;
        return;
    }
    if (threadIdx.x < 32) {
        int before;
        const int hi = select_bin(hist, count / 2, &before);
        if (threadIdx.x == 0) { sHigh = hi; sRank = count / 2 - before; }
    }
    __syncthreads();
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    __syncthreads();
    const int hi = sHigh;
    for (int i = threadIdx.x; i < nL; i += blockDim.x) {
        const int s = sad[i];
        if (s >= 0 && min(s >> 8, 255) == hi) atomicAdd(&hist[s & 255], 1);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        int before;
        const int lo = select_bin(hist, sRank, &before);
        if (threadIdx.x == 0) sMedian = (hi << 8) | lo;
    }
    __syncthreads();
    const float thDist = __fmul_rn(1.5f * 1.4f, (float)sMedian);
    int mine = 0;
    for (int i = threadIdx.x; i < nL; i += blockDim.x) {
        const int s = sad[i];
        if (s < 0) continue;
        if ((float)s < thDist) {
            ++mine;
        } else {
            uRight[i] = -1.0f;
            depth[i] = -1.0f;
            sad[i] = -1;
        }
    }
    if (mine) atomicAdd(&sKept, mine);
    __syncthreads();
    if (threadIdx.x == 0) *kept = sKept;
}

}  // namespace

int launch_stereo(const StereoParams& P, const orb_keypoint* dKeysL, const unsigned char* dDescL, int nLmax, const int* dNL,
                  const orb_keypoint* dKeysR, const unsigned char* dDescR, int nRmax, const int* dNR, float* dURight,
                  float* dDepth, int* dSad, int* dKept, cudaStream_t st, int* launches) {
    if (nLmax > 0)
        stereo_match_kernel<<<ceil_div(nLmax, kWarpsPerCta), kWarpsPerCta * 32, 0, st>>>(
            P, dKeysL, (const uint4*)dDescL, nLmax, dNL, dKeysR, (const uint4*)dDescR, nRmax, dNR, dURight, dDepth, dSad);
    stereo_median_kernel<<<1, 1024, 0, st>>>(nLmax, dNL, dURight, dDepth, dSad, dKept);
    *launches += 2;
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orbb
