// Windowed and vocabulary-gated searches of ORBmatcher on flat arrays, plus the Frame grid they read.
//
//   Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea     Frame.cc:574-589, 726-736, 671-724
//   ORBmatcher::SearchForInitialization                             ORBmatcher.cc:405-520
//   ORBmatcher::SearchByProjection(Frame&, const Frame&, th, mono)  ORBmatcher.cc:1341-1498
//   ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th)  ORBmatcher.cc:45-129
//   ORBmatcher::SearchForTriangulation (+CheckDistEpipolarLine)     ORBmatcher.cc:657-823, 140-157
//
// The reference loops are "parallel distances, sequential bookkeeping": a later query sees what earlier queries
// matched (vMatchedDistance / mvpMapPoints).  They are split accordingly:
//   phase 1 (one warp per query, whole GPU): enumerate the grid candidates in the reference's order (cell column,
//            cell row, bucket order), apply the static filters, compute the 256-bit Hamming distances;
//   phase 2 (one warp per search): replay the queries in order over the stored (index, distance) lists with the
//            dynamic skip rules; lanes split a query's candidates and merge (distance, order) keys by shuffle.
// SearchForTriangulation never sets vbMatched2 (ORBmatcher.cc:725), so its queries are independent: one thread each.
#include <cstring>
#include <vector>

#include "matcher.h"

namespace orbb {

struct FrameDev {
    int n;
    const orb_keypoint* keys;
    const uint4* desc;
    const int* cellStart;   // [64*48 + 1], cell id = ix*48 + iy
    const int* cellIdx;
    float minX, minY, maxX, maxY, invW, invH;
};

struct AreaQuery {       // one GetFeaturesInArea call + the static per-candidate stereo filter
    float x, y, r;
    int minLevel, maxLevel;
    int active;
    float stereoCenter, stereoTol;   // candidate with uRight > 0 is dropped if |stereoCenter - uRight| > stereoTol
};

constexpr int kCells = kGridCols * kGridRows;

// ------------------------------------------------------------------------------------------------ grid build
__global__ void grid_count_kernel(FrameDev f, int* cellOf, int* counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f.n) return;
    const int px = (int)roundf(__fmul_rn(__fsub_rn(f.keys[i].x, f.minX), f.invW));   // PosInGrid, Frame.cc:728-729
    const int py = (int)roundf(__fmul_rn(__fsub_rn(f.keys[i].y, f.minY), f.invH));
    int c = -1;
    if (px >= 0 && px < kGridCols && py >= 0 && py < kGridRows) {
        c = px * kGridRows + py;
        atomicAdd(&counts[c], 1);
    }
    cellOf[i] = c;
}

// exclusive scan of n ints by one block of 1024 threads (n is a few thousand)
__global__ void __launch_bounds__(1024) scan_kernel(const int* in, int* out, int n) {
    __shared__ int warpSums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n ? in[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) warpSums[threadIdx.x >> 5] = incl;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wbase += warpSums[w];
        const int c = carry;
        if (i < n) out[i] = c + wbase + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + wbase + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
}

// exclusive scan of a long array in three launches (the batched searches scan the candidate counts of all jobs at once:
// 512 k entries took 0.53 ms in the one-block kernel above): per-block sums, scan of the sums, per-block scan + offset
constexpr int kScanTile = 4096;      // entries per block (1024 threads x 4)
__global__ void __launch_bounds__(1024) scan_tile_sums_kernel(const int* __restrict__ in, int n, int* __restrict__ sums) {
    __shared__ int warpSums[32];
    const int base = blockIdx.x * kScanTile;
    int v = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = base + k * 1024 + threadIdx.x;
        if (i < n) v += in[i];
    }
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) warpSums[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        const int t = __reduce_add_sync(0xffffffffu, warpSums[threadIdx.x]);
        if (threadIdx.x == 0) sums[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(1024) scan_tiles_kernel(const int* __restrict__ in, int n, const int* __restrict__ tileOffsets,
                                                           int* __restrict__ out) {
    __shared__ int warpSums[32];
    const int base = blockIdx.x * kScanTile + threadIdx.x * 4;     // 4 consecutive entries per thread
    int v[4], t = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = base + k < n ? in[base + k] : 0; t += v[k]; }
    int incl = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += u;
    }
    if ((threadIdx.x & 31) == 31) warpSums[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        const int w = warpSums[threadIdx.x];
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, wi, o);
            if ((int)threadIdx.x >= o) wi += u;
        }
        warpSums[threadIdx.x] = wi - w;
    }
    __syncthreads();
    int run = tileOffsets[blockIdx.x] + warpSums[threadIdx.x >> 5] + incl - t;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 1023) out[n] = run;   // total (entries past n are zeros)
}

__global__ void grid_fill_kernel(int n, const int* cellOf, const int* cellStart, int* cursor, int* cellIdx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cellOf[i];
    if (c >= 0) cellIdx[cellStart[c] + atomicAdd(&cursor[c], 1)] = i;
}

// buckets were filled in arbitrary order; the reference pushes indices in ascending order (Frame.cc:581-588)
__global__ void grid_sort_kernel(const int* cellStart, int* cellIdx) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= kCells) return;
    const int a = cellStart[c], b = cellStart[c + 1];
    for (int i = a + 1; i < b; ++i) {
        const int v = cellIdx[i];
        int j = i - 1;
        while (j >= a && cellIdx[j] > v) { cellIdx[j + 1] = cellIdx[j]; --j; }
        cellIdx[j + 1] = v;
    }
}

// ------------------------------------------------------------------------------------------------ Frame post-extraction
// cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK) on CV_32FC2 points, as Frame::UndistortKeyPoints and
// Frame::ComputeImageBounds call it (Frame.cc:767, :793).  OpenCV widens K and the coefficients to double, normalises,
// runs five fixed-point iterations of the Brown model (TermCriteria(MAX_ITER, 5, 0.01)), applies P = K in double and
// rounds once to float.  Every double operation below is a separately rounded IEEE operation in OpenCV's order
// (its build has no FMA contraction); terms that are identically zero there (k4..k6, thin prism, tilt, R = I) are dropped.
struct CamDev {
    double fx, fy, cx, cy, ifx, ify;
    double k1, k2, p1, p2, k3;
    int distorted;   // mDistCoef.at<float>(0) != 0.0 (Frame.cc:750): otherwise mvKeysUn = mvKeys
};

__device__ __forceinline__ float2 undistort_point(const CamDev& c, float uf, float vf) {
    const double u = (double)uf, v = (double)vf;
    double x = __dmul_rn(__dsub_rn(u, c.cx), c.ifx), y = __dmul_rn(__dsub_rn(v, c.cy), c.ify);
    const double x0 = x, y0 = y;
    const double p1x2 = __dmul_rn(2.0, c.p1), p2x2 = __dmul_rn(2.0, c.p2);
    for (int j = 0; j < 5; ++j) {
        const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
        const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(c.k3, r2), c.k2), r2), c.k1), r2));
        const double icdist = __ddiv_rn(1.0, den);
        if (icdist < 0) {   // OpenCV's guard against a sign flip of the radial factor: keep the normalised input
            x = x0;
            y = y0;
            break;
        }
        const double deltaX = __dadd_rn(__dmul_rn(__dmul_rn(p1x2, x), y), __dmul_rn(c.p2, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x))));
        const double deltaY = __dadd_rn(__dmul_rn(c.p1, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))), __dmul_rn(__dmul_rn(p2x2, x), y));
        x = __dmul_rn(__dsub_rn(x0, deltaX), icdist);
        y = __dmul_rn(__dsub_rn(y0, deltaY), icdist);
    }
    return make_float2((float)__dadd_rn(__dmul_rn(c.fx, x), c.cx), (float)__dadd_rn(__dmul_rn(c.fy, y), c.cy));
}

__global__ void undistort_kernel(CamDev c, const float2* __restrict__ in, int n, float2* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = c.distorted ? undistort_point(c, in[i].x, in[i].y) : in[i];
}

// Frame::Frame after ExtractORB, for keypoints that never left the device: mvKeysUn (UndistortKeyPoints, Frame.cc:748-778),
// a private copy of the descriptors, and the first pass of AssignFeaturesToGrid (PosInGrid, Frame.cc:726-736).
__global__ void frame_import_kernel(CamDev c, const orb_keypoint* __restrict__ srcKeys, const uint4* __restrict__ srcDesc,
                                    int n, orb_keypoint* __restrict__ keys, uint4* __restrict__ desc, float minX, float minY,
                                    float invW, float invH, int* __restrict__ cellOf, int* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    orb_keypoint kp = srcKeys[i];
    if (c.distorted) {
        const float2 p = undistort_point(c, kp.x, kp.y);
        kp.x = p.x;
        kp.y = p.y;
    }
    keys[i] = kp;
    desc[2 * i] = srcDesc[2 * i];
    desc[2 * i + 1] = srcDesc[2 * i + 1];
    const int px = (int)roundf(__fmul_rn(__fsub_rn(kp.x, minX), invW));
    const int py = (int)roundf(__fmul_rn(__fsub_rn(kp.y, minY), invH));
    int cell = -1;
    if (px >= 0 && px < kGridCols && py >= 0 && py < kGridRows) {
        cell = px * kGridRows + py;
        atomicAdd(&counts[cell], 1);
    }
    cellOf[i] = cell;
}

// ------------------------------------------------------------------------------------------------ candidate enumeration
struct CellRange { int c0, c1, r0, r1; bool empty; };

__device__ __forceinline__ CellRange cell_range(const FrameDev& f, float x, float y, float r) {   // Frame.cc:676-690
    CellRange cr;
    cr.empty = true;
    cr.c0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, f.minX), r), f.invW)));
    if (cr.c0 >= kGridCols) return cr;
    cr.c1 = min(kGridCols - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, f.minX), r), f.invW)));
    if (cr.c1 < 0) return cr;
    cr.r0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, f.minY), r), f.invH)));
    if (cr.r0 >= kGridRows) return cr;
    cr.r1 = min(kGridRows - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, f.minY), r), f.invH)));
    if (cr.r1 < 0) return cr;
    cr.empty = false;
    return cr;
}

// One warp walks the query's cells in reference order; `emit(idx, position)` is called by the lane that owns the
// candidate, positions are the candidate's rank in the reference's vIndices. Returns the candidate count.
template <class Emit>
__device__ __forceinline__ int warp_enumerate(const FrameDev& f, const AreaQuery& q, const float* uRight, Emit emit) {
    const int lane = threadIdx.x & 31;
    const CellRange cr = cell_range(f, q.x, q.y, q.r);
    if (cr.empty) return 0;
    const bool checkLevels = (q.minLevel > 0) || (q.maxLevel >= 0);   // Frame.cc:692
    int base = 0;
    // cell id = ix * rows + iy and bucket members are stored cell after cell, so the cells (ix, r0..r1) of one grid
    // column are ONE contiguous run of cellIdx, already in the reference's (iy, bucket) order. The runs' bounds are
    // fetched 32 columns at a time.
    for (int ix0 = cr.c0; ix0 <= cr.c1; ix0 += 32) {
        int ra = 0, rb = 0;
        if (ix0 + lane <= cr.c1) {
            ra = f.cellStart[(ix0 + lane) * kGridRows + cr.r0];
            rb = f.cellStart[(ix0 + lane) * kGridRows + cr.r1 + 1];
        }
        const int ncol = min(32, cr.c1 - ix0 + 1);
        for (int j = 0; j < ncol; ++j) {
            const int a = __shfl_sync(0xffffffffu, ra, j), b = __shfl_sync(0xffffffffu, rb, j);
            for (int k0 = a; k0 < b; k0 += 32) {
                const int k = k0 + lane;
                bool ok = false;
                int idx = -1;
                if (k < b) {
                    idx = f.cellIdx[k];
                    const orb_keypoint kp = f.keys[idx];
                    ok = true;
                    if (checkLevels) {
                        if (kp.octave < q.minLevel) ok = false;
                        if (q.maxLevel >= 0 && kp.octave > q.maxLevel) ok = false;
                    }
                    const float dx = __fsub_rn(kp.x, q.x), dy = __fsub_rn(kp.y, q.y);
                    if (!(fabsf(dx) < q.r && fabsf(dy) < q.r)) ok = false;
                    if (ok && uRight) {
                        const float ur = uRight[idx];
                        if (ur > 0 && fabsf(__fsub_rn(q.stereoCenter, ur)) > q.stereoTol) ok = false;
                    }
                }
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                if (ok) emit(idx, base + __popc(m & ((1u << lane) - 1u)));
                base += __popc(m);
            }
        }
    }
    return base;
}

// count (cand == nullptr) or fill the per-query candidate lists: (train index, distance) in reference order.
// stereoInFill: the stereo filter is static, so it is applied here; the dynamic filters wait for phase 2.
__device__ __forceinline__ void candidates_body(const FrameDev& f, const AreaQuery* __restrict__ queries, const uint4* __restrict__ qdesc,
                                                int nq, const float* __restrict__ uRight, int* __restrict__ counts,
                                                const int* __restrict__ offsets, int2* __restrict__ cand, int qi) {
    if (qi >= nq) return;
    const AreaQuery q = queries[qi];
    int n = 0;
    if (q.active) {
        if (!cand) {
            n = warp_enumerate(f, q, uRight, [](int, int) {});
        } else {
            const uint4 qa = qdesc[2 * qi], qb = qdesc[2 * qi + 1];
            int2* out = cand + offsets[qi];
            n = warp_enumerate(f, q, uRight, [&](int idx, int pos) {
                out[pos] = make_int2(idx | (f.keys[idx].octave << 24), hamming256(qa, qb, f.desc[2 * idx], f.desc[2 * idx + 1]));
            });
        }
    }
    if (!cand && (threadIdx.x & 31) == 0) counts[qi] = n;
}

__global__ void __launch_bounds__(256)
candidates_kernel(FrameDev f, const AreaQuery* __restrict__ queries, const uint4* __restrict__ qdesc, int nq,
                  const float* __restrict__ uRight, int* __restrict__ counts, const int* __restrict__ offsets,
                  int2* __restrict__ cand) {
    candidates_body(f, queries, qdesc, nq, uRight, counts, offsets, cand, blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
}

// Batched searches: job j's queries are entries [qBase, qBase + nq) of ONE concatenated AreaQuery / counts / offsets array,
// so the whole batch shares a single scan and a single candidate buffer (offsets are global positions in it).
struct CandJob {
    FrameDev f;              // frame whose grid is searched
    const uint4* qdesc;      // the job's query descriptors
    const float* uRight;     // or nullptr
    int nq, qBase;
};

__global__ void __launch_bounds__(256)
candidates_batch_kernel(const CandJob* __restrict__ jobs, const AreaQuery* __restrict__ queries, int* __restrict__ counts,
                        const int* __restrict__ offsets, int2* __restrict__ cand) {
    const CandJob& j = jobs[blockIdx.y];
    candidates_body(j.f, queries + j.qBase, j.qdesc, j.nq, j.uRight, counts ? counts + j.qBase : nullptr,
                    offsets ? offsets + j.qBase : nullptr, cand, blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
}

// GetFeaturesInArea as an API of its own (tests, adapter)
__global__ void __launch_bounds__(256)
area_kernel(FrameDev f, const float* __restrict__ xyr, int nq, int minLevel, int maxLevel, int* __restrict__ idxOut,
            int cap, int* __restrict__ countOut) {
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= nq) return;
    AreaQuery q;
    q.x = xyr[3 * qi]; q.y = xyr[3 * qi + 1]; q.r = xyr[3 * qi + 2];
    q.minLevel = minLevel; q.maxLevel = maxLevel; q.active = 1; q.stereoCenter = 0; q.stereoTol = 0;
    int* out = idxOut + (size_t)qi * cap;
    const int n = warp_enumerate(f, q, nullptr, [&](int idx, int pos) { if (pos < cap) out[pos] = idx; });
    if ((threadIdx.x & 31) == 0) countOut[qi] = n;
}

// ------------------------------------------------------------------------------------------------ query builders
__device__ __forceinline__ AreaQuery init_query(const FrameDev& f1, const float* __restrict__ prevXY, int window, int i) {
    AreaQuery a;
    const int level1 = f1.keys[i].octave;
    a.x = prevXY[2 * i]; a.y = prevXY[2 * i + 1]; a.r = (float)window;
    a.minLevel = level1; a.maxLevel = level1;
    a.active = level1 > 0 ? 0 : 1;                        // ORBmatcher.cc:421-423
    a.stereoCenter = 0; a.stereoTol = 0;
    return a;
}

__global__ void init_queries_kernel(FrameDev f1, const float* __restrict__ prevXY, int window, AreaQuery* __restrict__ q) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < f1.n) q[i] = init_query(f1, prevXY, window, i);
}

__device__ __forceinline__ AreaQuery proj_query(const FrameDev& cur, const orbm_proj_query& p, const float* __restrict__ sf, float th,
                                                int mode, float mbf) {
    AreaQuery a;
    a.x = p.u; a.y = p.v;
    a.active = p.valid != 0;
    if (p.u < cur.minX || p.u > cur.maxX) a.active = 0;   // ORBmatcher.cc:1390-1393
    if (p.v < cur.minY || p.v > cur.maxY) a.active = 0;
    // a skipped query (valid == 0) may carry any octave: the header says the flag gates the whole entry, so the scale
    // table is only indexed for active ones (the host has range-checked those)
    const int o = a.active ? p.octave : 0;
    a.r = __fmul_rn(th, sf[o]);                           // :1398
    if (mode == 1) { a.minLevel = o; a.maxLevel = -1; }   // forward  (:1409)
    else if (mode == 2) { a.minLevel = 0; a.maxLevel = o; }   // backward (:1411)
    else if (mode == 3) { a.minLevel = o - 1; a.maxLevel = o; }   // SearchByProjection(KF, Scw, ...) (:369-371)
    else { a.minLevel = o - 1; a.maxLevel = o + 1; }      // :1413
    a.stereoCenter = __fsub_rn(p.u, __fmul_rn(mbf, p.invz));   // :1435
    a.stereoTol = a.r;
    return a;
}

__global__ void proj_queries_kernel(FrameDev cur, const orbm_proj_query* __restrict__ pq, int nq,
                                    const float* __restrict__ sf, float th, int mode, float mbf, AreaQuery* __restrict__ q) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) q[i] = proj_query(cur, pq[i], sf, th, mode, mbf);
}

// ORBmatcher.cc:1376-1393 on the device: x3Dc = Rcw * x3Dw + tcw is cv::gemm on CV_32F 3x3 . 3x1 + 3x1 operands, whose
// small-matrix path sums the three products in float in source order and adds the addend in double with one rounding
// (oracle/cvprims.cpp::gemm3_f32, pinned against cv2.gemm); invzc = 1.0 / z is a double division rounded once; u and v
// are float expressions in source order.  The image-bounds test (:1390-1393) stays in proj_query.
struct PoseDev { float R[9], t[3], fx, fy, cx, cy; };
__global__ void project_world_kernel(PoseDev P, const orbm_world_query* __restrict__ wq, int nq, orbm_proj_query* __restrict__ pq) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const orbm_world_query w = wq[i];
    float c[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float s = __fadd_rn(__fadd_rn(__fmul_rn(P.R[3 * r], w.x), __fmul_rn(P.R[3 * r + 1], w.y)), __fmul_rn(P.R[3 * r + 2], w.z));
        c[r] = (float)__dadd_rn((double)s, (double)P.t[r]);
    }
    const float invz = (float)__ddiv_rn(1.0, (double)c[2]);
    orbm_proj_query o;
    o.u = __fadd_rn(__fmul_rn(__fmul_rn(P.fx, c[0]), invz), P.cx);
    o.v = __fadd_rn(__fmul_rn(__fmul_rn(P.fy, c[1]), invz), P.cy);
    o.invz = invz;
    o.octave = w.octave;
    o.valid = (w.valid != 0 && !(invz < 0.f)) ? 1 : 0;    // :1383-1384; a NaN depth passes here and fails the bounds test
    o.obs_positive = w.obs_positive;
    o.angle = w.angle;
    pq[i] = o;
}

__global__ void point_queries_kernel(const orbm_point_query* __restrict__ pq, int nq, const float* __restrict__ sf, float th,
                                     AreaQuery* __restrict__ q) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const orbm_point_query p = pq[i];
    AreaQuery a;
    float r = ((double)p.view_cos > 0.998) ? 2.5f : 4.0f;   // RadiusByViewingCos, ORBmatcher.cc:131-137
    if (th != 1.0f) r = __fmul_rn(r, th);                    // bFactor (:49, 65-66)
    a.x = p.proj_x; a.y = p.proj_y;
    a.active = p.in_view != 0;
    const int level = a.active ? p.level : 0;                // a skipped entry's level is not looked at (see above)
    a.r = __fmul_rn(r, sf[level]);
    a.minLevel = level - 1; a.maxLevel = level;              // :69
    a.stereoCenter = p.proj_xr; a.stereoTol = a.r;           // :94-99
    q[i] = a;
}

// ------------------------------------------------------------------------------------------------ phase 2: replays
constexpr int kOrdShift = 22;
constexpr int kOrdMask = (1 << kOrdShift) - 1;
constexpr int kNone = 0x7fffffff;

constexpr int kCandIdxMask = 0x00ffffff;   // candidate.x = keypoint index | octave << 24

// Rotation-histogram tail shared by the searches (e.g. ORBmatcher.cc:473-512).  Pushes were recorded in order as
// (angle source A, angle source B, value to clear); bins are computed here, in parallel, because only their COUNTS
// matter until the end.  guardMatched: SearchForInitialization decrements only for still-matched entries (:503-507);
// the projection searches decrement per push, duplicates included (:1483-1491).
template <class AngleA, class AngleB>
__device__ void histogram_prune(const int* pushA, const int* pushB, const int* pushVal, int nPush, int* hist, int* target,
                                bool guardMatched, int* nmatches, AngleA angleA, AngleB angleB) {
    // whole CTA; the caller has synchronised after the last push
    const int tid = threadIdx.x, nthr = blockDim.x;
    __shared__ int keep[3];
    __shared__ int removed;
    if (tid == 0) removed = 0;
    for (int k = tid; k < nPush; k += nthr) atomicAdd(&hist[rotation_bin(angleA(pushA[k]), angleB(pushB[k]))], 1);
    __syncthreads();
    if (tid == 0) three_maxima(hist, keep[0], keep[1], keep[2]);
    __syncthreads();
    for (int k = tid; k < nPush; k += nthr) {
        const int b = rotation_bin(angleA(pushA[k]), angleB(pushB[k]));
        if (b == keep[0] || b == keep[1] || b == keep[2]) continue;
        const int v = pushVal[k];
        if (guardMatched) {
            if (target[v] >= 0) { target[v] = -1; atomicAdd(&removed, 1); }   // each value is pushed at most once here
        } else {
            target[v] = -1;
            atomicAdd(&removed, 1);
        }
    }
    __syncthreads();
    if (tid == 0) *nmatches -= removed;
    __syncthreads();
}

// Ordered walk over the queries by ONE CTA.  The order-dependent state (what earlier queries matched) lives in shared
// memory and is indexed by TARGET keypoint; a query reads it only at its own candidates and changes it only at the
// candidate it accepts.  That makes the sequential loop speculatable:
//   round: the next 16 unresolved queries are evaluated at once, one warp each (candidates spread over the lanes, best /
//          second by two warp reductions -- keys carry the candidate's position, hence the owning lane), all against
//          the state left by the committed queries.  Every tentative acceptor stamps its target with its rank in the
//          round (shared atomicMin).  State changes only ever REMOVE candidates, so a query's outcome can change only
//          if its best or second-best target goes: it is DIRTY if one of the two carries a smaller stamp than its own
//          rank.  The clean prefix of the round commits -- its results are the sequential ones -- and the next round
//          starts at the first dirty query.  The first query of a round is never dirty, so a round
//          retires at least one query and, with conflicts as rare as they are between 2000 keypoints, nearly all 16.
// Queries and their candidate lists (contiguous in the CSR) are staged in shared memory a chunk at a time with coalesced
// loads, so the per-round chain is shared-memory lookups, warp reductions and three barriers -- no global round trip.
//   meta(qi)                       small per-query integer needed by decide / commit (staged with the chunk)
//   skip(c)                        candidate c is ignored (dynamic state)
//   decide(qi, meta, best, second, bestX, secondX) -> accept?   pure; bestX/secondX = candidate.x of the winners
//                                  (kUsesSecond = false promises that it ignores second / secondX)
//   commit(qi, meta, best, bestX, pushPos) -> change of the match count; run by ONE thread per accepted query
constexpr int RP_WARPS = 16;                 // queries in flight per round
constexpr int RP_THREADS = RP_WARPS * 32;
constexpr int kChunkQueries = 256;           // queries staged per chunk
constexpr int kStageCand = 6144;             // candidates staged per chunk (48 KB)
constexpr int kReplayFixedInts = kStageCand * 2 + kChunkQueries * 4;   // staging area at the start of dynamic smem

template <bool kUsesSecond, class Meta, class Skip, class Decide, class Commit>
__device__ __forceinline__ void replay_queries(const AreaQuery* __restrict__ q, const int* __restrict__ offsets,
                                               const int2* __restrict__ cand, int nq, int* stamp, int nTargets, int& nPush,
                                               int* nMatchesShared, Meta meta, Skip skip, Decide decide, Commit commit) {
    extern __shared__ int dyn[];
    int2* stage = reinterpret_cast<int2*>(dyn);
    int4* qmeta = reinterpret_cast<int4*>(dyn + kStageCand * 2);   // candidate range [x, y), active, meta
    __shared__ int4 sRound[32];   // per query of the round: accepted?, best target, second-best target
    __shared__ int sVerdict[2];   // retired queries, mask of the committing ones
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // two stamp arrays, used by alternate rounds: a round's stamps are cleared during the next round, when nobody reads them
    for (int i = tid; i < 2 * nTargets; i += RP_THREADS) stamp[i] = INT_MAX;
    int base = 0, chunkBegin = 0, chunkEnd = 0, stageA = 0, parity = 0, prevTarget = -1, myMatches = 0;
    if (tid < 32) sRound[tid] = make_int4(0, -1, -1, 0);
    bool staged = true;

    // one round over queries [base, base + nr); `src[k]` is candidate k of the CSR. Returns the number of queries retired.
    auto round = [&](const int2* src, int nr) {
        int* S = stamp + parity * nTargets;
        if (lane == 0 && prevTarget >= 0) stamp[(parity ^ 1) * nTargets + prevTarget] = INT_MAX;
        prevTarget = -1;
        // A. speculative evaluation, one warp per query
        bool ok = false;
        int gb = kNone, bestX = -1, target = -1, target2 = -1;
        int4 mq = make_int4(0, 0, 0, 0);
        if (warp < nr) mq = qmeta[base - chunkBegin + warp];
        if (warp < nr && mq.z && mq.x < mq.y) {
            int bk = kNone, sk = kNone, bx = -1, sx = -1;   // this lane's best / second key and their candidate.x
            for (int k = mq.x + lane; k < mq.y; k += 32) {
                const int2 c = src[k];
                if (skip(c)) continue;
                const int key = (c.y << kOrdShift) | (k - mq.x);
                if (key < bk) { sk = bk; sx = bx; bk = key; bx = c.x; }
                else if (key < sk) { sk = key; sx = c.x; }
            }
            // keys are unique and carry the candidate's position, whose low five bits are the owning lane
            gb = __reduce_min_sync(0xffffffffu, bk);
            if (gb != kNone) {
                const int gs = __reduce_min_sync(0xffffffffu, bk == gb ? sk : bk);
                bestX = __shfl_sync(0xffffffffu, bx, gb & 31);
                const int secondX = __shfl_sync(0xffffffffu, bk == gs ? bx : sx, gs & 31);   // meaningful iff gs != kNone
                target = bestX & kCandIdxMask;
                if (kUsesSecond && gs != kNone) target2 = secondX & kCandIdxMask;
                ok = decide(base + warp, mq.w, gb, gs, bestX, secondX);
                if (ok && lane == 0) { atomicMin(&S[target], warp); prevTarget = target; }
            }
        }
        if (lane == 0) sRound[warp] = make_int4(ok, target, target2, 0);
        __syncthreads();

        // B. state only ever REMOVES candidates (occupied / matched flags are set, matched distances shrink), so a
        //    query's outcome can change only if an earlier query of the round takes its best or its second-best target.
        //    Warp 0 works this out for the whole round (lane = query).
        if (warp == 0) {
            const int4 r = sRound[lane];
            const bool in = lane < nr;
            const int s1 = S[max(r.y, 0)], s2 = S[max(r.z, 0)];
            const bool dirty = in && ((r.y >= 0 && s1 < lane) || (r.z >= 0 && s2 < lane));
            const unsigned dirtyMask = __ballot_sync(0xffffffffu, dirty);
            const int f = dirtyMask ? __ffs(dirtyMask) - 1 : nr;
            const unsigned okPrefix = __ballot_sync(0xffffffffu, lane < f && r.x);
            if (lane == 0) { sVerdict[0] = f; sVerdict[1] = (int)okPrefix; }
        }
        __syncthreads();

        // C. commit the clean prefix
        const int f = sVerdict[0];
        const unsigned okPrefix = (unsigned)sVerdict[1];
        if (ok && lane == 0 && warp < f) myMatches += commit(base + warp, mq.w, gb, bestX, nPush + __popc(okPrefix & ((1u << warp) - 1u)));
        nPush += __popc(okPrefix);
        parity ^= 1;
        __syncthreads();
        return f;
    };

    while (base < nq) {
        if (base == chunkEnd) {
            // stage the next chunk: as many queries (<= 256) as have their candidates fit the buffer
            const int first = offsets[base];
            const int qi = base + tid;
            int a = 0, b = 0;
            bool fits = false;
            if (tid < kChunkQueries && qi < nq) { a = offsets[qi]; b = offsets[qi + 1]; fits = b - first <= kStageCand; }
            const int cnt = __syncthreads_count(fits);     // offsets are monotone: the fitting queries are a prefix
            staged = cnt > 0;
            chunkBegin = base;
            chunkEnd = base + max(cnt, 1);
            stageA = first;
            if (tid < chunkEnd - base) qmeta[tid] = make_int4(a, b, q[qi].active, meta(qi));
            if (staged) {
                const int total = offsets[chunkEnd] - first;
#pragma unroll 4
                for (int k = tid; k < total; k += RP_THREADS) stage[k] = cand[first + k];
            }
            __syncthreads();
        }
        const int nr = min(RP_WARPS, chunkEnd - base);
        base += staged ? round(stage - stageA, nr) : round(cand, nr);
    }
    if (myMatches) atomicAdd(nMatchesShared, myMatches);
}

// SearchForInitialization replay (ORBmatcher.cc:417-517). m21 / vMatchedDistance in shared memory.
__device__ __forceinline__ void init_replay_body(const FrameDev& f1, const FrameDev& f2, const AreaQuery* __restrict__ q,
                                                 const int* __restrict__ offsets, const int2* __restrict__ cand, float ratio,
                                                 int checkOri, float* prevXY, int* m12, int* pushA, int* pushB, int* nmatchesOut) {
    extern __shared__ int dyn[];
    int* m21 = dyn + kReplayFixedInts;
    int* matchedDist = m21 + f2.n;
    int* stamp = m21 + 2 * f2.n;   // 2 * f2.n entries
    __shared__ int hist[kHistoLength];
    __shared__ int nmatches;
    const int tid = threadIdx.x;
    if (tid < kHistoLength) hist[tid] = 0;
    if (tid == 0) nmatches = 0;
    for (int i = tid; i < f1.n; i += RP_THREADS) m12[i] = -1;
    for (int i = tid; i < f2.n; i += RP_THREADS) { m21[i] = -1; matchedDist[i] = INT_MAX; }
    int nPush = 0;
    replay_queries<true>(q, offsets, cand, f1.n, stamp, f2.n, nPush, &nmatches, [](int) { return 0; },
        [&](const int2& c) { return matchedDist[c.x & kCandIdxMask] <= c.y; },               // :444
        [&](int, int, int best, int second, int, int) {
            const int bd = best >> kOrdShift;
            const int sd = second == kNone ? INT_MAX : second >> kOrdShift;
            return bd <= kThLow && (float)bd < __fmul_rn((float)sd, ratio);                   // :459-461
        },
        [&](int i1, int, int best, int bestX, int pos) {
            const int i2 = bestX & kCandIdxMask;
            int delta = 1;
            if (m21[i2] >= 0) { m12[m21[i2]] = -1; delta = 0; }                               // :463-467
            m12[i1] = i2; m21[i2] = i1; matchedDist[i2] = best >> kOrdShift;
            if (checkOri) { pushA[pos] = i1; pushB[pos] = i2; }
            return delta;
        });
    __threadfence_block();
    __syncthreads();
    if (checkOri)
        histogram_prune(pushA, pushB, pushA, nPush, hist, m12, true, &nmatches,
                        [&](int i1) { return f1.keys[i1].angle; }, [&](int i2) { return f2.keys[i2].angle; });
    if (tid == 0) *nmatchesOut = nmatches;
    for (int i1 = tid; i1 < f1.n; i1 += RP_THREADS)
        if (m12[i1] >= 0) {                                             // :515-517
            prevXY[2 * i1] = f2.keys[m12[i1]].x;
            prevXY[2 * i1 + 1] = f2.keys[m12[i1]].y;
        }
}

__global__ void __launch_bounds__(RP_THREADS)
init_replay_kernel(FrameDev f1, FrameDev f2, const AreaQuery* __restrict__ q, const int* __restrict__ offsets,
                   const int2* __restrict__ cand, float ratio, int checkOri, float* prevXY, int* m12, int* pushA,
                   int* pushB, int* nmatchesOut) {
    init_replay_body(f1, f2, q, offsets, cand, ratio, checkOri, prevXY, m12, pushA, pushB, nmatchesOut);
}

struct InitJob {           // one frame pair of orbm_search_for_initialization_batch (all pointers device memory)
    FrameDev f1, f2;
    float* prevXY;
    int* m12;
    int* pushA;
    int* pushB;
    int* nmatchesOut;
    int qBase;
};
__global__ void __launch_bounds__(RP_THREADS)
init_queries_batch_kernel(const InitJob* __restrict__ jobs, int window, AreaQuery* __restrict__ q) {
    const InitJob& j = jobs[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < j.f1.n) q[j.qBase + i] = init_query(j.f1, j.prevXY, window, i);
}
__global__ void __launch_bounds__(RP_THREADS)
init_replay_batch_kernel(const InitJob* __restrict__ jobs, const AreaQuery* __restrict__ q, const int* __restrict__ offsets,
                         const int2* __restrict__ cand, float ratio, int checkOri) {
    const InitJob j = jobs[blockIdx.x];
    init_replay_body(j.f1, j.f2, q + j.qBase, offsets + j.qBase, cand, ratio, checkOri, j.prevXY, j.m12, j.pushA, j.pushB, j.nmatchesOut);
}

// SearchByProjection(Frame, Frame) replay (ORBmatcher.cc:1363-1495), also the relocalisation overload (:1500-1627, where
// the acceptance threshold is ORBdist instead of TH_HIGH). Occupancy flags in shared memory.
__device__ __forceinline__ void proj_replay_body(const FrameDev& cur, const AreaQuery* __restrict__ q, const orbm_proj_query* __restrict__ pq,
                                                 int nq, const int* __restrict__ offsets, const int2* __restrict__ cand, int checkOri,
                                                 int maxDist, const unsigned char* __restrict__ occIn, int* curMatch, int* pushA,
                                                 int* pushB, int* nmatchesOut) {
    extern __shared__ int dyn[];
    int* stamp = dyn + kReplayFixedInts;
    unsigned char* occ = reinterpret_cast<unsigned char*>(stamp + 2 * cur.n);
    __shared__ int hist[kHistoLength];
    __shared__ int nmatches;
    const int tid = threadIdx.x;
    if (tid < kHistoLength) hist[tid] = 0;
    if (tid == 0) nmatches = 0;
    for (int i = tid; i < cur.n; i += RP_THREADS) { curMatch[i] = -1; occ[i] = occIn[i]; }
    int nPush = 0;
    replay_queries<false>(q, offsets, cand, nq, stamp, cur.n, nPush, &nmatches, [&](int qi) { return pq[qi].obs_positive; },
        [&](const int2& c) { return occ[c.x & kCandIdxMask] != 0; },                          // :1428-1430
        [&](int, int, int best, int, int, int) { return (best >> kOrdShift) <= maxDist; },    // :1453 / :1583
        [&](int i, int obs, int, int bestX, int pos) {
            const int i2 = bestX & kCandIdxMask;
            curMatch[i2] = i;
            occ[i2] = obs ? 1 : 0;
            if (checkOri) { pushA[pos] = i; pushB[pos] = i2; }
            return 1;
        });
    __threadfence_block();
    __syncthreads();
    if (checkOri)
        histogram_prune(pushA, pushB, pushB, nPush, hist, curMatch, false, &nmatches,
                        [&](int i) { return pq[i].angle; }, [&](int i2) { return cur.keys[i2].angle; });
    if (tid == 0) *nmatchesOut = nmatches;
}

__global__ void __launch_bounds__(RP_THREADS)
proj_replay_kernel(FrameDev cur, const AreaQuery* __restrict__ q, const orbm_proj_query* __restrict__ pq, int nq,
                   const int* __restrict__ offsets, const int2* __restrict__ cand, int checkOri, int maxDist,
                   const unsigned char* __restrict__ occIn, int* curMatch, int* pushA, int* pushB, int* nmatchesOut) {
    proj_replay_body(cur, q, pq, nq, offsets, cand, checkOri, maxDist, occIn, curMatch, pushA, pushB, nmatchesOut);
}

struct ProjJob {           // one (current frame, query set) of orbm_search_by_projection_batch (all pointers device memory)
    FrameDev cur;
    const orbm_proj_query* pq;
    const unsigned char* occIn;
    int* curMatch;
    int* pushA;
    int* pushB;
    int* nmatchesOut;
    int nq, qBase;
};
__global__ void proj_queries_batch_kernel(const ProjJob* __restrict__ jobs, const float* __restrict__ sf, float th, int mode, float mbf,
                                          AreaQuery* __restrict__ q) {
    const ProjJob& j = jobs[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < j.nq) q[j.qBase + i] = proj_query(j.cur, j.pq[i], sf, th, mode, mbf);
}
__global__ void __launch_bounds__(RP_THREADS)
proj_replay_batch_kernel(const ProjJob* __restrict__ jobs, const AreaQuery* __restrict__ q, const int* __restrict__ offsets,
                         const int2* __restrict__ cand, int checkOri, int maxDist) {
    const ProjJob j = jobs[blockIdx.x];
    proj_replay_body(j.cur, q + j.qBase, j.pq, j.nq, offsets + j.qBase, cand, checkOri, maxDist, j.occIn, j.curMatch, j.pushA, j.pushB,
                     j.nmatchesOut);
}

// SearchByProjection(Frame, MapPoints) replay (ORBmatcher.cc:51-126).
__global__ void __launch_bounds__(RP_THREADS)
point_replay_kernel(FrameDev f, const AreaQuery* __restrict__ q, const orbm_point_query* __restrict__ pq, int nq,
                    const int* __restrict__ offsets, const int2* __restrict__ cand, float ratio,
                    const unsigned char* __restrict__ occIn, int* match, int* nmatchesOut) {
    extern __shared__ int dyn[];
    int* stamp = dyn + kReplayFixedInts;
    unsigned char* occ = reinterpret_cast<unsigned char*>(stamp + 2 * f.n);
    __shared__ int nmatches;
    const int tid = threadIdx.x;
    if (tid == 0) nmatches = 0;
    for (int i = tid; i < f.n; i += RP_THREADS) { match[i] = -1; occ[i] = occIn[i]; }
    int nPush = 0;
    replay_queries<true>(q, offsets, cand, nq, stamp, f.n, nPush, &nmatches, [&](int qi) { return pq[qi].obs_positive; },
        [&](const int2& c) { return occ[c.x & kCandIdxMask] != 0; },                          // :84-86
        [&](int, int, int best, int second, int bestX, int secondX) {
            const int bd = best >> kOrdShift;
            if (bd > kThHigh) return false;                                                   // :115
            const int bestLevel = bestX >> 24;
            int sd = 256, secondLevel = -1;
            if (second != kNone) { sd = second >> kOrdShift; secondLevel = secondX >> 24; }
            return !(bestLevel == secondLevel && (float)bd > __fmul_rn(ratio, (float)sd));    // :118-119
        },
        [&](int i, int obs, int, int bestX, int) {
            const int bi = bestX & kCandIdxMask;
            match[bi] = i;
            occ[bi] = obs ? 1 : 0;
            return 1;
        });
    __syncthreads();
    if (tid == 0) *nmatchesOut = nmatches;
}

// ------------------------------------------------------------------------------------------------ triangulation
struct TriParams {
    FrameDev k1, k2;
    int nNodes1, nNodes2, nEntries1;
    const int *nodeId1, *start1, *idx1, *nodeId2, *start2, *idx2;
    const unsigned char *has1, *has2;
    const float *uR1, *uR2;
    float F[9];
    float ex, ey;
    const float *sf2, *sigma2;
    int onlyStereo, checkOri;
};

__device__ __forceinline__ bool epipolar_ok(const orb_keypoint& k1, const orb_keypoint& k2, const float* F, const float* sigma2) {
    // l = x1' F12 (ORBmatcher.cc:143-145), all float32, left to right, no contraction
    const float a = __fadd_rn(__fadd_rn(__fmul_rn(k1.x, F[0]), __fmul_rn(k1.y, F[3])), F[6]);
    const float b = __fadd_rn(__fadd_rn(__fmul_rn(k1.x, F[1]), __fmul_rn(k1.y, F[4])), F[7]);
    const float c = __fadd_rn(__fadd_rn(__fmul_rn(k1.x, F[2]), __fmul_rn(k1.y, F[5])), F[8]);
    const float num = __fadd_rn(__fadd_rn(__fmul_rn(a, k2.x), __fmul_rn(b, k2.y)), c);
    const float den = __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
    if (den == 0) return false;
    const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
    return (double)dsqr < __dmul_rn(3.84, (double)sigma2[k2.octave]);   // :156, compared in double
}

// one thread per feature-vector entry of KF1: node lookup by binary search replaces the map merge-walk (:691-789)
__global__ void tri_match_kernel(TriParams P, const int* __restrict__ entryNode, int* __restrict__ m12, int* hist) {
    const int p1 = blockIdx.x * blockDim.x + threadIdx.x;
    if (p1 >= P.nEntries1) return;
    const int node = P.nodeId1[entryNode[p1]];
    int lo = 0, hi = P.nNodes2 - 1, b = -1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1, v = P.nodeId2[mid];
        if (v == node) { b = mid; break; }
        if (v < node) lo = mid + 1; else hi = mid - 1;
    }
    if (b < 0) return;
    const int idx1 = P.idx1[p1];
    if (P.has1[idx1]) return;
    const bool stereo1 = P.uR1 ? (P.uR1[idx1] >= 0) : false;
    if (P.onlyStereo && !stereo1) return;
    const orb_keypoint kp1 = P.k1.keys[idx1];
    const uint4 da = P.k1.desc[2 * idx1], db = P.k1.desc[2 * idx1 + 1];
    int best = kThLow, bestIdx = -1;
    for (int p2 = P.start2[b]; p2 < P.start2[b + 1]; ++p2) {
        const int idx2 = P.idx2[p2];
        if (P.has2[idx2]) continue;
        const bool stereo2 = P.uR2 ? (P.uR2[idx2] >= 0) : false;
        if (P.onlyStereo && !stereo2) continue;
        const int dist = hamming256(da, db, P.k2.desc[2 * idx2], P.k2.desc[2 * idx2 + 1]);
        if (dist > kThLow || dist > best) continue;                    // :738 (ties replace)
        const orb_keypoint kp2 = P.k2.keys[idx2];
        if (!stereo1 && !stereo2) {
            const float dex = __fsub_rn(P.ex, kp2.x), dey = __fsub_rn(P.ey, kp2.y);
            if (__fadd_rn(__fmul_rn(dex, dex), __fmul_rn(dey, dey)) < __fmul_rn(100.f, P.sf2[kp2.octave])) continue;   // :747
        }
        if (epipolar_ok(kp1, kp2, P.F, P.sigma2)) { bestIdx = idx2; best = dist; }
    }
    if (bestIdx >= 0) {
        m12[idx1] = bestIdx;
        if (P.checkOri) atomicAdd(&hist[rotation_bin(kp1.angle, P.k2.keys[bestIdx].angle)], 1);
    }
}

__global__ void tri_entry_node_kernel(int nNodes, const int* start, int* entryNode) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= nNodes) return;
    for (int p = start[a]; p < start[a + 1]; ++p) entryNode[p] = a;
}

__global__ void __launch_bounds__(256) tri_finish_kernel(FrameDev k1, FrameDev k2, int checkOri, int* m12, int* hist, int* nmatchesOut) {
    __shared__ int keep[3];
    __shared__ int total;
    if (threadIdx.x == 0) {
        total = 0;
        if (checkOri) three_maxima(hist, keep[0], keep[1], keep[2]);
    }
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < k1.n; i += blockDim.x) {
        const int m = m12[i];
        if (m < 0) continue;
        if (checkOri) {
            const int bin = rotation_bin(k1.keys[i].angle, k2.keys[m].angle);
            if (bin != keep[0] && bin != keep[1] && bin != keep[2]) { m12[i] = -1; continue; }
        }
        ++mine;
    }
    atomicAdd(&total, mine);
    __syncthreads();
    if (threadIdx.x == 0) *nmatchesOut = total;
}

__global__ void fill_int_kernel(int* p, int n, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ------------------------------------------------------------------------------------------------ projected best match
// The independent search shared by Fuse (ORBmatcher.cc:892-944), Fuse with Sim3 (:1051-1075) and both directions of
// SearchBySim3 (:1191-1215, :1271-1295): no query sees another's result, so one warp per query, whole GPU.
__global__ void __launch_bounds__(256)
projected_best_kernel(FrameDev f, const orbm_best_query* __restrict__ queries, const uint4* __restrict__ qdesc, int nq,
                      int chi2Filter, const float* __restrict__ uRight, const float* __restrict__ invSigma2,
                      int* __restrict__ bestIdx, int* __restrict__ bestDist) {
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= nq) return;
    const int lane = threadIdx.x & 31;
    const orbm_best_query bq = queries[qi];
    int best = kNone, bx = -1;
    if (bq.valid) {
        AreaQuery a;
        a.x = bq.u; a.y = bq.v; a.r = bq.radius;
        a.minLevel = bq.level - 1; a.maxLevel = bq.level;   // the in-loop octave test of the reference (:908-911)
        a.active = 1; a.stereoCenter = 0; a.stereoTol = 0;
        const uint4 qa = qdesc[2 * qi], qb = qdesc[2 * qi + 1];
        warp_enumerate(f, a, nullptr, [&](int idx, int pos) {
            if (chi2Filter) {                                // :914-938
                const orb_keypoint kp = f.keys[idx];
                const float ex = __fsub_rn(bq.u, kp.x), ey = __fsub_rn(bq.v, kp.y);
                float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                double limit = 5.99;
                if (uRight && uRight[idx] >= 0) {
                    const float er = __fsub_rn(bq.ur, uRight[idx]);
                    e2 = __fadd_rn(e2, __fmul_rn(er, er));
                    limit = 7.8;
                }
                if ((double)__fmul_rn(e2, invSigma2[kp.octave]) > limit) return;
            }
            const int key = (hamming256(qa, qb, f.desc[2 * idx], f.desc[2 * idx + 1]) << kOrdShift) | pos;
            if (key < best) { best = key; bx = idx; }
        });
    }
    int g = best;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) g = min(g, __shfl_xor_sync(0xffffffffu, g, o));
    const unsigned owner = __ballot_sync(0xffffffffu, best == g && g != kNone);
    const int idx = owner ? __shfl_sync(0xffffffffu, bx, __ffs(owner) - 1) : -1;
    if (lane == 0) {
        bestIdx[qi] = idx;
        bestDist[qi] = g == kNone ? 256 : g >> kOrdShift;
    }
}

// ------------------------------------------------------------------------------------------------ SearchByBoW
// Both overloads (ORBmatcher.cc:159-288 KeyFrame->Frame, 522-655 KeyFrame->KeyFrame): queries are the feature-vector
// entries of frame 1 in map order (node id ascending, list order inside a node), candidates the entries of the same
// node in frame 2.  Phase 1 (one warp per entry): node lookup, static validity, distances.  Phase 2: the ordered
// one-warp replay with the "already matched" flags of frame 2 in shared memory.
struct BowParams {
    FrameDev k1, k2;
    int nNodes2, nEntries1;
    const int *nodeId1, *idx1, *nodeId2, *start2, *idx2;
    const unsigned char *valid1, *valid2;   // may be null: everything valid
};

__global__ void __launch_bounds__(256)
bow_candidates_kernel(BowParams P, const int* __restrict__ entryNode, AreaQuery* __restrict__ q, int* __restrict__ counts,
                      const int* __restrict__ offsets, int2* __restrict__ cand) {
    const int p1 = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p1 >= P.nEntries1) return;
    const int lane = threadIdx.x & 31;
    const int node = P.nodeId1[entryNode[p1]];
    int lo = 0, hi = P.nNodes2 - 1, b = -1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1, v = P.nodeId2[mid];
        if (v == node) { b = mid; break; }
        if (v < node) lo = mid + 1; else hi = mid - 1;
    }
    const int idx1 = P.idx1[p1];
    const bool active = b >= 0 && (!P.valid1 || P.valid1[idx1]);
    int n = 0;
    if (active) {
        const uint4 qa = P.k1.desc[2 * idx1], qb = P.k1.desc[2 * idx1 + 1];
        int2* out = cand ? cand + offsets[p1] : nullptr;
        for (int k0 = P.start2[b]; k0 < P.start2[b + 1]; k0 += 32) {
            const int k = k0 + lane;
            int idx2 = -1;
            bool ok = false;
            if (k < P.start2[b + 1]) {
                idx2 = P.idx2[k];
                ok = !P.valid2 || P.valid2[idx2];
            }
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (ok && out)
                out[n + __popc(m & ((1u << lane) - 1u))] =
                    make_int2(idx2 | (P.k2.keys[idx2].octave << 24), hamming256(qa, qb, P.k2.desc[2 * idx2], P.k2.desc[2 * idx2 + 1]));
            n += __popc(m);
        }
    }
    if (!cand && lane == 0) {
        counts[p1] = n;
        q[p1].active = active ? 1 : 0;
    }
}

__global__ void __launch_bounds__(RP_THREADS)
bow_replay_kernel(FrameDev k1, FrameDev k2, const AreaQuery* __restrict__ q, const int* __restrict__ idx1OfEntry, int nq,
                  const int* __restrict__ offsets, const int2* __restrict__ cand, float ratio, int checkOri, int strictLow,
                  int* m12, int* m21, int* pushA, int* pushB, int* nmatchesOut) {
    extern __shared__ int dyn[];
    int* stamp = dyn + kReplayFixedInts;
    unsigned char* matched2 = reinterpret_cast<unsigned char*>(stamp + 2 * k2.n);
    __shared__ int hist[kHistoLength];
    __shared__ int nmatches;
    const int tid = threadIdx.x;
    if (tid < kHistoLength) hist[tid] = 0;
    if (tid == 0) nmatches = 0;
    for (int i = tid; i < k1.n; i += RP_THREADS) m12[i] = -1;
    for (int i = tid; i < k2.n; i += RP_THREADS) { m21[i] = -1; matched2[i] = 0; }
    int nPush = 0;
    replay_queries<true>(q, offsets, cand, nq, stamp, k2.n, nPush, &nmatches, [&](int p1) { return idx1OfEntry[p1]; },
        [&](const int2& c) { return matched2[c.x & kCandIdxMask] != 0; },                     // :203-204 / :576
        [&](int, int, int best, int second, int, int) {
            const int bd = best >> kOrdShift;
            const int sd = second == kNone ? 256 : second >> kOrdShift;
            const bool low = strictLow ? bd < kThLow : bd <= kThLow;                          // :598 / :227
            return low && (float)bd < __fmul_rn(ratio, (float)sd);                            // :229 / :600
        },
        [&](int, int idx1, int, int bestX, int pos) {
            const int i2 = bestX & kCandIdxMask;
            m12[idx1] = i2; m21[i2] = idx1; matched2[i2] = 1;
            if (checkOri) { pushA[pos] = idx1; pushB[pos] = i2; }
            return 1;
        });
    __threadfence_block();
    __syncthreads();
    if (checkOri) {
        histogram_prune(pushA, pushB, pushA, nPush, hist, m12, false, &nmatches,
                        [&](int i1) { return k1.keys[i1].angle; }, [&](int i2) { return k2.keys[i2].angle; });
        __threadfence_block();
        __syncthreads();
        for (int k = tid; k < nPush; k += RP_THREADS)
            if (m12[pushA[k]] < 0) m21[pushB[k]] = -1;
    }
    if (tid == 0) *nmatchesOut = nmatches;
}

}  // namespace orbb

// =================================================================================================== host side
using namespace orbb;

struct orbm_frame_s {
    orbm_matcher* m = nullptr;
    int n = 0;
    DevBuf keys, desc, cellStart, cellIdx;
    float minX = 0, minY = 0, maxX = 0, maxY = 0, invW = 0, invH = 0;
    FrameDev dev() const {
        FrameDev f;
        f.n = n; f.keys = keys.as<orb_keypoint>(); f.desc = desc.as<uint4>();
        f.cellStart = cellStart.as<int>(); f.cellIdx = cellIdx.as<int>();
        f.minX = minX; f.minY = minY; f.maxX = maxX; f.maxY = maxY; f.invW = invW; f.invH = invH;
        return f;
    }
};

namespace {

constexpr size_t kReplayFixed = (size_t)kReplayFixedInts * 4;   // candidate / query staging of replay_queries
constexpr size_t kReplaySmemMax = 160 * 1024;   // shared-memory state of the replays (10k keypoints for init, 18k for the others)

void* stage_take(orbm_matcher* h, size_t bytes);

// phase 1 for nq queries already built in h->ws0 (AreaQuery[nq]); leaves offsets in ws1 and candidates in ws2
int run_candidates(orbm_matcher* h, const FrameDev& f, const uint4* dQdesc, int nq, const float* dURight, int* totalOut) {
    cudaStream_t st = h->stream;
    ORB_CHECK(h->out4.reserve((size_t)(nq + 1) * 4));
    ORB_CHECK(h->ws1.reserve((size_t)(nq + 2) * 4));
    int* counts = h->out4.as<int>();
    int* offsets = h->ws1.as<int>();
    const int wpb = 8, blocks = ceil_div(nq, wpb);
    candidates_kernel<<<blocks, wpb * 32, 0, st>>>(f, h->ws0.as<AreaQuery>(), dQdesc, nq, dURight, counts, nullptr, nullptr);
    scan_kernel<<<1, 1024, 0, st>>>(counts, offsets, nq);
    int total = 0;
    {
        int* pinnedTotal = static_cast<int*>(stage_take(h, 4));
        ORB_CUDA(cudaMemcpyAsync(pinnedTotal ? pinnedTotal : &total, offsets + nq, 4, cudaMemcpyDeviceToHost, st));
        ORB_CUDA(cudaStreamSynchronize(st));
        if (pinnedTotal) total = *pinnedTotal;
    }
    ORB_CHECK(h->ws2.reserve((size_t)(total + 1) * sizeof(int2)));
    candidates_kernel<<<blocks, wpb * 32, 0, st>>>(f, h->ws0.as<AreaQuery>(), dQdesc, nq, dURight, nullptr, offsets, h->ws2.as<int2>());
    h->launches += 3;
    ORB_CUDA(cudaGetLastError());
    *totalOut = total;
    return ORB_OK;
}

int upload(DevBuf& b, const void* src, size_t bytes, cudaStream_t st) {
    ORB_CHECK(b.reserve(bytes + 16));
    if (bytes) ORB_CUDA(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, st));
    return ORB_OK;
}

// ---- pinned staging of a single call's pageable inputs and outputs (matcher.h) -----------------------------------------
constexpr size_t kStageBytes = 4u << 20;
// a slot of the call's staging area, or nullptr when it is full (the caller then copies directly)
void* stage_take(orbm_matcher* h, size_t bytes) {
    if (!h->stage.p && h->stage.reserve(kStageBytes) != ORB_OK) return nullptr;
    const size_t need = (bytes + 255) & ~(size_t)255;
    if (h->stageOff + need > kStageBytes) return nullptr;
    void* p = (char*)h->stage.p + h->stageOff;
    h->stageOff += need;
    return p;
}
int stage_upload(orbm_matcher* h, DevBuf& b, const void* src, size_t bytes, cudaStream_t st) {
    ORB_CHECK(b.reserve(bytes + 16));
    if (!bytes) return ORB_OK;
    void* s = stage_take(h, bytes);
    if (s) std::memcpy(s, src, bytes);
    ORB_CUDA(cudaMemcpyAsync(b.p, s ? s : src, bytes, cudaMemcpyHostToDevice, st));
    return ORB_OK;
}
int stage_download(orbm_matcher* h, void* dst, const void* dsrc, size_t bytes, cudaStream_t st) {
    if (!bytes) return ORB_OK;
    void* s = h->nPending < 8 ? stage_take(h, bytes) : nullptr;
    ORB_CUDA(cudaMemcpyAsync(s ? s : dst, dsrc, bytes, cudaMemcpyDeviceToHost, st));
    if (s) h->pending[h->nPending++] = orbm_matcher::PendingOut{dst, s, bytes};
    return ORB_OK;
}
// the call's synchronisation; afterwards the staged outputs are handed to the caller's arrays
int stage_finish(orbm_matcher* h, cudaStream_t st) {
    ORB_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < h->nPending; ++i) std::memcpy(h->pending[i].dst, h->pending[i].src, h->pending[i].bytes);
    h->nPending = 0;
    return ORB_OK;
}

orbm_frame_s* new_frame(orbm_matcher* h, int n, float minX, float minY, float maxX, float maxY) {
    orbm_frame_s* f = new orbm_frame_s;
    f->m = h;
    f->n = n;
    f->minX = minX; f->minY = minY; f->maxX = maxX; f->maxY = maxY;
    f->invW = (float)kGridCols / (maxX - minX);   // Frame.cc:93-94
    f->invH = (float)kGridRows / (maxY - minY);
    return f;
}

// AssignFeaturesToGrid (Frame.cc:574-589) in four launches: cell of every key + counts (the caller's kernel), exclusive
// scan, bucket fill, per-bucket sort into the reference's push_back order.  cellOf lives in h->ws0.
int grid_workspace(orbm_matcher* h, orbm_frame_s* f, int** counts, int** cursor, cudaStream_t st) {
    ORB_CHECK(f->cellStart.reserve((kCells + 1) * 4));
    ORB_CHECK(f->cellIdx.reserve((size_t)(f->n + 1) * 4));
    ORB_CHECK(h->ws0.reserve((size_t)(f->n + 1) * 4));
    ORB_CHECK(h->ws1.reserve((size_t)(kCells + 1) * 4 * 2));
    *counts = h->ws1.as<int>();
    *cursor = *counts + kCells + 1;
    ORB_CUDA(cudaMemsetAsync(*counts, 0, (size_t)(kCells + 1) * 4 * 2, st));
    return ORB_OK;
}

int finish_grid(orbm_matcher* h, orbm_frame_s* f, int* counts, int* cursor, cudaStream_t st) {
    const int n = f->n;
    scan_kernel<<<1, 1024, 0, st>>>(counts, f->cellStart.as<int>(), kCells);
    if (n > 0) grid_fill_kernel<<<ceil_div(n, 256), 256, 0, st>>>(n, h->ws0.as<int>(), f->cellStart.as<int>(), cursor, f->cellIdx.as<int>());
    grid_sort_kernel<<<ceil_div(kCells, 256), 256, 0, st>>>(f->cellStart.as<int>(), f->cellIdx.as<int>());
    h->launches += 3;
    ORB_CUDA(cudaGetLastError());
    ORB_CUDA(cudaStreamSynchronize(st));
    return ORB_OK;
}

// mK / mDistCoef (CV_32F, Tracking.cc reads fx fy cx cy k1 k2 p1 p2 [k3] from the settings file) widened to double
int make_camera(const orb_camera* cam, CamDev* c, const char* who) {
    std::memset(c, 0, sizeof *c);
    if (!cam) return ORB_OK;
    if (!(cam->fx != 0.0f) || !(cam->fy != 0.0f)) return fail(ORB_ERR_INVALID, "%s: focal length is zero", who);
    c->fx = cam->fx; c->fy = cam->fy; c->cx = cam->cx; c->cy = cam->cy;
    c->ifx = 1.0 / c->fx;
    c->ify = 1.0 / c->fy;
    c->k1 = cam->k1; c->k2 = cam->k2; c->p1 = cam->p1; c->p2 = cam->p2; c->k3 = cam->k3;
    c->distorted = cam->k1 != 0.0f;
    return ORB_OK;
}

}  // namespace

extern "C" {

int orbm_frame_create(orbm_handle h, const orb_keypoint* keys, const uint8_t* desc, int n, float minX, float minY,
                      float maxX, float maxY, orbm_frame* out) {
    ORBM_ENTER(h);
    if (!out) return fail(ORB_ERR_INVALID, "orbm_frame_create: null out");
    *out = nullptr;
    if (n < 0 || (n > 0 && (!keys || !desc)) || !(maxX > minX) || !(maxY > minY))
        return fail(ORB_ERR_INVALID, "orbm_frame_create: bad arguments");
    if (n > kOrdMask) return fail(ORB_ERR_INVALID, "orbm_frame_create: more than %d keypoints", kOrdMask);
    orbm_frame_s* f = new_frame(h, n, minX, minY, maxX, maxY);
    cudaStream_t st = h->stream;
    auto body = [&]() -> int {
        ORB_CHECK(stage_upload(h, f->keys, keys, (size_t)n * sizeof(orb_keypoint), st));
        ORB_CHECK(stage_upload(h, f->desc, desc, (size_t)n * 32, st));
        int *counts, *cursor;
        ORB_CHECK(grid_workspace(h, f, &counts, &cursor, st));
        if (n > 0) grid_count_kernel<<<ceil_div(n, 256), 256, 0, st>>>(f->dev(), h->ws0.as<int>(), counts);
        h->launches += 1;
        return finish_grid(h, f, counts, cursor, st);
    };
    const int status = body();
    if (status != ORB_OK) {
        orbm_frame_destroy(f);
        return status;
    }
    *out = f;
    return ORB_OK;
}

int orbm_frame_create_device(orbm_handle h, const orb_keypoint* dKeys, const uint8_t* dDesc, const int* dCount, int capacity,
                             const orb_camera* cam, float minX, float minY, float maxX, float maxY, void* producerStream,
                             orbm_frame* out) {
    ORBM_ENTER(h);
    if (!out) return fail(ORB_ERR_INVALID, "orbm_frame_create_device: null out");
    *out = nullptr;
    if (!dKeys || !dDesc || !dCount || capacity < 1 || !(maxX > minX) || !(maxY > minY))
        return fail(ORB_ERR_INVALID, "orbm_frame_create_device: bad arguments");
    if (((uintptr_t)dDesc & 15) != 0) return fail(ORB_ERR_INVALID, "orbm_frame_create_device: descriptors must be 16-byte aligned");
    CamDev c;
    ORB_CHECK(make_camera(cam, &c, "orbm_frame_create_device"));
    // N = mvKeys.size() is host state of the Frame (it sizes mvpMapPoints, mvbOutlier): the one 4-byte read-back
    int n = 0;
    cudaStream_t ps = (cudaStream_t)producerStream;
    ORB_CUDA(cudaMemcpyAsync(&n, dCount, 4, cudaMemcpyDeviceToHost, ps));
    ORB_CUDA(cudaStreamSynchronize(ps));
    if (n < 0) return fail(ORB_ERR_INVALID, "orbm_frame_create_device: negative keypoint count on the device");
    if (n > capacity) return fail(ORB_ERR_CAPACITY, "orbm_frame_create_device: %d keypoints, capacity %d", n, capacity);
    if (n > kOrdMask) return fail(ORB_ERR_INVALID, "orbm_frame_create_device: more than %d keypoints", kOrdMask);
    orbm_frame_s* f = new_frame(h, n, minX, minY, maxX, maxY);
    cudaStream_t st = h->stream;
    auto body = [&]() -> int {
        ORB_CHECK(f->keys.reserve((size_t)n * sizeof(orb_keypoint) + 16));
        ORB_CHECK(f->desc.reserve((size_t)n * 32 + 16));
        int *counts, *cursor;
        ORB_CHECK(grid_workspace(h, f, &counts, &cursor, st));
        if (n > 0)
            frame_import_kernel<<<ceil_div(n, 128), 128, 0, st>>>(c, dKeys, (const uint4*)dDesc, n, f->keys.as<orb_keypoint>(),
                                                                  f->desc.as<uint4>(), f->minX, f->minY, f->invW, f->invH,
                                                                  h->ws0.as<int>(), counts);
        h->launches += 1;
        return finish_grid(h, f, counts, cursor, st);
    };
    const int status = body();
    if (status != ORB_OK) {
        orbm_frame_destroy(f);
        return status;
    }
    *out = f;
    return ORB_OK;
}

int orbm_frame_size(orbm_frame f, int* n) {
    if (!f || !n) return fail(ORB_ERR_INVALID, "orbm_frame_size: null argument");
    *n = f->n;
    return ORB_OK;
}

int orbm_frame_download(orbm_frame f, orb_keypoint* keysUn, uint8_t* desc) {
    if (!f) return fail(ORB_ERR_INVALID, "orbm_frame_download: null frame");
    DeviceGuard g(f->m->device);
    if (f->n == 0) return ORB_OK;
    if (keysUn) ORB_CUDA(cudaMemcpy(keysUn, f->keys.p, (size_t)f->n * sizeof(orb_keypoint), cudaMemcpyDeviceToHost));
    if (desc) ORB_CUDA(cudaMemcpy(desc, f->desc.p, (size_t)f->n * 32, cudaMemcpyDeviceToHost));
    return ORB_OK;
}

int orbm_undistort_points(orbm_handle h, const orb_camera* cam, const float* xy, int n, float* xyOut) {
    ORBM_ENTER(h);
    if (n < 0 || (n > 0 && (!xy || !xyOut))) return fail(ORB_ERR_INVALID, "orbm_undistort_points: bad arguments");
    CamDev c;
    ORB_CHECK(make_camera(cam, &c, "orbm_undistort_points"));
    if (n == 0) return ORB_OK;
    cudaStream_t st = h->stream;
    ORB_CHECK(upload(h->in0, xy, (size_t)n * 8, st));
    ORB_CHECK(h->out0.reserve((size_t)n * 8));
    undistort_kernel<<<ceil_div(n, 128), 128, 0, st>>>(c, h->in0.as<float2>(), n, h->out0.as<float2>());
    h->launches += 1;
    ORB_CUDA(cudaGetLastError());
    ORB_CUDA(cudaMemcpyAsync(xyOut, h->out0.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    return ORB_OK;
}

int orbm_image_bounds(orbm_handle h, const orb_camera* cam, int width, int height, float* bounds) {
    if (!bounds || width <= 0 || height <= 0) return fail(ORB_ERR_INVALID, "orbm_image_bounds: bad arguments");
    if (!cam || cam->k1 == 0.0f) {   // Frame.cc:803-808
        bounds[0] = 0.0f; bounds[1] = 0.0f; bounds[2] = (float)width; bounds[3] = (float)height;
        return ORB_OK;
    }
    const float corners[8] = {0.0f, 0.0f, (float)width, 0.0f, 0.0f, (float)height, (float)width, (float)height};
    float un[8];
    ORB_CHECK(orbm_undistort_points(h, cam, corners, 4, un));
    bounds[0] = fminf(un[0], un[4]);   // mnMinX = min(mat(0,0), mat(2,0))   Frame.cc:797-800
    bounds[2] = fmaxf(un[2], un[6]);   // mnMaxX = max(mat(1,0), mat(3,0))
    bounds[1] = fminf(un[1], un[3]);   // mnMinY = min(mat(0,1), mat(1,1))
    bounds[3] = fmaxf(un[5], un[7]);   // mnMaxY = max(mat(2,1), mat(3,1))
    return ORB_OK;
}

int orbm_frame_destroy(orbm_frame f) {
    if (!f) return ORB_OK;
    DeviceGuard g(f->m->device);
    f->keys.release(); f->desc.release(); f->cellStart.release(); f->cellIdx.release();
    delete f;
    return ORB_OK;
}

int orbm_frame_grid(orbm_frame f, int* cellStart, int* cellIdx) {
    if (!f || !cellStart || !cellIdx) return fail(ORB_ERR_INVALID, "orbm_frame_grid: null argument");
    DeviceGuard g(f->m->device);
    ORB_CUDA(cudaMemcpy(cellStart, f->cellStart.p, (kCells + 1) * 4, cudaMemcpyDeviceToHost));
    const int total = cellStart[kCells];
    if (total > 0) ORB_CUDA(cudaMemcpy(cellIdx, f->cellIdx.p, (size_t)total * 4, cudaMemcpyDeviceToHost));
    return ORB_OK;
}

int orbm_features_in_area(orbm_frame f, const float* xyr, int nq, int minLevel, int maxLevel, int* idxOut, int cap, int* countOut) {
    if (!f) return fail(ORB_ERR_INVALID, "orbm_features_in_area: null frame");
    orbm_matcher* h = f->m;
    ORBM_ENTER(h);
    if (nq < 0 || cap < 1 || (nq > 0 && (!xyr || !idxOut || !countOut))) return fail(ORB_ERR_INVALID, "orbm_features_in_area: bad arguments");
    if (nq == 0) return ORB_OK;
    cudaStream_t st = h->stream;
    ORB_CHECK(upload(h->in0, xyr, (size_t)nq * 12, st));
    ORB_CHECK(h->out0.reserve((size_t)nq * cap * 4));
    ORB_CHECK(h->out1.reserve((size_t)nq * 4));
    area_kernel<<<ceil_div(nq, 8), 256, 0, st>>>(f->dev(), h->in0.as<float>(), nq, minLevel, maxLevel, h->out0.as<int>(), cap, h->out1.as<int>());
    h->launches += 1;
    ORB_CUDA(cudaGetLastError());
    ORB_CUDA(cudaMemcpyAsync(idxOut, h->out0.p, (size_t)nq * cap * 4, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaMemcpyAsync(countOut, h->out1.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    return ORB_OK;
}

int orbm_search_for_initialization(orbm_handle h, orbm_frame f1, orbm_frame f2, float* prevXY, int* matches12,
                                   int windowSize, float ratio, int checkOri, int* nmatches) {
    ORBM_ENTER(h);
    if (!f1 || !f2 || !prevXY || !matches12 || !nmatches) return fail(ORB_ERR_INVALID, "orbm_search_for_initialization: null argument");
    *nmatches = 0;
    const int n1 = f1->n, n2 = f2->n;
    if (n1 == 0) return ORB_OK;
    cudaStream_t st = h->stream;
    ORB_CHECK(stage_upload(h, h->in0, prevXY, (size_t)n1 * 8, st));
    ORB_CHECK(h->ws0.reserve((size_t)n1 * sizeof(AreaQuery)));
    const FrameDev d1 = f1->dev(), d2 = f2->dev();
    init_queries_kernel<<<ceil_div(n1, 256), 256, 0, st>>>(d1, h->in0.as<float>(), windowSize, h->ws0.as<AreaQuery>());
    h->launches += 1;
    int total = 0;
    ORB_CHECK(run_candidates(h, d2, d1.desc, n1, nullptr, &total));
    ORB_CHECK(h->out0.reserve((size_t)(n1 + 1) * 4));          // m12
    ORB_CHECK(h->out2.reserve((size_t)(n1 + 1) * 4 * 2));      // pushBin, pushVal
    ORB_CHECK(h->out3.reserve(16));
    int* pushBin = h->out2.as<int>();
    const size_t replaySmem = (size_t)std::max(n2, 1) * 16 + 16;  // m21 + vMatchedDistance + two stamp arrays
    if (replaySmem > kReplaySmemMax) return fail(ORB_ERR_CAPACITY, "orbm_search_for_initialization: %d keypoints exceed the replay state (%d)", n2, (int)(kReplaySmemMax / 16));
    ORB_CUDA(cudaFuncSetAttribute(init_replay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kReplayFixed + kReplaySmemMax)));
    init_replay_kernel<<<1, RP_THREADS, kReplayFixed + replaySmem, st>>>(d1, d2, h->ws0.as<AreaQuery>(), h->ws1.as<int>(), h->ws2.as<int2>(), ratio, checkOri,
                                                  h->in0.as<float>(), h->out0.as<int>(), pushBin, pushBin + n1 + 1, h->out3.as<int>());
    h->launches += 1;
    ORB_CUDA(cudaGetLastError());
    ORB_CHECK(stage_download(h, matches12, h->out0.p, (size_t)n1 * 4, st));
    ORB_CHECK(stage_download(h, prevXY, h->in0.p, (size_t)n1 * 8, st));
    ORB_CHECK(stage_download(h, nmatches, h->out3.p, 4, st));
    return stage_finish(h, st);
}

int orbm_search_by_projection(orbm_handle h, orbm_frame cur, const float* sf, int nlevels, const float* uRight, float mbf,
                              const orbm_proj_query* queries, const uint8_t* qdesc, int nq, float th, int mode,
                              const uint8_t* occupied, int* curMatch, int checkOri, int* nmatches) {
    return orbm_search_by_projection_ex(h, cur, sf, nlevels, uRight, mbf, queries, qdesc, nq, th, mode, kThHigh, occupied,
                                        curMatch, checkOri, nmatches);
}

namespace {

// everything after the queries (device, h->in0) and their descriptors (h->in1) are in place
int projection_search_device(orbm_matcher* h, orbm_frame cur, const float* sf, int nlevels, const float* uRight, float mbf, int nq,
                             float th, int mode, int maxDist, const uint8_t* occupied, int* curMatch, int checkOri, int* nmatches) {
    const int n = cur->n;
    cudaStream_t st = h->stream;
    ORB_CHECK(stage_upload(h, h->in2, sf, (size_t)nlevels * 4, st));
    if (uRight) ORB_CHECK(stage_upload(h, h->in3, uRight, (size_t)n * 4, st));
    ORB_CHECK(h->in4.reserve((size_t)n + 16));
    if (occupied) ORB_CHECK(stage_upload(h, h->in4, occupied, (size_t)n, st));
    else ORB_CUDA(cudaMemsetAsync(h->in4.p, 0, (size_t)n, st));
    ORB_CHECK(h->ws0.reserve((size_t)nq * sizeof(AreaQuery)));
    const FrameDev d = cur->dev();
    proj_queries_kernel<<<ceil_div(nq, 256), 256, 0, st>>>(d, h->in0.as<orbm_proj_query>(), nq, h->in2.as<float>(), th, mode, mbf,
                                                           h->ws0.as<AreaQuery>());
    h->launches += 1;
    int total = 0;
    ORB_CHECK(run_candidates(h, d, h->in1.as<uint4>(), nq, uRight ? h->in3.as<float>() : nullptr, &total));
    ORB_CHECK(h->out0.reserve((size_t)(n + 1) * 4));
    ORB_CHECK(h->out2.reserve((size_t)(nq + 1) * 4 * 2));
    ORB_CHECK(h->out3.reserve(16));
    int* pushBin = h->out2.as<int>();
    if (9 * (size_t)n + 32 > kReplaySmemMax) return fail(ORB_ERR_CAPACITY, "orbm_search_by_projection: %d keypoints exceed the replay state", n);
    ORB_CUDA(cudaFuncSetAttribute(proj_replay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kReplayFixed + kReplaySmemMax)));
    proj_replay_kernel<<<1, RP_THREADS, kReplayFixed + 9 * (size_t)n + 32, st>>>(d, h->ws0.as<AreaQuery>(), h->in0.as<orbm_proj_query>(), nq, h->ws1.as<int>(),
                                                      h->ws2.as<int2>(), checkOri, maxDist, h->in4.as<unsigned char>(), h->out0.as<int>(),
                                                      pushBin, pushBin + nq + 1, h->out3.as<int>());
    h->launches += 1;
    ORB_CUDA(cudaGetLastError());
    ORB_CHECK(stage_download(h, curMatch, h->out0.p, (size_t)n * 4, st));
    ORB_CHECK(stage_download(h, nmatches, h->out3.p, 4, st));
    return stage_finish(h, st);
}

}  // namespace

int orbm_search_by_projection_ex(orbm_handle h, orbm_frame cur, const float* sf, int nlevels, const float* uRight, float mbf,
                                 const orbm_proj_query* queries, const uint8_t* qdesc, int nq, float th, int mode, int maxDist,
                                 const uint8_t* occupied, int* curMatch, int checkOri, int* nmatches) {
    ORBM_ENTER(h);
    if (!cur || !sf || nlevels < 1 || !curMatch || !nmatches || nq < 0 || (nq > 0 && (!queries || !qdesc)))
        return fail(ORB_ERR_INVALID, "orbm_search_by_projection: bad arguments");
    for (int i = 0; i < nq; ++i)
        if (queries[i].valid && (queries[i].octave < 0 || queries[i].octave >= nlevels))
            return fail(ORB_ERR_INVALID, "orbm_search_by_projection: query %d has octave %d outside 0..%d", i, queries[i].octave, nlevels - 1);
    *nmatches = 0;
    const int n = cur->n;
    for (int i = 0; i < n; ++i) curMatch[i] = -1;
    if (nq == 0 || n == 0) return ORB_OK;
    cudaStream_t st = h->stream;
    ORB_CHECK(stage_upload(h, h->in0, queries, (size_t)nq * sizeof(orbm_proj_query), st));
    ORB_CHECK(stage_upload(h, h->in1, qdesc, (size_t)nq * 32, st));
    return projection_search_device(h, cur, sf, nlevels, uRight, mbf, nq, th, mode, maxDist, occupied, curMatch, checkOri, nmatches);
}

int orbm_search_by_projection_world(orbm_handle h, orbm_frame cur, const float* sf, int nlevels, const float* uRight, float mbf,
                                    const orbm_pose* pose, const orbm_world_query* queries, const uint8_t* qdesc, int nq, float th,
                                    int mode, int maxDist, const uint8_t* occupied, int* curMatch, int checkOri, int* nmatches) {
    ORBM_ENTER(h);
    if (!cur || !sf || nlevels < 1 || !pose || !curMatch || !nmatches || nq < 0 || (nq > 0 && (!queries || !qdesc)))
        return fail(ORB_ERR_INVALID, "orbm_search_by_projection_world: bad arguments");
    for (int i = 0; i < nq; ++i)
        if (queries[i].valid && (queries[i].octave < 0 || queries[i].octave >= nlevels))
            return fail(ORB_ERR_INVALID, "orbm_search_by_projection_world: query %d has octave %d outside 0..%d", i, queries[i].octave,
                        nlevels - 1);
    *nmatches = 0;
    const int n = cur->n;
    for (int i = 0; i < n; ++i) curMatch[i] = -1;
    if (nq == 0 || n == 0) return ORB_OK;
    cudaStream_t st = h->stream;
    ORB_CHECK(stage_upload(h, h->in5, queries, (size_t)nq * sizeof(orbm_world_query), st));
    ORB_CHECK(stage_upload(h, h->in1, qdesc, (size_t)nq * 32, st));
    ORB_CHECK(h->in0.reserve((size_t)nq * sizeof(orbm_proj_query)));
    PoseDev P;
    for (int i = 0; i < 9; ++i) P.R[i] = pose->Rcw[i];
    for (int i = 0; i < 3; ++i) P.t[i] = pose->tcw[i];
    P.fx = pose->fx; P.fy = pose->fy; P.cx = pose->cx; P.cy = pose->cy;
    project_world_kernel<<<ceil_div(nq, 256), 256, 0, st>>>(P, h->in5.as<orbm_world_query>(), nq, h->in0.as<orbm_proj_query>());
    h->launches += 1;
    return projection_search_device(h, cur, sf, nlevels, uRight, mbf, nq, th, mode, maxDist, occupied, curMatch, checkOri, nmatches);
}

int orbm_project_points(orbm_handle h, const orbm_pose* pose, const float* xyz, int n, float* u, float* v, float* invz) {
    ORBM_ENTER(h);
    if (!pose || n < 0 || (n > 0 && (!xyz || !u || !v || !invz))) return fail(ORB_ERR_INVALID, "orbm_project_points: bad arguments");
    if (n == 0) return ORB_OK;
    cudaStream_t st = h->stream;
    std::vector<orbm_world_query> wq((size_t)n);
    for (int i = 0; i < n; ++i) {
        wq[i] = orbm_world_query();
        wq[i].x = xyz[3 * i]; wq[i].y = xyz[3 * i + 1]; wq[i].z = xyz[3 * i + 2];
        wq[i].valid = 1;
    }
    ORB_CHECK(upload(h->in5, wq.data(), (size_t)n * sizeof(orbm_world_query), st));
    ORB_CHECK(h->in0.reserve((size_t)n * sizeof(orbm_proj_query)));
    PoseDev P;
    for (int i = 0; i < 9; ++i) P.R[i] = pose->Rcw[i];
    for (int i = 0; i < 3; ++i) P.t[i] = pose->tcw[i];
    P.fx = pose->fx; P.fy = pose->fy; P.cx = pose->cx; P.cy = pose->cy;
    project_world_kernel<<<ceil_div(n, 256), 256, 0, st>>>(P, h->in5.as<orbm_world_query>(), n, h->in0.as<orbm_proj_query>());
    h->launches += 1;
    ORB_CUDA(cudaGetLastError());
    std::vector<orbm_proj_query> pq((size_t)n);
    ORB_CUDA(cudaMemcpyAsync(pq.data(), h->in0.p, (size_t)n * sizeof(orbm_proj_query), cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < n; ++i) { u[i] = pq[i].u; v[i] = pq[i].v; invz[i] = pq[i].invz; }
    return ORB_OK;
}

// ---- batched windowed searches: n independent jobs share ONE pass of every phase (query build, candidate count, scan,
// candidate fill, replay with one CTA per job), so a batch costs the launches and the single host synchronisation of one
// search while all 148 SMs work (the single-pair calls keep one CTA busy).  Results are those of n single calls.
namespace {

// phase 1 for a whole batch whose AreaQuery array (nqAll entries) is in h->ws0 and whose CandJob table is at dJobs
int run_candidates_batch(orbm_matcher* h, const CandJob* dJobs, int nJobs, int maxNq, int nqAll, int* totalOut) {
    cudaStream_t st = h->stream;
    ORB_CHECK(h->out4.reserve((size_t)(nqAll + 1) * 4));
    ORB_CHECK(h->ws1.reserve((size_t)(nqAll + 2) * 4));
    int* counts = h->out4.as<int>();
    int* offsets = h->ws1.as<int>();
    const int wpb = 8;
    const dim3 grid(ceil_div(maxNq, wpb), nJobs);
    candidates_batch_kernel<<<grid, wpb * 32, 0, st>>>(dJobs, h->ws0.as<AreaQuery>(), counts, nullptr, nullptr);
    const int tiles = ceil_div(nqAll, kScanTile);
    ORB_CHECK(h->ws3.reserve((size_t)(2 * tiles + 2) * 4));
    int* tileSums = h->ws3.as<int>();
    int* tileOffsets = tileSums + tiles;
    scan_tile_sums_kernel<<<tiles, 1024, 0, st>>>(counts, nqAll, tileSums);
    scan_kernel<<<1, 1024, 0, st>>>(tileSums, tileOffsets, tiles);
    scan_tiles_kernel<<<tiles, 1024, 0, st>>>(counts, nqAll, tileOffsets, offsets);
    h->launches += 2;
    int total = 0;
    ORB_CUDA(cudaMemcpyAsync(&total, offsets + nqAll, 4, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    ORB_CHECK(h->ws2.reserve((size_t)(total + 1) * sizeof(int2)));
    candidates_batch_kernel<<<grid, wpb * 32, 0, st>>>(dJobs, h->ws0.as<AreaQuery>(), nullptr, offsets, h->ws2.as<int2>());
    h->launches += 3;
    ORB_CUDA(cudaGetLastError());
    *totalOut = total;
    return ORB_OK;
}

}  // namespace

int orbm_search_by_projection_batch(orbm_handle h, orbm_projection_job* jobs, int nJobs, const float* sf, int nlevels, float mbf,
                                    float th, int mode, int maxDist, int checkOri, long long* candidatesOut) {
    ORBM_ENTER(h);
    if (nJobs < 0 || (nJobs > 0 && !jobs) || !sf || nlevels < 1) return fail(ORB_ERR_INVALID, "orbm_search_by_projection_batch: bad arguments");
    if (candidatesOut) *candidatesOut = 0;
    if (nJobs > 65535) return fail(ORB_ERR_INVALID, "orbm_search_by_projection_batch: at most 65535 jobs per call");
    size_t nqAll = 0, nCurAll = 0;
    int maxNq = 0, maxN = 0;
    bool anyRight = false;
    for (int j = 0; j < nJobs; ++j) {
        orbm_projection_job& J = jobs[j];
        if (!J.cur || !J.cur_match || J.nq < 0 || (J.nq > 0 && (!J.queries || !J.query_desc)))
            return fail(ORB_ERR_INVALID, "orbm_search_by_projection_batch: job %d has bad arguments", j);
        if (J.cur->m != h) return fail(ORB_ERR_INVALID, "orbm_search_by_projection_batch: job %d: frame belongs to another matcher", j);
        for (int i = 0; i < J.nq; ++i)
            if (J.queries[i].valid && (J.queries[i].octave < 0 || J.queries[i].octave >= nlevels))
                return fail(ORB_ERR_INVALID, "orbm_search_by_projection_batch: job %d query %d has octave %d outside 0..%d", j, i,
                            J.queries[i].octave, nlevels - 1);
        J.nmatches = 0;
        if (J.nq == 0) std::fill(J.cur_match, J.cur_match + J.cur->n, -1);   // a job that runs overwrites all of it
        nqAll += (size_t)J.nq;
        nCurAll += (size_t)J.cur->n;
        maxNq = std::max(maxNq, J.nq);
        maxN = std::max(maxN, J.cur->n);
        anyRight = anyRight || J.u_right;
    }
    if (nqAll == 0 || nCurAll == 0) return ORB_OK;
    if (nqAll > (size_t)1 << 30) return fail(ORB_ERR_INVALID, "orbm_search_by_projection_batch: too many queries in one call");
    if (9 * (size_t)maxN + 32 > kReplaySmemMax) return fail(ORB_ERR_CAPACITY, "orbm_search_by_projection_batch: %d keypoints exceed the replay state", maxN);
    cudaStream_t st = h->stream;
    // inputs concatenated in page-locked staging (kept by the handle): one memcpy per array and job, one full-rate H2D each
    ORB_CHECK(h->pin0.reserve(nqAll * sizeof(orbm_proj_query)));
    ORB_CHECK(h->pin1.reserve(nqAll * 32));
    ORB_CHECK(h->pin2.reserve(nCurAll));
    if (anyRight) ORB_CHECK(h->pin3.reserve(nCurAll * 4));
    orbm_proj_query* hq = h->pin0.as<orbm_proj_query>();
    uint8_t* hd = h->pin1.as<uint8_t>();
    uint8_t* hocc = h->pin2.as<uint8_t>();
    float* hur = h->pin3.as<float>();
    {
        size_t qo = 0, co = 0;
        for (int j = 0; j < nJobs; ++j) {
            const orbm_projection_job& J = jobs[j];
            if (J.nq) {
                std::memcpy(&hq[qo], J.queries, (size_t)J.nq * sizeof(orbm_proj_query));
                std::memcpy(&hd[qo * 32], J.query_desc, (size_t)J.nq * 32);
            }
            if (J.cur->n) {
                if (J.occupied) std::memcpy(&hocc[co], J.occupied, (size_t)J.cur->n);
                else std::memset(&hocc[co], 0, (size_t)J.cur->n);
                if (anyRight) {
                    if (J.u_right) std::memcpy(&hur[co], J.u_right, (size_t)J.cur->n * 4);
                    else std::fill(hur + co, hur + co + J.cur->n, -1.0f);
                }
            }
            qo += (size_t)J.nq;
            co += (size_t)J.cur->n;
        }
    }
    ORB_CHECK(upload(h->in0, hq, nqAll * sizeof(orbm_proj_query), st));
    ORB_CHECK(upload(h->in1, hd, nqAll * 32, st));
    ORB_CHECK(upload(h->in2, sf, (size_t)nlevels * 4, st));
    if (anyRight) ORB_CHECK(upload(h->in3, hur, nCurAll * 4, st));
    ORB_CHECK(upload(h->in4, hocc, nCurAll, st));
    ORB_CHECK(h->ws0.reserve(nqAll * sizeof(AreaQuery)));
    ORB_CHECK(h->out0.reserve((nCurAll + 1) * 4));
    ORB_CHECK(h->out2.reserve((nqAll + (size_t)nJobs) * 4 * 2));
    ORB_CHECK(h->out3.reserve((size_t)nJobs * 4 + 16));
    // job tables
    std::vector<ProjJob> pj(nJobs);
    std::vector<CandJob> cj(nJobs);
    {
        size_t qo = 0, co = 0;
        for (int j = 0; j < nJobs; ++j) {
            const orbm_projection_job& J = jobs[j];
            ProjJob& P = pj[j];
            P.cur = J.cur->dev();
            P.pq = h->in0.as<orbm_proj_query>() + qo;
            P.occIn = h->in4.as<unsigned char>() + co;
            P.curMatch = h->out0.as<int>() + co;
            P.pushA = h->out2.as<int>() + 2 * (qo + j);
            P.pushB = P.pushA + J.nq + 1;
            P.nmatchesOut = h->out3.as<int>() + j;
            P.nq = J.nq;
            P.qBase = (int)qo;
            CandJob& C = cj[j];
            C.f = P.cur;
            C.qdesc = h->in1.as<uint4>() + 2 * qo;
            C.uRight = (anyRight && J.u_right) ? h->in3.as<float>() + co : nullptr;
            C.nq = J.nq;
            C.qBase = (int)qo;
            qo += (size_t)J.nq;
            co += (size_t)J.cur->n;
        }
    }
    ORB_CHECK(h->in5.reserve((size_t)nJobs * (sizeof(ProjJob) + sizeof(CandJob)) + 64));
    ProjJob* dPj = h->in5.as<ProjJob>();
    CandJob* dCj = reinterpret_cast<CandJob*>(dPj + nJobs);
    ORB_CUDA(cudaMemcpyAsync(dPj, pj.data(), (size_t)nJobs * sizeof(ProjJob), cudaMemcpyHostToDevice, st));
    ORB_CUDA(cudaMemcpyAsync(dCj, cj.data(), (size_t)nJobs * sizeof(CandJob), cudaMemcpyHostToDevice, st));
    proj_queries_batch_kernel<<<dim3(ceil_div(maxNq, 256), nJobs), 256, 0, st>>>(dPj, h->in2.as<float>(), th, mode, mbf, h->ws0.as<AreaQuery>());
    h->launches += 1;
    int total = 0;
    ORB_CHECK(run_candidates_batch(h, dCj, nJobs, maxNq, (int)nqAll, &total));
    if (candidatesOut) *candidatesOut = total;
    ORB_CUDA(cudaFuncSetAttribute(proj_replay_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kReplayFixed + kReplaySmemMax)));
    proj_replay_batch_kernel<<<nJobs, RP_THREADS, kReplayFixed + 9 * (size_t)maxN + 32, st>>>(dPj, h->ws0.as<AreaQuery>(), h->ws1.as<int>(), h->ws2.as<int2>(),
                                                                                      checkOri, maxDist);
    h->launches += 1;
    ORB_CUDA(cudaGetLastError());
    ORB_CHECK(h->pin4.reserve((nCurAll + (size_t)nJobs) * 4));
    int* hm = h->pin4.as<int>();
    int* hn = hm + nCurAll;
    ORB_CUDA(cudaMemcpyAsync(hm, h->out0.p, nCurAll * 4, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaMemcpyAsync(hn, h->out3.p, (size_t)nJobs * 4, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    size_t co = 0;
    for (int j = 0; j < nJobs; ++j) {
        orbm_projection_job& J = jobs[j];
        if (J.nq > 0 && J.cur->n > 0) {
            std::memcpy(J.cur_match, &hm[co], (size_t)J.cur->n * 4);
            J.nmatches = hn[j];
        }
        co += (size_t)J.cur->n;
    }
    return ORB_OK;
}

int orbm_search_for_initialization_batch(orbm_handle h, orbm_init_job* jobs, int nJobs, int windowSize, float ratio, int checkOri,
                                         long long* candidatesOut) {
    ORBM_ENTER(h);
    if (nJobs < 0 || (nJobs > 0 && !jobs)) return fail(ORB_ERR_INVALID, "orbm_search_for_initialization_batch: bad arguments");
    if (candidatesOut) *candidatesOut = 0;
    if (nJobs > 65535) return fail(ORB_ERR_INVALID, "orbm_search_for_initialization_batch: at most 65535 jobs per call");
    size_t n1All = 0;
    int maxN1 = 0, maxN2 = 0;
    for (int j = 0; j < nJobs; ++j) {
        orbm_init_job& J = jobs[j];
        if (!J.f1 || !J.f2 || !J.prev_xy || !J.matches12) return fail(ORB_ERR_INVALID, "orbm_search_for_initialization_batch: job %d has a null argument", j);
        if (J.f1->m != h || J.f2->m != h) return fail(ORB_ERR_INVALID, "orbm_search_for_initialization_batch: job %d: frame belongs to another matcher", j);
        J.nmatches = 0;
        n1All += (size_t)J.f1->n;
        maxN1 = std::max(maxN1, J.f1->n);
        maxN2 = std::max(maxN2, J.f2->n);
    }
    if (n1All == 0) return ORB_OK;
    const size_t replaySmem = (size_t)std::max(maxN2, 1) * 16 + 16;
    if (replaySmem > kReplaySmemMax) return fail(ORB_ERR_CAPACITY, "orbm_search_for_initialization_batch: %d keypoints exceed the replay state", maxN2);
    cudaStream_t st = h->stream;
    std::vector<float> hprev(n1All * 2);
    {
        size_t o = 0;
        for (int j = 0; j < nJobs; ++j) {
            if (jobs[j].f1->n) std::memcpy(&hprev[o * 2], jobs[j].prev_xy, (size_t)jobs[j].f1->n * 8);
            o += (size_t)jobs[j].f1->n;
        }
    }
    ORB_CHECK(upload(h->in0, hprev.data(), n1All * 8, st));
    ORB_CHECK(h->ws0.reserve(n1All * sizeof(AreaQuery)));
    ORB_CHECK(h->out0.reserve((n1All + 1) * 4));
    ORB_CHECK(h->out2.reserve((n1All + (size_t)nJobs) * 4 * 2));
    ORB_CHECK(h->out3.reserve((size_t)nJobs * 4 + 16));
    std::vector<InitJob> ij(nJobs);
    std::vector<CandJob> cj(nJobs);
    {
        size_t o = 0;
        for (int j = 0; j < nJobs; ++j) {
            const orbm_init_job& J = jobs[j];
            InitJob& I = ij[j];
            I.f1 = J.f1->dev();
            I.f2 = J.f2->dev();
            I.prevXY = h->in0.as<float>() + 2 * o;
            I.m12 = h->out0.as<int>() + o;
            I.pushA = h->out2.as<int>() + 2 * (o + j);
            I.pushB = I.pushA + J.f1->n + 1;
            I.nmatchesOut = h->out3.as<int>() + j;
            I.qBase = (int)o;
            CandJob& C = cj[j];
            C.f = I.f2;
            C.qdesc = I.f1.desc;
            C.uRight = nullptr;
            C.nq = J.f1->n;
            C.qBase = (int)o;
            o += (size_t)J.f1->n;
        }
    }
    ORB_CHECK(h->in5.reserve((size_t)nJobs * (sizeof(InitJob) + sizeof(CandJob)) + 64));
    InitJob* dIj = h->in5.as<InitJob>();
    CandJob* dCj = reinterpret_cast<CandJob*>(dIj + nJobs);
    ORB_CUDA(cudaMemcpyAsync(dIj, ij.data(), (size_t)nJobs * sizeof(InitJob), cudaMemcpyHostToDevice, st));
    ORB_CUDA(cudaMemcpyAsync(dCj, cj.data(), (size_t)nJobs * sizeof(CandJob), cudaMemcpyHostToDevice, st));
    init_queries_batch_kernel<<<dim3(ceil_div(maxN1, RP_THREADS), nJobs), RP_THREADS, 0, st>>>(dIj, windowSize, h->ws0.as<AreaQuery>());
    h->launches += 1;
    int total = 0;
    ORB_CHECK(run_candidates_batch(h, dCj, nJobs, maxN1, (int)n1All, &total));
    if (candidatesOut) *candidatesOut = total;
    ORB_CUDA(cudaFuncSetAttribute(init_replay_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kReplayFixed + kReplaySmemMax)));
    init_replay_batch_kernel<<<nJobs, RP_THREADS, kReplayFixed + replaySmem, st>>>(dIj, h->ws0.as<AreaQuery>(), h->ws1.as<int>(), h->ws2.as<int2>(), ratio,
                                                                               checkOri);
    h->launches += 1;
    ORB_CUDA(cudaGetLastError());
    std::vector<int> hm(n1All), hn(nJobs);
    ORB_CUDA(cudaMemcpyAsync(hm.data(), h->out0.p, n1All * 4, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaMemcpyAsync(hprev.data(), h->in0.p, n1All * 8, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaMemcpyAsync(hn.data(), h->out3.p, (size_t)nJobs * 4, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    size_t o = 0;
    for (int j = 0; j < nJobs; ++j) {
        orbm_init_job& J = jobs[j];
        if (J.f1->n) {
            std::memcpy(J.matches12, &hm[o], (size_t)J.f1->n * 4);
            std::memcpy(J.prev_xy, &hprev[o * 2], (size_t)J.f1->n * 8);
        }
        J.nmatches = hn[j];
        o += (size_t)J.f1->n;
    }
    return ORB_OK;
}

int orbm_search_by_projection_points(orbm_handle h, orbm_frame f, const float* sf, int nlevels, const float* uRight,
                                     const orbm_point_query* queries, const uint8_t* qdesc, int nq, float th, float ratio,
                                     const uint8_t* occupied, int* match, int* nmatches) {
    ORBM_ENTER(h);
    if (!f || !sf || nlevels < 1 || !match || !nmatches || nq < 0 || (nq > 0 && (!queries || !qdesc)))
        return fail(ORB_ERR_INVALID, "orbm_search_by_projection_points: bad arguments");
    for (int i = 0; i < nq; ++i)
        if (queries[i].in_view && (queries[i].level < 0 || queries[i].level >= nlevels))
            return fail(ORB_ERR_INVALID, "orbm_search_by_projection_points: query %d has level %d outside 0..%d", i, queries[i].level, nlevels - 1);
    *nmatches = 0;
    const int n = f->n;
    for (int i = 0; i < n; ++i) match[i] = -1;
    if (nq == 0 || n == 0) return ORB_OK;
    cudaStream_t st = h->stream;
    ORB_CHECK(stage_upload(h, h->in0, queries, (size_t)nq * sizeof(orbm_point_query), st));
    ORB_CHECK(stage_upload(h, h->in1, qdesc, (size_t)nq * 32, st));
    ORB_CHECK(stage_upload(h, h->in2, sf, (size_t)nlevels * 4, st));
    if (uRight) ORB_CHECK(stage_upload(h, h->in3, uRight, (size_t)n * 4, st));
    ORB_CHECK(h->in4.reserve((size_t)n + 16));
    if (occupied) ORB_CHECK(stage_upload(h, h->in4, occupied, (size_t)n, st));
    else ORB_CUDA(cudaMemsetAsync(h->in4.p, 0, (size_t)n, st));
    ORB_CHECK(h->ws0.reserve((size_t)nq * sizeof(AreaQuery)));
    const FrameDev d = f->dev();
    point_queries_kernel<<<ceil_div(nq, 256), 256, 0, st>>>(h->in0.as<orbm_point_query>(), nq, h->in2.as<float>(), th, h->ws0.as<AreaQuery>());
    h->launches += 1;
    int total = 0;
    ORB_CHECK(run_candidates(h, d, h->in1.as<uint4>(), nq, uRight ? h->in3.as<float>() : nullptr, &total));
    ORB_CHECK(h->out0.reserve((size_t)(n + 1) * 4));
    ORB_CHECK(h->out3.reserve(16));
    if (9 * (size_t)n + 32 > kReplaySmemMax) return fail(ORB_ERR_CAPACITY, "orbm_search_by_projection_points: %d keypoints exceed the replay state", n);
    ORB_CUDA(cudaFuncSetAttribute(point_replay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kReplayFixed + kReplaySmemMax)));
    point_replay_kernel<<<1, RP_THREADS, kReplayFixed + 9 * (size_t)n + 32, st>>>(d, h->ws0.as<AreaQuery>(), h->in0.as<orbm_point_query>(), nq, h->ws1.as<int>(),
                                                       h->ws2.as<int2>(), ratio, h->in4.as<unsigned char>(), h->out0.as<int>(),
                                                       h->out3.as<int>());
    h->launches += 1;
    ORB_CUDA(cudaGetLastError());
    ORB_CHECK(stage_download(h, match, h->out0.p, (size_t)n * 4, st));
    ORB_CHECK(stage_download(h, nmatches, h->out3.p, 4, st));
    return stage_finish(h, st);
}

int orbm_search_for_triangulation(orbm_handle h, orbm_frame k1, orbm_frame k2, int nNodes1, const int* nodeId1,
                                  const int* start1, const int* idx1, int nNodes2, const int* nodeId2, const int* start2,
                                  const int* idx2, const uint8_t* has1, const uint8_t* has2, const float* uR1, const float* uR2,
                                  const float* f12, float ex, float ey, const float* sf2, const float* sigma2, int nlevels,
                                  int onlyStereo, int checkOri, int* matches12, int* nmatches) {
    ORBM_ENTER(h);
    if (!k1 || !k2 || !f12 || !sf2 || !sigma2 || !matches12 || !nmatches || nNodes1 < 0 || nNodes2 < 0 || nlevels < 1 ||
        (nNodes1 > 0 && (!nodeId1 || !start1 || !idx1)) || (nNodes2 > 0 && (!nodeId2 || !start2 || !idx2)) || !has1 || !has2)
        return fail(ORB_ERR_INVALID, "orbm_search_for_triangulation: bad arguments");
    *nmatches = 0;
    const int n1 = k1->n, n2 = k2->n;
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    if (nNodes1 == 0 || nNodes2 == 0 || n1 == 0 || n2 == 0) return ORB_OK;
    const int e1 = start1[nNodes1], e2 = start2[nNodes2];
    for (int a = 1; a < nNodes1; ++a)
        if (nodeId1[a] <= nodeId1[a - 1]) return fail(ORB_ERR_INVALID, "orbm_search_for_triangulation: node ids of KF1 not ascending");
    for (int a = 1; a < nNodes2; ++a)
        if (nodeId2[a] <= nodeId2[a - 1]) return fail(ORB_ERR_INVALID, "orbm_search_for_triangulation: node ids of KF2 not ascending");
    cudaStream_t st = h->stream;
    // one staging buffer: ints first, then floats, then bytes
    std::vector<int> ints;
    auto putInts = [&](const int* p, int n) { size_t o = ints.size(); ints.insert(ints.end(), p, p + n); return o; };
    const size_t oId1 = putInts(nodeId1, nNodes1), oS1 = putInts(start1, nNodes1 + 1), oI1 = putInts(idx1, e1);
    const size_t oId2 = putInts(nodeId2, nNodes2), oS2 = putInts(start2, nNodes2 + 1), oI2 = putInts(idx2, e2);
    ORB_CHECK(stage_upload(h, h->in0, ints.data(), ints.size() * 4, st));
    std::vector<float> fl;
    fl.insert(fl.end(), sf2, sf2 + nlevels);
    fl.insert(fl.end(), sigma2, sigma2 + nlevels);
    const size_t oU1 = fl.size();
    if (uR1) fl.insert(fl.end(), uR1, uR1 + n1);
    const size_t oU2 = fl.size();
    if (uR2) fl.insert(fl.end(), uR2, uR2 + n2);
    ORB_CHECK(stage_upload(h, h->in1, fl.data(), fl.size() * 4, st));
    ORB_CHECK(stage_upload(h, h->in2, has1, (size_t)n1, st));
    ORB_CHECK(stage_upload(h, h->in3, has2, (size_t)n2, st));
    ORB_CHECK(h->ws0.reserve((size_t)(e1 + 1) * 4));              // entryNode
    ORB_CHECK(h->out0.reserve((size_t)(n1 + 1) * 4));             // m12
    ORB_CHECK(h->out1.reserve((kHistoLength + 2) * 4));           // hist, nmatches
    TriParams P;
    P.k1 = k1->dev(); P.k2 = k2->dev();
    P.nNodes1 = nNodes1; P.nNodes2 = nNodes2; P.nEntries1 = e1;
    const int* di = h->in0.as<int>();
    P.nodeId1 = di + oId1; P.start1 = di + oS1; P.idx1 = di + oI1;
    P.nodeId2 = di + oId2; P.start2 = di + oS2; P.idx2 = di + oI2;
    P.has1 = h->in2.as<unsigned char>(); P.has2 = h->in3.as<unsigned char>();
    const float* df = h->in1.as<float>();
    P.sf2 = df; P.sigma2 = df + nlevels;
    P.uR1 = uR1 ? df + oU1 : nullptr; P.uR2 = uR2 ? df + oU2 : nullptr;
    for (int i = 0; i < 9; ++i) P.F[i] = f12[i];
    P.ex = ex; P.ey = ey; P.onlyStereo = onlyStereo; P.checkOri = checkOri;
    int* hist = h->out1.as<int>();
    ORB_CUDA(cudaMemsetAsync(hist, 0, (kHistoLength + 2) * 4, st));
    fill_int_kernel<<<ceil_div(n1, 256), 256, 0, st>>>(h->out0.as<int>(), n1, -1);
    tri_entry_node_kernel<<<ceil_div(nNodes1, 128), 128, 0, st>>>(nNodes1, P.start1, h->ws0.as<int>());
    if (e1 > 0) tri_match_kernel<<<ceil_div(e1, 128), 128, 0, st>>>(P, h->ws0.as<int>(), h->out0.as<int>(), hist);
    tri_finish_kernel<<<1, 256, 0, st>>>(P.k1, P.k2, checkOri, h->out0.as<int>(), hist, hist + kHistoLength);
    h->launches += 4;
    ORB_CUDA(cudaGetLastError());
    ORB_CHECK(stage_download(h, matches12, h->out0.p, (size_t)n1 * 4, st));
    ORB_CHECK(stage_download(h, nmatches, hist + kHistoLength, 4, st));
    return stage_finish(h, st);
}

int orbm_search_by_bow(orbm_handle h, orbm_frame k1, orbm_frame k2, int nNodes1, const int* nodeId1, const int* start1,
                       const int* idx1, int nNodes2, const int* nodeId2, const int* start2, const int* idx2,
                       const uint8_t* valid1, const uint8_t* valid2, float ratio, int checkOri, int strictLow,
                       int* matches12, int* matches21, int* nmatches) {
    ORBM_ENTER(h);
    if (!k1 || !k2 || !matches12 || !matches21 || !nmatches || nNodes1 < 0 || nNodes2 < 0 ||
        (nNodes1 > 0 && (!nodeId1 || !start1 || !idx1)) || (nNodes2 > 0 && (!nodeId2 || !start2 || !idx2)))
        return fail(ORB_ERR_INVALID, "orbm_search_by_bow: bad arguments");
    *nmatches = 0;
    const int n1 = k1->n, n2 = k2->n;
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    for (int i = 0; i < n2; ++i) matches21[i] = -1;
    if (nNodes1 == 0 || nNodes2 == 0 || n1 == 0 || n2 == 0) return ORB_OK;
    const int e1 = start1[nNodes1], e2 = start2[nNodes2];
    for (int a = 1; a < nNodes1; ++a)
        if (nodeId1[a] <= nodeId1[a - 1]) return fail(ORB_ERR_INVALID, "orbm_search_by_bow: node ids of frame 1 not ascending");
    for (int a = 1; a < nNodes2; ++a)
        if (nodeId2[a] <= nodeId2[a - 1]) return fail(ORB_ERR_INVALID, "orbm_search_by_bow: node ids of frame 2 not ascending");
    if (e1 == 0 || e2 == 0) return ORB_OK;
    if (9 * (size_t)n2 + 32 > kReplaySmemMax) return fail(ORB_ERR_CAPACITY, "orbm_search_by_bow: %d keypoints exceed the replay state", n2);
    cudaStream_t st = h->stream;
    std::vector<int> ints;
    auto putInts = [&](const int* p, int n) { size_t o = ints.size(); ints.insert(ints.end(), p, p + n); return o; };
    const size_t oId1 = putInts(nodeId1, nNodes1), oS1 = putInts(start1, nNodes1 + 1), oI1 = putInts(idx1, e1);
    const size_t oId2 = putInts(nodeId2, nNodes2), oS2 = putInts(start2, nNodes2 + 1), oI2 = putInts(idx2, e2);
    ORB_CHECK(stage_upload(h, h->in0, ints.data(), ints.size() * 4, st));
    if (valid1) ORB_CHECK(stage_upload(h, h->in2, valid1, (size_t)n1, st));
    if (valid2) ORB_CHECK(stage_upload(h, h->in3, valid2, (size_t)n2, st));
    ORB_CHECK(h->in5.reserve((size_t)(e1 + 1) * 4));                     // entry -> node position
    ORB_CHECK(h->ws0.reserve((size_t)e1 * sizeof(AreaQuery)));
    ORB_CHECK(h->out4.reserve((size_t)(e1 + 1) * 4));
    ORB_CHECK(h->ws1.reserve((size_t)(e1 + 2) * 4));
    const int* di = h->in0.as<int>();
    BowParams P;
    P.k1 = k1->dev(); P.k2 = k2->dev();
    P.nNodes2 = nNodes2; P.nEntries1 = e1;
    P.nodeId1 = di + oId1; P.idx1 = di + oI1; P.nodeId2 = di + oId2; P.start2 = di + oS2; P.idx2 = di + oI2;
    P.valid1 = valid1 ? h->in2.as<unsigned char>() : nullptr;
    P.valid2 = valid2 ? h->in3.as<unsigned char>() : nullptr;
    int* counts = h->out4.as<int>();
    int* offsets = h->ws1.as<int>();
    tri_entry_node_kernel<<<ceil_div(nNodes1, 128), 128, 0, st>>>(nNodes1, di + oS1, h->in5.as<int>());
    const int wpb = 8, blocks = ceil_div(e1, wpb);
    bow_candidates_kernel<<<blocks, wpb * 32, 0, st>>>(P, h->in5.as<int>(), h->ws0.as<AreaQuery>(), counts, nullptr, nullptr);
    scan_kernel<<<1, 1024, 0, st>>>(counts, offsets, e1);
    int total = 0;
    ORB_CUDA(cudaMemcpyAsync(&total, offsets + e1, 4, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    ORB_CHECK(h->ws2.reserve((size_t)(total + 1) * sizeof(int2)));
    bow_candidates_kernel<<<blocks, wpb * 32, 0, st>>>(P, h->in5.as<int>(), h->ws0.as<AreaQuery>(), counts, offsets, h->ws2.as<int2>());
    ORB_CHECK(h->out0.reserve((size_t)(n1 + 1) * 4));
    ORB_CHECK(h->out1.reserve((size_t)(n2 + 1) * 4));
    ORB_CHECK(h->out2.reserve((size_t)(e1 + 1) * 4 * 2));
    ORB_CHECK(h->out3.reserve(16));
    int* pushA = h->out2.as<int>();
    ORB_CUDA(cudaFuncSetAttribute(bow_replay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kReplayFixed + kReplaySmemMax)));
    bow_replay_kernel<<<1, RP_THREADS, kReplayFixed + 9 * (size_t)n2 + 32, st>>>(P.k1, P.k2, h->ws0.as<AreaQuery>(), P.idx1, e1, offsets, h->ws2.as<int2>(), ratio,
                                                      checkOri, strictLow, h->out0.as<int>(), h->out1.as<int>(), pushA,
                                                      pushA + e1 + 1, h->out3.as<int>());
    h->launches += 5;
    ORB_CUDA(cudaGetLastError());
    ORB_CHECK(stage_download(h, matches12, h->out0.p, (size_t)n1 * 4, st));
    ORB_CHECK(stage_download(h, matches21, h->out1.p, (size_t)n2 * 4, st));
    ORB_CHECK(stage_download(h, nmatches, h->out3.p, 4, st));
    return stage_finish(h, st);
}

int orbm_search_projected_best(orbm_handle h, orbm_frame kf, const orbm_best_query* queries, const uint8_t* qdesc, int nq,
                               int chi2Filter, const float* uRight, const float* invSigma2, int nlevels, int* bestIdx,
                               int* bestDist) {
    ORBM_ENTER(h);
    if (!kf || nq < 0 || (nq > 0 && (!queries || !qdesc || !bestIdx || !bestDist)) || (chi2Filter && (!invSigma2 || nlevels < 1)))
        return fail(ORB_ERR_INVALID, "orbm_search_projected_best: bad arguments");
    if (nq == 0) return ORB_OK;
    const int n = kf->n;
    if (n == 0) {
        for (int i = 0; i < nq; ++i) { bestIdx[i] = -1; bestDist[i] = 256; }
        return ORB_OK;
    }
    cudaStream_t st = h->stream;
    ORB_CHECK(stage_upload(h, h->in0, queries, (size_t)nq * sizeof(orbm_best_query), st));
    ORB_CHECK(stage_upload(h, h->in1, qdesc, (size_t)nq * 32, st));
    if (chi2Filter) ORB_CHECK(stage_upload(h, h->in2, invSigma2, (size_t)nlevels * 4, st));
    if (chi2Filter && uRight) ORB_CHECK(stage_upload(h, h->in3, uRight, (size_t)n * 4, st));
    ORB_CHECK(h->out0.reserve((size_t)nq * 4));
    ORB_CHECK(h->out1.reserve((size_t)nq * 4));
    projected_best_kernel<<<ceil_div(nq, 8), 256, 0, st>>>(kf->dev(), h->in0.as<orbm_best_query>(), h->in1.as<uint4>(), nq, chi2Filter,
                                                           (chi2Filter && uRight) ? h->in3.as<float>() : nullptr,
                                                           chi2Filter ? h->in2.as<float>() : nullptr, h->out0.as<int>(), h->out1.as<int>());
    h->launches += 1;
    ORB_CUDA(cudaGetLastError());
    ORB_CHECK(stage_download(h, bestIdx, h->out0.p, (size_t)nq * 4, st));
    ORB_CHECK(stage_download(h, bestDist, h->out1.p, (size_t)nq * 4, st));
    return stage_finish(h, st);
}

}  // extern "C"
