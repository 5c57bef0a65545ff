// Hamming matcher kernels (north-star kernel 7).
//
//   bf_scan_kernel      every query of a pair against every train descriptor: best / second / index
//                       (inner loop of ORBmatcher.cc:432-457, distance = ORBmatcher.cc:1675-1691)
//   bf_accept_kernel    TH_LOW + ratio test and rotation-histogram pruning (ORBmatcher.cc:459-512)
//   allpairs_kernel     SearchByBoW(KF,KF) inner loop (ORBmatcher.cc:566-618, 634-652) for keyframe x keyframe tiles
//   distance_kernel     DescriptorDistance for independent pairs
//   popc_peak_kernel    POPC-pipe microbenchmark: the roofline denominator for matching
//
// Work shape: each thread keeps QPT query descriptors in registers (8 x u32 each) and streams the train descriptors
// of its pair from a shared-memory tile; all lanes read the same 32 B (two LDS.128 broadcasts), so per compare the
// SM issues 8 LOP3 + 8 POPC + 4 IADD3 + 4 min/max-type ops and no per-thread memory traffic.  No tensor cores:
// binary MMA is not a tcgen05 path.  The POPC pipe (8 POPC32 per 256-bit compare) is the bound.
//
// Order rule: the reference scans trains in index order with strict '<', so the lowest index wins ties for best and
// the second-best is the second smallest distance counted with multiplicity.  Both fall out of min/second-min over the
// packed key (distance << 22 | index), which is why nt is limited to 2^22.
#include "common.cuh"
#include "matcher.h"

namespace orbb {

constexpr int kKeyShift = 22;
constexpr int kKeyIdxMask = (1 << kKeyShift) - 1;
constexpr int kNoKey = 0x7fffffff;

__device__ __forceinline__ void key_update(int key, int& best, int& second) {
    second = min(second, max(key, best));
    best = min(best, key);
}

// ------------------------------------------------------------------------------------------------ brute force
constexpr int BF_THREADS = 128;
constexpr int BF_QPT = 4;                       // queries per thread
constexpr int BF_QTILE = BF_THREADS * BF_QPT;   // 512 queries per block
constexpr int BF_TTILE = 512;                   // train descriptors per shared-memory tile (16 KB)

__global__ void __launch_bounds__(BF_THREADS)
bf_scan_kernel(const uint4* __restrict__ queries, const uint4* __restrict__ trains, int nq, int nt,
               int* __restrict__ bestKey, int* __restrict__ secondKey) {
    __shared__ uint4 tile[BF_TTILE * 2];
    const int pair = blockIdx.y;
    const uint4* q = queries + (size_t)pair * nq * 2;
    const uint4* t = trains + (size_t)pair * nt * 2;
    const int q0 = blockIdx.x * BF_QTILE + threadIdx.x;   // thread's queries: q0 + k*BF_THREADS (coalesced output)

    uint4 qa[BF_QPT], qb[BF_QPT];
    int best[BF_QPT], second[BF_QPT];
#pragma unroll
    for (int k = 0; k < BF_QPT; ++k) {
        const int qi = min(q0 + k * BF_THREADS, nq - 1);
        qa[k] = __ldg(q + 2 * (size_t)qi);
        qb[k] = __ldg(q + 2 * (size_t)qi + 1);
        best[k] = kNoKey;
        second[k] = kNoKey;
    }

    for (int base = 0; base < nt; base += BF_TTILE) {
        const int cnt = min(BF_TTILE, nt - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * 2; i += BF_THREADS) tile[i] = __ldg(t + 2 * (size_t)base + i);
        __syncthreads();
#pragma unroll 2
        for (int j = 0; j < cnt; ++j) {
            const uint4 ta = tile[2 * j], tb = tile[2 * j + 1];
            const int jj = base + j;
#pragma unroll
            for (int k = 0; k < BF_QPT; ++k) {
                const int d = hamming256(qa[k], qb[k], ta, tb);
                key_update((d << kKeyShift) + jj, best[k], second[k]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < BF_QPT; ++k) {
        const int qi = q0 + k * BF_THREADS;
        if (qi < nq) {
            bestKey[(size_t)pair * nq + qi] = best[k];
            secondKey[(size_t)pair * nq + qi] = second[k];
        }
    }
}

// One block per pair. Reads packed keys, writes the reference-visible outputs.
__global__ void __launch_bounds__(256)
bf_accept_kernel(const int* __restrict__ bestKey, const int* __restrict__ secondKey, const float* __restrict__ qAngle,
                 const float* __restrict__ tAngle, int nq, int nt, float ratio, int checkOri, int* __restrict__ best,
                 int* __restrict__ second, int* __restrict__ idx, int* __restrict__ matches12,
                 int* __restrict__ nmatches) {
    __shared__ int hist[kHistoLength];
    __shared__ int keep[3];
    __shared__ int total;
    const int pair = blockIdx.x;
    const size_t off = (size_t)pair * nq;
    if (threadIdx.x < kHistoLength) hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < nq; i += blockDim.x) {
        const int bk = bestKey[off + i], sk = secondKey[off + i];
        const int bd = bk == kNoKey ? INT_MAX : (bk >> kKeyShift);
        const int sd = sk == kNoKey ? INT_MAX : (sk >> kKeyShift);
        const int bi = bk == kNoKey ? -1 : (bk & kKeyIdxMask);
        if (best) best[off + i] = bd;
        if (second) second[off + i] = sd;
        if (idx) idx[off + i] = bi;
        const bool ok = bd <= kThLow && (float)bd < __fmul_rn((float)sd, ratio);
        matches12[off + i] = ok ? bi : -1;
        if (ok) {
            atomicAdd(&total, 1);
            if (checkOri) atomicAdd(&hist[rotation_bin(qAngle[off + i], tAngle[(size_t)pair * nt + bi])], 1);
        }
    }
    __syncthreads();
    if (checkOri) {
        if (threadIdx.x == 0) three_maxima(hist, keep[0], keep[1], keep[2]);
        __syncthreads();
        for (int i = threadIdx.x; i < nq; i += blockDim.x) {
            const int m = matches12[off + i];
            if (m < 0) continue;
            const int bin = rotation_bin(qAngle[off + i], tAngle[(size_t)pair * nt + m]);
            if (bin != keep[0] && bin != keep[1] && bin != keep[2]) {
                matches12[off + i] = -1;
                atomicSub(&total, 1);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) nmatches[pair] = total;
}

// ------------------------------------------------------------------------------------------------ all pairs
// Block = one query keyframe x a run of db keyframes. 256 threads x 4 queries in registers; the db keyframe's
// descriptors (n_desc x 32 B) sit in shared memory, double buffered with cp.async so the next keyframe streams in
// while the POPC loop runs.  After each keyframe, warp 0 replays the reference's sequential one-to-one rule.
constexpr int AP_THREADS = 256;
constexpr int AP_QPT = 4;
constexpr int AP_MAXQ = AP_THREADS * AP_QPT;   // 1024 descriptors per keyframe at most

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

struct ApShared {
    uint4 db[2][AP_MAXQ * 2];     // 2 x 32 KB
    int bestKey[AP_MAXQ];
    int secondKey[AP_MAXQ];
    unsigned char matched2[AP_MAXQ];
    int hist[kHistoLength];
    int nCand;
};

// db keyframe d of the launch -> row of `table` and column of `counts`.  n == 0: row = d, column = colOffset + d; otherwise
// d runs over the concatenation of n segments (the other ranks' parts of one gathered chunk of the sharded table).
__device__ __forceinline__ void ap_locate(const ApSegments& sg, int d, int colOffset, int& row, int& col) {
    row = d;
    col = colOffset + d;
    for (int s = 0; s < sg.n; ++s)
        if (d >= sg.start[s] && d < sg.start[s + 1]) {
            row = sg.row[s] + d - sg.start[s];
            col = sg.col[s] + d - sg.start[s];
        }
}

__global__ void __launch_bounds__(AP_THREADS)
allpairs_kernel(const uint4* __restrict__ qTable, const float* __restrict__ qAngles, const uint4* __restrict__ table,
                const float* __restrict__ angles, int nDesc, int qBegin, int dbBegin, int dbEnd, int dbPerBlock, int nKfTotal,
                int colOffset, const __grid_constant__ ApSegments segs, float ratio, int checkOri, int* __restrict__ counts) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    ApShared& S = *reinterpret_cast<ApShared*>(smemRaw);
    const int qkf = qBegin + blockIdx.x;
    const int j0 = dbBegin + blockIdx.y * dbPerBlock;
    const int j1 = min(j0 + dbPerBlock, dbEnd);
    if (j0 >= j1) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const uint4* q = qTable + (size_t)qkf * nDesc * 2;
    uint4 qa[AP_QPT], qb[AP_QPT];
#pragma unroll
    for (int k = 0; k < AP_QPT; ++k) {
        const int qi = min(tid + k * AP_THREADS, nDesc - 1);
        qa[k] = __ldg(q + 2 * (size_t)qi);
        qb[k] = __ldg(q + 2 * (size_t)qi + 1);
    }
    auto prefetch = [&](int j, int buf) {
        int row, col;
        ap_locate(segs, j, colOffset, row, col);
        const uint4* src = table + (size_t)row * nDesc * 2;
        for (int i = tid; i < nDesc * 2; i += AP_THREADS) cp_async16(&S.db[buf][i], src + i);
        cp_async_commit();
    };
    prefetch(j0, 0);
    for (int j = j0; j < j1; ++j) {
        const int buf = (j - j0) & 1;
        int dbRow, dbCol;
        ap_locate(segs, j, colOffset, dbRow, dbCol);
        if (j + 1 < j1) { prefetch(j + 1, buf ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();

        int best[AP_QPT], second[AP_QPT];
#pragma unroll
        for (int k = 0; k < AP_QPT; ++k) best[k] = second[k] = kNoKey;
        const uint4* db = S.db[buf];
#pragma unroll 2
        for (int t = 0; t < nDesc; ++t) {
            const uint4 ta = db[2 * t], tb = db[2 * t + 1];
#pragma unroll
            for (int k = 0; k < AP_QPT; ++k) {
                const int d = hamming256(qa[k], qb[k], ta, tb);
                key_update((d << kKeyShift) + t, best[k], second[k]);
            }
        }
        // candidates: best < TH_LOW (ORBmatcher.cc:598); most keyframe pairs have none and finish here
        int mine = 0;
#pragma unroll
        for (int k = 0; k < AP_QPT; ++k) {
            const int qi = tid + k * AP_THREADS;
            const bool c = qi < nDesc && (best[k] >> kKeyShift) < kThLow;
            S.bestKey[qi] = c ? best[k] : kNoKey;
            S.secondKey[qi] = second[k];
            mine += c;
        }
        const int any = __syncthreads_count(mine);
        int result = 0;
        if (any) {
            for (int i = tid; i < AP_MAXQ; i += AP_THREADS) S.matched2[i] = 0;
            if (tid < kHistoLength) S.hist[tid] = 0;
            __syncthreads();
            if (warp == 0) {
                // sequential replay in query order (i1 ascending); lanes cooperate on the rare recomputation
                int nMatched = 0, accepted = 0;
                const float* a1 = qAngles + (size_t)qkf * nDesc;
                const float* a2 = angles + (size_t)dbRow * nDesc;
                for (int qbase = 0; qbase < nDesc; qbase += 32) {
                    const int qi = qbase + lane;
                    unsigned cand = __ballot_sync(0xffffffffu, qi < nDesc && S.bestKey[qi] != kNoKey);
                    while (cand) {
                        const int b = __ffs(cand) - 1;
                        cand &= cand - 1;
                        const int i1 = qbase + b;
                        int bk = S.bestKey[i1], sk = S.secondKey[i1];
                        if (nMatched > 0) {
                            // earlier matches removed trains from the pool (vbMatched2): rescan the unmatched ones
                            const uint4 xa = __ldg(q + 2 * (size_t)i1), xb = __ldg(q + 2 * (size_t)i1 + 1);
                            bk = sk = kNoKey;
                            for (int t = lane; t < nDesc; t += 32) {
                                if (S.matched2[t]) continue;
                                const int d = hamming256(xa, xb, db[2 * t], db[2 * t + 1]);
                                key_update((d << kKeyShift) + t, bk, sk);
                            }
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) {
                                const int ob = __shfl_xor_sync(0xffffffffu, bk, o);
                                const int os = __shfl_xor_sync(0xffffffffu, sk, o);
                                sk = min(min(sk, os), max(bk, ob));
                                bk = min(bk, ob);
                            }
                        }
                        // 256 = "no candidate" in the reference (bestDist1/2 start at 256)
                        const int bd = bk == kNoKey ? 256 : (bk >> kKeyShift);
                        const int sd = sk == kNoKey ? 256 : (sk >> kKeyShift);
                        if (bd < kThLow && (float)bd < __fmul_rn(ratio, (float)sd)) {
                            const int i2 = bk & kKeyIdxMask;
                            __syncwarp();   // all lanes are done reading matched2 for this query
                            if (lane == 0) {
                                S.matched2[i2] = 1;
                                if (checkOri) S.hist[rotation_bin(a1[i1], a2[i2])] += 1;
                            }
                            __syncwarp();
                            ++nMatched;
                            ++accepted;
                        }
                    }
                }
                if (lane == 0) {
                    int r = accepted;
                    if (checkOri) {
                        int i1, i2, i3;
                        three_maxima(S.hist, i1, i2, i3);
                        r = 0;
                        for (int b = 0; b < kHistoLength; ++b)
                            if (b == i1 || b == i2 || b == i3) r += S.hist[b];
                    }
                    S.nCand = r;
                }
            }
            __syncthreads();
            result = S.nCand;
        }
        if (tid == 0) counts[(size_t)blockIdx.x * nKfTotal + dbCol] = result;
        __syncthreads();   // db[buf] and the key arrays are reused two iterations later / next iteration
    }
}

// ------------------------------------------------------------------------------------------------ small kernels
__global__ void distance_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, int n, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = hamming256(a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1]);
}

constexpr int POPC_CHAINS = 8;
constexpr int POPC_UNROLL = 16;
__global__ void __launch_bounds__(256) popc_peak_kernel(unsigned* out, int iters, unsigned seed) {
    unsigned x[POPC_CHAINS];
#pragma unroll
    for (int c = 0; c < POPC_CHAINS; ++c) x[c] = seed * (threadIdx.x + 1) + c * 0x9e3779b9u + blockIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < POPC_UNROLL; ++u)
#pragma unroll
            for (int c = 0; c < POPC_CHAINS; ++c) asm volatile("popc.b32 %0, %0;" : "+r"(x[c]));
    }
    unsigned s = 0;
#pragma unroll
    for (int c = 0; c < POPC_CHAINS; ++c) s += x[c];
    if (s == 0xffffffffu) out[0] = s;   // never true (popc <= 32): keeps the chains live
}

// ------------------------------------------------------------------------------------------------ launchers
int launch_bruteforce(const uint8_t* dq, const float* dqa, int nq, const uint8_t* dt, const float* dta, int nt,
                      int nPairs, float ratio, int checkOri, int* dBestKey, int* dSecondKey, int* dBest, int* dSecond,
                      int* dIdx, int* dM12, int* dN, cudaStream_t st, int* launches) {
    if (nq <= 0 || nPairs <= 0) return ORB_OK;
    if (nt > kKeyIdxMask) return fail(ORB_ERR_INVALID, "bruteforce: nt=%d exceeds %d", nt, kKeyIdxMask);
    if (nPairs > 65535) return fail(ORB_ERR_INVALID, "bruteforce: n_pairs=%d exceeds 65535 per call", nPairs);
    dim3 grid(ceil_div(nq, BF_QTILE), nPairs);
    bf_scan_kernel<<<grid, BF_THREADS, 0, st>>>((const uint4*)dq, (const uint4*)dt, nq, nt, dBestKey, dSecondKey);
    bf_accept_kernel<<<nPairs, 256, 0, st>>>(dBestKey, dSecondKey, dqa, dta, nq, nt, ratio, checkOri, dBest, dSecond,
                                             dIdx, dM12, dN);
    if (launches) *launches += 2;
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

// Query keyframes [qBegin, qEnd) of (dQTable, dQAngles) against db keyframes [dbBegin, dbEnd) of (dTable, dAngles);
// counts[(q - qBegin) * ldCounts + colOffset + db].  The two tables may be the same one (single GPU) or the local block
// and a gathered chunk of another rank's block (sharded all-pairs).
int launch_allpairs_ex(const uint8_t* dQTable, const float* dQAngles, int qBegin, int qEnd, const uint8_t* dTable,
                       const float* dAngles, int dbBegin, int dbEnd, int nDesc, int ldCounts, int colOffset, float ratio,
                       int checkOri, int* dCounts, cudaStream_t st, int* launches, const ApSegments* segs) {
    if (nDesc < 1 || nDesc > AP_MAXQ) return fail(ORB_ERR_INVALID, "allpairs: n_desc=%d must be in 1..%d", nDesc, AP_MAXQ);
    if (qBegin < 0 || dbBegin < 0 || qBegin > qEnd || dbBegin > dbEnd) return fail(ORB_ERR_INVALID, "allpairs: bad ranges");
    const int nQ = qEnd - qBegin, nDb = dbEnd - dbBegin;
    if (nQ == 0 || nDb == 0) return ORB_OK;
    // per device and cheap; set every time so that matchers on different GPUs of one process all get it
    ORB_CUDA(cudaFuncSetAttribute(allpairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ApShared)));
    // long enough runs that the query registers are amortised, and at least ~8 waves of blocks (3 resident per SM): a
    // launch's last wave is only partly filled, which cost 10 % of a 64 x 512 keyframe launch at 2.3 waves
    int dbPerBlock = 64;
    while (dbPerBlock > 8 && (long long)nQ * ceil_div(nDb, dbPerBlock) < 148 * 3 * 8) dbPerBlock >>= 1;
    const int chunks = ceil_div(nDb, dbPerBlock);
    if (chunks > 65535) return fail(ORB_ERR_INVALID, "allpairs: too many db chunks");
    dim3 grid(nQ, chunks);
    ApSegments sg;
    if (segs) sg = *segs;
    else sg.n = 0;
    allpairs_kernel<<<grid, AP_THREADS, sizeof(ApShared), st>>>((const uint4*)dQTable, dQAngles, (const uint4*)dTable, dAngles,
                                                                nDesc, qBegin, dbBegin, dbEnd, dbPerBlock, ldCounts, colOffset,
                                                                sg, ratio, checkOri, dCounts);
    if (launches) *launches += 1;
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

int launch_allpairs(const uint8_t* dTable, const float* dAngles, int nKf, int nDesc, int qBegin, int qEnd, int dbBegin,
                    int dbEnd, float ratio, int checkOri, int* dCounts, cudaStream_t st, int* launches) {
    if (qEnd > nKf || dbEnd > nKf) return fail(ORB_ERR_INVALID, "allpairs: bad ranges");
    return launch_allpairs_ex(dTable, dAngles, qBegin, qEnd, dTable, dAngles, dbBegin, dbEnd, nDesc, nKf, 0, ratio, checkOri,
                              dCounts, st, launches, nullptr);
}

int launch_distance(const uint8_t* da, const uint8_t* db, int n, int* dOut, cudaStream_t st, int* launches) {
    if (n <= 0) return ORB_OK;
    distance_kernel<<<ceil_div(n, 256), 256, 0, st>>>((const uint4*)da, (const uint4*)db, n, dOut);
    if (launches) *launches += 1;
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

// ------------------------------------------------------------------------------------------------ distinctive descriptor
// MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:257-322) for a batch of map points.  The reference fills an N x N
// float matrix, sorts every row and takes element (size_t)(0.5*(N-1)).  Distances are integers in [0, 256], so the median
// of a row is read off a 257-bin histogram: CTA per map point, warp per row; lanes stride over the columns and count into
// the warp's shared histogram, then a warp scan over the bins (9 per lane) finds the first bin whose cumulative count
// exceeds the median rank.  (median, row) keys are min-reduced per CTA: the first row with the smallest median wins.
constexpr int kDistWarps = 8;
constexpr int kDistBins = 288;   // 257 bins padded to 9 per lane

__global__ void __launch_bounds__(kDistWarps * 32)
distinctive_kernel(const uint4* __restrict__ desc, const int* __restrict__ start, int nPoints, int* __restrict__ best,
                   int* __restrict__ bestMedian) {
    __shared__ int hist[kDistWarps][kDistBins];
    __shared__ unsigned bestKey;
    const int p = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s0 = start[p], n = start[p + 1] - s0;
    if (threadIdx.x == 0) bestKey = 0xffffffffu;
    __syncthreads();
    if (n > 0) {
        const uint4* d = desc + 2 * (size_t)s0;
        const int rank = (int)(0.5 * (double)(n - 1));          // vDists[0.5*(N-1)], :311
        for (int i = warp; i < n; i += kDistWarps) {
            for (int b = lane; b < kDistBins; b += 32) hist[warp][b] = 0;
            __syncwarp();
            const uint4 a0 = __ldg(&d[2 * i]), a1 = __ldg(&d[2 * i + 1]);
            for (int j = lane; j < n; j += 32)
                atomicAdd(&hist[warp][hamming256(a0, a1, __ldg(&d[2 * j]), __ldg(&d[2 * j + 1]))], 1);
            __syncwarp();
            int c[9], sum = 0;
#pragma unroll
            for (int k = 0; k < 9; ++k) { c[k] = hist[warp][lane * 9 + k]; sum += c[k]; }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            int before = incl - sum, median = 0x7fff;
#pragma unroll
            for (int k = 0; k < 9; ++k) {                       // first bin with cumulative count > rank
                if (median == 0x7fff && before + c[k] > rank) median = lane * 9 + k;
                before += c[k];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) median = min(median, __shfl_xor_sync(0xffffffffu, median, o));
            if (lane == 0) atomicMin(&bestKey, ((unsigned)median << 16) | (unsigned)i);
            __syncwarp();
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        best[p] = n > 0 ? (int)(bestKey & 0xffffu) : -1;
        if (bestMedian) bestMedian[p] = n > 0 ? (int)(bestKey >> 16) : 0x7fffffff;
    }
}

int launch_distinctive(const uint8_t* dDesc, const int* dStart, int nPoints, int* dBest, int* dBestMedian, cudaStream_t st,
                       int* launches) {
    if (nPoints <= 0) return ORB_OK;
    distinctive_kernel<<<nPoints, kDistWarps * 32, 0, st>>>((const uint4*)dDesc, dStart, nPoints, dBest, dBestMedian);
    if (launches) *launches += 1;
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

int measure_popc_peak(cudaStream_t st, double* popcPerS) {
    unsigned* d = nullptr;
    ORB_CUDA(cudaMalloc(&d, 4));
    const int blocks = 148 * 8, iters = 4096;
    cudaEvent_t e0, e1;
    ORB_CUDA(cudaEventCreate(&e0));
    ORB_CUDA(cudaEventCreate(&e1));
    double bestRate = 0;
    for (int rep = 0; rep < 5; ++rep) {
        ORB_CUDA(cudaEventRecord(e0, st));
        popc_peak_kernel<<<blocks, 256, 0, st>>>(d, iters, 12345u + rep);
        ORB_CUDA(cudaEventRecord(e1, st));
        ORB_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        ORB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double n = (double)blocks * 256 * iters * POPC_CHAINS * POPC_UNROLL;
        if (rep > 0) bestRate = fmax(bestRate, n / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *popcPerS = bestRate;
    return ORB_OK;
}

}  // namespace orbb
