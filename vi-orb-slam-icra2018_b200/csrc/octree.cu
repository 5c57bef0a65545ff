// Quadtree keypoint selection (north-star kernel 3):
//   ORBextractor::DistributeOctTree  ORBextractor.cc:541-765   (ExtractorNode::DivideNode :483-539)
//
// One CTA per (frame, level).  The reference manipulates a std::list sequentially; the same result is obtained in
// level-synchronous rounds because of three facts (checked against a literal list simulation, tests/test_octree*):
//   1. A breadth-first pass splits every node that holds >1 key, pushes the non-empty children n1..n4 to the list
//      FRONT and erases the parent, so afterwards the list is  reverse(children in processing order) ++ (unsplit
//      nodes in their old order).  Positions therefore come from two prefix scans.
//   2. A key's child is two integer comparisons against the parent's midpoint (x0+ceil(w/2), y0+ceil(h/2)); keys keep
//      their relative order inside a child, so "first key with the maximal response" == lowest original index.
//   3. The careful phase (:675-740) walks the nodes created in the previous round by (size desc, address desc) and
//      stops as soon as the list reaches N nodes: a prefix scan of (children-1) over that order finds the cut.
//      ADDRESS ORDER IS DEFINED AS CREATION ORDER (the reference leaves it to the allocator, :686); a later-created
//      node sits nearer the list front, so ties on size resolve toward the smaller list index.
// Candidate keys come from the per-cell FAST slots in (cell row, cell col, y, x) order and live in shared memory
// (global spill space if a level has more candidates than fit).  The selected keys get their orientation in the
// descriptor kernel (brief.cu), which has a warp per keypoint anyway.
#include "extractor.h"

namespace orbb {

constexpr int OT_THREADS = 256;
constexpr int OT_WARPS = OT_THREADS / 32;

struct OtLayout {   // byte offsets into dynamic shared memory
    int boxA, boxB, cntA, cntB, newA, newB, child, t0, t1, t2, t3, cellOff, keyVal, keyNode, total;
};

__host__ __device__ inline OtLayout ot_layout(int nodeCap, int cellCap, int keyCap) {
    OtLayout o;
    int p = 0;
    auto take = [&](int bytes) { int r = p; p += (bytes + 15) & ~15; return r; };
    o.boxA = take(nodeCap * 8); o.boxB = take(nodeCap * 8);
    o.cntA = take(nodeCap * 4); o.cntB = take(nodeCap * 4);
    o.newA = take(nodeCap); o.newB = take(nodeCap);
    o.child = take(nodeCap * 16);
    o.t0 = take(nodeCap * 4); o.t1 = take(nodeCap * 4); o.t2 = take(nodeCap * 4); o.t3 = take(nodeCap * 4);
    o.cellOff = take((cellCap + 1) * 4);
    o.keyVal = take(keyCap * 4);
    o.keyNode = take(keyCap * 2);
    o.total = p;
    return o;
}

__device__ __forceinline__ int key_x(unsigned int k) { return (int)(k >> 20); }
__device__ __forceinline__ int key_y(unsigned int k) { return (int)((k >> 8) & 0xfffu); }
__device__ __forceinline__ int key_s(unsigned int k) { return (int)(k & 0xffu); }

// child of a node for a key: 0 = n1 (upper-left), 1 = n2 (upper-right), 2 = n3 (lower-left), 3 = n4 (lower-right):
//   mx = b.x + ((b.z - b.x + 1) >> 1)  (UL.x + ceil((UR.x-UL.x)/2)), my likewise;  (key_x < mx ? 0 : 1) + (key_y < my ? 0 : 2)
// The same decision on the packed words: the box is two words (x | y << 16), (z | w << 16) of non-negative shorts, so one
// three-input add gives (x + z + 1) | (y + w + 1) << 16, and mx = (x + z + 1) >> 1, my = (y + w + 1) >> 1 are the midpoints
// above (b.x + ((b.z - b.x + 1) >> 1) == (b.x + b.z + 1) >> 1).  A key holds x in bits 20..31 and y in bits 8..19, so
//   key_x(key) < mx  <=>  key < (mx << 20)   and   key_y(key) < my  <=>  (key & 0xfff00) < (my << 8):
// no field extraction and no sign extension per (key, round); child_of was 18 % of the kernel's instructions
// (2.23 -> 2.04 ms per 4096 frames with the thresholds kept per node in shared memory, which cost levels with many keys
// their place there; this form needs no table).
__device__ __forceinline__ int child_of(short4 b, unsigned int key) {
    const uint2 w = *reinterpret_cast<const uint2*>(&b);
    const unsigned int s = w.x + w.y + 0x00010001u;
    const unsigned int tx = (s & 0xfffeu) << 19;          // ((s & 0xffff) >> 1) << 20
    const unsigned int ty = (s >> 9) & 0xffffff00u;       // (s >> 17) << 8
    return (key < tx ? 0 : 1) + ((key & 0xfff00u) < ty ? 0 : 2);
}
__device__ __forceinline__ short4 child_box(short4 b, int c) {
    const short mx = (short)(b.x + ((b.z - b.x + 1) >> 1));
    const short my = (short)(b.y + ((b.w - b.y + 1) >> 1));
    short4 r;
    r.x = (c & 1) ? mx : b.x;
    r.z = (c & 1) ? b.z : mx;
    r.y = (c & 2) ? my : b.y;
    r.w = (c & 2) ? b.w : my;
    return r;
}

// Exclusive scan of a[0..n) in shared memory by the whole block; returns the total. `tmp` holds OT_WARPS+1 ints.
__device__ int block_scan_exclusive(int* a, int n, int* tmp) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + OT_THREADS - 1) / OT_THREADS;
    const int i0 = min(tid * per, n), i1 = min(i0 + per, n);
    int sum = 0;
    for (int i = i0; i < i1; ++i) sum += a[i];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    __syncthreads();   // tmp may still be read from a previous call
    if (lane == 31) tmp[warp] = incl;
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < OT_WARPS; ++w) {
        const int s = tmp[w];
        if (w < warp) base += s;
        total += s;
    }
    int run = base + incl - sum;
    for (int i = i0; i < i1; ++i) {
        const int v = a[i];
        a[i] = run;
        run += v;
    }
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(OT_THREADS, 4)
octree_kernel(const __grid_constant__ ExtractParams P, int nodeCap, int cellCap, int keyCap) {
    extern __shared__ __align__(16) unsigned char sm[];
    __shared__ int scanTmp[OT_WARPS + 1];
    __shared__ int ctl[8];   // 0: nToExpand, 1: cut, 2: scratch
    const int frame = blockIdx.x, level = blockIdx.y;   // level-major dispatch: the long level-0 CTAs of all frames start first
    const LevelGeom& L = P.lv[level];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int* selCountOut = P.selCount + (size_t)frame * P.nLevels + level;

    const OtLayout lay = ot_layout(nodeCap, cellCap, keyCap);
    short4* box = reinterpret_cast<short4*>(sm + lay.boxA);
    short4* boxN = reinterpret_cast<short4*>(sm + lay.boxB);
    int* cnt = reinterpret_cast<int*>(sm + lay.cntA);
    int* cntN = reinterpret_cast<int*>(sm + lay.cntB);
    unsigned char* isNew = sm + lay.newA;
    unsigned char* isNewN = sm + lay.newB;
    int* child = reinterpret_cast<int*>(sm + lay.child);
    int* t0 = reinterpret_cast<int*>(sm + lay.t0);   // children per processed node -> position base
    int* t1 = reinterpret_cast<int*>(sm + lay.t1);   // processing rank per node (-1: not processed)
    int* t2 = reinterpret_cast<int*>(sm + lay.t2);   // survivor index per node
    int* t3 = reinterpret_cast<int*>(sm + lay.t3);   // node by processing rank
    int* cellOff = reinterpret_cast<int*>(sm + lay.cellOff);

    // ---- 1. candidate list = concatenation of the level's cell slots in (row, col) order
    const int nCells = L.nCells;
    for (int c = tid; c < nCells; c += OT_THREADS) cellOff[c] = P.cellCount[(size_t)frame * P.nCellsTotal + L.cellBase + c];
    __syncthreads();
    const int n = nCells > 0 ? block_scan_exclusive(cellOff, nCells, scanTmp) : 0;
    if (n == 0) {
        if (tid == 0) *selCountOut = 0;
        return;
    }
    unsigned int* kv;
    unsigned short* kn;
    if (n <= keyCap) {
        kv = reinterpret_cast<unsigned int*>(sm + lay.keyVal);
        kn = reinterpret_cast<unsigned short*>(sm + lay.keyNode);
    } else {
        kv = P.keyWs + (size_t)frame * P.keyWsFrameEntries + L.keyWsOff;
        kn = reinterpret_cast<unsigned short*>(kv + L.keyWsCap);
    }
    if (tid == 0) cellOff[nCells] = n;
    __syncthreads();
    {
        // One thread per key: its cell is the last c with cellOff[c] <= k (binary search in shared memory; an empty cell
        // shares its offset with its successor and is never picked), its slot follows from the cell index (the level's
        // slots are laid out cell after cell, extractor.cu::configure).  Every global load is independent of every other:
        // the warp-per-cell walk this replaces was a chain of ~70 dependent loads per warp on level 0, 17 % of the
        // kernel's stall samples and most of a single frame's quadtree latency.
        const unsigned int* slots = P.slots + (size_t)frame * P.slotFrameEntries + L.slotBase;
        const int slotCap = L.slotCap;
        for (int k = tid; k < n; k += OT_THREADS) {
            int lo = 0, hi = nCells;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (cellOff[mid] <= k) lo = mid; else hi = mid;
            }
            kv[k] = slots[(size_t)lo * slotCap + (k - cellOff[lo])];
        }
    }

    // ---- 2. roots (:545-587)
    const int N = L.nFeatures;
    const int nIni = L.nIni;
    for (int r = tid; r < nIni; r += OT_THREADS) {
        short4 b;
        b.x = (short)(int)__fmul_rn(L.hX, (float)r);
        b.z = (short)(int)__fmul_rn(L.hX, (float)(r + 1));
        b.y = 0;
        b.w = (short)L.winH;
        boxN[r] = b;
        cntN[r] = 0;
    }
    __syncthreads();
    for (int k = tid; k < n; k += OT_THREADS) {
        const int r = (int)__fdiv_rn((float)key_x(kv[k]), L.hX);
        kn[k] = (unsigned short)r;
        atomicAdd(&cntN[r], 1);
    }
    __syncthreads();
    if (tid == 0) {   // drop empty roots, keep order (nIni is a handful)
        int s = 0;
        for (int r = 0; r < nIni; ++r) {
            t2[r] = s;
            if (cntN[r] > 0) { box[s] = boxN[r]; cnt[s] = cntN[r]; isNew[s] = 0; ++s; }
        }
        ctl[2] = s;
    }
    __syncthreads();
    int S = ctl[2];
    for (int k = tid; k < n; k += OT_THREADS) kn[k] = (unsigned short)t2[kn[k]];
    __syncthreads();

    // ---- 3. rounds
    bool careful = false;
    while (true) {
        for (int i = tid; i < 4 * S; i += OT_THREADS) child[i] = 0;
        if (tid == 0) { ctl[0] = 0; ctl[1] = 0x7fffffff; }
        __syncthreads();
        // children sizes of every candidate node
        for (int k = tid; k < n; k += OT_THREADS) {
            const int i = kn[k];
            if (cnt[i] > 1 && (!careful || isNew[i])) atomicAdd(&child[4 * i + child_of(box[i], kv[k])], 1);
        }
        __syncthreads();
        // t0 = 1 for candidates (scanned below), t1 = -1
        for (int i = tid; i < S; i += OT_THREADS) {
            t0[i] = (cnt[i] > 1 && (!careful || isNew[i])) ? 1 : 0;
            t1[i] = -1;
        }
        __syncthreads();
        const int nCand = block_scan_exclusive(t0, S, scanTmp);   // t0[i] = index among candidates (list order)
        int nProc = nCand;
        if (!careful) {
            for (int i = tid; i < S; i += OT_THREADS)
                if (cnt[i] > 1) { t3[t0[i]] = i; t1[i] = t0[i]; }
        } else {
            // compact candidates into t2, then rank by (size desc, list index asc) == sort(size,address) walked from the back
            for (int i = tid; i < S; i += OT_THREADS)
                if (cnt[i] > 1 && isNew[i]) t2[t0[i]] = i;
            __syncthreads();
            for (int a = tid; a < nCand; a += OT_THREADS) {
                const int ia = t2[a], sa = cnt[ia];
                int rank = 0;
                for (int b = 0; b < nCand; ++b) {
                    const int sb = cnt[t2[b]];
                    rank += (sb > sa) || (sb == sa && b < a);
                }
                t3[rank] = ia;
            }
            __syncthreads();
            // running list size after each split; first rank that reaches N ends the round (:732-733)
            for (int r = tid; r < nCand; r += OT_THREADS) {
                const int i = t3[r];
                t0[r] = (child[4 * i] > 0) + (child[4 * i + 1] > 0) + (child[4 * i + 2] > 0) + (child[4 * i + 3] > 0) - 1;
            }
            __syncthreads();
            block_scan_exclusive(t0, nCand, scanTmp);
            for (int r = tid; r < nCand; r += OT_THREADS) {
                const int i = t3[r];
                const int gain = (child[4 * i] > 0) + (child[4 * i + 1] > 0) + (child[4 * i + 2] > 0) + (child[4 * i + 3] > 0) - 1;
                if (S + t0[r] + gain >= N) atomicMin(&ctl[1], r);
            }
            __syncthreads();
            if (ctl[1] != 0x7fffffff) nProc = ctl[1] + 1;
            for (int r = tid; r < nProc; r += OT_THREADS) t1[t3[r]] = r;
        }
        __syncthreads();
        // position base of each processed node's children in creation order
        for (int r = tid; r < nProc; r += OT_THREADS) {
            const int i = t3[r];
            t0[r] = (child[4 * i] > 0) + (child[4 * i + 1] > 0) + (child[4 * i + 2] > 0) + (child[4 * i + 3] > 0);
        }
        __syncthreads();
        const int E = nProc > 0 ? block_scan_exclusive(t0, nProc, scanTmp) : 0;
        // survivors keep their order behind the new nodes
        for (int i = tid; i < S; i += OT_THREADS) t2[i] = t1[i] < 0 ? 1 : 0;
        __syncthreads();
        block_scan_exclusive(t2, S, scanTmp);
        // build the new list
        for (int i = tid; i < S; i += OT_THREADS) {
            const int r = t1[i];
            if (r < 0) {
                const int ni = E + t2[i];
                boxN[ni] = box[i]; cntN[ni] = cnt[i]; isNewN[ni] = 0;
                t2[i] = ni;
            } else {
                int pos = t0[r], expandable = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int m = child[4 * i + c];
                    if (m > 0) {
                        const int ni = E - 1 - pos;
                        boxN[ni] = child_box(box[i], c); cntN[ni] = m; isNewN[ni] = 1;
                        child[4 * i + c] = ni;
                        expandable += m > 1;
                        ++pos;
                    }
                }
                if (expandable) atomicAdd(&ctl[0], expandable);
            }
        }
        __syncthreads();
        for (int k = tid; k < n; k += OT_THREADS) {
            const int i = kn[k];
            kn[k] = (unsigned short)(t1[i] < 0 ? t2[i] : child[4 * i + child_of(box[i], kv[k])]);
        }
        const int newS = E + (S - nProc);
        const int nToExpand = ctl[0];
        __syncthreads();
        { short4* tb = box; box = boxN; boxN = tb; }
        { int* tc = cnt; cnt = cntN; cntN = tc; }
        { unsigned char* tn = isNew; isNew = isNewN; isNewN = tn; }
        const int oldS = S;
        S = newS;
        if (newS >= N || newS == oldS) break;
        if (!careful && newS + 3 * nToExpand > N) careful = true;
    }

    // ---- 4. best key per node: max response, lowest original index on ties (:744-762)
    for (int i = tid; i < S; i += OT_THREADS) { t0[i] = -1; t1[i] = 0x7fffffff; }
    __syncthreads();
    for (int k = tid; k < n; k += OT_THREADS) atomicMax(&t0[kn[k]], key_s(kv[k]));
    __syncthreads();
    for (int k = tid; k < n; k += OT_THREADS)
        if (key_s(kv[k]) == t0[kn[k]]) atomicMin(&t1[kn[k]], k);
    __syncthreads();

    // ---- 5. emit in list order; the orientation is computed where the descriptor is (brief.cu)
    const int nOut = min(S, L.selCap);
    SelKey* out = P.sel + (size_t)frame * P.selPerFrame + L.selBase;
    for (int s = tid; s < nOut; s += OT_THREADS) {
        const unsigned int key = kv[t1[s]];
        SelKey k;
        k.x = (float)(key_x(key) + 16); k.y = (float)(key_y(key) + 16);   // + minBorderX/Y (:843-844)
        k.response = (float)key_s(key);
        k.angle = 0.f;
        out[s] = k;
    }
    if (tid == 0) *selCountOut = nOut;
}

int octree_smem_plan(int nodeCap, int cellCap, int* smemBytes, int* keyCapSmem) {
    // Occupancy matters more than capacity: the kernel is latency-bound (scans, barriers, shared atomics), so aim for
    // 4 CTAs per SM (~55 KB each). Levels with more candidates than fit spill their key list to global memory.
    const int hardBudget = 200 * 1024;
    const OtLayout fixed = ot_layout(nodeCap, cellCap, 0);
    if (fixed.total > hardBudget - 6 * 1024)
        return fail(ORB_ERR_INVALID, "quadtree tables (%d B for %d nodes, %d cells) exceed shared memory", fixed.total, nodeCap, cellCap);
    const int budget = fixed.total + 24 * 1024 < 55 * 1024 ? 55 * 1024 : (fixed.total + 24 * 1024 < hardBudget ? fixed.total + 24 * 1024 : hardBudget);
    int keyCap = ((budget - fixed.total - 64) / 6) & ~7;
    *keyCapSmem = keyCap;
    *smemBytes = ot_layout(nodeCap, cellCap, keyCap).total;
    return ORB_OK;
}

int launch_octree(const ExtractParams& P, int smemBytes, int keyCapSmem, int nodeCap, int cellCap, cudaStream_t st,
                  int* launches) {
    // per device and cheap; set every time so that handles on different GPUs of one process all get it
    ORB_CUDA(cudaFuncSetAttribute(octree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes));
    dim3 grid(P.nFrames, P.nLevels);
    octree_kernel<<<grid, OT_THREADS, smemBytes, st>>>(P, nodeCap, cellCap, keyCapSmem);
    ++*launches;
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orbb
