// Vocabulary-tree descent behind Frame::ComputeBoW / KeyFrame::ComputeBoW (Frame.cc:736-745, KeyFrame.cc:393-401):
//   TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup)   Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1443-1485
//   FORB::distance                                                             Thirdparty/DBoW2/DBoW2/FORB.cpp:84-105
// and the host-side assembly of the two containers the reference fills from it:
//   BowVector::addWeight / normalize(L1), FeatureVector::addFeature            BowVector.cpp:31-43, 59-81; FeatureVector.cpp:30-44
//
// The tree lives in HBM as flat arrays (node descriptors as 2 x uint4, child lists as CSR, word id and idf weight per
// node).  One warp per feature: at every level the lanes take one child each (strided when a node has more than 32),
// (distance << 16 | child position) is min-reduced by shuffle -- the reference's strict '<' keeps the first of equal
// children, which is the smallest position -- and the warp moves to that child until it stands on a leaf.  k = 10, L = 6
// is 60 compares per feature: the kernel is latency-bound on the dependent node loads, not on POPC.
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

#include "matcher.h"

using namespace orbb;

struct orbm_vocabulary_s {
    orbm_matcher* m = nullptr;
    int nNodes = 0, L = 0, depth = 0;
    DevBuf desc, childStart, children, wordId, weight;
};

namespace orbb {

struct VocabDev {
    const uint4* desc;
    const int* childStart;
    const int* children;
    const int* wordId;
    const double* weight;
    int L;
};

constexpr int kBowWarps = 8;

__global__ void __launch_bounds__(kBowWarps * 32)
bow_descend_kernel(VocabDev V, const uint4* __restrict__ feat, int n, int levelsup, int* __restrict__ word,
                   double* __restrict__ weight, int* __restrict__ node) {
    const int i = blockIdx.x * kBowWarps + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const uint4 a0 = feat[2 * i], a1 = feat[2 * i + 1];
    const int nidLevel = V.L - levelsup;          // TemplatedVocabulary.h:1452-1453
    int nid = 0, cur = 0, level = 0;
    int c0 = V.childStart[0], c1 = V.childStart[1];
    while (c1 > c0) {                             // do { ... } while (!isLeaf()); the root of a non-empty tree has children
        ++level;
        unsigned best = 0xffffffffu;
        for (int c = c0 + lane; c < c1; c += 32) {
            const int id = V.children[c];
            const unsigned d = (unsigned)hamming256(a0, a1, V.desc[2 * id], V.desc[2 * id + 1]);
            best = min(best, (d << 16) | (unsigned)(c - c0));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        cur = V.children[c0 + (int)(best & 0xffffu)];
        if (level == nidLevel) nid = cur;
        c0 = V.childStart[cur];
        c1 = V.childStart[cur + 1];
    }
    if (lane == 0) {
        word[i] = V.wordId[cur];
        weight[i] = V.weight[cur];
        node[i] = nid;
    }
}

}  // namespace orbb

namespace {

// m_nodes must be a tree rooted at node 0: every other node reachable exactly once.  Returns its depth, or -1.
int tree_depth(int nNodes, const int* childStart, const int* children) {
    if (childStart[0] != 0) return -1;
    for (int i = 0; i < nNodes; ++i)
        if (childStart[i + 1] < childStart[i]) return -1;
    const int nEdges = childStart[nNodes];
    if (nEdges != nNodes - 1) return -1;
    std::vector<char> seen(nNodes, 0);
    std::vector<int> frontier(1, 0), next;
    seen[0] = 1;
    int depth = 0, visited = 1;
    while (!frontier.empty()) {
        next.clear();
        for (int u : frontier)
            for (int c = childStart[u]; c < childStart[u + 1]; ++c) {
                const int v = children[c];
                if (v <= 0 || v >= nNodes || seen[v]) return -1;
                seen[v] = 1;
                ++visited;
                next.push_back(v);
            }
        if (!next.empty()) ++depth;
        frontier.swap(next);
    }
    return visited == nNodes ? depth : -1;
}

}  // namespace

extern "C" {

int orbm_vocabulary_create(orbm_handle h, int nNodes, int L, const uint8_t* nodeDesc, const int* childStart,
                           const int* children, const int* wordId, const double* weight, orbm_vocabulary* out) {
    ORBM_ENTER(h);
    if (!out) return fail(ORB_ERR_INVALID, "orbm_vocabulary_create: null out");
    *out = nullptr;
    if (nNodes < 2 || L < 1 || !nodeDesc || !childStart || !children || !wordId || !weight)
        return fail(ORB_ERR_INVALID, "orbm_vocabulary_create: bad arguments (an empty vocabulary transforms to nothing)");
    const int depth = tree_depth(nNodes, childStart, children);
    if (depth < 1) return fail(ORB_ERR_INVALID, "orbm_vocabulary_create: the child lists do not form a tree rooted at node 0");
    for (int i = 0; i < nNodes; ++i)
        if (childStart[i + 1] - childStart[i] > 65535)
            return fail(ORB_ERR_INVALID, "orbm_vocabulary_create: node %d has more than 65535 children", i);
    orbm_vocabulary_s* v = new orbm_vocabulary_s;
    v->m = h; v->nNodes = nNodes; v->L = L; v->depth = depth;
    cudaStream_t st = h->stream;
    auto body = [&]() -> int {
        const size_t nE = (size_t)childStart[nNodes];
        ORB_CHECK(v->desc.reserve((size_t)nNodes * 32));
        ORB_CHECK(v->childStart.reserve((size_t)(nNodes + 1) * 4));
        ORB_CHECK(v->children.reserve(nE * 4));
        ORB_CHECK(v->wordId.reserve((size_t)nNodes * 4));
        ORB_CHECK(v->weight.reserve((size_t)nNodes * 8));
        ORB_CUDA(cudaMemcpyAsync(v->desc.p, nodeDesc, (size_t)nNodes * 32, cudaMemcpyHostToDevice, st));
        ORB_CUDA(cudaMemcpyAsync(v->childStart.p, childStart, (size_t)(nNodes + 1) * 4, cudaMemcpyHostToDevice, st));
        ORB_CUDA(cudaMemcpyAsync(v->children.p, children, nE * 4, cudaMemcpyHostToDevice, st));
        ORB_CUDA(cudaMemcpyAsync(v->wordId.p, wordId, (size_t)nNodes * 4, cudaMemcpyHostToDevice, st));
        ORB_CUDA(cudaMemcpyAsync(v->weight.p, weight, (size_t)nNodes * 8, cudaMemcpyHostToDevice, st));
        ORB_CUDA(cudaStreamSynchronize(st));
        return ORB_OK;
    };
    const int status = body();
    if (status != ORB_OK) {
        orbm_vocabulary_destroy(v);
        return status;
    }
    *out = v;
    return ORB_OK;
}

int orbm_vocabulary_destroy(orbm_vocabulary v) {
    if (!v) return ORB_OK;
    DeviceGuard g(v->m->device);
    v->desc.release(); v->childStart.release(); v->children.release(); v->wordId.release(); v->weight.release();
    delete v;
    return ORB_OK;
}

int orbm_bow_transform_device(orbm_handle h, orbm_vocabulary v, const uint8_t* dDesc, int n, int levelsup, int* dWord,
                              double* dWeight, int* dNode, void* stream) {
    ORBM_ENTER(h);
    if (!v || v->m != h) return fail(ORB_ERR_INVALID, "orbm_bow_transform_device: vocabulary of another matcher (or null)");
    if (n < 0 || (n > 0 && (!dDesc || !dWord || !dWeight || !dNode)))
        return fail(ORB_ERR_INVALID, "orbm_bow_transform_device: bad arguments");
    if (n == 0) return ORB_OK;
    if (((uintptr_t)dDesc & 15) != 0) return fail(ORB_ERR_INVALID, "orbm_bow_transform_device: descriptors must be 16-byte aligned");
    VocabDev V{v->desc.as<uint4>(), v->childStart.as<int>(), v->children.as<int>(), v->wordId.as<int>(), v->weight.as<double>(), v->L};
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    bow_descend_kernel<<<ceil_div(n, kBowWarps), kBowWarps * 32, 0, st>>>(V, (const uint4*)dDesc, n, levelsup, dWord, dWeight, dNode);
    h->launches += 1;
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

int orbm_bow_transform(orbm_handle h, orbm_vocabulary v, const uint8_t* desc, int n, int levelsup, int* word, double* weight,
                       int* node, int* bowWord, double* bowValue, int* nWords, int* fvNode, int* fvStart, int* fvIdx,
                       int* nNodes) {
    ORBM_ENTER(h);
    if (!v || v->m != h) return fail(ORB_ERR_INVALID, "orbm_bow_transform: vocabulary of another matcher (or null)");
    if (n < 0 || (n > 0 && !desc)) return fail(ORB_ERR_INVALID, "orbm_bow_transform: bad arguments");
    const bool wantBow = bowWord || bowValue || nWords, wantFv = fvNode || fvStart || fvIdx || nNodes;
    if ((wantBow && !(bowWord && bowValue && nWords)) || (wantFv && !(fvNode && fvStart && fvIdx && nNodes)))
        return fail(ORB_ERR_INVALID, "orbm_bow_transform: give all of bow_word/bow_value/n_words (or none), likewise the feature vector");
    if (nWords) *nWords = 0;
    if (nNodes) *nNodes = 0;
    if (fvStart) fvStart[0] = 0;
    if (n == 0) return ORB_OK;
    cudaStream_t st = h->stream;
    ORB_CHECK(h->in0.reserve((size_t)n * 32));
    ORB_CHECK(h->out0.reserve((size_t)n * 4));
    ORB_CHECK(h->out1.reserve((size_t)n * 8));
    ORB_CHECK(h->out2.reserve((size_t)n * 4));
    ORB_CUDA(cudaMemcpyAsync(h->in0.p, desc, (size_t)n * 32, cudaMemcpyHostToDevice, st));
    ORB_CHECK(orbm_bow_transform_device(h, v, h->in0.as<uint8_t>(), n, levelsup, h->out0.as<int>(), h->out1.as<double>(),
                                        h->out2.as<int>(), st));
    std::vector<int> w(n), nd(n);
    std::vector<double> wt(n);
    ORB_CUDA(cudaMemcpyAsync(w.data(), h->out0.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaMemcpyAsync(wt.data(), h->out1.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaMemcpyAsync(nd.data(), h->out2.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    if (word) std::memcpy(word, w.data(), (size_t)n * 4);
    if (weight) std::memcpy(weight, wt.data(), (size_t)n * 8);
    if (node) std::memcpy(node, nd.data(), (size_t)n * 4);
    // The containers are host objects in the reference (std::map); what it does to them, in ascending feature order
    // (this fork races four threads here, TemplatedVocabulary.h:1198-1215: only the order inside a node differs).
    if (wantBow) {
        std::map<unsigned, double> bv;
        for (int i = 0; i < n; ++i)
            if (wt[i] > 0) bv[(unsigned)w[i]] += wt[i];              // addWeight: insert(id, w) or second += w
        double norm = 0.0;
        for (auto& kv : bv) norm += std::fabs(kv.second);            // normalize(L1)
        int k = 0;
        for (auto& kv : bv) { bowWord[k] = (int)kv.first; bowValue[k] = norm > 0.0 ? kv.second / norm : kv.second; ++k; }
        *nWords = k;
    }
    if (wantFv) {
        std::map<unsigned, std::vector<int>> fv;
        for (int i = 0; i < n; ++i)
            if (wt[i] > 0) fv[(unsigned)nd[i]].push_back(i);          // addFeature
        int k = 0, t = 0;
        for (auto& kv : fv) {
            fvNode[k] = (int)kv.first;
            for (int i : kv.second) fvIdx[t++] = i;
            fvStart[++k] = t;
        }
        *nNodes = k;
    }
    return ORB_OK;
}

}  // extern "C"
