// TEST INFRASTRUCTURE ONLY -- see bow_oracle.h
#include "bow_oracle.h"

#include <cmath>
#include <map>

#include "match_oracle.h"

namespace orbo {

void bow_transform_feature(const VocabArrays& V, const uint8_t* feature, int levelsup, int* wordId, double* weight,
                           int* nodeId) {
    const int nidLevel = V.L - levelsup;                       // TemplatedVocabulary.h:1452
    if (nidLevel <= 0 && nodeId) *nodeId = 0;                  // root
    int finalId = 0, currentLevel = 0;
    do {
        ++currentLevel;
        const int* kids = V.children + V.childStart[finalId];
        const int nKids = V.childStart[finalId + 1] - V.childStart[finalId];
        finalId = kids[0];
        double bestD = descriptor_distance(feature, V.desc + 32 * (size_t)finalId);   // FORB::distance
        for (int c = 1; c < nKids; ++c) {
            const int id = kids[c];
            const double d = descriptor_distance(feature, V.desc + 32 * (size_t)id);
            if (d < bestD) { bestD = d; finalId = id; }
        }
        if (nodeId && currentLevel == nidLevel) *nodeId = finalId;
    } while (V.childStart[finalId + 1] != V.childStart[finalId]);   // !isLeaf()
    *wordId = V.wordId[finalId];
    *weight = V.weight[finalId];
}

void bow_transform(const VocabArrays& V, const uint8_t* desc, int n, int levelsup, std::vector<int>& bowWord,
                   std::vector<double>& bowValue, std::vector<int>& fvNode, std::vector<int>& fvStart,
                   std::vector<int>& fvIdx, int* perFeatureWord, double* perFeatureWeight, int* perFeatureNode) {
    std::map<unsigned, double> v;                       // BowVector
    std::map<unsigned, std::vector<unsigned>> fv;       // FeatureVector
    for (int i = 0; i < n; ++i) {
        int id = 0, nid = 0;
        double w = 0;
        bow_transform_feature(V, desc + 32 * (size_t)i, levelsup, &id, &w, &nid);
        if (perFeatureWord) perFeatureWord[i] = id;
        if (perFeatureWeight) perFeatureWeight[i] = w;
        if (perFeatureNode) perFeatureNode[i] = nid;
        if (w > 0) {                                    // not stopped, :1300-1305
            v[(unsigned)id] += w;                       // addWeight: insert(id, w) or second += w (0.0 + w == w)
            fv[(unsigned)nid].push_back((unsigned)i);   // addFeature
        }
    }
    double norm = 0.0;                                  // normalize(L1), BowVector.cpp:59-81
    for (auto& kv : v) norm += std::fabs(kv.second);
    if (norm > 0.0)
        for (auto& kv : v) kv.second /= norm;
    bowWord.clear(); bowValue.clear(); fvNode.clear(); fvStart.assign(1, 0); fvIdx.clear();
    for (auto& kv : v) { bowWord.push_back((int)kv.first); bowValue.push_back(kv.second); }
    for (auto& kv : fv) {
        fvNode.push_back((int)kv.first);
        for (unsigned i : kv.second) fvIdx.push_back((int)i);
        fvStart.push_back((int)fvIdx.size());
    }
}

}  // namespace orbo
