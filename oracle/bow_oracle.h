// TEST INFRASTRUCTURE ONLY -- CPU oracle of the vocabulary-tree descent behind Frame::ComputeBoW / KeyFrame::ComputeBoW
// (Frame.cc:736-745: mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4)).
//
// Restates (all in /root/reference/Thirdparty/DBoW2/DBoW2, vendored in the reference):
//   TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup)   TemplatedVocabulary.h:1443-1485
//   TemplatedVocabulary::transform(features, v, fv, levelsup)                 TemplatedVocabulary.h:1166-1262 (TF_IDF, L1)
//   FORB::distance                                                             FORB.cpp:84-105
//   BowVector::addWeight / normalize(L1)                                       BowVector.cpp:31-43, 59-81
//   FeatureVector::addFeature                                                  FeatureVector.cpp:30-44
// Pinned: tests/test_bow_ref.py runs this against those functions' own text compiled in place (oracle/_ref/libbow_ref.so).
// One definition is ours: this fork fills v and fv from four racing threads (TemplatedVocabulary.h:1198-1215), so the
// order of the feature indices inside one FeatureVector node depends on thread timing there; the canonical order here is
// ascending feature index, i.e. what its own single-threaded variant (transformForMultiThread_1, :1313-1346) produces.
// The vocabulary FILE (Vocabulary/ORBvoc.bin) is absent from the checkout; trees are given as flat arrays.
#pragma once
#include <cstdint>
#include <vector>

namespace orbo {

struct VocabArrays {            // m_nodes of a TemplatedVocabulary<FORB::TDescriptor, FORB>, node 0 = root
    int nNodes = 0;
    int L = 0;                  // m_L: depth levels
    const uint8_t* desc = nullptr;     // nNodes x 32  (Node::descriptor; the root's is unused)
    const int* childStart = nullptr;   // nNodes + 1   (Node::children as CSR, in vector order)
    const int* children = nullptr;
    const int* wordId = nullptr;       // Node::word_id (meaningful for leaves)
    const double* weight = nullptr;    // Node::weight  (idf of a word)
};

// one feature down the tree: closest child at every level (strict '<': the first child wins a tie), until a leaf
void bow_transform_feature(const VocabArrays& V, const uint8_t* feature, int levelsup, int* wordId, double* weight,
                           int* nodeId);

// the whole transform with TF_IDF weighting and L1 scoring (what ORBvoc uses): BowVector as (ascending word id, value)
// and FeatureVector as node-sorted CSR with ascending feature indices.  perFeature* (may be null) receive the descent of
// every feature.
void bow_transform(const VocabArrays& V, const uint8_t* desc, int n, int levelsup, std::vector<int>& bowWord,
                   std::vector<double>& bowValue, std::vector<int>& fvNode, std::vector<int>& fvStart,
                   std::vector<int>& fvIdx, int* perFeatureWord, double* perFeatureWeight, int* perFeatureNode);

}  // namespace orbo
