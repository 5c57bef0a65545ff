// Host-side mirror of ORB_SLAM2::ORBVocabulary (reference: include/ORBVocabulary.h:30-31, a typedef of
// DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>) for the one thing the front-end needs of it:
// transform(features, BowVector, FeatureVector, levelsup) as Frame::ComputeBoW calls it (Frame.cc:736-745), plus the
// file formats System.cc:334-339 loads it from.
//
//   loadFromTextFile   / saveToTextFile     Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1563-1651, :1654-1676
//   loadFromBinaryFile / saveToBinaryFile   Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1679-1724, :1726-1745
//
// The tree is kept as the flat arrays orbm_vocabulary_create takes (node 0 = root; children as CSR in the reference's
// push_back order, i.e. ascending node id per parent), not as a vector of Node objects.  The readers reproduce what the
// reference's readers build, including two artefacts of their `while(!f.eof())` loops, because those nodes take part in
// the descent: a text file that ends with a newline (every file saveToTextFile writes) yields one extra node made from
// the empty last line (a sibling of the last real node, with its leaf flag, weight 0, no descriptor bytes); the binary
// reader stores the last record twice.  Where the
// reference would touch memory out of bounds or read uninitialised bytes the result is defined instead: a parent id that
// does not name an earlier node, or more records than the header announces, make the reader return false; the descriptor
// of the empty-line node is all zero.
#ifndef ORBB200_ADAPTER_ORBVOCABULARY_H
#define ORBB200_ADAPTER_ORBVOCABULARY_H

#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/orbb200.h"

namespace ORB_SLAM2 {

class ORBVocabulary {
public:
    enum { DESC_BYTES = 32 };                       // FORB::L
    int m_k = 0, m_L = 0, m_scoring = 0, m_weighting = 0;   // header fields (DBoW2 ScoringType / WeightingType as ints)
    // flat m_nodes
    std::vector<unsigned char> desc;                // nodes x 32
    std::vector<int> parent, wordId, childStart, children;
    std::vector<double> weight;
    std::vector<char> isWord;                       // the node is in m_words (was flagged as a leaf in the file)

    ORBVocabulary() {}
    ~ORBVocabulary() { release(); }
    ORBVocabulary(const ORBVocabulary&) = delete;
    ORBVocabulary& operator=(const ORBVocabulary&) = delete;

    bool empty() const { return nWords_ == 0; }     // TemplatedVocabulary::empty(): m_words.empty()
    unsigned int size() const { return (unsigned)nWords_; }
    int nodes() const { return (int)parent.size(); }

    bool loadFromTextFile(const std::string& filename) {
        clear();
        std::ifstream f(filename.c_str());
        std::string line;
        std::getline(f, line);
        int scoring = 0, weighting = 0;
        {
            std::stringstream ss;
            ss << line;
            ss >> m_k; ss >> m_L; ss >> scoring; ss >> weighting;
        }
        if (m_k < 0 || m_k > 20 || m_L < 1 || m_L > 10 || scoring < 0 || scoring > 5 || weighting < 0 || weighting > 3) {
            std::cerr << "Vocabulary loading failure: This is not a correct text file!" << std::endl;
            return false;
        }
        m_scoring = scoring;
        m_weighting = weighting;
        addNode(0, nullptr, 0.0, false);            // the root
        // One node per line; the empty line after a final '\n' counts too.  On that empty line the extractions fail before
        // any digit is looked at, which leaves their targets untouched: the reference's (uninitialised) locals then still
        // hold the previous line's parent id and leaf flag in the builds we could run (GCC), so that is the definition here.
        int pid = 0, leaf = 0;
        while (!f.eof()) {
            std::getline(f, line);
            std::stringstream ls;
            ls << line;
            ls >> pid;
            ls >> leaf;
            std::stringstream bytes;                // the 32 byte tokens, re-parsed as FORB::fromString does
            for (int i = 0; i < DESC_BYTES; ++i) {
                std::string tok;
                ls >> tok;
                bytes << tok << " ";
            }
            unsigned char d[DESC_BYTES];
            std::memset(d, 0, sizeof d);
            {
                std::stringstream ps(bytes.str());
                for (int i = 0; i < DESC_BYTES; ++i) {
                    int n;
                    ps >> n;
                    if (!ps.fail()) d[i] = (unsigned char)n;
                }
            }
            double w = 0;
            ls >> w;
            if (pid < 0 || pid >= nodes()) return fail();
            addNode(pid, d, w, leaf > 0);
        }
        finish();
        return true;
    }

    bool loadFromBinaryFile(const std::string& filename) {
        clear();
        std::ifstream f(filename.c_str(), std::ios_base::in | std::ios::binary);
        unsigned int nbNodes = 0, sizeNode = 0;
        int hdr[4] = {0, 0, 0, 0};
        f.read((char*)&nbNodes, 4);
        f.read((char*)&sizeNode, 4);
        f.read((char*)hdr, 16);
        if (!f || sizeNode < 4 + DESC_BYTES + 4 + 1 || sizeNode > 4096) return fail();
        m_k = hdr[0]; m_L = hdr[1]; m_scoring = hdr[2]; m_weighting = hdr[3];
        addNode(0, nullptr, 0.0, false);
        std::vector<char> buf(sizeNode, 0);
        while (!f.eof()) {                          // the read that hits the end of the file leaves buf as it was:
            f.read(buf.data(), sizeNode);           // the last record is stored a second time
            if ((unsigned)nodes() > nbNodes) return fail();
            int pid;
            float w;
            std::memcpy(&pid, buf.data(), 4);
            std::memcpy(&w, buf.data() + 4 + DESC_BYTES, 4);
            if (pid < 0 || pid >= nodes()) return fail();
            addNode(pid, (const unsigned char*)buf.data() + 4, (double)w, buf[8 + DESC_BYTES] != 0);
        }
        // the reference sizes m_nodes to nb_nodes + 1 up front: nodes no record reached stay default-constructed
        while ((unsigned)nodes() < nbNodes + 1) addNode(0, nullptr, 0.0, false, /*linkToParent=*/false);
        finish();
        return true;
    }

    void saveToTextFile(const std::string& filename) const {
        std::fstream f(filename.c_str(), std::ios_base::out);
        f << m_k << " " << m_L << " " << " " << m_scoring << " " << m_weighting << std::endl;
        for (int i = 1; i < nodes(); ++i) {
            f << parent[i] << " " << (isLeaf(i) ? 1 : 0) << " ";
            for (int b = 0; b < DESC_BYTES; ++b) f << (int)desc[(size_t)i * DESC_BYTES + b] << " ";
            f << " " << weight[i] << std::endl;
        }
    }

    void saveToBinaryFile(const std::string& filename) const {
        std::fstream f(filename.c_str(), std::ios_base::out | std::ios::binary);
        const unsigned int nbNodes = (unsigned)nodes(), sizeNode = 4 + DESC_BYTES + 4 + 1;
        f.write((const char*)&nbNodes, 4);
        f.write((const char*)&sizeNode, 4);
        const int hdr[4] = {m_k, m_L, m_scoring, m_weighting};
        f.write((const char*)hdr, 16);
        for (int i = 1; i < nodes(); ++i) {
            const float w = (float)weight[i];
            const char leaf = isLeaf(i) ? 1 : 0;
            f.write((const char*)&parent[i], 4);
            f.write((const char*)&desc[(size_t)i * DESC_BYTES], DESC_BYTES);
            f.write((const char*)&w, 4);
            f.write(&leaf, 1);
        }
    }

    bool isLeaf(int node) const { return childStart[node + 1] == childStart[node]; }   // Node::isLeaf(): no children

    // ---- device side -------------------------------------------------------------------------------------------
    // copies the tree to the matcher's device (once); transform() then runs the descent on the GPU
    void upload(orbm_handle matcher) {
        release();
        check(orbm_vocabulary_create(matcher, nodes(), m_L, desc.data(), childStart.data(), children.data(), wordId.data(),
                                     weight.data(), &dev_));
        matcher_ = matcher;
    }
    orbm_vocabulary handle() const { return dev_; }

    // transform(features, v, fv, levelsup): BowVector as ascending (word, value) arrays, FeatureVector as node-sorted CSR
    void transform(const std::vector<unsigned char>& features, std::vector<int>& bowWord, std::vector<double>& bowValue,
                   std::vector<int>& fvNode, std::vector<int>& fvStart, std::vector<int>& fvIdx, int levelsup) const {
        const int n = (int)(features.size() / DESC_BYTES);
        bowWord.assign(n + 1, 0); bowValue.assign(n + 1, 0.0);
        fvNode.assign(n + 1, 0); fvStart.assign(n + 2, 0); fvIdx.assign(n + 1, 0);
        int nWords = 0, nNodes = 0;
        if (!dev_) throw std::runtime_error("orbb200: ORBVocabulary::transform before upload()");
        check(orbm_bow_transform(matcher_, dev_, features.data(), n, levelsup, nullptr, nullptr, nullptr, bowWord.data(),
                                 bowValue.data(), &nWords, fvNode.data(), fvStart.data(), fvIdx.data(), &nNodes));
        bowWord.resize(nWords); bowValue.resize(nWords);
        fvNode.resize(nNodes); fvStart.resize(nNodes + 1);
        fvIdx.resize(fvStart[nNodes]);
    }

    // The reference's call shape, Frame.cc:743-744: transform(vCurrentDesc, mBowVec, mFeatVec, 4) with DBoW2::BowVector
    // (std::map<WordId, WordValue>) and DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned int>>).  Templated on
    // the two container types so that this header does not depend on DBoW2's; any map-like pair works.
    template <class BowVectorT, class FeatureVectorT>
    void transform(const std::vector<unsigned char>& features, BowVectorT& v, FeatureVectorT& fv, int levelsup) const {
        std::vector<int> bw, fn, fs, fi;
        std::vector<double> bv;
        transform(features, bw, bv, fn, fs, fi, levelsup);
        v.clear();
        fv.clear();
        for (size_t i = 0; i < bw.size(); ++i) v[bw[i]] = bv[i];
        for (size_t k = 0; k < fn.size(); ++k) {
            typename FeatureVectorT::mapped_type& idx = fv[fn[k]];
            idx.assign(fi.begin() + fs[k], fi.begin() + fs[k + 1]);
        }
    }

private:
    std::vector<std::vector<int> > kids_;           // while loading; flattened by finish()
    int nWords_ = 0;
    orbm_vocabulary dev_ = nullptr;
    orbm_handle matcher_ = nullptr;

    void clear() {
        m_k = m_L = m_scoring = m_weighting = 0;
        desc.clear(); parent.clear(); wordId.clear(); childStart.clear(); children.clear(); weight.clear(); isWord.clear();
        kids_.clear();
        nWords_ = 0;
    }
    bool fail() { clear(); return false; }
    void addNode(int pid, const unsigned char* d, double w, bool word, bool linkToParent = true) {
        const int id = nodes();
        parent.push_back(pid);
        desc.insert(desc.end(), DESC_BYTES, 0);
        if (d) std::memcpy(&desc[(size_t)id * DESC_BYTES], d, DESC_BYTES);
        weight.push_back(w);
        wordId.push_back(word ? nWords_ : 0);       // Node::word_id defaults to 0
        isWord.push_back(word ? 1 : 0);
        if (word) ++nWords_;
        kids_.push_back(std::vector<int>());
        if (id > 0 && linkToParent) kids_[pid].push_back(id);
    }
    void finish() {
        childStart.assign(nodes() + 1, 0);
        children.clear();
        for (int i = 0; i < nodes(); ++i) {
            childStart[i] = (int)children.size();
            children.insert(children.end(), kids_[i].begin(), kids_[i].end());
        }
        childStart[nodes()] = (int)children.size();
        kids_.clear();
    }
    void release() {
        if (dev_) orbm_vocabulary_destroy(dev_);
        dev_ = nullptr;
    }
    static void check(int st) {
        if (st != ORB_OK) throw std::runtime_error(std::string("orbb200: ") + orb_last_error());
    }
};

}  // namespace ORB_SLAM2

#endif
