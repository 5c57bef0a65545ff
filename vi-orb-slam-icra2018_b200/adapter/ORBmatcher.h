// Host-side mirror of ORB_SLAM2::ORBmatcher (reference: include/ORBmatcher.h:41-83) over the orbb200 C ABI.
//
// The reference's search methods take Frame / KeyFrame / MapPoint object graphs.  The data those loops actually read
// are a handful of arrays (undistorted keypoints, descriptors, image bounds, scale tables, per-point flags), so the
// adapter's native surface is `FrameView` + the query structs of orbb200.h; with -DORBB200_WITH_ORBSLAM (ORB-SLAM2
// headers on the include path) the reference's exact signatures are provided as thin overloads that flatten the
// objects, run the search on the GPU and write the results back into the objects the way the reference does.
// Same constants (TH_LOW, TH_HIGH, HISTO_LENGTH), same constructor (nnratio, checkOri), same return values.
#ifndef ORBB200_ADAPTER_ORBMATCHER_H
#define ORBB200_ADAPTER_ORBMATCHER_H

#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/orbb200.h"

namespace ORB_SLAM2 {

class ORBextractor;

// What a search reads of a Frame or KeyFrame. Owns the device copy + 64x48 grid (Frame.cc:574-589).
class FrameView {
public:
    FrameView(orbm_handle m, const orb_keypoint* keysUn, const unsigned char* descriptors, int n, float minX, float minY,
              float maxX, float maxY)
        : n_(n) {
        if (orbm_frame_create(m, keysUn, descriptors, n, minX, minY, maxX, maxY, &f_) != ORB_OK)
            throw std::runtime_error(std::string("orbb200: ") + orb_last_error());
    }
    // adopts a frame that already lives on the device (orbm_frame_create_device)
    FrameView(orbm_frame f, int n) : f_(f), n_(n) {}
    ~FrameView() { orbm_frame_destroy(f_); }
    FrameView(const FrameView&) = delete;
    FrameView& operator=(const FrameView&) = delete;
    FrameView(FrameView&& o) : f_(o.f_), n_(o.n_) { o.f_ = nullptr; }
    orbm_frame get() const { return f_; }
    int size() const { return n_; }

private:
    orbm_frame f_ = nullptr;
    int n_ = 0;
};

namespace orbb_detail {

// The reference constructs an ORBmatcher on the stack at every call site (Tracking.cc:1644, 2084, ...; LocalMapping.cc,
// LoopClosing.cc), and a Frame / KeyFrame is searched many times over its life.  Two process-wide pieces keep that cheap:
//   * a pool of matcher handles per device (stream + workspaces are created once, an ORBmatcher borrows one);
//   * per handle, an LRU cache of device-resident frames keyed by (Frame | KeyFrame, mnId, N): the upload of mvKeysUn /
//     mDescriptors and the grid build happen on the first search that sees the object, not on every call.  A Frame's
//     keypoints never change after construction and copies of a Frame keep its mnId (Frame.cc:38-72).
struct ViewRef {
    std::shared_ptr<FrameView> p;
    operator const FrameView&() const { return *p; }
};

class FrameCache {
public:
    struct Key {
        int kind;
        unsigned long id;
        int n;
        bool operator<(const Key& o) const { return kind != o.kind ? kind < o.kind : id != o.id ? id < o.id : n < o.n; }
    };
    std::shared_ptr<FrameView> find(const Key& k) {
        std::map<Key, std::list<Entry>::iterator>::iterator it = index_.find(k);
        if (it == index_.end()) { ++misses; return std::shared_ptr<FrameView>(); }
        lru_.splice(lru_.begin(), lru_, it->second);
        ++hits;
        return it->second->view;
    }
    void insert(const Key& k, const std::shared_ptr<FrameView>& v) {
        std::map<Key, std::list<Entry>::iterator>::iterator it = index_.find(k);
        if (it != index_.end()) { lru_.erase(it->second); index_.erase(it); }
        lru_.push_front(Entry{k, v});
        index_[k] = lru_.begin();
        while (lru_.size() > capacity) { index_.erase(lru_.back().key); lru_.pop_back(); }
    }
    void clear() { lru_.clear(); index_.clear(); }
    size_t capacity = 256;
    long hits = 0, misses = 0;

private:
    struct Entry { Key key; std::shared_ptr<FrameView> view; };
    std::list<Entry> lru_;
    std::map<Key, std::list<Entry>::iterator> index_;
};

struct MatcherSlot {
    orbm_handle h = nullptr;
    int device = 0;
    FrameCache cache;
};

class HandlePool {
public:
    static HandlePool& get() { static HandlePool p; return p; }
    MatcherSlot* acquire(int device) {
        {
            std::lock_guard<std::mutex> g(m_);
            std::vector<MatcherSlot*>& f = free_[device];
            if (!f.empty()) { MatcherSlot* s = f.back(); f.pop_back(); return s; }   // LIFO: a thread gets its last handle back
        }
        MatcherSlot* s = new MatcherSlot;
        s->device = device;
        if (orbm_create(device, &s->h) != ORB_OK) { delete s; throw std::runtime_error(std::string("orbb200: ") + orb_last_error()); }
        std::lock_guard<std::mutex> g(m_);
        all_.push_back(s);
        return s;
    }
    void release(MatcherSlot* s) {
        std::lock_guard<std::mutex> g(m_);
        free_[s->device].push_back(s);
    }
    // cache statistics over all handles (tests, tuning)
    void stats(long* hits, long* misses) {
        std::lock_guard<std::mutex> g(m_);
        *hits = *misses = 0;
        for (size_t i = 0; i < all_.size(); ++i) { *hits += all_[i]->cache.hits; *misses += all_[i]->cache.misses; }
    }
    // drops every cached frame and idle handle (call before the CUDA context goes away, or to bound memory)
    void clear() {
        std::lock_guard<std::mutex> g(m_);
        for (size_t i = 0; i < all_.size(); ++i) all_[i]->cache.clear();
    }

private:
    std::mutex m_;
    std::map<int, std::vector<MatcherSlot*> > free_;
    std::vector<MatcherSlot*> all_;
};

}  // namespace orbb_detail

class ORBmatcher {
public:
    static const int TH_LOW = 50;
    static const int TH_HIGH = 100;
    static const int HISTO_LENGTH = 30;

    ORBmatcher(float nnratio = 0.6, bool checkOri = true, int device = 0) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {
        slot_ = orbb_detail::HandlePool::get().acquire(device);
        h_ = slot_->h;
    }
    ~ORBmatcher() { orbb_detail::HandlePool::get().release(slot_); }
    ORBmatcher(const ORBmatcher&) = delete;
    ORBmatcher& operator=(const ORBmatcher&) = delete;
    orbm_handle handle() { return h_; }

    // Computes the Hamming distance between two ORB descriptors (ORBmatcher.cc:1675-1691). One pair per call is a poor
    // use of a GPU; DescriptorDistances() takes n pairs.  Kept for drop-in completeness.
    int DescriptorDistance(const unsigned char* a, const unsigned char* b) {
        int d = 0;
        check(orbm_distance(h_, a, b, 1, &d));
        return d;
    }
    void DescriptorDistances(const unsigned char* a, const unsigned char* b, int n, int* out) { check(orbm_distance(h_, a, b, n, out)); }

    // SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)   ORBmatcher.cc:405-520
    int SearchForInitialization(const FrameView& F1, const FrameView& F2, std::vector<float>& vbPrevMatchedXY,
                                std::vector<int>& vnMatches12, int windowSize = 10) {
        vnMatches12.assign(F1.size(), -1);
        vbPrevMatchedXY.resize((size_t)F1.size() * 2);
        int n = 0;
        check(orbm_search_for_initialization(h_, F1.get(), F2.get(), vbPrevMatchedXY.data(), vnMatches12.data(), windowSize,
                                             mfNNratio, mbCheckOrientation, &n));
        return n;
    }

    // SearchByProjection(CurrentFrame, LastFrame, th, bMono)   ORBmatcher.cc:1341-1498
    // queries: one per LastFrame keypoint (projection done by the caller); curMatch[i2] = query index or -1
    int SearchByProjection(const FrameView& Current, const std::vector<float>& scaleFactors, const float* uRight, float mbf,
                           const std::vector<orbm_proj_query>& queries, const unsigned char* queryDescriptors, float th, int mode,
                           const unsigned char* occupied, std::vector<int>& curMatch) {
        curMatch.assign(Current.size(), -1);
        int n = 0;
        check(orbm_search_by_projection(h_, Current.get(), scaleFactors.data(), (int)scaleFactors.size(), uRight, mbf,
                                        queries.data(), queryDescriptors, (int)queries.size(), th, mode, occupied,
                                        curMatch.data(), mbCheckOrientation, &n));
        return n;
    }

    // the same with world points + current pose: the projection of ORBmatcher.cc:1376-1393 runs on the device
    int SearchByProjection(const FrameView& Current, const std::vector<float>& scaleFactors, const float* uRight, float mbf,
                           const orbm_pose& pose, const std::vector<orbm_world_query>& queries, const unsigned char* queryDescriptors,
                           float th, int mode, const unsigned char* occupied, std::vector<int>& curMatch) {
        curMatch.assign(Current.size(), -1);
        int n = 0;
        check(orbm_search_by_projection_world(h_, Current.get(), scaleFactors.data(), (int)scaleFactors.size(), uRight, mbf, &pose,
                                              queries.data(), queryDescriptors, (int)queries.size(), th, mode, TH_HIGH, occupied,
                                              curMatch.data(), mbCheckOrientation, &n));
        return n;
    }

    // SearchByProjection(F, vpMapPoints, th)   ORBmatcher.cc:45-129
    int SearchByProjection(const FrameView& F, const std::vector<float>& scaleFactors, const float* uRight,
                           const std::vector<orbm_point_query>& queries, const unsigned char* queryDescriptors, float th,
                           const unsigned char* occupied, std::vector<int>& match) {
        match.assign(F.size(), -1);
        int n = 0;
        check(orbm_search_by_projection_points(h_, F.get(), scaleFactors.data(), (int)scaleFactors.size(), uRight, queries.data(),
                                               queryDescriptors, (int)queries.size(), th, mfNNratio, occupied, match.data(), &n));
        return n;
    }

    // SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo)   ORBmatcher.cc:657-823
    struct FeatureVectorCSR { std::vector<int> nodeId, start, idx; };   // DBoW2::FeatureVector flattened, ids ascending
    int SearchForTriangulation(const FrameView& KF1, const FrameView& KF2, const FeatureVectorCSR& fv1, const FeatureVectorCSR& fv2,
                               const unsigned char* hasMapPoint1, const unsigned char* hasMapPoint2, const float* uRight1,
                               const float* uRight2, const float F12[9], float ex, float ey, const std::vector<float>& scaleFactors2,
                               const std::vector<float>& levelSigma2_2, std::vector<std::pair<size_t, size_t> >& vMatchedPairs,
                               bool bOnlyStereo) {
        std::vector<int> m12(KF1.size(), -1);
        int n = 0;
        check(orbm_search_for_triangulation(h_, KF1.get(), KF2.get(), (int)fv1.nodeId.size(), fv1.nodeId.data(), fv1.start.data(),
                                            fv1.idx.data(), (int)fv2.nodeId.size(), fv2.nodeId.data(), fv2.start.data(), fv2.idx.data(),
                                            hasMapPoint1, hasMapPoint2, uRight1, uRight2, F12, ex, ey, scaleFactors2.data(),
                                            levelSigma2_2.data(), (int)scaleFactors2.size(), bOnlyStereo, mbCheckOrientation,
                                            m12.data(), &n));
        vMatchedPairs.clear();
        vMatchedPairs.reserve(n);
        for (size_t i = 0; i < m12.size(); ++i)
            if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, (size_t)m12[i]));   // ORBmatcher.cc:812-820
        return n;
    }

    // SearchByBoW(pKF, F, vpMapPointMatches) ORBmatcher.cc:159-288 (strictLow = false) and
    // SearchByBoW(pKF1, pKF2, vpMatches12)   ORBmatcher.cc:522-655 (strictLow = true).
    // matches12[i1] = i2, matches21[i2] = i1 (or -1); the Frame overload's vpMapPointMatches[i2] is pKF's point at matches21[i2].
    int SearchByBoW(const FrameView& KF1, const FrameView& F2, const FeatureVectorCSR& fv1, const FeatureVectorCSR& fv2,
                    const unsigned char* valid1, const unsigned char* valid2, bool strictLow, std::vector<int>& matches12,
                    std::vector<int>& matches21) {
        matches12.assign(KF1.size(), -1);
        matches21.assign(F2.size(), -1);
        int n = 0;
        check(orbm_search_by_bow(h_, KF1.get(), F2.get(), (int)fv1.nodeId.size(), fv1.nodeId.data(), fv1.start.data(), fv1.idx.data(),
                                 (int)fv2.nodeId.size(), fv2.nodeId.data(), fv2.start.data(), fv2.idx.data(), valid1, valid2,
                                 mfNNratio, mbCheckOrientation, strictLow, matches12.data(), matches21.data(), &n));
        return n;
    }

    // SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) ORBmatcher.cc:1500-1627 (mode 0, maxDistance = ORBdist)
    // and SearchByProjection(pKF, Scw, vpPoints, vpMatched, th) ORBmatcher.cc:290-403 (mode 3, maxDistance = TH_LOW, no
    // orientation check: construct the matcher with checkOri = false): the caller predicts the level into query.octave.
    int SearchByProjection(const FrameView& Current, const std::vector<float>& scaleFactors, const std::vector<orbm_proj_query>& queries,
                           const unsigned char* queryDescriptors, float th, int mode, int maxDistance, const unsigned char* occupied,
                           std::vector<int>& curMatch) {
        curMatch.assign(Current.size(), -1);
        int n = 0;
        check(orbm_search_by_projection_ex(h_, Current.get(), scaleFactors.data(), (int)scaleFactors.size(), nullptr, 0.f,
                                           queries.data(), queryDescriptors, (int)queries.size(), th, mode, maxDistance, occupied,
                                           curMatch.data(), mbCheckOrientation, &n));
        return n;
    }

    // The search inside Fuse (ORBmatcher.cc:892-944, chi2 = true; :1051-1075) and SearchBySim3 (:1191-1215, :1271-1295):
    // closest descriptor per projected point; Replace / AddObservation / the mutual-agreement test stay with the caller.
    void ProjectedBest(const FrameView& KF, const std::vector<orbm_best_query>& queries, const unsigned char* queryDescriptors,
                       bool chi2, const float* uRight, const std::vector<float>& invLevelSigma2, std::vector<int>& bestIdx,
                       std::vector<int>& bestDist) {
        bestIdx.assign(queries.size(), -1);
        bestDist.assign(queries.size(), 256);
        check(orbm_search_projected_best(h_, KF.get(), queries.data(), queryDescriptors, (int)queries.size(), chi2, uRight,
                                         invLevelSigma2.data(), (int)invLevelSigma2.size(), bestIdx.data(), bestDist.data()));
    }

#ifdef ORBB200_WITH_ORBSLAM
    // ---- the reference's own signatures (need Frame.h / KeyFrame.h / MapPoint.h of the host project) -------------------
    // Declared here, defined in adapter/ORBmatcher_orbslam.inl, which flattens the objects exactly as documented in
    // INTEGRATION.md section 3.  Not compiled in this repository: the host project's headers need Eigen, DBoW2 and g2o.
    int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12,
                                int windowSize = 10);
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);
    int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3);
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t> >& vMatchedPairs,
                               const bool bOnlyStereo);
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
    int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12);
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th,
                           const int ORBdist);
    int SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched,
                           int th);
    int Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th = 3.0);
    int Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th, std::vector<MapPoint*>& vpReplacePoint);
    int SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12,
                     const cv::Mat& t12, const float th);
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
    // Hand-over from the extractor without a second upload: call right after ExtractORB in Frame's constructor, once
    // mnId, N, mK, mDistCoef and the image bounds are set (Frame.cc:75-109).  The Frame's device copy (undistorted
    // keypoints, descriptors, grid) is built from what `extractor`'s last call left on the device and cached under the
    // Frame's mnId; the searches then find it instead of uploading mvKeysUn / mDescriptors.  frameIndex: 0 = left image.
    static void RegisterFrame(const Frame& F, ORBextractor& extractor, int frameIndex = 0, int device = 0);
#endif

protected:
    float mfNNratio;
    bool mbCheckOrientation;

private:
    orbm_handle h_ = nullptr;
    orbb_detail::MatcherSlot* slot_ = nullptr;
#ifdef ORBB200_WITH_ORBSLAM
    template <class F> orbb_detail::ViewRef view_of(const F& f);
#endif
    static void check(int st) {
        if (st != ORB_OK) throw std::runtime_error(std::string("orbb200: ") + orb_last_error());
    }
};

}  // namespace ORB_SLAM2

#endif
