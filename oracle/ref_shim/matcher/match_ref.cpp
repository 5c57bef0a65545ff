// TEST INFRASTRUCTURE ONLY.
// oracle/_ref/libmatch_ref.so: the reference's own ORBmatcher.cc, compiled where it lies (REF_ORBMATCHER_CC, never
// copied) against the stand-ins of slam_shim.hpp, behind the same flat-array C signatures as the oracle's orbo_search_*
// entry points -- so a test can hand identical inputs to the oracle restatement and to the reference's code and demand
// identical outputs.  The searches whose inputs the C ABI takes as already-projected coordinates are driven with an
// identity camera (R = I, t = 0, fx = fy = 1, cx = cy = 0, depth 1), for which the reference's own projection arithmetic
// returns the given coordinates bit for bit.
#include <chrono>

#include "slam_shim.hpp"

using namespace std;   // the reference's headers lean on a using-directive leaked by the headers replaced here

#include REF_ORBMATCHER_CC

using namespace ORB_SLAM2;

namespace {

// wall time of the reference's search call alone (the stand-in objects are built outside of it)
double g_last_search_ms = 0;
struct SearchTimer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    ~SearchTimer() { g_last_search_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

struct RefFrame {
    std::vector<cv::KeyPoint> keys;
    std::vector<uint8_t> desc;
    float minX, minY, maxX, maxY;
};

void fill(FrameData& f, const RefFrame* r) {
    f.set(r->keys.data(), r->desc.data(), (int)r->keys.size(), r->minX, r->minY, r->maxX, r->maxY);
}

cv::Mat row32(const uint8_t* d) {
    cv::Mat m(1, 32, CV_8U);
    std::memcpy(m.ptr<uchar>(0), d, 32);
    return m;
}

cv::Mat vec3(float x, float y, float z) {
    cv::Mat m(3, 1, CV_32F);
    m.at<float>(0) = x; m.at<float>(1) = y; m.at<float>(2) = z;
    return m;
}

void set_featvec(DBoW2::FeatureVector& fv, int nNodes, const int* nodeId, const int* start, const int* idx) {
    for (int k = 0; k < nNodes; ++k)
        fv[(DBoW2::NodeId)nodeId[k]] = std::vector<unsigned int>(idx + start[k], idx + start[k + 1]);
}

// index of p inside pool[0..n) or -1
int index_of(const MapPoint* p, const std::vector<MapPoint>& pool) {
    if (!p || pool.empty() || p < &pool[0] || p > &pool[pool.size() - 1]) return -1;
    return (int)(p - &pool[0]);
}

}  // namespace

extern "C" {

int refm_distance(const uint8_t* a, const uint8_t* b) { return ORBmatcher::DescriptorDistance(row32(a), row32(b)); }

void* refm_frame_create(const cv::KeyPoint* keysUn, const uint8_t* desc, int n, float minX, float minY, float maxX, float maxY) {
    RefFrame* f = new RefFrame;
    f->keys.assign(keysUn, keysUn + n);
    f->desc.assign(desc, desc + (size_t)n * 32);
    f->minX = minX; f->minY = minY; f->maxX = maxX; f->maxY = maxY;
    return f;
}
void refm_frame_destroy(void* f) { delete (RefFrame*)f; }
double refm_last_search_ms() { return g_last_search_ms; }

// ORBmatcher::SearchForInitialization  (ORBmatcher.cc:405)
int refm_search_init(void* f1, void* f2, float* prevXY, int* m12, int window, float ratio, int checkOri) {
    Frame F1, F2;
    fill(F1, (RefFrame*)f1);
    fill(F2, (RefFrame*)f2);
    std::vector<cv::Point2f> prev(F1.N);
    for (int i = 0; i < F1.N; ++i) prev[i] = cv::Point2f(prevXY[2 * i], prevXY[2 * i + 1]);
    std::vector<int> v12;
    ORBmatcher m(ratio, checkOri != 0);
    int n;
    { SearchTimer timer; n = m.SearchForInitialization(F1, F2, prev, v12, window); }
    for (int i = 0; i < F1.N; ++i) { m12[i] = v12[i]; prevXY[2 * i] = prev[i].x; prevXY[2 * i + 1] = prev[i].y; }
    return n;
}

// ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono)  (ORBmatcher.cc:1341); needs q[i].invz == 1
int refm_search_projection(void* cur, const float* sf, int nLevels, const float* uRight, float mbf, const orbo::ProjQuery* q,
                           const uint8_t* qdesc, int nq, float th, int mode, const uint8_t* occupied, int* curMatch,
                           int checkOri) {
    Frame C, L;
    fill(C, (RefFrame*)cur);
    C.mvScaleFactors.assign(sf, sf + nLevels);
    if (uRight) C.mvuRight.assign(uRight, uRight + C.N);
    C.mbf = mbf;
    C.mb = 1.f;
    MapPoint taken;
    taken.nObs = 1;
    C.mvpMapPoints.assign(C.N, (MapPoint*)NULL);
    for (int i = 0; i < C.N; ++i)
        if (occupied && occupied[i]) C.mvpMapPoints[i] = &taken;
    std::vector<MapPoint> mp(nq);
    L.N = nq;
    L.mvKeys.resize(nq);
    L.mvKeysUn.resize(nq);
    L.mvpMapPoints.assign(nq, (MapPoint*)NULL);
    L.mvbOutlier.assign(nq, false);
    for (int i = 0; i < nq; ++i) {
        if (q[i].valid && q[i].invz != 1.f) return -2;
        mp[i].worldPos = vec3(q[i].u, q[i].v, 1.f);
        mp[i].descriptor = row32(qdesc + (size_t)i * 32);
        mp[i].nObs = q[i].obsPositive ? 1 : 0;
        L.mvKeys[i].octave = q[i].octave;
        L.mvKeysUn[i].octave = q[i].octave;
        L.mvKeysUn[i].angle = q[i].angle;
        if (q[i].valid) L.mvpMapPoints[i] = &mp[i];
    }
    // tlc = Rlw * twc + tlw decides forward / backward against CurrentFrame.mb (:1361-1364)
    L.mTcw.at<float>(2, 3) = mode == 1 ? 2.f : mode == 2 ? -2.f : 0.f;
    ORBmatcher m(0.9f, checkOri != 0);
    int n;
    { SearchTimer timer; n = m.SearchByProjection(C, L, th, false); }
    for (int i = 0; i < C.N; ++i) curMatch[i] = index_of(C.mvpMapPoints[i], mp);
    return n;
}

// ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th)  (ORBmatcher.cc:45)
int refm_search_points(void* Fp, const float* sf, int nLevels, const float* uRight, const orbo::MapPointQuery* q,
                       const uint8_t* qdesc, int nq, float th, float ratio, const uint8_t* occupied, int* match) {
    Frame F;
    fill(F, (RefFrame*)Fp);
    F.mvScaleFactors.assign(sf, sf + nLevels);
    if (uRight) F.mvuRight.assign(uRight, uRight + F.N);
    MapPoint taken;
    taken.nObs = 1;
    F.mvpMapPoints.assign(F.N, (MapPoint*)NULL);
    for (int i = 0; i < F.N; ++i)
        if (occupied && occupied[i]) F.mvpMapPoints[i] = &taken;
    std::vector<MapPoint> mp(nq);
    std::vector<MapPoint*> vp(nq);
    for (int i = 0; i < nq; ++i) {
        mp[i].mbTrackInView = q[i].inView != 0;
        mp[i].mTrackProjX = q[i].projX; mp[i].mTrackProjY = q[i].projY; mp[i].mTrackProjXR = q[i].projXR;
        mp[i].mTrackViewCos = q[i].viewCos;
        mp[i].mnTrackScaleLevel = q[i].level;
        mp[i].nObs = q[i].obsPositive ? 1 : 0;
        mp[i].descriptor = row32(qdesc + (size_t)i * 32);
        vp[i] = &mp[i];
    }
    ORBmatcher m(ratio, true);
    int n;
    { SearchTimer timer; n = m.SearchByProjection(F, vp, th); }
    for (int i = 0; i < F.N; ++i) match[i] = index_of(F.mvpMapPoints[i], mp);
    return n;
}

// ORBmatcher::SearchForTriangulation  (ORBmatcher.cc:657)
int refm_search_triangulation(void* k1, void* k2, int nNodes1, const int* nodeId1, const int* start1, const int* idx1,
                              int nNodes2, const int* nodeId2, const int* start2, const int* idx2, const uint8_t* has1,
                              const uint8_t* has2, const float* uR1, const float* uR2, const float* F12, float ex, float ey,
                              const float* sf2, const float* sigma2_2, int nLevels, int onlyStereo, int checkOri, int* m12) {
    KeyFrame K1, K2;
    fill(K1, (RefFrame*)k1);
    fill(K2, (RefFrame*)k2);
    set_featvec(K1.mFeatVec, nNodes1, nodeId1, start1, idx1);
    set_featvec(K2.mFeatVec, nNodes2, nodeId2, start2, idx2);
    MapPoint some;
    K1.mvpMapPoints.assign(K1.N, (MapPoint*)NULL);
    K2.mvpMapPoints.assign(K2.N, (MapPoint*)NULL);
    for (int i = 0; i < K1.N; ++i) if (has1 && has1[i]) K1.mvpMapPoints[i] = &some;
    for (int i = 0; i < K2.N; ++i) if (has2 && has2[i]) K2.mvpMapPoints[i] = &some;
    if (uR1) K1.mvuRight.assign(uR1, uR1 + K1.N);
    if (uR2) K2.mvuRight.assign(uR2, uR2 + K2.N);
    K2.mvScaleFactors.assign(sf2, sf2 + nLevels);
    K2.mvLevelSigma2.assign(sigma2_2, sigma2_2 + nLevels);
    K1.Ow = vec3(ex, ey, 1.f);                 // C2 = R2w * Cw + t2w = (ex, ey, 1): the epipole in K2 is (ex, ey) (:664-670)
    cv::Mat F(3, 3, CV_32F);
    for (int i = 0; i < 9; ++i) F.at<float>(i / 3, i % 3) = F12[i];
    std::vector<std::pair<size_t, size_t> > pairs;
    ORBmatcher m(0.6f, checkOri != 0);
    int n;
    { SearchTimer timer; n = m.SearchForTriangulation(&K1, &K2, F, pairs, onlyStereo != 0); }
    for (int i = 0; i < K1.N; ++i) m12[i] = -1;
    for (size_t i = 0; i < pairs.size(); ++i) m12[pairs[i].first] = (int)pairs[i].second;
    return n;
}

// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) (ORBmatcher.cc:159, strictLow = 0) and
// ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, ...) (ORBmatcher.cc:522, strictLow = 1)
int refm_search_bow(void* k1, void* k2, int nNodes1, const int* nodeId1, const int* start1, const int* idx1, int nNodes2,
                    const int* nodeId2, const int* start2, const int* idx2, const uint8_t* valid1, const uint8_t* valid2,
                    float ratio, int checkOri, int strictLow, int* m12, int* m21) {
    KeyFrame K1;
    fill(K1, (RefFrame*)k1);
    set_featvec(K1.mFeatVec, nNodes1, nodeId1, start1, idx1);
    std::vector<MapPoint> mp1(K1.N);
    K1.mvpMapPoints.assign(K1.N, (MapPoint*)NULL);
    for (int i = 0; i < K1.N; ++i)
        if (!valid1 || valid1[i]) K1.mvpMapPoints[i] = &mp1[i];
    ORBmatcher m(ratio, checkOri != 0);
    int n;
    if (!strictLow) {
        Frame F2;
        fill(F2, (RefFrame*)k2);
        set_featvec(F2.mFeatVec, nNodes2, nodeId2, start2, idx2);
        std::vector<MapPoint*> found;
        { SearchTimer timer; n = m.SearchByBoW(&K1, F2, found); }
        for (int i = 0; i < K1.N; ++i) m12[i] = -1;
        for (int i2 = 0; i2 < F2.N; ++i2) {
            m21[i2] = index_of(found[i2], mp1);
            if (m21[i2] >= 0) m12[m21[i2]] = i2;
        }
    } else {
        KeyFrame K2;
        fill(K2, (RefFrame*)k2);
        set_featvec(K2.mFeatVec, nNodes2, nodeId2, start2, idx2);
        std::vector<MapPoint> mp2(K2.N);
        K2.mvpMapPoints.assign(K2.N, (MapPoint*)NULL);
        for (int i = 0; i < K2.N; ++i)
            if (!valid2 || valid2[i]) K2.mvpMapPoints[i] = &mp2[i];
        std::vector<MapPoint*> found;
        { SearchTimer timer; n = m.SearchByBoW(&K1, &K2, found); }
        for (int i = 0; i < K2.N; ++i) m21[i] = -1;
        for (int i1 = 0; i1 < K1.N; ++i1) {
            m12[i1] = index_of(found[i1], mp2);
            if (m12[i1] >= 0) m21[m12[i1]] = i1;
        }
    }
    return n;
}

// ORBmatcher::SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, th, ORBdist)  (ORBmatcher.cc:1500, relocalisation).
// q[i].octave is the level the reference predicts with MapPoint::PredictScale; q[i].valid = 0 covers NULL / bad / already found.
int refm_search_projection_kf(void* cur, const float* sf, int nLevels, const orbo::ProjQuery* q, const uint8_t* qdesc, int nq,
                              float th, int orbDist, const uint8_t* occupied, int* curMatch, int checkOri) {
    Frame C;
    fill(C, (RefFrame*)cur);
    C.mvScaleFactors.assign(sf, sf + nLevels);
    MapPoint taken;
    C.mvpMapPoints.assign(C.N, (MapPoint*)NULL);
    for (int i = 0; i < C.N; ++i)
        if (occupied && occupied[i]) C.mvpMapPoints[i] = &taken;
    KeyFrame K;
    std::vector<MapPoint> mp(nq);
    K.N = nq;
    K.mvKeysUn.resize(nq);
    K.mvpMapPoints.assign(nq, (MapPoint*)NULL);
    for (int i = 0; i < nq; ++i) {
        if (q[i].valid && q[i].invz != 1.f) return -2;
        mp[i].worldPos = vec3(q[i].u, q[i].v, 1.f);
        mp[i].descriptor = row32(qdesc + (size_t)i * 32);
        mp[i].predictedLevel = q[i].octave;
        K.mvKeysUn[i].angle = q[i].angle;
        if (q[i].valid) K.mvpMapPoints[i] = &mp[i];
    }
    std::set<MapPoint*> none;
    ORBmatcher m(0.9f, checkOri != 0);
    int n;
    { SearchTimer timer; n = m.SearchByProjection(C, &K, none, th, orbDist); }
    for (int i = 0; i < C.N; ++i) curMatch[i] = index_of(C.mvpMapPoints[i], mp);
    return n;
}

// ORBmatcher::SearchByProjection(KeyFrame*, cv::Mat Scw, vpPoints, vpMatched, th)  (ORBmatcher.cc:290, loop closing)
int refm_search_projection_sim3(void* kf, const float* sf, int nLevels, const orbo::ProjQuery* q, const uint8_t* qdesc, int nq,
                                int th, const uint8_t* occupied, int* match) {
    KeyFrame K;
    fill(K, (RefFrame*)kf);
    K.mvScaleFactors.assign(sf, sf + nLevels);
    MapPoint taken;
    std::vector<MapPoint*> vpMatched(K.N, (MapPoint*)NULL);
    for (int i = 0; i < K.N; ++i)
        if (occupied && occupied[i]) vpMatched[i] = &taken;
    std::vector<MapPoint> mp(nq);
    std::vector<MapPoint*> vp(nq);
    for (int i = 0; i < nq; ++i) {
        if (q[i].valid && q[i].invz != 1.f) return -2;
        mp[i].worldPos = vec3(q[i].u, q[i].v, 1.f);
        mp[i].normal = mp[i].worldPos;                       // passes the viewing-angle test (:352-355)
        mp[i].descriptor = row32(qdesc + (size_t)i * 32);
        mp[i].predictedLevel = q[i].octave;
        mp[i].bad = !q[i].valid;
        vp[i] = &mp[i];
    }
    ORBmatcher m(0.75f, true);
    int n;
    { SearchTimer timer; n = m.SearchByProjection(&K, cv::Mat::eye(4, 4, CV_32F), vp, vpMatched, th); }
    for (int i = 0; i < K.N; ++i) match[i] = index_of(vpMatched[i], mp);
    return n;
}

// ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th) (ORBmatcher.cc:825; scw = 0) and
// ORBmatcher::Fuse(KeyFrame*, cv::Mat Scw, vpPoints, th, vpReplacePoint) (ORBmatcher.cc:977; scw = 1).
// The keyframe holds no map points, so every accepted point ends in AddObservation(pKF, bestIdx): fusedIdx[i] = bestIdx or -1.
int refm_fuse(void* kf, const float* sf, const float* invSigma2, int nLevels, const float* uRight, float bf,
              const orbo::BestQuery* q, const uint8_t* qdesc, int nq, float th, int scw, int* fusedIdx) {
    KeyFrame K;
    fill(K, (RefFrame*)kf);
    K.mvScaleFactors.assign(sf, sf + nLevels);
    K.mvInvLevelSigma2.assign(invSigma2, invSigma2 + nLevels);
    if (uRight) K.mvuRight.assign(uRight, uRight + K.N);
    K.mbf = bf;
    K.mvpMapPoints.assign(K.N, (MapPoint*)NULL);
    std::vector<MapPoint> mp(nq);
    std::vector<MapPoint*> vp(nq);
    for (int i = 0; i < nq; ++i) {
        mp[i].worldPos = vec3(q[i].u, q[i].v, 1.f);
        mp[i].normal = mp[i].worldPos;
        mp[i].descriptor = row32(qdesc + (size_t)i * 32);
        mp[i].predictedLevel = q[i].level;
        mp[i].bad = !q[i].valid;
        vp[i] = &mp[i];
    }
    ORBmatcher m(0.6f, true);
    int n;
    if (!scw) {
        { SearchTimer timer; n = m.Fuse(&K, vp, th); }
    } else {
        std::vector<MapPoint*> repl(nq, (MapPoint*)NULL);
        { SearchTimer timer; n = m.Fuse(&K, cv::Mat::eye(4, 4, CV_32F), vp, th, repl); }
    }
    for (int i = 0; i < nq; ++i) fusedIdx[i] = mp[i].added.empty() ? -1 : (int)mp[i].added[0].second;
    return n;
}

// ORBmatcher::SearchBySim3  (ORBmatcher.cc:1102) with s12 = 1, R12 = I, t12 = 0 and identity keyframe poses: map point i of a
// keyframe sits at (u, v, 1), i.e. projects to (u, v) in the other one.  level = what PredictScale returns for it.
int refm_search_sim3(void* k1, void* k2, const float* sf1, const float* sf2, int nLevels, const float* uv1, const int* level1,
                     const uint8_t* has1, const float* uv2, const int* level2, const uint8_t* has2, float th, int* m12) {
    KeyFrame K1, K2;
    fill(K1, (RefFrame*)k1);
    fill(K2, (RefFrame*)k2);
    K1.mvScaleFactors.assign(sf1, sf1 + nLevels);
    K2.mvScaleFactors.assign(sf2, sf2 + nLevels);
    std::vector<MapPoint> mp1(K1.N), mp2(K2.N);
    K1.mvpMapPoints.assign(K1.N, (MapPoint*)NULL);
    K2.mvpMapPoints.assign(K2.N, (MapPoint*)NULL);
    for (int i = 0; i < K1.N; ++i) {
        mp1[i].worldPos = vec3(uv1[2 * i], uv1[2 * i + 1], 1.f);
        mp1[i].predictedLevel = level1[i];
        mp1[i].descriptor = row32(K1.mDescriptors.ptr<uchar>(i));
        if (has1[i]) K1.mvpMapPoints[i] = &mp1[i];
    }
    for (int i = 0; i < K2.N; ++i) {
        mp2[i].worldPos = vec3(uv2[2 * i], uv2[2 * i + 1], 1.f);
        mp2[i].predictedLevel = level2[i];
        mp2[i].descriptor = row32(K2.mDescriptors.ptr<uchar>(i));
        if (has2[i]) K2.mvpMapPoints[i] = &mp2[i];
    }
    std::vector<MapPoint*> v12(K1.N, (MapPoint*)NULL);
    ORBmatcher m(0.75f, true);
    int n;
    { SearchTimer timer; n = m.SearchBySim3(&K1, &K2, v12, 1.f, cv::Mat::eye(3, 3, CV_32F), cv::Mat(3, 1, CV_32F), th); }
    for (int i = 0; i < K1.N; ++i) m12[i] = index_of(v12[i], mp2);
    return n;
}

}  // extern "C"
