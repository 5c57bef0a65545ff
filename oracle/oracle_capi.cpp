// TEST INFRASTRUCTURE ONLY -- flat C entry points of the CPU oracle for ctypes (tests/, bench.py cpu_baseline leg,
// __graft_entry__.smoke()).  Nothing under vi-orb-slam-icra2018_b200/ links or loads this library.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "match_oracle.h"
#include "bow_oracle.h"
#include "orb_oracle.h"

using namespace orbo;

extern "C" {

// ---- primitives -------------------------------------------------------------------------------
void orbo_resize(const uint8_t* s, int sw, int sh, int ss, uint8_t* d, int dw, int dh, int ds) {
    resize_linear_u8(s, sw, sh, ss, d, dw, dh, ds);
}
void orbo_resize_table(int ssize, int dsize, int* ofs, int16_t* coef) { resize_axis_table(ssize, dsize, ofs, coef); }
void orbo_border(const uint8_t* s, int w, int h, int ss, uint8_t* d, int ds, int b) {
    copy_make_border_reflect101(s, w, h, ss, d, ds, b);
}
void orbo_undistort(const float* xy, int n, const float* k4, const float* dist, int nDist, float* out) {
    undistort_points(xy, n, k4[0], k4[1], k4[2], k4[3], dist, nDist, out);
}
int orbo_fast(const uint8_t* img, int w, int h, int stride, int th, KeyPoint* out, int cap) {
    std::vector<KeyPoint> v;
    fast9_16(img, w, h, stride, th, true, v);
    for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
    return (int)v.size();
}
int orbo_fast_score(const uint8_t* p, int stride) { return fast_score(p, stride); }
void orbo_blur(const uint8_t* s, int w, int h, int ss, uint8_t* d, int ds) { gaussian_blur_7x7_s2(s, w, h, ss, d, ds); }
float orbo_atan2(float y, float x) { return fast_atan2(y, x); }
void orbo_atan2_n(const float* y, const float* x, float* out, int n) {
    for (int i = 0; i < n; ++i) out[i] = fast_atan2(y[i], x[i]);
}
float orbo_cosf(float a) { return cosf(a); }
float orbo_sinf(float a) { return sinf(a); }
void orbo_sincosf_n(const float* a, float* s, float* c, int n) {
    for (int i = 0; i < n; ++i) { s[i] = sinf(a[i]); c[i] = cosf(a[i]); }
}
int orbo_round(float v) { return cv_round(v); }
const int8_t* orbo_pattern() { return brief_pattern(); }

int orbo_distribute(const KeyPoint* keys, int n, int minX, int maxX, int minY, int maxY, int N, KeyPoint* out, int cap) {
    std::vector<KeyPoint> in(keys, keys + n);
    std::vector<KeyPoint> r = distribute_octree(in, minX, maxX, minY, maxY, N);
    for (int i = 0; i < (int)r.size() && i < cap; ++i) out[i] = r[i];
    return (int)r.size();
}
float orbo_ic_angle(const uint8_t* center, int stride, const int* umax16) {
    std::vector<int> u(umax16, umax16 + 16);
    return ic_angle(center, stride, u);
}
void orbo_descriptor(float angleDeg, const uint8_t* center, int stride, uint8_t* desc32) {
    orb_descriptor(angleDeg, center, stride, desc32);
}

// ---- extractor --------------------------------------------------------------------------------
void* orbo_create(int nf, float sf, int nl, int ini, int mn) { return new Extractor(nf, sf, nl, ini, mn); }
void orbo_destroy(void* h) { delete (Extractor*)h; }
// returns number of keypoints (even if > cap; only cap are written), -1 on unsupported geometry
int orbo_extract(void* h, const uint8_t* img, int w, int hgt, int stride, KeyPoint* kps, uint8_t* desc, int cap) {
    Extractor* e = (Extractor*)h;
    std::vector<KeyPoint> k;
    std::vector<uint8_t> d;
    if (!e->extract(img, w, hgt, stride, k, d)) return -1;
    const int n = std::min((int)k.size(), cap);
    if (n > 0) {
        std::memcpy(kps, k.data(), (size_t)n * sizeof(KeyPoint));
        std::memcpy(desc, d.data(), (size_t)n * 32);
    }
    return (int)k.size();
}
void orbo_tables(void* h, float* scale, float* inv, float* s2, float* is2, int* perLevel, int* umax16) {
    Extractor* e = (Extractor*)h;
    for (int i = 0; i < e->nlevels; ++i) {
        scale[i] = e->scale[i]; inv[i] = e->invScale[i]; s2[i] = e->sigma2[i]; is2[i] = e->invSigma2[i];
        perLevel[i] = e->featuresPerLevel[i];
    }
    for (int i = 0; i < 16; ++i) umax16[i] = e->umax[i];
}
void orbo_level_size(void* h, int l, int* w, int* hg) {
    Extractor* e = (Extractor*)h;
    *w = e->pyramid[l].w; *hg = e->pyramid[l].h;
}
void orbo_level_padded(void* h, int l, uint8_t* out) {   // (w+38)*(h+38) bytes
    Extractor* e = (Extractor*)h;
    std::memcpy(out, e->pyramid[l].padded.data(), e->pyramid[l].padded.size());
}
int orbo_level_blurred(void* h, int l, uint8_t* out) {   // w*h bytes; returns 0 if the level was not blurred
    Extractor* e = (Extractor*)h;
    if (l >= (int)e->blurred.size() || e->blurred[l].empty()) return 0;
    std::memcpy(out, e->blurred[l].data(), e->blurred[l].size());
    return 1;
}
int orbo_level_candidates(void* h, int l, KeyPoint* out, int cap) {
    Extractor* e = (Extractor*)h;
    const std::vector<KeyPoint>& v = e->candidates[l];
    for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
    return (int)v.size();
}
int orbo_level_selected(void* h, int l, KeyPoint* out, int cap) {
    Extractor* e = (Extractor*)h;
    const std::vector<KeyPoint>& v = e->selected[l];
    for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
    return (int)v.size();
}
void orbo_stage_ms(void* h, double* out3) {
    Extractor* e = (Extractor*)h;
    out3[0] = e->msPyramid; out3[1] = e->msKeypoints; out3[2] = e->msDescriptors;
}

// ---- matcher ----------------------------------------------------------------------------------
int orbo_distance(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }
void orbo_three_maxima(const int* sizes, int L, int* out3) { three_maxima(sizes, L, out3[0], out3[1], out3[2]); }
int orbo_rotation_bin(float a1, float a2) { return rotation_bin(a1, a2); }

struct OrboFrame {
    FrameArrays fa;
    std::vector<KeyPoint> keys;
    std::vector<uint8_t> desc;
};
void* orbo_frame_create(const KeyPoint* keysUn, const uint8_t* desc, int n, float minX, float minY, float maxX,
                        float maxY) {
    OrboFrame* f = new OrboFrame;
    f->keys.assign(keysUn, keysUn + n);
    f->desc.assign(desc, desc + (size_t)n * 32);
    f->fa.n = n;
    f->fa.keysUn = f->keys.data();
    f->fa.desc = f->desc.data();
    f->fa.minX = minX; f->fa.minY = minY; f->fa.maxX = maxX; f->fa.maxY = maxY;
    f->fa.invW = (float)GRID_COLS / (maxX - minX);   // Frame.cc:93-94
    f->fa.invH = (float)GRID_ROWS / (maxY - minY);
    f->fa.buildGrid();
    return f;
}
void orbo_frame_destroy(void* f) { delete (OrboFrame*)f; }
void orbo_frame_grid(void* f, int* cellStart /* 3073 */, int* cellIdx /* n */) {
    OrboFrame* F = (OrboFrame*)f;
    std::copy(F->fa.cellStart.begin(), F->fa.cellStart.end(), cellStart);
    std::copy(F->fa.cellIdx.begin(), F->fa.cellIdx.end(), cellIdx);
}
int orbo_frame_area(void* f, float x, float y, float r, int minLevel, int maxLevel, int* out, int cap) {
    std::vector<int> v;
    ((OrboFrame*)f)->fa.featuresInArea(x, y, r, minLevel, maxLevel, v);
    for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
    return (int)v.size();
}
int orbo_search_init(void* f1, void* f2, float* prevXY, int* m12, int window, float ratio, int checkOri) {
    return search_for_initialization(((OrboFrame*)f1)->fa, ((OrboFrame*)f2)->fa, prevXY, m12, window, ratio,
                                     checkOri != 0);
}
int orbo_search_projection(void* cur, const float* sf, const float* uRight, float mbf, const ProjQuery* q,
                           const uint8_t* qdesc, int nq, float th, int mode, const uint8_t* occupied, int* curMatch,
                           int checkOri) {
    return search_by_projection_frame(((OrboFrame*)cur)->fa, sf, uRight, mbf, q, qdesc, nq, th, mode, occupied,
                                      curMatch, checkOri != 0);
}
int orbo_search_projection_ex(void* cur, const float* sf, const float* uRight, float mbf, const ProjQuery* q,
                              const uint8_t* qdesc, int nq, float th, int mode, int maxDist, const uint8_t* occupied,
                              int* curMatch, int checkOri) {
    return search_by_projection_frame(((OrboFrame*)cur)->fa, sf, uRight, mbf, q, qdesc, nq, th, mode, occupied,
                                      curMatch, checkOri != 0, maxDist);
}
int orbo_search_points(void* F, const float* sf, const float* uRight, const MapPointQuery* q, const uint8_t* qdesc,
                       int nq, float th, float ratio, const uint8_t* occupied, int* match) {
    return search_by_projection_points(((OrboFrame*)F)->fa, sf, uRight, q, qdesc, nq, th, ratio, occupied, match);
}
int orbo_search_triangulation(void* k1, void* k2, int nNodes1, const int* nodeId1, const int* start1, const int* idx1,
                              int nNodes2, const int* nodeId2, const int* start2, const int* idx2,
                              const uint8_t* has1, const uint8_t* has2, const float* uR1, const float* uR2,
                              const float* F12, float ex, float ey, const float* sf2, const float* sigma2_2,
                              int onlyStereo, int checkOri, int* m12) {
    FeatVec a{nNodes1, nodeId1, start1, idx1}, b{nNodes2, nodeId2, start2, idx2};
    EpiParams ep;
    std::memcpy(ep.F12, F12, sizeof(ep.F12));
    ep.ex = ex; ep.ey = ey; ep.scaleFactors2 = sf2; ep.levelSigma2_2 = sigma2_2;
    return search_for_triangulation(((OrboFrame*)k1)->fa, ((OrboFrame*)k2)->fa, a, b, has1, has2, uR1, uR2, ep,
                                    onlyStereo != 0, checkOri != 0, m12);
}
int orbo_search_bow(void* k1, void* k2, int nNodes1, const int* nodeId1, const int* start1, const int* idx1, int nNodes2,
                    const int* nodeId2, const int* start2, const int* idx2, const uint8_t* valid1, const uint8_t* valid2,
                    float ratio, int checkOri, int strictLow, int* m12, int* m21) {
    FeatVec a{nNodes1, nodeId1, start1, idx1}, b{nNodes2, nodeId2, start2, idx2};
    return search_by_bow(((OrboFrame*)k1)->fa, ((OrboFrame*)k2)->fa, a, b, valid1, valid2, ratio, checkOri != 0,
                         strictLow != 0, m12, m21);
}
void orbo_search_best(void* k, const BestQuery* q, const uint8_t* qdesc, int nq, int chi2, const float* uRight,
                      const float* invSigma2, int* bestIdx, int* bestDist) {
    search_projected_best(((OrboFrame*)k)->fa, q, qdesc, nq, chi2 != 0, uRight, invSigma2, bestIdx, bestDist);
}
int orbo_bruteforce(const uint8_t* q, const float* qa, int nq, const uint8_t* t, const float* ta, int nt, float ratio,
                    int checkOri, int* best, int* second, int* idx, int* m12) {
    return bruteforce_match(q, qa, nq, t, ta, nt, ratio, checkOri != 0, best, second, idx, m12);
}
// vocabulary-tree descent; the CSR outputs are written into caller arrays sized n (+1), the counts come back through nOut[3]
void orbo_bow_transform(int nNodes, int L, const uint8_t* ndesc, const int* childStart, const int* children, const int* wordId,
                        const double* weight, const uint8_t* desc, int n, int levelsup, int* pfWord, double* pfWeight,
                        int* pfNode, int* bowWord, double* bowValue, int* fvNode, int* fvStart, int* fvIdx, int* nOut) {
    VocabArrays V;
    V.nNodes = nNodes; V.L = L; V.desc = ndesc; V.childStart = childStart; V.children = children; V.wordId = wordId;
    V.weight = weight;
    std::vector<int> bw, fn, fs, fi;
    std::vector<double> bv;
    bow_transform(V, desc, n, levelsup, bw, bv, fn, fs, fi, pfWord, pfWeight, pfNode);
    std::copy(bw.begin(), bw.end(), bowWord);
    std::copy(bv.begin(), bv.end(), bowValue);
    std::copy(fn.begin(), fn.end(), fvNode);
    std::copy(fs.begin(), fs.end(), fvStart);
    std::copy(fi.begin(), fi.end(), fvIdx);
    nOut[0] = (int)bw.size(); nOut[1] = (int)fn.size(); nOut[2] = (int)fi.size();
}
void orbo_gemm3(const float* A, const float* b, const float* c, float* d) { gemm3_f32(A, b, c, d); }
void orbo_project(const float* Rcw, const float* tcw, const float* k4, const float* bounds, const float* xyz, int n, float* u,
                  float* v, float* invz, int32_t* valid) {
    project_points(Rcw, tcw, k4[0], k4[1], k4[2], k4[3], bounds, xyz, n, u, v, invz, valid);
}
void orbo_distinctive(const uint8_t* desc, const int* start, int nPoints, int* best, int* bestMedian) {
    for (int p = 0; p < nPoints; ++p)
        best[p] = distinctive_descriptor(desc + 32 * (size_t)start[p], start[p + 1] - start[p], bestMedian + p);
}
int orbo_stereo(const KeyPoint* keysL, const uint8_t* descL, int nL, const KeyPoint* keysR, const uint8_t* descR, int nR,
                const uint8_t* const* paddedL, const uint8_t* const* paddedR, const int* cols, const int* rows, int nLevels,
                const float* sf, const float* isf, float mb, float mbf, float* uRight, float* depth, int* sad) {
    std::vector<StereoLevel> L(nLevels), R(nLevels);
    for (int l = 0; l < nLevels; ++l) {
        L[l] = StereoLevel{paddedL[l], cols[l] + 38, cols[l], rows[l]};
        R[l] = StereoLevel{paddedR[l], cols[l] + 38, cols[l], rows[l]};
    }
    return compute_stereo_matches(keysL, descL, nL, keysR, descR, nR, L.data(), R.data(), nLevels, sf, isf, mb, mbf, uRight,
                                  depth, sad);
}
int orbo_kf_pair(const uint8_t* d1, const float* a1, int n1, const uint8_t* d2, const float* a2, int n2, float ratio,
                 int checkOri, int* m12) {
    return kf_pair_match_count(d1, a1, n1, d2, a2, n2, ratio, checkOri != 0, m12);
}

}  // extern "C"
