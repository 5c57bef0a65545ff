// TEST INFRASTRUCTURE ONLY -- CPU oracle (see match_oracle.h). Build: -O2 -ffp-contract=off.
#include "match_oracle.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>

namespace orbo {

// ORBmatcher.cc:1675-1691 -- SWAR bit count over 8 int32 words
int descriptor_distance(const uint8_t* a, const uint8_t* b) {
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t wa, wb;
        std::memcpy(&wa, a + 4 * i, 4);
        std::memcpy(&wb, b + 4 * i, 4);
        uint32_t v = wa ^ wb;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0x0F0F0F0Fu) * 0x01010101u) >> 24);
    }
    return dist;
}

// ORBmatcher.cc:1629-1670
void three_maxima(const int* sz, int L, int& i1, int& i2, int& i3) {
    int m1 = 0, m2 = 0, m3 = 0;
    i1 = i2 = i3 = -1;
    for (int i = 0; i < L; ++i) {
        const int s = sz[i];
        if (s > m1) { m3 = m2; m2 = m1; m1 = s; i3 = i2; i2 = i1; i1 = i; }
        else if (s > m2) { m3 = m2; m2 = s; i3 = i2; i2 = i; }
        else if (s > m3) { m3 = s; i3 = i; }
    }
    if ((float)m2 < 0.1f * (float)m1) { i2 = -1; i3 = -1; }
    else if ((float)m3 < 0.1f * (float)m1) { i3 = -1; }
}

// e.g. ORBmatcher.cc:475-481 ; factor = 1.0f/HISTO_LENGTH is the upstream quirk (bins are 30 degrees wide)
int rotation_bin(float a1, float a2) {
    const float factor = 1.0f / HISTO_LENGTH;
    float rot = a1 - a2;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)std::round(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

// Frame.cc:574-589 + 726-736
void FrameArrays::buildGrid() {
    std::vector<std::vector<int>> cells(GRID_COLS * GRID_ROWS);
    for (int i = 0; i < n; ++i) {
        const int px = (int)std::round((keysUn[i].x - minX) * invW);
        const int py = (int)std::round((keysUn[i].y - minY) * invH);
        if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
        cells[px * GRID_ROWS + py].push_back(i);
    }
    cellStart.assign(GRID_COLS * GRID_ROWS + 1, 0);
    cellIdx.clear();
    for (int c = 0; c < GRID_COLS * GRID_ROWS; ++c) {
        cellStart[c] = (int)cellIdx.size();
        cellIdx.insert(cellIdx.end(), cells[c].begin(), cells[c].end());
    }
    cellStart[GRID_COLS * GRID_ROWS] = (int)cellIdx.size();
}

// Frame.cc:671-724
void FrameArrays::featuresInArea(float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& out) const {
    out.clear();
    const int c0 = std::max(0, (int)std::floor((x - minX - r) * invW));
    if (c0 >= GRID_COLS) return;
    const int c1 = std::min(GRID_COLS - 1, (int)std::ceil((x - minX + r) * invW));
    if (c1 < 0) return;
    const int r0 = std::max(0, (int)std::floor((y - minY - r) * invH));
    if (r0 >= GRID_ROWS) return;
    const int r1 = std::min(GRID_ROWS - 1, (int)std::ceil((y - minY + r) * invH));
    if (r1 < 0) return;
    const bool checkLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = c0; ix <= c1; ++ix)
        for (int iy = r0; iy <= r1; ++iy) {
            const int c = ix * GRID_ROWS + iy;
            for (int k = cellStart[c]; k < cellStart[c + 1]; ++k) {
                const KeyPoint& kp = keysUn[cellIdx[k]];
                if (checkLevels) {
                    if (kp.octave < minLevel) continue;
                    if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                }
                const float dx = kp.x - x, dy = kp.y - y;
                if (std::fabs(dx) < r && std::fabs(dy) < r) out.push_back(cellIdx[k]);
            }
        }
}

// Shared tail of every search: prune matches whose rotation bin is not one of the three dominant ones.
namespace {
struct RotHist {
    std::vector<int> bins[HISTO_LENGTH];
    void add(float a1, float a2, int v) { bins[rotation_bin(a1, a2)].push_back(v); }
    template <class F> void pruneMinor(F drop) const {
        int sz[HISTO_LENGTH], i1, i2, i3;
        for (int i = 0; i < HISTO_LENGTH; ++i) sz[i] = (int)bins[i].size();
        three_maxima(sz, HISTO_LENGTH, i1, i2, i3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == i1 || i == i2 || i == i3) continue;
            for (int v : bins[i]) drop(v);
        }
    }
};
}  // namespace

// ORBmatcher.cc:405-520
int search_for_initialization(const FrameArrays& F1, const FrameArrays& F2, float* prevXY, int* m12,
                              int windowSize, float nnratio, bool checkOri) {
    int nmatches = 0;
    std::fill(m12, m12 + F1.n, -1);
    RotHist hist;
    std::vector<int> matchedDist(F2.n, INT_MAX), m21(F2.n, -1), cand;
    for (int i1 = 0; i1 < F1.n; ++i1) {
        const int level1 = F1.keysUn[i1].octave;
        if (level1 > 0) continue;
        F2.featuresInArea(prevXY[2 * i1], prevXY[2 * i1 + 1], (float)windowSize, level1, level1, cand);
        if (cand.empty()) continue;
        const uint8_t* d1 = F1.desc + (size_t)i1 * 32;
        int best = INT_MAX, second = INT_MAX, bestIdx = -1;
        for (int i2 : cand) {
            const int dist = descriptor_distance(d1, F2.desc + (size_t)i2 * 32);
            if (matchedDist[i2] <= dist) continue;
            if (dist < best) { second = best; best = dist; bestIdx = i2; }
            else if (dist < second) second = dist;
        }
        if (best <= TH_LOW && (float)best < (float)second * nnratio) {
            if (m21[bestIdx] >= 0) { m12[m21[bestIdx]] = -1; --nmatches; }
            m12[i1] = bestIdx;
            m21[bestIdx] = i1;
            matchedDist[bestIdx] = best;
            ++nmatches;
            if (checkOri) hist.add(F1.keysUn[i1].angle, F2.keysUn[bestIdx].angle, i1);
        }
    }
    if (checkOri)
        hist.pruneMinor([&](int i1) { if (m12[i1] >= 0) { m12[i1] = -1; --nmatches; } });
    for (int i1 = 0; i1 < F1.n; ++i1)
        if (m12[i1] >= 0) {
            prevXY[2 * i1] = F2.keysUn[m12[i1]].x;
            prevXY[2 * i1 + 1] = F2.keysUn[m12[i1]].y;
        }
    return nmatches;
}

// ORBmatcher.cc:1341-1498 (projection itself, :1376-1388, stays with the caller)
int search_by_projection_frame(const FrameArrays& cur, const float* sf, const float* uRight, float mbf,
                               const ProjQuery* q, const uint8_t* qdesc, int nq, float th, int mode,
                               const uint8_t* curOccupied, int* curMatch, bool checkOri, int maxDist) {
    int nmatches = 0;
    std::vector<uint8_t> occ(curOccupied, curOccupied + cur.n);
    std::fill(curMatch, curMatch + cur.n, -1);
    RotHist hist;
    std::vector<int> cand;
    for (int i = 0; i < nq; ++i) {
        if (!q[i].valid) continue;
        const float u = q[i].u, v = q[i].v;
        if (u < cur.minX || u > cur.maxX) continue;
        if (v < cur.minY || v > cur.maxY) continue;
        const int oct = q[i].octave;
        const float radius = th * sf[oct];
        if (mode == 1) cur.featuresInArea(u, v, radius, oct, -1, cand);
        else if (mode == 2) cur.featuresInArea(u, v, radius, 0, oct, cand);
        else if (mode == 3) cur.featuresInArea(u, v, radius, oct - 1, oct, cand);   // SearchByProjection(KF, Scw, ..): :362-381
        else cur.featuresInArea(u, v, radius, oct - 1, oct + 1, cand);
        if (cand.empty()) continue;
        const uint8_t* d = qdesc + (size_t)i * 32;
        int best = 256, bestIdx = -1;
        for (int i2 : cand) {
            if (occ[i2]) continue;
            if (uRight && uRight[i2] > 0) {
                const float ur = u - mbf * q[i].invz;
                const float er = std::fabs(ur - uRight[i2]);
                if (er > radius) continue;
            }
            const int dist = descriptor_distance(d, cur.desc + (size_t)i2 * 32);
            if (dist < best) { best = dist; bestIdx = i2; }
        }
        if (best <= maxDist) {   // TH_HIGH (:1453); ORBdist in the relocalisation overload (:1583)
            curMatch[bestIdx] = i;
            occ[bestIdx] = q[i].obsPositive ? 1 : 0;
            ++nmatches;
            if (checkOri) hist.add(q[i].angle, cur.keysUn[bestIdx].angle, bestIdx);
        }
    }
    if (checkOri)
        hist.pruneMinor([&](int i2) { curMatch[i2] = -1; --nmatches; });
    return nmatches;
}

// ORBmatcher.cc:45-129 (+131-137)
int search_by_projection_points(const FrameArrays& F, const float* sf, const float* uRight, const MapPointQuery* q,
                                const uint8_t* qdesc, int nq, float th, float nnratio, const uint8_t* occupied,
                                int* match) {
    int nmatches = 0;
    std::vector<uint8_t> occ(occupied, occupied + F.n);
    std::fill(match, match + F.n, -1);
    const bool bFactor = th != 1.0;
    std::vector<int> cand;
    for (int i = 0; i < nq; ++i) {
        if (!q[i].inView) continue;
        const int lvl = q[i].level;
        float r = ((double)q[i].viewCos > 0.998) ? 2.5f : 4.0f;
        if (bFactor) r *= th;
        F.featuresInArea(q[i].projX, q[i].projY, r * sf[lvl], lvl - 1, lvl, cand);
        if (cand.empty()) continue;
        const uint8_t* d = qdesc + (size_t)i * 32;
        int best = 256, bestLevel = -1, second = 256, secondLevel = -1, bestIdx = -1;
        for (int idx : cand) {
            if (occ[idx]) continue;
            if (uRight && uRight[idx] > 0) {
                const float er = std::fabs(q[i].projXR - uRight[idx]);
                if (er > r * sf[lvl]) continue;
            }
            const int dist = descriptor_distance(d, F.desc + (size_t)idx * 32);
            if (dist < best) {
                second = best; best = dist; secondLevel = bestLevel;
                bestLevel = F.keysUn[idx].octave; bestIdx = idx;
            } else if (dist < second) {
                secondLevel = F.keysUn[idx].octave; second = dist;
            }
        }
        if (best <= TH_HIGH) {
            if (bestLevel == secondLevel && (float)best > nnratio * (float)second) continue;
            match[bestIdx] = i;
            occ[bestIdx] = q[i].obsPositive ? 1 : 0;
            ++nmatches;
        }
    }
    return nmatches;
}

// ORBmatcher.cc:140-157
static bool epipolar_ok(const KeyPoint& k1, const KeyPoint& k2, const EpiParams& ep) {
    const float* F = ep.F12;  // row-major 3x3
    const float a = k1.x * F[0] + k1.y * F[3] + F[6];
    const float b = k1.x * F[1] + k1.y * F[4] + F[7];
    const float c = k1.x * F[2] + k1.y * F[5] + F[8];
    const float num = a * k2.x + b * k2.y + c;
    const float den = a * a + b * b;
    if (den == 0) return false;
    const float dsqr = num * num / den;
    return (double)dsqr < 3.84 * (double)ep.levelSigma2_2[k2.octave];
}

// ORBmatcher.cc:657-823
int search_for_triangulation(const FrameArrays& K1, const FrameArrays& K2, const FeatVec& fv1, const FeatVec& fv2,
                             const uint8_t* has1, const uint8_t* has2, const float* uR1, const float* uR2,
                             const EpiParams& ep, bool onlyStereo, bool checkOri, int* m12) {
    int nmatches = 0;
    std::fill(m12, m12 + K1.n, -1);
    RotHist hist;
    int a = 0, b = 0;
    while (a < fv1.nNodes && b < fv2.nNodes) {
        if (fv1.nodeId[a] < fv2.nodeId[b]) { ++a; continue; }   // lower_bound on a sorted map == skip ahead
        if (fv1.nodeId[a] > fv2.nodeId[b]) { ++b; continue; }
        for (int p1 = fv1.start[a]; p1 < fv1.start[a + 1]; ++p1) {
            const int idx1 = fv1.idx[p1];
            if (has1[idx1]) continue;
            const bool stereo1 = uR1 ? (uR1[idx1] >= 0) : false;
            if (onlyStereo && !stereo1) continue;
            const KeyPoint& kp1 = K1.keysUn[idx1];
            const uint8_t* d1 = K1.desc + (size_t)idx1 * 32;
            int best = TH_LOW, bestIdx = -1;
            for (int p2 = fv2.start[b]; p2 < fv2.start[b + 1]; ++p2) {
                const int idx2 = fv2.idx[p2];
                if (has2[idx2]) continue;   // vbMatched2 is never set in the reference (:725)
                const bool stereo2 = uR2 ? (uR2[idx2] >= 0) : false;
                if (onlyStereo && !stereo2) continue;
                const int dist = descriptor_distance(d1, K2.desc + (size_t)idx2 * 32);
                if (dist > TH_LOW || dist > best) continue;
                const KeyPoint& kp2 = K2.keysUn[idx2];
                if (!stereo1 && !stereo2) {
                    const float dex = ep.ex - kp2.x, dey = ep.ey - kp2.y;
                    if (dex * dex + dey * dey < 100 * ep.scaleFactors2[kp2.octave]) continue;
                }
                if (epipolar_ok(kp1, kp2, ep)) { bestIdx = idx2; best = dist; }
            }
            if (bestIdx >= 0) {
                m12[idx1] = bestIdx;
                ++nmatches;
                if (checkOri) hist.add(kp1.angle, K2.keysUn[bestIdx].angle, idx1);
            }
        }
        ++a; ++b;
    }
    if (checkOri)
        hist.pruneMinor([&](int i1) { m12[i1] = -1; --nmatches; });
    return nmatches;
}

// ORBmatcher.cc:159-288 and 522-655
int search_by_bow(const FrameArrays& K1, const FrameArrays& K2, const FeatVec& fv1, const FeatVec& fv2,
                  const uint8_t* valid1, const uint8_t* valid2, float nnratio, bool checkOri, bool strictLow, int* m12,
                  int* m21) {
    int nmatches = 0;
    std::fill(m12, m12 + K1.n, -1);
    std::fill(m21, m21 + K2.n, -1);
    RotHist hist;
    int a = 0, b = 0;
    while (a < fv1.nNodes && b < fv2.nNodes) {
        if (fv1.nodeId[a] < fv2.nodeId[b]) { ++a; continue; }
        if (fv1.nodeId[a] > fv2.nodeId[b]) { ++b; continue; }
        for (int p1 = fv1.start[a]; p1 < fv1.start[a + 1]; ++p1) {
            const int idx1 = fv1.idx[p1];
            if (valid1 && !valid1[idx1]) continue;
            const uint8_t* d1 = K1.desc + (size_t)idx1 * 32;
            int best = 256, second = 256, bestIdx = -1;
            for (int p2 = fv2.start[b]; p2 < fv2.start[b + 1]; ++p2) {
                const int idx2 = fv2.idx[p2];
                if (m21[idx2] >= 0) continue;                    // vbMatched2 / vpMapPointMatches[realIdxF]
                if (valid2 && !valid2[idx2]) continue;
                const int dist = descriptor_distance(d1, K2.desc + (size_t)idx2 * 32);
                if (dist < best) { second = best; best = dist; bestIdx = idx2; }
                else if (dist < second) second = dist;
            }
            const bool low = strictLow ? best < TH_LOW : best <= TH_LOW;
            if (low && (float)best < nnratio * (float)second) {
                m12[idx1] = bestIdx;
                m21[bestIdx] = idx1;
                ++nmatches;
                if (checkOri) hist.add(K1.keysUn[idx1].angle, K2.keysUn[bestIdx].angle, idx1);
            }
        }
        ++a; ++b;
    }
    if (checkOri)
        hist.pruneMinor([&](int i1) { m21[m12[i1]] = -1; m12[i1] = -1; --nmatches; });
    return nmatches;
}

// ORBmatcher.cc:892-944 (chi2Filter), 1051-1075, 1191-1215, 1271-1295
void search_projected_best(const FrameArrays& K, const BestQuery* q, const uint8_t* qdesc, int nq, bool chi2Filter,
                           const float* uRight, const float* invSigma2, int* bestIdx, int* bestDist) {
    std::vector<int> cand;
    for (int i = 0; i < nq; ++i) {
        bestIdx[i] = -1;
        bestDist[i] = 256;
        if (!q[i].valid) continue;
        K.featuresInArea(q[i].u, q[i].v, q[i].radius, -1, -1, cand);     // KeyFrame overload: no level filter
        const uint8_t* d = qdesc + (size_t)i * 32;
        for (int idx : cand) {
            const KeyPoint& kp = K.keysUn[idx];
            if (kp.octave < q[i].level - 1 || kp.octave > q[i].level) continue;
            if (chi2Filter) {
                const float ex = q[i].u - kp.x, ey = q[i].v - kp.y;
                if (uRight && uRight[idx] >= 0) {
                    const float er = q[i].ur - uRight[idx];
                    const float e2 = ex * ex + ey * ey + er * er;
                    if (e2 * invSigma2[kp.octave] > 7.8) continue;
                } else {
                    const float e2 = ex * ex + ey * ey;
                    if (e2 * invSigma2[kp.octave] > 5.99) continue;
                }
            }
            const int dist = descriptor_distance(d, K.desc + (size_t)idx * 32);
            if (dist < bestDist[i]) { bestDist[i] = dist; bestIdx[i] = idx; }
        }
    }
}

int bruteforce_match(const uint8_t* q, const float* qAngle, int nq, const uint8_t* t, const float* tAngle, int nt,
                     float nnratio, bool checkOri, int* bestDist, int* secondDist, int* bestIdx, int* m12) {
    int nmatches = 0;
    RotHist hist;
    for (int i = 0; i < nq; ++i) {
        int best = INT_MAX, second = INT_MAX, idx = -1;
        for (int j = 0; j < nt; ++j) {
            const int dist = descriptor_distance(q + (size_t)i * 32, t + (size_t)j * 32);
            if (dist < best) { second = best; best = dist; idx = j; }
            else if (dist < second) second = dist;
        }
        bestDist[i] = best; secondDist[i] = second; bestIdx[i] = idx;
        m12[i] = -1;
        if (best <= TH_LOW && (float)best < (float)second * nnratio) {
            m12[i] = idx;
            ++nmatches;
            if (checkOri) hist.add(qAngle[i], tAngle[idx], i);
        }
    }
    if (checkOri)
        hist.pruneMinor([&](int i1) { if (m12[i1] >= 0) { m12[i1] = -1; --nmatches; } });
    return nmatches;
}

// ORBmatcher.cc:566-618 + 634-652 with every index in one shared node and every map point valid
int kf_pair_match_count(const uint8_t* d1, const float* a1, int n1, const uint8_t* d2, const float* a2, int n2,
                        float nnratio, bool checkOri, int* m12out) {
    int nmatches = 0;
    std::vector<uint8_t> matched2(n2, 0);
    std::vector<int> m12(n1, -1);
    RotHist hist;
    for (int i1 = 0; i1 < n1; ++i1) {
        int best = 256, second = 256, bestIdx = -1;
        for (int i2 = 0; i2 < n2; ++i2) {
            if (matched2[i2]) continue;
            const int dist = descriptor_distance(d1 + (size_t)i1 * 32, d2 + (size_t)i2 * 32);
            if (dist < best) { second = best; best = dist; bestIdx = i2; }
            else if (dist < second) second = dist;
        }
        if (best < TH_LOW && (float)best < nnratio * (float)second) {
            m12[i1] = bestIdx;
            matched2[bestIdx] = 1;
            if (checkOri) hist.add(a1[i1], a2[bestIdx], i1);
            ++nmatches;
        }
    }
    if (checkOri)
        hist.pruneMinor([&](int i1) { m12[i1] = -1; --nmatches; });
    if (m12out) std::copy(m12.begin(), m12.end(), m12out);
    return nmatches;
}

// ---------------------------------------------------------------------------------------------
// projection of map points into the current frame, ORBmatcher.cc:1376-1393
// ---------------------------------------------------------------------------------------------
void project_points(const float* Rcw, const float* tcw, float fx, float fy, float cx, float cy, const float* bounds,
                    const float* xyzWorld, int n, float* u, float* v, float* invz, int32_t* valid) {
    for (int i = 0; i < n; ++i) {
        float x3Dc[3];
        gemm3_f32(Rcw, xyzWorld + 3 * i, tcw, x3Dc);                  // :1377
        const float xc = x3Dc[0], yc = x3Dc[1];
        const float invzc = (float)(1.0 / (double)x3Dc[2]);           // :1381 (1.0 is a double literal)
        u[i] = fx * xc * invzc + cx;                                  // :1387-1388
        v[i] = fy * yc * invzc + cy;
        invz[i] = invzc;
        valid[i] = !(invzc < 0) && !(u[i] < bounds[0] || u[i] > bounds[2]) && !(v[i] < bounds[1] || v[i] > bounds[3]);
    }
}

// ---------------------------------------------------------------------------------------------
// MapPoint::ComputeDistinctiveDescriptors, MapPoint.cc:257-322
// ---------------------------------------------------------------------------------------------
int distinctive_descriptor(const uint8_t* desc, int n, int* bestMedian) {
    if (bestMedian) *bestMedian = INT_MAX;
    if (n <= 0) return -1;                                                    // :271-272, :285-286
    std::vector<float> D((size_t)n * n);                                      // float Distances[N][N], :291
    for (int i = 0; i < n; ++i) {
        D[(size_t)i * n + i] = 0;
        for (int j = i + 1; j < n; ++j) {
            const int d = descriptor_distance(desc + 32 * (size_t)i, desc + 32 * (size_t)j);
            D[(size_t)i * n + j] = (float)d;
            D[(size_t)j * n + i] = (float)d;
        }
    }
    int best = INT_MAX, bestIdx = 0;                                          // :305-318
    for (int i = 0; i < n; ++i) {
        std::vector<int> v(D.begin() + (size_t)i * n, D.begin() + (size_t)(i + 1) * n);
        std::sort(v.begin(), v.end());
        const int median = v[(size_t)(0.5 * (n - 1))];
        if (median < best) { best = median; bestIdx = i; }
    }
    if (bestMedian) *bestMedian = best;
    return bestIdx;
}

// ---------------------------------------------------------------------------------------------
// Frame::ComputeStereoMatches, Frame.cc:810-984
// ---------------------------------------------------------------------------------------------
int compute_stereo_matches(const KeyPoint* keysL, const uint8_t* descL, int nL, const KeyPoint* keysR, const uint8_t* descR,
                           int nR, const StereoLevel* levelsL, const StereoLevel* levelsR, int nLevels,
                           const float* scaleFactors, const float* invScaleFactors, float mb, float mbf,
                           float* uRight, float* depth, int* sad) {
    (void)nLevels;
    for (int i = 0; i < nL; ++i) { uRight[i] = -1.0f; depth[i] = -1.0f; sad[i] = -1; }   // :812-813
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;                                        // :815
    const int nRows = levelsL[0].rows;                                                   // :817
    std::vector<std::vector<int>> rowIndices(nRows);                                     // :820-838
    for (int iR = 0; iR < nR; ++iR) {
        const float kpY = keysR[iR].y;
        const float r = 2.0f * scaleFactors[keysR[iR].octave];
        const int maxr = (int)std::ceil(kpY + r);
        const int minr = (int)std::floor(kpY - r);
        for (int yi = std::max(minr, 0); yi <= std::min(maxr, nRows - 1); ++yi) rowIndices[yi].push_back(iR);
    }
    const float minZ = mb, minD = 0, maxD = mbf / minZ;                                  // :841-843
    std::vector<std::pair<int, int>> distIdx;
    for (int iL = 0; iL < nL; ++iL) {
        const KeyPoint& kpL = keysL[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.y, uL = kpL.x;
        const long row = (long)vL;                                                       // vRowIndices[vL], :856
        if (row < 0 || row >= nRows) continue;
        const std::vector<int>& cand = rowIndices[row];
        if (cand.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH;
        int bestIdxR = 0;
        const uint8_t* dL = descL + 32 * (size_t)iL;
        for (int iR : cand) {                                                            // :873-895
            const KeyPoint& kpR = keysR[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = descriptor_distance(dL, descR + 32 * (size_t)iR);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (!(bestDist < thOrbDist)) continue;                                           // :898
        const float uR0 = keysR[bestIdxR].x;
        const float scaleFactor = invScaleFactors[kpL.octave];
        const float scaleduL = std::round(kpL.x * scaleFactor);
        const float scaledvL = std::round(kpL.y * scaleFactor);
        const float scaleduR0 = std::round(uR0 * scaleFactor);
        const int w = 5, L = 5;
        const StereoLevel& PL = levelsL[kpL.octave];
        const StereoLevel& PR = levelsR[kpL.octave];
        // IL = level(rows v-w..v+w, cols u-w..u+w) as float minus its centre (:908-910); rowRange/colRange take ints
        const int v0 = (int)(scaledvL - w), u0 = (int)(scaleduL - w);
        // cv::Mat::rowRange / colRange assert that the window lies inside the level (the reference would throw): no match
        if (u0 < 0 || u0 + 2 * w + 1 > PL.cols || v0 < 0 || v0 + 2 * w + 1 > PL.rows) continue;
        if ((int)scaleduR0 - L - w < 0) continue;
        float IL[11][11];
        auto pixL = [&](int y, int x) { return (float)PL.padded[(size_t)(y + 19) * PL.stride + (x + 19)]; };
        auto pixR = [&](int y, int x) { return (float)PR.padded[(size_t)(y + 19) * PR.stride + (x + 19)]; };
        const float cL = pixL(v0 + w, u0 + w);
        for (int y = 0; y < 11; ++y)
            for (int x = 0; x < 11; ++x) IL[y][x] = pixL(v0 + y, u0 + x) - cL;
        int bestSad = INT_MAX, bestincR = 0;
        float dists[2 * 5 + 1];
        const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;              // :918-921
        if (iniu < 0 || endu >= PR.cols) continue;
        for (int incR = -L; incR <= +L; ++incR) {                                        // :923-937
            const int ur0 = (int)(scaleduR0 + incR - w);
            const float cR = pixR(v0 + w, ur0 + w);
            double acc = 0;                                                              // cv::norm(NORM_L1) sums in double
            for (int y = 0; y < 11; ++y)
                for (int x = 0; x < 11; ++x) acc += (double)std::fabs(IL[y][x] - (pixR(v0 + y, ur0 + x) - cR));
            const float dist = (float)acc;
            if (dist < bestSad) { bestSad = (int)dist; bestincR = incR; }
            dists[L + incR] = dist;
        }
        if (bestincR == -L || bestincR == L) continue;                                   // :939
        const float dist1 = dists[L + bestincR - 1], dist2 = dists[L + bestincR], dist3 = dists[L + bestincR + 1];
        const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));  // :947
        if (deltaR < -1 || deltaR > 1) continue;
        float bestuR = scaleFactors[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);   // :953
        float disparity = (uL - bestuR);
        if (disparity >= minD && disparity < maxD) {
            if (disparity <= 0) {
                disparity = 0.01;
                bestuR = uL - 0.01;                                                      // double arithmetic, then float
            }
            depth[iL] = mbf / disparity;
            uRight[iL] = bestuR;
            sad[iL] = bestSad;
            distIdx.push_back(std::pair<int, int>(bestSad, iL));
        }
    }
    if (distIdx.empty()) return 0;                                                       // (reference: undefined)
    std::sort(distIdx.begin(), distIdx.end());                                           // :971-984
    const float median = distIdx[distIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    int kept = (int)distIdx.size();
    for (int i = (int)distIdx.size() - 1; i >= 0; --i) {
        if (distIdx[i].first < thDist) break;
        uRight[distIdx[i].second] = -1;
        depth[distIdx[i].second] = -1;
        sad[distIdx[i].second] = -1;
        --kept;
    }
    return kept;
}

}  // namespace orbo
