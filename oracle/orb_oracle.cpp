// TEST INFRASTRUCTURE ONLY -- CPU oracle (see orb_oracle.h). Build: -O2 -ffp-contract=off, no -march=native.
#include "orb_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <list>
#include <utility>

namespace orbo {

static const int8_t kPattern[1024] = {
#include "brief_pattern.inc"
};
const int8_t* brief_pattern() { return kPattern; }

// -------------------------------------------------------------------------------------------------
// ctor tables -- ORBextractor.cc:412-472
// -------------------------------------------------------------------------------------------------
Extractor::Extractor(int nf, float sf, int nl, int ini, int mn)
    : nfeatures(nf), nlevels(nl), iniTh(ini), minTh(mn), scaleFactor((double)sf) {
    scale.assign(nl, 1.f);
    sigma2.assign(nl, 1.f);
    for (int i = 1; i < nl; ++i) {
        scale[i] = (float)((double)scale[i - 1] * scaleFactor);  // float * double member (:423)
        sigma2[i] = scale[i] * scale[i];
    }
    invScale.resize(nl);
    invSigma2.resize(nl);
    for (int i = 0; i < nl; ++i) {
        invScale[i] = 1.0f / scale[i];
        invSigma2[i] = 1.0f / sigma2[i];
    }
    featuresPerLevel.assign(nl, 0);
    const float factor = (float)(1.0 / scaleFactor);  // 1.0f / double (:437)
    float desired = (float)nf * (1.f - factor) / (1.f - (float)std::pow((double)factor, (double)nl));
    int sum = 0;
    for (int l = 0; l < nl - 1; ++l) {
        featuresPerLevel[l] = cv_round(desired);
        sum += featuresPerLevel[l];
        desired *= factor;
    }
    featuresPerLevel[nl - 1] = std::max(nf - sum, 0);

    // circular patch row extents (:456-471)
    umax.assign(kHalfPatch + 1, 0);
    const int vmax = cv_floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    const int vmin = cv_ceil(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (int v = 0; v <= vmax; ++v) umax[v] = cv_round_d(std::sqrt(hp2 - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
    pyramid.resize(nl);
}

// -------------------------------------------------------------------------------------------------
// ComputePyramid -- :1128-1153.  Level l is resized from level l-1 (chained), then framed by 19 px.
// -------------------------------------------------------------------------------------------------
void Extractor::computePyramid(const uint8_t* img, int w, int h, int stride) {
    for (int l = 0; l < nlevels; ++l) {
        PyramidLevel& L = pyramid[l];
        L.w = cv_round((float)w * invScale[l]);
        L.h = cv_round((float)h * invScale[l]);
        L.stride = L.w + 2 * kEdge;
        L.padded.assign((size_t)L.stride * (L.h + 2 * kEdge), 0);
        std::vector<uint8_t> plain((size_t)L.w * L.h);
        if (l == 0) {
            for (int y = 0; y < h; ++y) std::memcpy(&plain[(size_t)y * w], img + (size_t)y * stride, w);
        } else {
            const PyramidLevel& P = pyramid[l - 1];
            resize_linear_u8(P.roi(), P.w, P.h, P.stride, plain.data(), L.w, L.h, L.w);
        }
        copy_make_border_reflect101(plain.data(), L.w, L.h, L.w, L.padded.data(), L.stride, kEdge);
    }
}

// -------------------------------------------------------------------------------------------------
// Quadtree -- :483-539 (DivideNode) and :541-765 (DistributeOctTree), literal std::list simulation.
// -------------------------------------------------------------------------------------------------
namespace {
struct Node {
    int ulx = 0, uly = 0, urx = 0, ury = 0, blx = 0, bly = 0, brx = 0, bry = 0;
    std::vector<KeyPoint> keys;
    std::list<Node>::iterator self;
    bool leaf = false;   // bNoMore
    long serial = 0;     // creation order; stands in for the heap address compared at :686
};

void split(const Node& p, Node c[4]) {
    const int halfX = (int)std::ceil((float)(p.urx - p.ulx) / 2);
    const int halfY = (int)std::ceil((float)(p.bry - p.uly) / 2);
    const int mx = p.ulx + halfX, my = p.uly + halfY;
    // n1: upper-left, n2: upper-right, n3: lower-left, n4: lower-right
    c[0].ulx = p.ulx; c[0].uly = p.uly; c[0].urx = mx;    c[0].ury = p.uly;
    c[0].blx = p.ulx; c[0].bly = my;    c[0].brx = mx;    c[0].bry = my;
    c[1].ulx = mx;    c[1].uly = p.uly; c[1].urx = p.urx; c[1].ury = p.ury;
    c[1].blx = mx;    c[1].bly = my;    c[1].brx = p.urx; c[1].bry = my;
    c[2].ulx = p.ulx; c[2].uly = my;    c[2].urx = mx;    c[2].ury = my;
    c[2].blx = p.blx; c[2].bly = p.bly; c[2].brx = mx;    c[2].bry = p.bly;
    c[3].ulx = mx;    c[3].uly = my;    c[3].urx = p.urx; c[3].ury = my;
    c[3].blx = mx;    c[3].bly = p.bly; c[3].brx = p.brx; c[3].bry = p.bry;
    for (const KeyPoint& k : p.keys) {
        const bool left = k.x < (float)mx, top = k.y < (float)my;   // float vs int compare (:517-527)
        c[left ? (top ? 0 : 2) : (top ? 1 : 3)].keys.push_back(k);
    }
    for (int i = 0; i < 4; ++i) c[i].leaf = c[i].keys.size() == 1;
}
}  // namespace

std::vector<KeyPoint> distribute_octree(const std::vector<KeyPoint>& keys, int minX, int maxX, int minY,
                                        int maxY, int N) {
    std::vector<KeyPoint> result;
    const int nIni = (int)std::round((float)(maxX - minX) / (float)(maxY - minY));
    if (nIni < 1) return result;  // reference: division by zero / out-of-range index
    const float hX = (float)(maxX - minX) / (float)nIni;

    std::list<Node> nodes;
    long serial = 0;
    std::vector<Node*> roots(nIni);
    for (int i = 0; i < nIni; ++i) {
        Node n;
        n.ulx = (int)(hX * (float)i);
        n.urx = (int)(hX * (float)(i + 1));
        n.uly = n.ury = 0;
        n.blx = n.ulx; n.brx = n.urx;
        n.bly = n.bry = maxY - minY;
        n.serial = serial++;
        nodes.push_back(n);
        roots[i] = &nodes.back();
    }
    for (const KeyPoint& k : keys) roots[(size_t)(k.x / hX)]->keys.push_back(k);

    for (auto it = nodes.begin(); it != nodes.end();) {
        if (it->keys.size() == 1) { it->leaf = true; ++it; }
        else if (it->keys.empty()) it = nodes.erase(it);
        else ++it;
    }

    typedef std::pair<int, Node*> SizedNode;
    auto bySizeThenAge = [](const SizedNode& a, const SizedNode& b) {
        return a.first != b.first ? a.first < b.first : a.second->serial < b.second->serial;
    };
    // pushes the non-empty children to the list front in the order n1..n4; returns those with >1 key
    auto emitChildren = [&](Node c[4], std::vector<SizedNode>& expandable) {
        int n = 0;
        for (int i = 0; i < 4; ++i) {
            if (c[i].keys.empty()) continue;
            c[i].serial = serial++;
            nodes.push_front(c[i]);
            if (c[i].keys.size() > 1) {
                ++n;
                expandable.push_back(SizedNode((int)c[i].keys.size(), &nodes.front()));
                nodes.front().self = nodes.begin();
            }
        }
        return n;
    };

    bool done = false;
    std::vector<SizedNode> expandable;
    while (!done) {
        const int before = (int)nodes.size();
        int nToExpand = 0;
        expandable.clear();
        for (auto it = nodes.begin(); it != nodes.end();) {
            if (it->leaf) { ++it; continue; }
            Node c[4];
            split(*it, c);
            nToExpand += emitChildren(c, expandable);
            it = nodes.erase(it);
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == before) {
            done = true;
        } else if ((int)nodes.size() + nToExpand * 3 > N) {
            while (!done) {
                const int before2 = (int)nodes.size();
                std::vector<SizedNode> prev = expandable;
                expandable.clear();
                std::sort(prev.begin(), prev.end(), bySizeThenAge);
                for (int j = (int)prev.size() - 1; j >= 0; --j) {
                    Node c[4];
                    split(*prev[j].second, c);
                    emitChildren(c, expandable);
                    nodes.erase(prev[j].second->self);
                    if ((int)nodes.size() >= N) break;
                }
                if ((int)nodes.size() >= N || (int)nodes.size() == before2) done = true;
            }
        }
    }

    result.reserve(nodes.size());
    for (const Node& n : nodes) {
        const KeyPoint* best = &n.keys[0];
        for (size_t k = 1; k < n.keys.size(); ++k)
            if (n.keys[k].response > best->response) best = &n.keys[k];
        result.push_back(*best);
    }
    return result;
}

// -------------------------------------------------------------------------------------------------
// IC_Angle -- :79-106
// -------------------------------------------------------------------------------------------------
float ic_angle(const uint8_t* c, int stride, const std::vector<int>& umax) {
    int m01 = 0, m10 = 0;
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
    for (int v = 1; v <= kHalfPatch; ++v) {
        int vsum = 0;
        const int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            const int lo = c[u + v * stride], hi = c[u - v * stride];
            vsum += lo - hi;
            m10 += u * (lo + hi);
        }
        m01 += v * vsum;
    }
    return fast_atan2((float)m01, (float)m10);
}

// -------------------------------------------------------------------------------------------------
// computeOrbDescriptor -- :110-149.  cos/sin of a float under `using namespace std` are cosf/sinf.
// -------------------------------------------------------------------------------------------------
void orb_descriptor(float angleDeg, const uint8_t* center, int stride, uint8_t* desc) {
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float angle = angleDeg * factorPI;
    const float a = cosf(angle), b = sinf(angle);
    const int8_t* p = kPattern;
    for (int i = 0; i < 32; ++i) {
        int val = 0;
        for (int k = 0; k < 8; ++k, p += 4) {
            const float x0 = (float)p[0], y0 = (float)p[1], x1 = (float)p[2], y1 = (float)p[3];
            const int t0 = center[cv_round(x0 * b + y0 * a) * stride + cv_round(x0 * a - y0 * b)];
            const int t1 = center[cv_round(x1 * b + y1 * a) * stride + cv_round(x1 * a - y1 * b)];
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

// -------------------------------------------------------------------------------------------------
// ComputeKeyPointsOctTree -- :767-855
// -------------------------------------------------------------------------------------------------
bool Extractor::computeKeyPoints() {
    candidates.assign(nlevels, {});
    selected.assign(nlevels, {});
    const float W = 30;
    for (int l = 0; l < nlevels; ++l) {
        const PyramidLevel& L = pyramid[l];
        const int minBX = kEdge - 3, minBY = minBX;
        const int maxBX = L.w - kEdge + 3, maxBY = L.h - kEdge + 3;
        const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
        const int nCols = (int)(width / W), nRows = (int)(height / W);
        if (nCols < 1 || nRows < 1) return false;
        const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
        std::vector<KeyPoint>& cand = candidates[l];
        std::vector<KeyPoint> cell;
        for (int i = 0; i < nRows; ++i) {
            const float iniY = (float)(minBY + i * hCell);
            float maxY = iniY + hCell + 6;
            if (iniY >= maxBY - 3) continue;
            if (maxY > maxBY) maxY = (float)maxBY;
            for (int j = 0; j < nCols; ++j) {
                const float iniX = (float)(minBX + j * wCell);
                float maxX = iniX + wCell + 6;
                if (iniX >= maxBX - 6) continue;
                if (maxX > maxBX) maxX = (float)maxBX;
                const uint8_t* roi = L.roi() + (size_t)(int)iniY * L.stride + (int)iniX;
                const int rw = (int)maxX - (int)iniX, rh = (int)maxY - (int)iniY;
                fast9_16(roi, rw, rh, L.stride, iniTh, true, cell);
                if (cell.empty()) fast9_16(roi, rw, rh, L.stride, minTh, true, cell);
                for (KeyPoint& k : cell) {
                    k.x += (float)(j * wCell);
                    k.y += (float)(i * hCell);
                    cand.push_back(k);
                }
            }
        }
        if ((int)std::round(width / height) < 1) return false;
        std::vector<KeyPoint>& sel = selected[l];
        sel = distribute_octree(cand, minBX, maxBX, minBY, maxBY, featuresPerLevel[l]);
        const int scaledPatch = (int)((float)kPatchSize * scale[l]);
        for (KeyPoint& k : sel) {
            k.x += (float)minBX;
            k.y += (float)minBY;
            k.octave = l;
            k.size = (float)scaledPatch;
        }
    }
    for (int l = 0; l < nlevels; ++l) {
        const PyramidLevel& L = pyramid[l];
        for (KeyPoint& k : selected[l])
            k.angle = ic_angle(L.roi() + (size_t)cv_round(k.y) * L.stride + cv_round(k.x), L.stride, umax);
    }
    return true;
}

// -------------------------------------------------------------------------------------------------
// operator() -- :1045-1126
// -------------------------------------------------------------------------------------------------
bool Extractor::extract(const uint8_t* img, int w, int h, int stride, std::vector<KeyPoint>& kps,
                        std::vector<uint8_t>& desc) {
    typedef std::chrono::steady_clock clk;
    kps.clear();
    desc.clear();
    if (!img || w <= 0 || h <= 0) return true;  // empty image: outputs untouched in the reference (:1048)
    const auto t0 = clk::now();
    computePyramid(img, w, h, stride);
    const auto t1 = clk::now();
    if (!computeKeyPoints()) return false;
    const auto t2 = clk::now();

    size_t total = 0;
    for (int l = 0; l < nlevels; ++l) total += selected[l].size();
    desc.assign(total * 32, 0);
    kps.reserve(total);
    blurred.assign(nlevels, {});
    size_t off = 0;
    for (int l = 0; l < nlevels; ++l) {
        const std::vector<KeyPoint>& sel = selected[l];
        if (sel.empty()) continue;
        const PyramidLevel& L = pyramid[l];
        std::vector<uint8_t>& B = blurred[l];
        B.resize((size_t)L.w * L.h);
        gaussian_blur_7x7_s2(L.roi(), L.w, L.h, L.stride, B.data(), L.w);
        for (size_t i = 0; i < sel.size(); ++i)
            orb_descriptor(sel[i].angle, &B[(size_t)cv_round(sel[i].y) * L.w + cv_round(sel[i].x)], L.w,
                           &desc[(off + i) * 32]);
        off += sel.size();
        for (KeyPoint k : sel) {
            if (l != 0) { k.x *= scale[l]; k.y *= scale[l]; }
            kps.push_back(k);
        }
    }
    const auto t3 = clk::now();
    msPyramid = std::chrono::duration<double, std::milli>(t1 - t0).count();
    msKeypoints = std::chrono::duration<double, std::milli>(t2 - t1).count();
    msDescriptors = std::chrono::duration<double, std::milli>(t3 - t2).count();
    return true;
}

}  // namespace orbo
