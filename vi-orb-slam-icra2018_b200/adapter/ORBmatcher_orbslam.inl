// Definitions of the reference-signature overloads of ORB_SLAM2::ORBmatcher (declared in adapter/ORBmatcher.h under
// ORBB200_WITH_ORBSLAM).  Include this file from ONE translation unit of the host project after Frame.h, KeyFrame.h and
// MapPoint.h.  The host project's own headers need OpenCV, Eigen, DBoW2 and g2o, none of which is installed here; in this
// repository the file is compiled and run against the stand-ins of oracle/ref_shim/matcher (test infrastructure) and
// compared with the reference's own ORBmatcher.cc on the same object graphs (tests/test_adapter_gpu.py).  Each function
// states which reference lines it stands in for; INTEGRATION.md section 3 lists what is read from and written back to the
// objects.  Everything between "flatten" and "write back" runs on the GPU.
#ifdef ORBB200_WITH_ORBSLAM

namespace ORB_SLAM2 {
namespace orbb_detail {

inline std::vector<orb_keypoint> keys_of(const std::vector<cv::KeyPoint>& v) {
    static_assert(sizeof(cv::KeyPoint) == sizeof(orb_keypoint), "cv::KeyPoint layout");
    std::vector<orb_keypoint> k(v.size());
    if (!v.empty()) std::memcpy(k.data(), v.data(), v.size() * sizeof(orb_keypoint));
    return k;
}
inline std::vector<unsigned char> rows_of(const cv::Mat& d) {
    std::vector<unsigned char> out((size_t)d.rows * 32);
    for (int i = 0; i < d.rows; ++i) std::memcpy(&out[(size_t)i * 32], d.ptr(i), 32);
    return out;
}
inline int frame_kind(const Frame&) { return 0; }
inline int frame_kind(const KeyFrame&) { return 1; }

inline ORBmatcher::FeatureVectorCSR flatten(const DBoW2::FeatureVector& fv) {
    ORBmatcher::FeatureVectorCSR c;
    c.start.push_back(0);
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {   // std::map: ascending node ids
        c.nodeId.push_back((int)it->first);
        for (size_t k = 0; k < it->second.size(); ++k) c.idx.push_back((int)it->second[k]);
        c.start.push_back((int)c.idx.size());
    }
    return c;
}

// A map point carried into a keyframe's image the way Fuse / SearchBySim3 / SearchByProjection(KF, Scw) do it on the
// host (e.g. ORBmatcher.cc:853-890): p = R*X + t, positive depth, u = fx*(x/z) + cx, inside the image, distance inside the
// point's scale-invariance range, optionally the 60-degree viewing-angle gate.  `floatReciprocal` keeps the one spelling
// difference between the overloads (1/z in float at :861 and :331, 1.0/z in double elsewhere).
struct Projected { bool ok; float u, v, invz, dist; };
template <class KF>
inline Projected project_into(MapPoint* pMP, const cv::Mat& X, const cv::Mat& p3Dc, const cv::Mat& PO, KF* kf, bool angleGate,
                              bool floatReciprocal) {
    Projected r = {false, 0.f, 0.f, 0.f, 0.f};
    if (p3Dc.at<float>(2) < 0.0f) return r;
    const float invz = floatReciprocal ? 1 / p3Dc.at<float>(2) : (float)(1.0 / p3Dc.at<float>(2));
    const float x = p3Dc.at<float>(0) * invz, y = p3Dc.at<float>(1) * invz;
    r.u = kf->fx * x + kf->cx;
    r.v = kf->fy * y + kf->cy;
    r.invz = invz;
    if (!kf->IsInImage(r.u, r.v)) return r;
    r.dist = cv::norm(PO);
    if (r.dist < pMP->GetMinDistanceInvariance() || r.dist > pMP->GetMaxDistanceInvariance()) return r;
    if (angleGate && PO.dot(pMP->GetNormal()) < 0.5 * r.dist) return r;
    (void)X;
    r.ok = true;
    return r;
}

}  // namespace orbb_detail

// The device copy of a Frame / KeyFrame: from this handle's cache, else uploaded (mvKeysUn, mDescriptors, grid) and cached.
template <class F> inline orbb_detail::ViewRef ORBmatcher::view_of(const F& f) {
    const orbb_detail::FrameCache::Key key = {orbb_detail::frame_kind(f), (unsigned long)f.mnId, (int)f.mvKeysUn.size()};
    orbb_detail::ViewRef r;
    r.p = slot_->cache.find(key);
    if (!r.p) {
        r.p = std::make_shared<FrameView>(h_, orbb_detail::keys_of(f.mvKeysUn).data(), orbb_detail::rows_of(f.mDescriptors).data(),
                                          (int)f.mvKeysUn.size(), f.mnMinX, f.mnMinY, f.mnMaxX, f.mnMaxY);
        slot_->cache.insert(key, r.p);
    }
    return r;
}

inline void ORBmatcher::RegisterFrame(const Frame& F, ORBextractor& extractor, int frameIndex, int device) {
    const orb_keypoint* dKeys = nullptr;
    const unsigned char* dDesc = nullptr;
    const int* dCount = nullptr;
    int capacity = 0;
    void* stream = nullptr;
    check(orbx_last_device_outputs(extractor.handle(), frameIndex, &dKeys, &dDesc, &dCount, &capacity, &stream));
    orb_camera cam = orb_camera();
    cam.fx = F.fx; cam.fy = F.fy; cam.cx = F.cx; cam.cy = F.cy;
    const int nd = F.mDistCoef.rows * F.mDistCoef.cols;
    if (nd > 0) cam.k1 = F.mDistCoef.at<float>(0);
    if (nd > 1) cam.k2 = F.mDistCoef.at<float>(1);
    if (nd > 2) cam.p1 = F.mDistCoef.at<float>(2);
    if (nd > 3) cam.p2 = F.mDistCoef.at<float>(3);
    if (nd > 4) cam.k3 = F.mDistCoef.at<float>(4);
    orbb_detail::MatcherSlot* slot = orbb_detail::HandlePool::get().acquire(device);
    orbm_frame f = nullptr;
    const int st = orbm_frame_create_device(slot->h, dKeys, dDesc, dCount, capacity, &cam, F.mnMinX, F.mnMinY, F.mnMaxX, F.mnMaxY,
                                            stream, &f);
    if (st == ORB_OK) {
        int n = 0;
        orbm_frame_size(f, &n);
        const orbb_detail::FrameCache::Key key = {0, (unsigned long)F.mnId, n};
        slot->cache.insert(key, std::make_shared<FrameView>(f, n));
    }
    orbb_detail::HandlePool::get().release(slot);
    check(st);
}

// ORBmatcher.cc:1675-1691. One pair is host work; batches go through DescriptorDistances().
inline int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
    const unsigned int* pa = a.ptr<unsigned int>();
    const unsigned int* pb = b.ptr<unsigned int>();
    int dist = 0;
    for (int i = 0; i < 8; ++i) dist += __builtin_popcount(pa[i] ^ pb[i]);
    return dist;
}

// ORBmatcher.cc:405-520
inline int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched,
                                               std::vector<int>& vnMatches12, int windowSize) {
    orbb_detail::ViewRef v1 = view_of(F1), v2 = view_of(F2);
    std::vector<float> prev(vbPrevMatched.size() * 2);
    for (size_t i = 0; i < vbPrevMatched.size(); ++i) { prev[2 * i] = vbPrevMatched[i].x; prev[2 * i + 1] = vbPrevMatched[i].y; }
    const int n = SearchForInitialization(v1, v2, prev, vnMatches12, windowSize);
    for (size_t i = 0; i < vbPrevMatched.size(); ++i) vbPrevMatched[i] = cv::Point2f(prev[2 * i], prev[2 * i + 1]);   // :515-517
    return n;
}

// ORBmatcher.cc:1341-1498. The per-point projection (:1376-1393) runs on the device (orbm_search_by_projection_world, the
// arithmetic of cv::gemm's CV_32F 3x3 path restated); only the once-per-call forward / backward decision (:1352-1364) is
// cv::Mat arithmetic on the host.
inline int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono) {
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat twc = -Rcw.t() * tcw;
    const cv::Mat Rlw = LastFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tlw = LastFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat tlc = Rlw * twc + tlw;
    const bool bForward = tlc.at<float>(2) > CurrentFrame.mb && !bMono;
    const bool bBackward = -tlc.at<float>(2) > CurrentFrame.mb && !bMono;

    orbm_pose pose;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) pose.Rcw[3 * r + c] = Rcw.at<float>(r, c);
        pose.tcw[r] = tcw.at<float>(r);
    }
    pose.fx = CurrentFrame.fx; pose.fy = CurrentFrame.fy; pose.cx = CurrentFrame.cx; pose.cy = CurrentFrame.cy;
    std::vector<orbm_world_query> q(LastFrame.N);
    std::vector<unsigned char> qdesc((size_t)LastFrame.N * 32, 0);
    for (int i = 0; i < LastFrame.N; ++i) {
        orbm_world_query& a = q[i];
        a = orbm_world_query();
        MapPoint* pMP = LastFrame.mvpMapPoints[i];
        if (!pMP || LastFrame.mvbOutlier[i]) continue;
        const cv::Mat x3Dw = pMP->GetWorldPos();
        a.x = x3Dw.at<float>(0); a.y = x3Dw.at<float>(1); a.z = x3Dw.at<float>(2);
        a.octave = LastFrame.mvKeys[i].octave;
        a.valid = 1;
        a.obs_positive = pMP->Observations() > 0;
        a.angle = LastFrame.mvKeysUn[i].angle;
        std::memcpy(&qdesc[(size_t)i * 32], pMP->GetDescriptor().ptr(0), 32);
    }
    std::vector<unsigned char> occupied(CurrentFrame.N, 0);
    for (int i = 0; i < CurrentFrame.N; ++i)
        occupied[i] = CurrentFrame.mvpMapPoints[i] && CurrentFrame.mvpMapPoints[i]->Observations() > 0;
    orbb_detail::ViewRef cur = view_of(CurrentFrame);
    std::vector<int> curMatch;
    const int n = SearchByProjection(cur, CurrentFrame.mvScaleFactors, CurrentFrame.mvuRight.data(), CurrentFrame.mbf, pose, q,
                                     qdesc.data(), th, bForward ? 1 : (bBackward ? 2 : 0), occupied.data(), curMatch);
    for (int i2 = 0; i2 < CurrentFrame.N; ++i2)
        if (curMatch[i2] >= 0) CurrentFrame.mvpMapPoints[i2] = LastFrame.mvpMapPoints[curMatch[i2]];   // :1455
    return n;
}

// ORBmatcher.cc:45-129
inline int ORBmatcher::SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th) {
    std::vector<orbm_point_query> q(vpMapPoints.size());
    std::vector<unsigned char> qdesc(vpMapPoints.size() * 32, 0);
    for (size_t i = 0; i < vpMapPoints.size(); ++i) {
        MapPoint* pMP = vpMapPoints[i];
        orbm_point_query& a = q[i];
        a = orbm_point_query();
        if (!pMP->mbTrackInView || pMP->isBad()) continue;
        a.in_view = 1;
        a.proj_x = pMP->mTrackProjX; a.proj_y = pMP->mTrackProjY; a.proj_xr = pMP->mTrackProjXR;
        a.view_cos = pMP->mTrackViewCos;
        a.level = pMP->mnTrackScaleLevel;
        a.obs_positive = pMP->Observations() > 0;
        std::memcpy(&qdesc[i * 32], pMP->GetDescriptor().ptr(0), 32);
    }
    std::vector<unsigned char> occupied(F.N, 0);
    for (int i = 0; i < F.N; ++i) occupied[i] = F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations() > 0;
    orbb_detail::ViewRef v = view_of(F);
    std::vector<int> match;
    const int n = SearchByProjection(v, F.mvScaleFactors, F.mvuRight.data(), q, qdesc.data(), th, occupied.data(), match);
    for (int i = 0; i < F.N; ++i)
        if (match[i] >= 0) F.mvpMapPoints[i] = vpMapPoints[match[i]];   // :121
    return n;
}

// ORBmatcher.cc:657-823
inline int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12,
                                              std::vector<std::pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo) {
    cv::Mat Cw = pKF1->GetCameraCenter();
    cv::Mat C2 = pKF2->GetRotation() * Cw + pKF2->GetTranslation();
    const float invz = 1.0f / C2.at<float>(2);
    const float ex = pKF2->fx * C2.at<float>(0) * invz + pKF2->cx;   // :668-670
    const float ey = pKF2->fy * C2.at<float>(1) * invz + pKF2->cy;
    const FeatureVectorCSR fv1 = orbb_detail::flatten(pKF1->mFeatVec), fv2 = orbb_detail::flatten(pKF2->mFeatVec);
    std::vector<unsigned char> has1(pKF1->N), has2(pKF2->N);
    for (int i = 0; i < pKF1->N; ++i) has1[i] = pKF1->GetMapPoint(i) != NULL;
    for (int i = 0; i < pKF2->N; ++i) has2[i] = pKF2->GetMapPoint(i) != NULL;
    float f12[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) f12[3 * r + c] = F12.at<float>(r, c);
    orbb_detail::ViewRef v1 = view_of(*pKF1), v2 = view_of(*pKF2);
    return SearchForTriangulation(v1, v2, fv1, fv2, has1.data(), has2.data(), pKF1->mvuRight.data(), pKF2->mvuRight.data(), f12,
                                  ex, ey, pKF2->mvScaleFactors, pKF2->mvLevelSigma2, vMatchedPairs, bOnlyStereo);
}


// ORBmatcher.cc:159-288
inline int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches) {
    const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    vpMapPointMatches = std::vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));
    std::vector<unsigned char> valid1(vpMapPointsKF.size());
    for (size_t i = 0; i < vpMapPointsKF.size(); ++i) valid1[i] = vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad();
    orbb_detail::ViewRef v1 = view_of(*pKF), v2 = view_of(F);
    std::vector<int> m12, m21;
    const int n = SearchByBoW(v1, v2, orbb_detail::flatten(pKF->mFeatVec), orbb_detail::flatten(F.mFeatVec), valid1.data(), NULL,
                              false, m12, m21);
    for (int i2 = 0; i2 < F.N; ++i2)
        if (m21[i2] >= 0) vpMapPointMatches[i2] = vpMapPointsKF[m21[i2]];   // :229-233
    return n;
}

// ORBmatcher.cc:522-655
inline int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12) {
    const std::vector<MapPoint*> vp1 = pKF1->GetMapPointMatches(), vp2 = pKF2->GetMapPointMatches();
    vpMatches12 = std::vector<MapPoint*>(vp1.size(), static_cast<MapPoint*>(NULL));
    std::vector<unsigned char> valid1(vp1.size()), valid2(vp2.size());
    for (size_t i = 0; i < vp1.size(); ++i) valid1[i] = vp1[i] && !vp1[i]->isBad();
    for (size_t i = 0; i < vp2.size(); ++i) valid2[i] = vp2[i] && !vp2[i]->isBad();
    orbb_detail::ViewRef v1 = view_of(*pKF1), v2 = view_of(*pKF2);
    std::vector<int> m12, m21;
    const int n = SearchByBoW(v1, v2, orbb_detail::flatten(pKF1->mFeatVec), orbb_detail::flatten(pKF2->mFeatVec), valid1.data(),
                              valid2.data(), true, m12, m21);
    for (size_t i1 = 0; i1 < vp1.size(); ++i1)
        if (m12[i1] >= 0) vpMatches12[i1] = vp2[m12[i1]];   // :600-604
    return n;
}

// ORBmatcher.cc:1500-1627 (relocalisation). The level comes from MapPoint::PredictScale, the acceptance bound is ORBdist.
inline int ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound,
                                          const float th, const int ORBdist) {
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat Ow = -Rcw.t() * tcw;
    const std::vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();
    std::vector<orbm_proj_query> q(vpMPs.size());
    std::vector<unsigned char> qdesc(vpMPs.size() * 32, 0);
    for (size_t i = 0; i < vpMPs.size(); ++i) {
        orbm_proj_query& a = q[i];
        a = orbm_proj_query();
        MapPoint* pMP = vpMPs[i];
        if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;
        cv::Mat x3Dw = pMP->GetWorldPos();
        cv::Mat x3Dc = Rcw * x3Dw + tcw;
        const float xc = x3Dc.at<float>(0), yc = x3Dc.at<float>(1);
        const float invzc = 1.0 / x3Dc.at<float>(2);
        a.u = CurrentFrame.fx * xc * invzc + CurrentFrame.cx;     // the image-bounds test (:1535-1538) runs with the query
        a.v = CurrentFrame.fy * yc * invzc + CurrentFrame.cy;
        cv::Mat PO = x3Dw - Ow;
        const float dist3D = cv::norm(PO);
        if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
        a.invz = 1.f;                                             // no stereo test in this overload
        a.octave = pMP->PredictScale(dist3D, &CurrentFrame);
        a.valid = 1;
        a.obs_positive = 1;                                       // any assigned keypoint is taken afterwards (:1566)
        a.angle = pKF->mvKeysUn[i].angle;
        std::memcpy(&qdesc[i * 32], pMP->GetDescriptor().ptr(0), 32);
    }
    std::vector<unsigned char> occupied(CurrentFrame.N, 0);
    for (int i = 0; i < CurrentFrame.N; ++i) occupied[i] = CurrentFrame.mvpMapPoints[i] != NULL;
    orbb_detail::ViewRef cur = view_of(CurrentFrame);
    std::vector<int> curMatch;
    const int n = SearchByProjection(cur, CurrentFrame.mvScaleFactors, q, qdesc.data(), th, 0, ORBdist, occupied.data(), curMatch);
    for (int i2 = 0; i2 < CurrentFrame.N; ++i2)
        if (curMatch[i2] >= 0) CurrentFrame.mvpMapPoints[i2] = vpMPs[curMatch[i2]];   // :1585
    return n;
}

// ORBmatcher.cc:290-403 (loop closing): levels [predicted-1, predicted], bound TH_LOW, no orientation check.
inline int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints,
                                          std::vector<MapPoint*>& vpMatched, int th) {
    cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
    const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
    cv::Mat Rcw = sRcw / scw;
    cv::Mat tcw = Scw.rowRange(0, 3).col(3) / scw;
    cv::Mat Ow = -Rcw.t() * tcw;
    std::set<MapPoint*> found(vpMatched.begin(), vpMatched.end());
    found.erase(static_cast<MapPoint*>(NULL));
    std::vector<orbm_proj_query> q(vpPoints.size());
    std::vector<unsigned char> qdesc(vpPoints.size() * 32, 0);
    for (size_t i = 0; i < vpPoints.size(); ++i) {
        orbm_proj_query& a = q[i];
        a = orbm_proj_query();
        MapPoint* pMP = vpPoints[i];
        if (pMP->isBad() || found.count(pMP)) continue;
        cv::Mat p3Dw = pMP->GetWorldPos();
        cv::Mat p3Dc = Rcw * p3Dw + tcw;
        const orbb_detail::Projected pr = orbb_detail::project_into(pMP, p3Dw, p3Dc, p3Dw - Ow, pKF, true, true);
        if (!pr.ok) continue;
        a.u = pr.u; a.v = pr.v; a.invz = 1.f;
        a.octave = pMP->PredictScale(pr.dist, pKF);
        a.valid = 1;
        a.obs_positive = 1;
        std::memcpy(&qdesc[i * 32], pMP->GetDescriptor().ptr(0), 32);
    }
    std::vector<unsigned char> occupied(pKF->N, 0);
    for (int i = 0; i < pKF->N; ++i) occupied[i] = vpMatched[i] != NULL;
    orbb_detail::ViewRef kf = view_of(*pKF);
    std::vector<int> match(pKF->N, -1);
    int n = 0;
    check(orbm_search_by_projection_ex(h_, kf.p->get(), pKF->mvScaleFactors.data(), (int)pKF->mvScaleFactors.size(), NULL, 0.f,
                                       q.data(), qdesc.data(), (int)q.size(), (float)th, 3, TH_LOW, occupied.data(), match.data(),
                                       0, &n));
    for (int i = 0; i < pKF->N; ++i)
        if (match[i] >= 0) vpMatched[i] = vpPoints[match[i]];   // :394-398
    return n;
}

// ORBmatcher.cc:825-975. The projected search runs on the GPU; what a hit does to the map stays here, in point order.
inline int ORBmatcher::Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th) {
    cv::Mat Rcw = pKF->GetRotation(), tcw = pKF->GetTranslation(), Ow = pKF->GetCameraCenter();
    std::vector<orbm_best_query> q(vpMapPoints.size());
    std::vector<unsigned char> qdesc(vpMapPoints.size() * 32, 0);
    for (size_t i = 0; i < vpMapPoints.size(); ++i) {
        orbm_best_query& a = q[i];
        a = orbm_best_query();
        MapPoint* pMP = vpMapPoints[i];
        if (!pMP || pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
        cv::Mat p3Dw = pMP->GetWorldPos();
        const orbb_detail::Projected pr = orbb_detail::project_into(pMP, p3Dw, Rcw * p3Dw + tcw, p3Dw - Ow, pKF, true, true);
        if (!pr.ok) continue;
        a.u = pr.u; a.v = pr.v;
        a.ur = pr.u - pKF->mbf * pr.invz;
        a.level = pMP->PredictScale(pr.dist, pKF);
        a.radius = th * pKF->mvScaleFactors[a.level];
        a.valid = 1;
        std::memcpy(&qdesc[i * 32], pMP->GetDescriptor().ptr(0), 32);
    }
    orbb_detail::ViewRef kf = view_of(*pKF);
    std::vector<int> bestIdx, bestDist;
    ProjectedBest(kf, q, qdesc.data(), true, pKF->mvuRight.data(), pKF->mvInvLevelSigma2, bestIdx, bestDist);
    int nFused = 0;
    for (size_t i = 0; i < vpMapPoints.size(); ++i) {
        if (!q[i].valid || bestIdx[i] < 0 || bestDist[i] > TH_LOW) continue;
        MapPoint* pMP = vpMapPoints[i];
        // the reference tests these at the top of every iteration (:849), i.e. AFTER the Replace / AddMapPoint of earlier
        // iterations: a repeated pointer, or a point an earlier Replace made bad, is skipped
        if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
        MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx[i]);
        if (pMPinKF) {                                            // :949-958
            if (!pMPinKF->isBad()) {
                if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                else pMPinKF->Replace(pMP);
            }
        } else {
            pMP->AddObservation(pKF, bestIdx[i]);
            pKF->AddMapPoint(pMP, bestIdx[i]);
        }
        ++nFused;
    }
    return nFused;
}

// ORBmatcher.cc:977-1099
inline int ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th,
                            std::vector<MapPoint*>& vpReplacePoint) {
    cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
    const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
    cv::Mat Rcw = sRcw / scw;
    cv::Mat tcw = Scw.rowRange(0, 3).col(3) / scw;
    cv::Mat Ow = -Rcw.t() * tcw;
    const std::set<MapPoint*> found = pKF->GetMapPoints();
    std::vector<orbm_best_query> q(vpPoints.size());
    std::vector<unsigned char> qdesc(vpPoints.size() * 32, 0);
    for (size_t i = 0; i < vpPoints.size(); ++i) {
        orbm_best_query& a = q[i];
        a = orbm_best_query();
        MapPoint* pMP = vpPoints[i];
        if (pMP->isBad() || found.count(pMP)) continue;
        cv::Mat p3Dw = pMP->GetWorldPos();
        const orbb_detail::Projected pr = orbb_detail::project_into(pMP, p3Dw, Rcw * p3Dw + tcw, p3Dw - Ow, pKF, true, false);
        if (!pr.ok) continue;
        a.u = pr.u; a.v = pr.v; a.ur = -1.f;
        a.level = pMP->PredictScale(pr.dist, pKF);
        a.radius = th * pKF->mvScaleFactors[a.level];
        a.valid = 1;
        std::memcpy(&qdesc[i * 32], pMP->GetDescriptor().ptr(0), 32);
    }
    orbb_detail::ViewRef kf = view_of(*pKF);
    std::vector<int> bestIdx, bestDist;
    ProjectedBest(kf, q, qdesc.data(), false, NULL, pKF->mvInvLevelSigma2, bestIdx, bestDist);
    int nFused = 0;
    for (size_t i = 0; i < vpPoints.size(); ++i) {
        if (!q[i].valid || bestIdx[i] < 0 || bestDist[i] > TH_LOW) continue;
        if (vpPoints[i]->isBad()) continue;                       // re-evaluated per iteration as the reference does (:1005)
        MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx[i]);
        if (pMPinKF) {                                            // :1082-1086
            if (!pMPinKF->isBad()) vpReplacePoint[i] = pMPinKF;
        } else {
            vpPoints[i]->AddObservation(pKF, bestIdx[i]);
            pKF->AddMapPoint(vpPoints[i], bestIdx[i]);
        }
        ++nFused;
    }
    return nFused;
}

// ORBmatcher.cc:1102-1321: two projected searches (GPU) and the mutual-agreement pass (host).
inline int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12,
                                    const cv::Mat& R12, const cv::Mat& t12, const float th) {
    cv::Mat R1w = pKF1->GetRotation(), t1w = pKF1->GetTranslation();
    cv::Mat R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
    cv::Mat sR12 = s12 * R12;
    cv::Mat sR21 = (1.0 / s12) * R12.t();
    cv::Mat t21 = -sR21 * t12;
    const std::vector<MapPoint*> vp1 = pKF1->GetMapPointMatches(), vp2 = pKF2->GetMapPointMatches();
    const int N1 = (int)vp1.size(), N2 = (int)vp2.size();
    std::vector<bool> done1(N1, false), done2(N2, false);
    for (int i = 0; i < N1; ++i) {
        MapPoint* pMP = vpMatches12[i];
        if (!pMP) continue;
        done1[i] = true;
        const int idx2 = pMP->GetIndexInKeyFrame(pKF2);
        if (idx2 >= 0 && idx2 < N2) done2[idx2] = true;
    }
    // points of `from` carried into `into`: camera coordinates there are A * (Rw * X + tw) + b
    struct Side {
        static void queries(const std::vector<MapPoint*>& vp, const std::vector<bool>& done, const cv::Mat& Rw, const cv::Mat& tw,
                            const cv::Mat& A, const cv::Mat& b, KeyFrame* into, float th, std::vector<orbm_best_query>& q,
                            std::vector<unsigned char>& qdesc) {
            q.assign(vp.size(), orbm_best_query());
            qdesc.assign(vp.size() * 32, 0);
            for (size_t i = 0; i < vp.size(); ++i) {
                MapPoint* pMP = vp[i];
                if (!pMP || done[i] || pMP->isBad()) continue;
                cv::Mat p3Dw = pMP->GetWorldPos();
                cv::Mat pOther = A * (Rw * p3Dw + tw) + b;
                const orbb_detail::Projected pr = orbb_detail::project_into(pMP, p3Dw, pOther, pOther, into, false, false);
                if (!pr.ok) continue;
                orbm_best_query& a = q[i];
                a.u = pr.u; a.v = pr.v; a.ur = -1.f;
                a.level = pMP->PredictScale(pr.dist, into);
                a.radius = th * into->mvScaleFactors[a.level];
                a.valid = 1;
                std::memcpy(&qdesc[i * 32], pMP->GetDescriptor().ptr(0), 32);
            }
        }
    };
    std::vector<orbm_best_query> q1, q2;
    std::vector<unsigned char> d1, d2;
    Side::queries(vp1, done1, R1w, t1w, sR21, t21, pKF2, th, q1, d1);   // :1141-1219
    Side::queries(vp2, done2, R2w, t2w, sR12, t12, pKF1, th, q2, d2);   // :1222-1299
    orbb_detail::ViewRef v1 = view_of(*pKF1), v2 = view_of(*pKF2);
    std::vector<int> b1, dist1, b2, dist2;
    ProjectedBest(v2, q1, d1.data(), false, NULL, pKF2->mvInvLevelSigma2, b1, dist1);
    ProjectedBest(v1, q2, d2.data(), false, NULL, pKF1->mvInvLevelSigma2, b2, dist2);
    int nFound = 0;
    for (int i1 = 0; i1 < N1; ++i1) {                              // :1302-1318
        if (!q1[i1].valid || b1[i1] < 0 || dist1[i1] > TH_HIGH) continue;
        const int idx2 = b1[i1];
        if (q2[idx2].valid && b2[idx2] == i1 && dist2[idx2] <= TH_HIGH) {
            vpMatches12[i1] = vp2[idx2];
            ++nFound;
        }
    }
    return nFound;
}

}  // namespace ORB_SLAM2

#endif  // ORBB200_WITH_ORBSLAM
