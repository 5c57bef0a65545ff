// Definitions of the reference-signature overloads of ORB_SLAM2::ORBmatcher (declared in adapter/ORBmatcher.h under
// ORBB200_WITH_ORBSLAM).  Include this file from ONE translation unit of the host project after Frame.h, KeyFrame.h and
// MapPoint.h.  It is NOT compiled in this repository (those headers need OpenCV, Eigen, DBoW2 and g2o, none of which is
// installed here); each function states which reference lines it stands in for, and INTEGRATION.md section 3 lists what
// is read from and written back to the objects.  Everything between "flatten" and "write back" runs on the GPU.
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
template <class F> inline FrameView view_of(orbm_handle h, const F& f) {
    return FrameView(h, keys_of(f.mvKeysUn).data(), rows_of(f.mDescriptors).data(), (int)f.mvKeysUn.size(), f.mnMinX, f.mnMinY,
                     f.mnMaxX, f.mnMaxY);
}

}  // namespace orbb_detail

// ORBmatcher.cc:405-520
inline int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched,
                                               std::vector<int>& vnMatches12, int windowSize) {
    FrameView v1 = orbb_detail::view_of(h_, F1), v2 = orbb_detail::view_of(h_, F2);
    std::vector<float> prev(vbPrevMatched.size() * 2);
    for (size_t i = 0; i < vbPrevMatched.size(); ++i) { prev[2 * i] = vbPrevMatched[i].x; prev[2 * i + 1] = vbPrevMatched[i].y; }
    const int n = SearchForInitialization(v1, v2, prev, vnMatches12, windowSize);
    for (size_t i = 0; i < vbPrevMatched.size(); ++i) vbPrevMatched[i] = cv::Point2f(prev[2 * i], prev[2 * i + 1]);   // :515-517
    return n;
}

// ORBmatcher.cc:1341-1498. The projection (:1352-1388) is cv::Mat arithmetic on the host, as in the reference.
inline int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono) {
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat twc = -Rcw.t() * tcw;
    const cv::Mat Rlw = LastFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tlw = LastFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat tlc = Rlw * twc + tlw;
    const bool bForward = tlc.at<float>(2) > CurrentFrame.mb && !bMono;
    const bool bBackward = -tlc.at<float>(2) > CurrentFrame.mb && !bMono;

    std::vector<orbm_proj_query> q(LastFrame.N);
    std::vector<unsigned char> qdesc((size_t)LastFrame.N * 32, 0);
    for (int i = 0; i < LastFrame.N; ++i) {
        orbm_proj_query& a = q[i];
        a = orbm_proj_query();
        MapPoint* pMP = LastFrame.mvpMapPoints[i];
        if (!pMP || LastFrame.mvbOutlier[i]) continue;
        cv::Mat x3Dc = Rcw * pMP->GetWorldPos() + tcw;
        const float xc = x3Dc.at<float>(0), yc = x3Dc.at<float>(1);
        const float invzc = 1.0 / x3Dc.at<float>(2);
        if (invzc < 0) continue;
        a.u = CurrentFrame.fx * xc * invzc + CurrentFrame.cx;
        a.v = CurrentFrame.fy * yc * invzc + CurrentFrame.cy;
        a.invz = invzc;
        a.octave = LastFrame.mvKeys[i].octave;
        a.valid = 1;
        a.obs_positive = pMP->Observations() > 0;
        a.angle = LastFrame.mvKeysUn[i].angle;
        std::memcpy(&qdesc[(size_t)i * 32], pMP->GetDescriptor().ptr(0), 32);
    }
    std::vector<unsigned char> occupied(CurrentFrame.N, 0);
    for (int i = 0; i < CurrentFrame.N; ++i)
        occupied[i] = CurrentFrame.mvpMapPoints[i] && CurrentFrame.mvpMapPoints[i]->Observations() > 0;
    FrameView cur = orbb_detail::view_of(h_, CurrentFrame);
    std::vector<int> curMatch;
    const int n = SearchByProjection(cur, CurrentFrame.mvScaleFactors, CurrentFrame.mvuRight.data(), CurrentFrame.mbf, q,
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
    FrameView v = orbb_detail::view_of(h_, F);
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
    auto flatten = [](const DBoW2::FeatureVector& fv) {
        FeatureVectorCSR c;
        c.start.push_back(0);
        for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {   // std::map: ascending node ids
            c.nodeId.push_back((int)it->first);
            for (size_t k = 0; k < it->second.size(); ++k) c.idx.push_back((int)it->second[k]);
            c.start.push_back((int)c.idx.size());
        }
        return c;
    };
    const FeatureVectorCSR fv1 = flatten(pKF1->mFeatVec), fv2 = flatten(pKF2->mFeatVec);
    std::vector<unsigned char> has1(pKF1->N), has2(pKF2->N);
    for (int i = 0; i < pKF1->N; ++i) has1[i] = pKF1->GetMapPoint(i) != NULL;
    for (int i = 0; i < pKF2->N; ++i) has2[i] = pKF2->GetMapPoint(i) != NULL;
    float f12[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) f12[3 * r + c] = F12.at<float>(r, c);
    FrameView v1 = orbb_detail::view_of(h_, *pKF1), v2 = orbb_detail::view_of(h_, *pKF2);
    return SearchForTriangulation(v1, v2, fv1, fv2, has1.data(), has2.data(), pKF1->mvuRight.data(), pKF2->mvuRight.data(), f12,
                                  ex, ey, pKF2->mvScaleFactors, pKF2->mvLevelSigma2, vMatchedPairs, bOnlyStereo);
}

}  // namespace ORB_SLAM2

#endif  // ORBB200_WITH_ORBSLAM
