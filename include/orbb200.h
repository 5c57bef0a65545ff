/* orbb200 -- B200-native (sm_100a) ORB feature extraction and binary-descriptor Hamming matching.
 *
 * C ABI of the drop-in replacement for the data-parallel front-end of hwb0314/VI-ORB-SLAM-ICRA2018:
 *   ORB_SLAM2::ORBextractor            /root/reference/include/ORBextractor.h:44-131, src/ORBextractor.cc
 *   ORB_SLAM2::ORBmatcher (searches)   /root/reference/include/ORBmatcher.h:41-83,    src/ORBmatcher.cc
 *   Frame grid (candidate lists)       /root/reference/src/Frame.cc:574-589, 671-736
 *
 * Plain pointers and sizes only; no C++ or torch types cross this boundary.  Every entry point returns an
 * orb_status; orb_last_error() gives the message of the last failure on the calling thread.  All compute runs in
 * CUDA kernels: there is no CPU fallback, and a missing/failed device is an error, never a silent slow path.
 * "host" entry points take host pointers and are synchronous; "_device" entry points take device pointers, enqueue on
 * the handle's stream (or the one given) and return without synchronising.
 *
 * The C++ adapter with the reference's own class names and signatures is adapter/ORBextractor.h, adapter/ORBmatcher.h.
 */
#ifndef ORBB200_H
#define ORBB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBB200_VERSION 100

typedef enum {
    ORB_OK = 0,
    ORB_ERR_INVALID = 1,      /* bad argument (null pointer, size out of range, geometry the reference cannot handle) */
    ORB_ERR_CAPACITY = 2,     /* an output or workspace bound given at create time would be exceeded */
    ORB_ERR_CUDA = 3,         /* a CUDA call failed; see orb_last_error() */
    ORB_ERR_UNSUPPORTED = 4
} orb_status;

/* cv::KeyPoint, byte for byte (28 B): what ORBextractor::operator() fills (ORBextractor.cc:839-849, 1112-1121) */
typedef struct {
    float x, y;        /* pt, level-0 pixel coordinates */
    float size;        /* (int)(31 * mvScaleFactor[octave]) */
    float angle;       /* IC_Angle, degrees in [0, 360) */
    float response;    /* FAST score */
    int32_t octave;
    int32_t class_id;  /* -1 */
} orb_keypoint;

const char* orb_last_error(void);
int orb_version(void);
/* number of visible CUDA devices, or a negative value when the runtime cannot be initialised */
int orb_device_count(void);

/* ================================================================================================ extractor */
typedef struct orbx_extractor* orbx_handle;

/* ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)   (ORBextractor.cc:412-472)
 * max_width/max_height/max_batch size the device arena once; extract calls never allocate.              */
int orbx_create(int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast,
                int max_width, int max_height, int max_batch, int device, orbx_handle* out);
int orbx_destroy(orbx_handle h);

/* Upper bound on keypoints per frame (sum over levels of max(N_l + 2, 4*nIni)); size outputs with it. */
int orbx_keypoint_capacity(orbx_handle h, int* cap);

/* ORBextractor::operator()(image, mask, keypoints, descriptors)   (ORBextractor.cc:1045-1126; mask is ignored there)
 * image: 8-bit grayscale, `stride` bytes per row. keypoints[capacity], descriptors[capacity*32]; *n_out = count.
 * An empty image (width or height 0, or null pointer) returns ORB_OK with *n_out = 0 (ORBextractor.cc:1048). */
int orbx_extract(orbx_handle h, const uint8_t* image, int width, int height, int stride,
                 orb_keypoint* keypoints, uint8_t* descriptors, int capacity, int* n_out);

/* The same operator over n_frames independent frames of one size (frame f at images + f*frame_stride).
 * Outputs for frame f start at keypoints + f*capacity and descriptors + f*capacity*32; n_out[f] = count.
 * Host buffers (pinned for full speed): the batch is cut into pipeline chunks of growing size so that the H2D copy of
 * one chunk, the kernels of the previous ones (two compute streams, alternating) and the D2H copy of the one before
 * overlap; towards the end of the batch the chunks shrink again so that little is left to do after the last copy.
 * Environment variables read as tuning / debugging aids: ORBB_PIPE_CHUNK = size of the first chunk in frames,
 * ORBB_PIPE_TAIL = smallest chunk at the end, ORBB_PIPE_STREAMS = 1 or 2, ORBB_PIPE_TRACE = print the per-chunk timeline. */
int orbx_extract_batch(orbx_handle h, const uint8_t* images, int n_frames, int width, int height, int stride,
                       size_t frame_stride, orb_keypoint* keypoints, uint8_t* descriptors, int capacity, int* n_out);

/* Device-resident variant: all pointers are device pointers, work is enqueued on `stream` (a cudaStream_t, NULL =
 * the handle's own stream) without synchronising; n_frames <= max_batch.                                           */
int orbx_extract_batch_device(orbx_handle h, const uint8_t* d_images, int n_frames, int width, int height, int stride,
                              size_t frame_stride, orb_keypoint* d_keypoints, uint8_t* d_descriptors, int capacity,
                              int* d_n_out, void* stream);
int orbx_synchronize(orbx_handle h);

/* Getters of ORBextractor.h:81-101 (+ mnFeaturesPerLevel and umax for tests). Arrays hold nlevels (umax: 16) values. */
int orbx_get_levels(orbx_handle h, int* nlevels);
int orbx_get_scale_tables(orbx_handle h, float* scale_factors, float* inv_scale_factors, float* level_sigma2,
                          float* inv_level_sigma2, int* features_per_level, int* umax);
/* mvImagePyramid[level] of frame `frame` of the last call (ORBextractor.h:103): the padded (w+38)x(h+38) buffer is
 * copied to `out` (may be NULL to query the size); *w,*h are the level size without the 19-px frame.           */
int orbx_get_level(orbx_handle h, int frame, int level, uint8_t* out, int* w, int* hgt);
/* GetTimeOfComputePyramid / ...KeyPointsOctTree / ...Descriptor (ORBextractor.h:51-53), milliseconds measured with CUDA
 * events on the handle's stream.  Host calls of up to 4 frames replay a captured CUDA graph from their second call of an
 * image size on (events cannot be read out of a graph): they leave the times of the last un-captured call in place;
 * ORBB_NO_GRAPH=1 in the environment keeps every call un-captured.                                                */
int orbx_stage_times(orbx_handle h, double* ms3);

/* Frame::ComputeStereoMatches (Frame.cc:810-984), the one consumer of the padded mvImagePyramid: for every left keypoint
 * the closest right descriptor on its row band, refined by the 11x11 SAD search and the parabola fit on the pyramid
 * level of the left keypoint, then the 1.5*1.4*median SAD filter.  The pyramids are read in place from frame
 * `left_frame` / `right_frame` of the LAST extract call of the two handles (same device, same image size and ctor
 * arguments; the host call waits for the right handle's stream).  Only that call's frames are resident: an index
 * past its frame count is ORB_ERR_INVALID, and of a call with more frames than max_batch only the last max_batch-sized
 * part is (frame 0 = the first frame of that part).  keys_* are mvKeys / mvKeysRight (not undistorted),
 * mb = baseline in metres, mbf = baseline * fx.  u_right[n_left], depth[n_left]: mvuRight, mvDepth (-1 = no match).
 * Where the reference has undefined behaviour the result is defined: a search window that leaves the level image is
 * no match (cv::Mat::colRange would assert), and with no match at all the median filter is skipped.                */
int orbx_compute_stereo_matches(orbx_handle left, int left_frame, orbx_handle right, int right_frame,
                                const orb_keypoint* keys_left, const uint8_t* desc_left, int n_left,
                                const orb_keypoint* keys_right, const uint8_t* desc_right, int n_right,
                                float mb, float mbf, float* u_right, float* depth, int* n_matches);
/* Device-resident variant: keypoints, descriptors and counts where orbx_extract_batch_device left them (d_n_* may be
 * NULL: then cap_* is the count); outputs are device arrays of cap_left entries (d_sad: SAD of each kept match or -1),
 * *d_kept the number of matches.  Enqueues on `stream` (NULL = the left handle's stream) without synchronising; the
 * caller orders it after both extractions.                                                                         */
int orbx_compute_stereo_matches_device(orbx_handle left, int left_frame, orbx_handle right, int right_frame,
                                       const orb_keypoint* d_keys_left, const uint8_t* d_desc_left, const int* d_n_left,
                                       int cap_left, const orb_keypoint* d_keys_right, const uint8_t* d_desc_right,
                                       const int* d_n_right, int cap_right, float mb, float mbf, float* d_u_right,
                                       float* d_depth, int* d_sad, int* d_kept, void* stream);

/* Introspection for stage-by-stage parity tests (not part of the reference surface). */
int orbx_debug_candidates(orbx_handle h, int frame, int level, orb_keypoint* out, int cap, int* n);
int orbx_debug_blurred(orbx_handle h, int frame, int level, uint8_t* out /* w*h */);
/* number of kernel launches issued by the last extract call */
/* Device copies of what the last HOST extract call returned (orbx_extract / orbx_extract_batch): keypoints, descriptors
 * and count of frame `frame`, valid until the next call on the handle; *stream = the handle's stream (already synchronised
 * when the host call returned).  With orbm_frame_create_device the Frame constructor's remaining work (UndistortKeyPoints,
 * AssignFeaturesToGrid, Frame.cc:75-109) runs on these without a second upload.                                     */
int orbx_last_device_outputs(orbx_handle h, int frame, const orb_keypoint** d_keys, const uint8_t** d_descriptors,
                             const int** d_count, int* capacity, void** stream);
int orbx_last_launch_count(orbx_handle h, int* n);
/* Per-kernel device timing for benchmarks: when on, every extract call brackets its kernels with CUDA events on the
 * launching stream (no synchronisation added).  orbx_kernel_times synchronises, returns the milliseconds summed over
 * the calls since the last query for [0] pyramid (all levels) [1] FAST cells [2] quadtree+orientation [3] blur
 * [4] BRIEF+assembly, the number of calls in *n_calls, and resets the accumulation.                              */
int orbx_set_profiling(orbx_handle h, int on);
int orbx_kernel_times(orbx_handle h, double* ms5, int* n_calls);

/* ================================================================================================== matcher */
typedef struct orbm_matcher* orbm_handle;
typedef struct orbm_frame_s* orbm_frame;

int orbm_create(int device, orbm_handle* out);
int orbm_destroy(orbm_handle h);
int orbm_synchronize(orbm_handle h);
int orbm_last_launch_count(orbm_handle h, int* n);

/* ORBmatcher::DescriptorDistance (ORBmatcher.cc:1675-1691) for n independent pairs a[i], b[i] (32 B each). */
int orbm_distance(orbm_handle h, const uint8_t* a, const uint8_t* b, int n, int* dist_out);

/* What the search loops read of a Frame / KeyFrame: undistorted keypoints (pt, angle, octave), descriptors and the
 * static image bounds mnMinX.. (Frame.cc:782-808).  The 64x48 grid is built on the device exactly as
 * AssignFeaturesToGrid / PosInGrid do (Frame.cc:574-589, 726-736).                                             */
int orbm_frame_create(orbm_handle h, const orb_keypoint* keys_un, const uint8_t* descriptors, int n,
                      float min_x, float min_y, float max_x, float max_y, orbm_frame* out);
int orbm_frame_destroy(orbm_frame f);

/* mK and mDistCoef as Frame holds them (CV_32F; Tracking.cc reads Camera.fx .. Camera.k3 from the settings file). */
typedef struct {
    float fx, fy, cx, cy;
    float k1, k2, p1, p2, k3;   /* k3 = 0 for the four-coefficient model */
} orb_camera;
/* cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK) on n (x, y) pairs, bit for bit as the installed OpenCV
 * computes it (five iterations in double, one rounding to float): what Frame::UndistortKeyPoints (Frame.cc:748-778) and
 * Frame::ComputeImageBounds (Frame.cc:780-808) call.  cam == NULL or k1 == 0 copies the input (Frame.cc:750-754).  */
int orbm_undistort_points(orbm_handle h, const orb_camera* cam, const float* xy, int n, float* xy_out);
/* Frame::ComputeImageBounds: bounds[4] = mnMinX, mnMinY, mnMaxX, mnMaxY of a width x height image. */
int orbm_image_bounds(orbm_handle h, const orb_camera* cam, int width, int height, float* bounds);
/* The Frame constructor's work after ExtractORB (Frame.cc:75-109: UndistortKeyPoints, AssignFeaturesToGrid) for
 * keypoints and descriptors that are still where orbx_extract_batch_device left them: d_keys / d_descriptors / d_count
 * are DEVICE pointers of one frame (capacity entries), producer_stream is the stream that extraction was enqueued on
 * (the call waits for it before reading the count; pass NULL only if the producer has already been synchronised --
 * the extractor's own stream is non-blocking, so the legacy default stream does not order against it).
 * The frame keeps its own undistorted copy; only the 4-byte count crosses PCIe (N is host state of a Frame).      */
int orbm_frame_create_device(orbm_handle h, const orb_keypoint* d_keys, const uint8_t* d_descriptors, const int* d_count,
                             int capacity, const orb_camera* cam, float min_x, float min_y, float max_x, float max_y,
                             void* producer_stream, orbm_frame* out);
/* N, and mvKeysUn / mDescriptors copied back to the host (either pointer may be NULL) */
int orbm_frame_size(orbm_frame f, int* n);
int orbm_frame_download(orbm_frame f, orb_keypoint* keys_un, uint8_t* descriptors);
/* mGrid as CSR: cell id = ix*48+iy; cell_start[3073], cell_idx[n] */
int orbm_frame_grid(orbm_frame f, int* cell_start, int* cell_idx);
/* Frame::GetFeaturesInArea (Frame.cc:671-724) for nq queries (x,y,r triples); min_level = max_level = -1 gives the
 * KeyFrame overload (KeyFrame.cc:1138-1177). idx_out[q*cap ...], count_out[q] (count may exceed cap: truncated). */
int orbm_features_in_area(orbm_frame f, const float* xyr, int nq, int min_level, int max_level, int* idx_out, int cap,
                          int* count_out);

/* ORBmatcher(nnratio, checkOri).SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)
 * (ORBmatcher.cc:405-520). prev_matched_xy: n1 x 2 floats, updated in place; matches12: n1 ints.                  */
int orbm_search_for_initialization(orbm_handle h, orbm_frame f1, orbm_frame f2, float* prev_matched_xy, int* matches12,
                                   int window_size, float nnratio, int check_orientation, int* nmatches);

/* One LastFrame keypoint that carries a map point (ORBmatcher.cc:1365-1410). The host projects (cv::Mat GEMM). */
typedef struct {
    float u, v;            /* projection into the current frame */
    float invz;            /* 1/z in the current camera (stereo check) */
    int32_t octave;        /* LastFrame.mvKeys[i].octave */
    int32_t valid;         /* pMP && !outlier && invz >= 0 */
    int32_t obs_positive;  /* pMP->Observations() > 0 */
    float angle;           /* LastFrame.mvKeysUn[i].angle */
} orbm_proj_query;
/* SearchByProjection(Frame& Current, const Frame& Last, th, bMono)  (ORBmatcher.cc:1341-1498).
 * mode 0: octave window [o-1,o+1]; 1: forward (>= o); 2: backward (0..o).  u_right may be NULL (monocular).
 * occupied[n_cur] != 0: keypoint already holds a map point with observations.  cur_match[n_cur]: query index or -1. */
int orbm_search_by_projection(orbm_handle h, orbm_frame cur, const float* scale_factors, int nlevels,
                              const float* u_right, float mbf, const orbm_proj_query* queries,
                              const uint8_t* query_desc, int nq, float th, int mode, const uint8_t* occupied,
                              int* cur_match, int check_orientation, int* nmatches);

/* The same search with the acceptance threshold as a parameter: SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th,
 * ORBdist) used by relocalisation (ORBmatcher.cc:1500-1627; ORBdist = 100 or 64, Tracking.cc:2677, 2691).  There the host
 * predicts the level (query.octave = nPredictedLevel), mode is 0, and every assigned keypoint counts as occupied
 * (obs_positive = 1, occupied[i2] = CurrentFrame.mvpMapPoints[i2] != NULL).                                          */
int orbm_search_by_projection_ex(orbm_handle h, orbm_frame cur, const float* scale_factors, int nlevels,
                                 const float* u_right, float mbf, const orbm_proj_query* queries,
                                 const uint8_t* query_desc, int nq, float th, int mode, int max_distance,
                                 const uint8_t* occupied, int* cur_match, int check_orientation, int* nmatches);

/* The same search with the projection of ORBmatcher.cc:1376-1393 done on the device: the caller hands over world points
 * and the current pose instead of projected coordinates (no per-point host loop, no cv::Mat temporaries).  The arithmetic
 * is the reference's: x3Dc = Rcw * x3Dw + tcw as OpenCV's CV_32F 3x3 . 3x1 + 3x1 gemm evaluates it (products summed in
 * float in source order, addend joined in double, one rounding), invzc = 1.0 / z in double, u = fx * xc * invzc + cx in
 * float.  mode as above: the caller derives it from tlc = Rlw * twc + tlw against CurrentFrame.mb (:1361-1364).      */
typedef struct {
    float x, y, z;         /* pMP->GetWorldPos() */
    int32_t octave;        /* LastFrame.mvKeys[i].octave */
    int32_t valid;         /* pMP && !LastFrame.mvbOutlier[i] */
    int32_t obs_positive;  /* pMP->Observations() > 0 */
    float angle;           /* LastFrame.mvKeysUn[i].angle */
} orbm_world_query;
typedef struct {
    float Rcw[9];          /* CurrentFrame.mTcw.rowRange(0,3).colRange(0,3), row-major */
    float tcw[3];          /* CurrentFrame.mTcw.rowRange(0,3).col(3) */
    float fx, fy, cx, cy;  /* CurrentFrame.fx .. cy */
} orbm_pose;
int orbm_search_by_projection_world(orbm_handle h, orbm_frame cur, const float* scale_factors, int nlevels,
                                    const float* u_right, float mbf, const orbm_pose* pose,
                                    const orbm_world_query* queries, const uint8_t* query_desc, int nq, float th, int mode,
                                    int max_distance, const uint8_t* occupied, int* cur_match, int check_orientation,
                                    int* nmatches);

/* The projection alone (ORBmatcher.cc:1376-1388): u, v, invz of n world points (xyz: n x 3 floats) in the given pose. */
int orbm_project_points(orbm_handle h, const orbm_pose* pose, const float* xyz, int n, float* u, float* v, float* invz);

/* Batched forms: n independent searches in one call.  Every phase (query construction, candidate enumeration, the
 * order-dependent bookkeeping with one CTA per job) runs once for the whole batch, so the call costs the launches and the
 * one host synchronisation of a single search while the whole GPU works; each job's results are exactly those of the
 * single call with the same arguments.  All frames must belong to matcher h.  *candidates (may be NULL) receives the number
 * of (query, keypoint) pairs whose descriptors were compared.                                                       */
typedef struct {
    orbm_frame cur;                  /* CurrentFrame */
    const orbm_proj_query* queries;  /* host, nq entries */
    const uint8_t* query_desc;       /* host, nq x 32 */
    int nq;
    const float* u_right;            /* host, cur's mvuRight or NULL */
    const uint8_t* occupied;         /* host, cur's occupancy flags or NULL */
    int* cur_match;                  /* host out: cur's n entries */
    int nmatches;                    /* out */
} orbm_projection_job;
int orbm_search_by_projection_batch(orbm_handle h, orbm_projection_job* jobs, int n_jobs, const float* scale_factors,
                                    int nlevels, float mbf, float th, int mode, int max_distance, int check_orientation,
                                    long long* candidates);
typedef struct {
    orbm_frame f1, f2;
    float* prev_xy;                  /* host in/out: vbPrevMatched of f1 */
    int* matches12;                  /* host out: f1's n entries */
    int nmatches;                    /* out */
} orbm_init_job;
int orbm_search_for_initialization_batch(orbm_handle h, orbm_init_job* jobs, int n_jobs, int window_size, float nnratio,
                                         int check_orientation, long long* candidates);

typedef struct {
    float proj_x, proj_y, proj_xr;
    float view_cos;
    int32_t level;         /* mnTrackScaleLevel */
    int32_t in_view;       /* mbTrackInView && !isBad() */
    int32_t obs_positive;
} orbm_point_query;
/* SearchByProjection(Frame& F, const vector<MapPoint*>&, th)  (ORBmatcher.cc:45-129) */
int orbm_search_by_projection_points(orbm_handle h, orbm_frame f, const float* scale_factors, int nlevels,
                                     const float* u_right, const orbm_point_query* queries, const uint8_t* query_desc,
                                     int nq, float th, float nnratio, const uint8_t* occupied, int* match,
                                     int* nmatches);

/* SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo)  (ORBmatcher.cc:657-823, 140-157).
 * Feature vectors (DBoW2 node -> keypoint indices) as node-sorted CSR.  has_point*: keypoint already has a map point.
 * f12: row-major 3x3; (ex,ey): epipole in image 2 (host computes it, ORBmatcher.cc:664-670).
 * matches12[n1] = index in KF2 or -1 (the reference's pair list is the entries >= 0 in ascending order).          */
int orbm_search_for_triangulation(orbm_handle h, orbm_frame kf1, orbm_frame kf2,
                                  int n_nodes1, const int* node_id1, const int* node_start1, const int* node_idx1,
                                  int n_nodes2, const int* node_id2, const int* node_start2, const int* node_idx2,
                                  const uint8_t* has_point1, const uint8_t* has_point2,
                                  const float* u_right1, const float* u_right2,
                                  const float* f12, float ex, float ey,
                                  const float* scale_factors2, const float* level_sigma2_2, int nlevels,
                                  int only_stereo, int check_orientation, int* matches12, int* nmatches);

/* SearchByBoW, both overloads: (KeyFrame*, Frame&, vpMapPointMatches) ORBmatcher.cc:159-288 with strict_low = 0
 * (accept best <= TH_LOW, valid2 = NULL) and (KeyFrame*, KeyFrame*, vpMatches12) ORBmatcher.cc:522-655 with strict_low = 1
 * (accept best < TH_LOW).  valid1 / valid2: the keypoint's map point exists and is not bad.  Feature vectors as in
 * orbm_search_for_triangulation.  matches12[n1] / matches21[n2]: partner index or -1 (the Frame overload's
 * vpMapPointMatches[i2] is pKF's map point at matches21[i2]).                                                       */
int orbm_search_by_bow(orbm_handle h, orbm_frame kf1, orbm_frame f2,
                       int n_nodes1, const int* node_id1, const int* node_start1, const int* node_idx1,
                       int n_nodes2, const int* node_id2, const int* node_start2, const int* node_idx2,
                       const uint8_t* valid1, const uint8_t* valid2, float nnratio, int check_orientation,
                       int strict_low, int* matches12, int* matches21, int* nmatches);

/* The independent projected search inside Fuse (ORBmatcher.cc:892-944, chi2_filter = 1), Fuse with a Sim3 (:1051-1075) and
 * both directions of SearchBySim3 (:1191-1215, :1271-1295) (chi2_filter = 0): per query KeyFrame::GetFeaturesInArea(u, v,
 * radius), candidates with octave in [level-1, level], optionally the reprojection chi-square test against
 * inv_level_sigma2 (5.99 mono / 7.8 when u_right[idx] >= 0), then the closest descriptor (strict '<', first wins).
 * best_idx[q] = keypoint index or -1, best_dist[q] = distance or 256.  What happens to a hit (Replace / AddObservation /
 * mutual-agreement check) is pointer-graph work and stays with the caller.                                            */
typedef struct {
    float u, v, radius;    /* projection and search radius th * mvScaleFactors[level] */
    float ur;              /* u - bf/z, only read by the stereo chi-square test */
    int32_t level;         /* nPredictedLevel */
    int32_t valid;         /* the point passed the caller's visibility tests */
} orbm_best_query;
int orbm_search_projected_best(orbm_handle h, orbm_frame kf, const orbm_best_query* queries, const uint8_t* query_desc,
                               int nq, int chi2_filter, const float* u_right, const float* inv_level_sigma2, int nlevels,
                               int* best_idx, int* best_dist);

/* Brute-force matching of n_pairs independent (query set, train set) pairs: every query against every train
 * descriptor, best / second-best / index with the reference's strict '<' (first wins), acceptance
 * best <= TH_LOW && best < (float)second * nnratio, rotation-histogram pruning (ORBmatcher.cc:432-461, 473-512).
 * Layout: queries[pair][nq][32], trains[pair][nt][32], angles [pair][n]; outputs [pair][nq]; nmatches[pair].
 * best/second/idx may be NULL.                                                                                   */
int orbm_bruteforce(orbm_handle h, const uint8_t* queries, const float* q_angles, int nq, const uint8_t* trains,
                    const float* t_angles, int nt, int n_pairs, float nnratio, int check_orientation,
                    int* best, int* second, int* idx, int* matches12, int* nmatches);
int orbm_bruteforce_device(orbm_handle h, const uint8_t* d_queries, const float* d_q_angles, int nq,
                           const uint8_t* d_trains, const float* d_t_angles, int nt, int n_pairs, float nnratio,
                           int check_orientation, int* d_best, int* d_second, int* d_idx, int* d_matches12,
                           int* d_nmatches, void* stream);

/* All-pairs keyframe matching: for query keyframes [q_begin, q_end) against all n_kf keyframes of the table
 * (n_kf x n_desc x 32 B, angles n_kf x n_desc), the number of matches under SearchByBoW(KF,KF) semantics without the
 * vocabulary gating (ORBmatcher.cc:566-618, 634-652).  counts: (q_end-q_begin) x n_kf int32, device memory.
 * db_begin/db_end restrict the db keyframes (columns) computed by this call.                                       */
int orbm_allpairs_device(orbm_handle h, const uint8_t* d_table, const float* d_angles, int n_kf, int n_desc,
                         int q_begin, int q_end, int db_begin, int db_end, float nnratio, int check_orientation,
                         int* d_counts, void* stream);

/* Config 5, multi-GPU: all-pairs keyframe matching sharded by query block, one process per GPU.  The only exchange of the
 * path is the all-gather of the keyframe descriptor table (NCCL over NVLink); it is cut into chunks of chunk_kf keyframes
 * per rank so that matching against the chunks that have landed overlaps the transfer of the next ones, and the rank's
 * own block (no communication) is matched first.  NCCL is loaded at run time (libnccl.so.2); single-GPU callers never
 * touch it.  Semantics of every (query keyframe, db keyframe) count as in orbm_allpairs_device.
 *   orbm_comm_unique_id: rank 0 makes the 128-byte ncclUniqueId, the caller hands it to the other ranks (MPI, a file,
 *                        torch.distributed.broadcast, ...)
 *   orbm_comm_create:    ncclCommInitRank on `device` (collective: every rank calls it)
 *   orbm_allpairs_sharded: d_local_desc = this rank's block [kf_per_rank[rank]][n_desc][32], d_local_angles likewise;
 *                        kf_per_rank[world] (host) = block sizes, blocks are contiguous in rank order;
 *                        q_count = how many keyframes from the start of the local block are queries (-1 = all of them:
 *                        the full matrix; a smaller number matches new keyframes against the whole map);
 *                        d_counts = [q_count][sum(kf_per_rank)] int32.  Enqueues on `stream` (NULL = the matcher's
 *                        stream) and on the communicator's own copy stream; returns without synchronising.
 *   orbm_comm_last_gather: time the copy stream spent in the last call's gathers, bytes received, chunk count.       */
typedef struct orbm_comm_s* orbm_comm;
int orbm_comm_unique_id(uint8_t* id128);
int orbm_comm_create(const uint8_t* id128, int rank, int world, int device, orbm_comm* out);
int orbm_comm_destroy(orbm_comm c);
int orbm_comm_info(orbm_comm c, int* rank, int* world, int* nccl_version);
int orbm_allpairs_sharded(orbm_handle h, orbm_comm c, const uint8_t* d_local_desc, const float* d_local_angles,
                          const int* kf_per_rank, int n_desc, int q_count, int chunk_kf, float nnratio,
                          int check_orientation, int* d_counts, void* stream);
int orbm_comm_last_gather(orbm_comm c, double* ms, double* bytes_received, int* chunks);

/* MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:257-322) for n_points map points in one call (LocalMapping runs it
 * for every point a new keyframe touches): point p is observed by descriptors[start[p] .. start[p+1]) -- rows of 32 B in
 * mObservations order, bad keyframes already dropped; start[0] = 0.  best[p] = index INSIDE the point's run of the
 * descriptor whose median distance to all of the run (itself included, rank (size_t)(0.5*(N-1))) is smallest, the first
 * such row winning; -1 for an empty run.  best_median (may be NULL) receives that median.                          */
int orbm_distinctive_descriptors(orbm_handle h, const uint8_t* descriptors, const int* start, int n_points, int* best,
                                 int* best_median);

/* Frame::ComputeBoW / KeyFrame::ComputeBoW (Frame.cc:736-745: mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4)).
 * The vocabulary is DBoW2's m_nodes as flat arrays (node 0 = root): node_desc[n_nodes][32], Node::children as CSR
 * (child_start[n_nodes+1], children[] in vector order), Node::word_id and Node::weight (idf) per node, depth_l = m_L.
 * adapter/ORBVocabulary.h reads Vocabulary/ORBvoc.txt / .bin (DBoW2's two file formats) into exactly these arrays.      */
typedef struct orbm_vocabulary_s* orbm_vocabulary;
int orbm_vocabulary_create(orbm_handle h, int n_nodes, int depth_l, const uint8_t* node_desc, const int* child_start,
                           const int* children, const int* word_id, const double* weight, orbm_vocabulary* out);
int orbm_vocabulary_destroy(orbm_vocabulary v);
/* TemplatedVocabulary::transform(features, v, fv, levelsup) with TF_IDF weighting and L1 scoring (what ORBvoc uses;
 * Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1166-1262, :1443-1485).  Per feature (any of them may be NULL): word_id[n],
 * weight[n] (the word's idf; 0 = stopped word, left out of both vectors), node_id[n] (ancestor `levelsup` levels above
 * the leaves; the root when the tree is shallower, and when a leaf sits above that level).  BowVector (all three or
 * none): bow_word[<=n] ascending, bow_value L1-normalised, *n_words.  FeatureVector (all four or none) as the node-sorted
 * CSR that orbm_search_by_bow / orbm_search_for_triangulation take: fv_node[<=n], fv_start[<=n+1], fv_idx[<=n], *n_nodes;
 * feature indices ascend inside a node (this fork fills them from four racing threads, so its order there is timing-
 * dependent; ascending is what its single-threaded variant produces).                                              */
int orbm_bow_transform(orbm_handle h, orbm_vocabulary v, const uint8_t* descriptors, int n, int levelsup,
                       int* word_id, double* weight, int* node_id, int* bow_word, double* bow_value, int* n_words,
                       int* fv_node, int* fv_start, int* fv_idx, int* n_nodes);
/* The descent alone on device-resident descriptors (e.g. where orbx_extract_batch_device left them); enqueues only. */
int orbm_bow_transform_device(orbm_handle h, orbm_vocabulary v, const uint8_t* d_descriptors, int n, int levelsup,
                              int* d_word_id, double* d_weight, int* d_node_id, void* stream);

/* Measured POPC-pipe throughput of this GPU (the roofline denominator for matching): 32-bit POPC per second. */
int orbm_popc_peak(orbm_handle h, double* popc_per_s);

#ifdef __cplusplus
}
#endif
#endif /* ORBB200_H */
