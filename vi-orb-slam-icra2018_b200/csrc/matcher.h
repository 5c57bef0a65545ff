// Internal declarations shared by the matcher translation units.
#pragma once
#include "common.cuh"

namespace orbb {

int launch_bruteforce(const uint8_t* dq, const float* dqa, int nq, const uint8_t* dt, const float* dta, int nt,
                      int nPairs, float ratio, int checkOri, int* dBestKey, int* dSecondKey, int* dBest, int* dSecond,
                      int* dIdx, int* dM12, int* dN, cudaStream_t st, int* launches);
int launch_allpairs(const uint8_t* dTable, const float* dAngles, int nKf, int nDesc, int qBegin, int qEnd, int dbBegin,
                    int dbEnd, float ratio, int checkOri, int* dCounts, cudaStream_t st, int* launches);
// db keyframes of one all-pairs launch as a concatenation of up to 16 segments: dense index [start[s], start[s+1]) ->
// rows row[s].. of the db table, columns col[s].. of the count matrix
struct ApSegments {
    int n;
    int start[17], row[16], col[16];
};
int launch_allpairs_ex(const uint8_t* dQTable, const float* dQAngles, int qBegin, int qEnd, const uint8_t* dTable,
                       const float* dAngles, int dbBegin, int dbEnd, int nDesc, int ldCounts, int colOffset, float ratio,
                       int checkOri, int* dCounts, cudaStream_t st, int* launches, const ApSegments* segs = nullptr);
int launch_distance(const uint8_t* da, const uint8_t* db, int n, int* dOut, cudaStream_t st, int* launches);
int launch_distinctive(const uint8_t* dDesc, const int* dStart, int nPoints, int* dBest, int* dBestMedian, cudaStream_t st,
                       int* launches);
int measure_popc_peak(cudaStream_t st, double* popcPerS);

}  // namespace orbb

// One matcher = one device, one stream, growable workspaces; not re-entrant per handle (ORBmatcher itself is a
// stateless value type constructed per call site, ORBmatcher.h:41).
struct orbm_matcher {
    int device = 0;
    cudaStream_t stream = nullptr;
    int launches = 0;
    orbb::DevBuf in0, in1, in2, in3, in4, in5, out0, out1, out2, out3, out4, ws0, ws1, ws2, ws3;
    orbb::PinnedBuf pin0, pin1, pin2, pin3, pin4;   // staging of the batched searches
    // Single calls take pageable host arrays, for which every cudaMemcpyAsync is a synchronous staged copy inside the driver.
    // stage_upload / stage_download go through this pinned area instead (bump-allocated per call, reset by ORBM_ENTER):
    // one memcpy + a truly asynchronous copy in; asynchronous copies out, handed to the caller's arrays by stage_finish()
    // after the call's one synchronisation.
    orbb::PinnedBuf stage;
    size_t stageOff = 0;
    struct PendingOut { void* dst; const void* src; size_t bytes; };
    PendingOut pending[8];
    int nPending = 0;
};

// Prologue of every matcher entry point: null check, device selection for the duration of the call, launch counter reset.
#define ORBM_ENTER(h)                                                                                 \
    if (!(h)) return ::orbb::fail(ORB_ERR_INVALID, "%s: null matcher handle", __func__);              \
    ::orbb::DeviceGuard guard__((h)->device);                                                         \
    if (!guard__.ok) return ::orbb::fail(ORB_ERR_CUDA, "%s: cannot select device %d", __func__, (h)->device); \
    (h)->launches = 0;                                                                                \
    (h)->stageOff = 0;                                                                                \
    (h)->nPending = 0;
