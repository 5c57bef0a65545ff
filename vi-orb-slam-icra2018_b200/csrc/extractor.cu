// Extractor handle: constructor tables, per-image-size geometry, device arena, launch sequencing, C-ABI entry points.
//   ORBextractor::ORBextractor   ORBextractor.cc:412-472     (scale tables, features per level, umax)
//   ORBextractor::operator()     ORBextractor.cc:1045-1126   (pyramid -> keypoints -> blur + descriptors)
// Everything per pixel / per keypoint runs in the kernels of pyramid.cu, fast_warp.cu, octree.cu, blur.cu, brief.cu; the
// host code here only derives sizes and tables (the same float expressions as the reference, compiled without FMA
// contraction) and enqueues 12 launches per batch on one stream.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "extractor.h"

using namespace orbb;

namespace {

inline int cv_round_f(float v) { return (int)lrintf(v); }
inline int round_up(int v, int a) { return (v + a - 1) / a * a; }
inline long long round_up_ll(long long v, long long a) { return (v + a - 1) / a * a; }

// cv::resize INTER_LINEAR per-axis table (imgproc/resize.cpp): source index and 11-bit coefficient pair
void resize_axis_table(int ssize, int dsize, int* ofs, short2* coef) {
    const double scale = 1.0 / ((double)dsize / (double)ssize);
    for (int d = 0; d < dsize; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
        ofs[d] = s;
        coef[d].x = (short)cv_round_f((1.f - f) * 2048.f);
        coef[d].y = (short)cv_round_f(f * 2048.f);
    }
}

}  // namespace

struct orbx_extractor {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t streamIn = nullptr, streamOut = nullptr;   // H2D / D2H of the host entry points, overlapped with compute
    std::vector<cudaEvent_t> pipeEvents;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t streamSide = nullptr;      // small calls: the blur (needs only the pyramid) runs beside FAST + quadtree
    cudaStream_t stream2 = nullptr;         // host pipeline: odd chunks' kernels (their launch gaps and tails hide under the even ones')
    PinnedBuf pinIn, pinOut;                // small calls: staging of pageable images / outputs
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    // ctor state
    int nfeatures = 0, nlevels = 0, iniTh = 0, minTh = 0;
    double scaleFactor = 1;
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> perLevel;
    int umax[16];
    int maxW = 0, maxH = 0, maxBatch = 1;
    // geometry of the current image size
    int curW = -1, curH = -1;
    ExtractParams P;
    std::vector<Cell> cells;
    std::vector<BlurTile> tiles;
    int kpCapacity = 0;       // upper bound on keypoints per frame
    int otSmem = 0, otKeyCap = 0, otNodeCap = 0, otCellCap = 0;
    // device memory
    DevBuf pyr, blur, slots, cellCount, sel, selCount, keyWs, dCells, dTiles, dTabOfs, dTabCoef, dPyCol, dPyRow, dPyBand, dPyBarrier, dFastMaps, dBriefMaps, dFastScratch, dFastCounters;
    DevBuf dImages, dKps, dDesc, dCount;
    DevBuf stKeysL, stDescL, stKeysR, stDescR, stOut;   // staging of orbx_compute_stereo_matches
    int lastFrames = 0, lastCapacity = 0;   // arena capacity in frames / caller capacity of the last call
    int residentFrames = 0;                 // frames of the last call whose pyramids are in the arena (slots 0 .. residentFrames-1)
    bool lastWasHost = false;               // the last call was a host call: its outputs are still in dKps / dDesc / dCount
    int launches = 0;
    double stageMs[3] = {0, 0, 0};
    bool timed = false;
    // the launch sequence of a small host call (<= 4 frames), captured once per (size, buffers) and replayed as a CUDA graph:
    // a live SLAM loop extracts one frame at a time, where 12 launches' host overhead is a third of the latency
    struct GraphKey {
        int w = 0, h = 0, stride = 0, nFrames = 0, capacity = 0;
        const void *img = nullptr, *kps = nullptr, *desc = nullptr, *cnt = nullptr, *pyr = nullptr;
        const void *pinIn = nullptr, *pinOut = nullptr;     // staged calls: the copies in and out are nodes of the graph
        bool operator==(const GraphKey& o) const {
            return w == o.w && h == o.h && stride == o.stride && nFrames == o.nFrames && capacity == o.capacity && img == o.img &&
                   kps == o.kps && desc == o.desc && cnt == o.cnt && pyr == o.pyr && pinIn == o.pinIn && pinOut == o.pinOut;
        }
    } graphKey, graphSeen;
    cudaGraphExec_t graphExec = nullptr;
    int graphLaunches = 0;
    // optional per-kernel event sets (orbx_set_profiling)
    bool profiling = false;
    std::vector<cudaEvent_t> profEvents;   // 6 per profiled call
    int profCalls = 0;
};

namespace {

int configure(orbx_extractor* e, int w, int h, int nFrames) {
    if (e->curW == w && e->curH == h && nFrames <= e->lastFrames) return ORB_OK;
    const int nl = e->nlevels;
    ExtractParams& P = e->P;
    std::vector<LevelGeom> lv(nl);
    long long pyrOff = 0, blurOff = 0, slotOff = 0, keyOff = 0;
    int tabOff = 0, selOff = 0, cellBase = 0;
    std::vector<Cell> cells;
    std::vector<BlurTile> tiles;
    int nodeCap = 8, cellCap = 1;
    for (int l = 0; l < nl; ++l) {
        LevelGeom& L = lv[l];
        L.w = cv_round_f((float)w * e->invScale[l]);   // :1133
        L.h = cv_round_f((float)h * e->invScale[l]);
        if (L.w < 62 || L.h < 62)
            return fail(ORB_ERR_INVALID, "level %d is %dx%d: the reference's cell grid needs at least 62x62 (ORBextractor.cc:783-789)",
                        l, L.w, L.h);
        if (L.w > 4096 + 32 || L.h > 4096 + 32) return fail(ORB_ERR_INVALID, "level %d is %dx%d: larger than 4128", l, L.w, L.h);
        L.pitch = round_up(kPadLeft + L.w + kEdge, 16);
        L.bpitch = round_up(L.w, 16);
        L.pyrOff = pyrOff;
        pyrOff = round_up_ll(pyrOff + (long long)L.pitch * (L.h + 2 * kEdge), 256);
        L.blurOff = blurOff;
        blurOff = round_up_ll(blurOff + (long long)L.bpitch * L.h, 256);
        L.xTab = tabOff;
        tabOff += L.w;
        L.yTab = tabOff;
        tabOff += L.h;
        // cell grid (:775-789)
        const int minB = kEdge - 3;
        const int maxBX = L.w - kEdge + 3, maxBY = L.h - kEdge + 3;
        const float width = (float)(maxBX - minB), height = (float)(maxBY - minB);
        const int nCols = (int)(width / 30.f), nRows = (int)(height / 30.f);
        const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
        if (wCell > kCellMax || hCell > kCellMax) return fail(ORB_ERR_INVALID, "level %d: cell %dx%d too large", l, wCell, hCell);
        L.slotCap = ((wCell + 1) / 2) * ((hCell + 1) / 2);
        L.slotBase = slotOff;
        L.cellBase = cellBase;
        for (int i = 0; i < nRows; ++i)
            for (int j = 0; j < nCols; ++j) {
                Cell c;
                c.level = (short)l;
                c.x0 = (short)(kEdge + j * wCell);
                c.y0 = (short)(kEdge + i * hCell);
                c.cw = (short)std::min(wCell, (L.w - kEdge) - c.x0);
                c.ch = (short)std::min(hCell, (L.h - kEdge) - c.y0);
                if (c.cw <= 0 || c.ch <= 0) continue;   // the reference skips it or FAST sees a ROI thinner than 7 px
                c.pad = 0;
                c.slot = (int)slotOff;
                slotOff += L.slotCap;
                cells.push_back(c);
            }
        L.nCells = (int)cells.size() - cellBase;
        cellBase = (int)cells.size();
        cellCap = std::max(cellCap, L.nCells);
        // quadtree (:545-560)
        L.nFeatures = e->perLevel[l];
        L.winW = maxBX - minB;
        L.winH = maxBY - minB;
        L.nIni = (int)roundf(width / height);
        if (L.nIni < 1) return fail(ORB_ERR_INVALID, "level %d: aspect ratio %dx%d gives no quadtree root (ORBextractor.cc:545)", l, L.w, L.h);
        L.hX = width / L.nIni;
        L.selCap = std::max(L.nFeatures + 3, 4 * L.nIni) + 1;
        L.selBase = selOff;
        selOff += L.selCap;
        nodeCap = std::max(nodeCap, L.selCap);
        L.keyWsCap = L.nCells * L.slotCap + 2;
        L.keyWsOff = keyOff;
        keyOff += L.keyWsCap + (L.keyWsCap + 1) / 2 + 2;
        L.scale = e->scale[l];
        L.patchSize = (float)(int)((float)kPatch * e->scale[l]);   // :837, 848
        for (int c = 0; c < blur_cta_count(L.w, L.h); ++c) tiles.push_back(BlurTile{l, c});
    }
    if (nodeCap > 65535) return fail(ORB_ERR_INVALID, "nfeatures too large for the quadtree kernel");
    ORB_CHECK(octree_smem_plan(nodeCap, cellCap, &e->otSmem, &e->otKeyCap));
    e->otNodeCap = nodeCap;
    e->otCellCap = cellCap;

    // resize tables
    std::vector<int> tabOfs(tabOff);
    std::vector<short2> tabCoef(tabOff);
    for (int l = 1; l < nl; ++l) {
        resize_axis_table(lv[l - 1].w, lv[l].w, &tabOfs[lv[l].xTab], &tabCoef[lv[l].xTab]);
        resize_axis_table(lv[l - 1].h, lv[l].h, &tabOfs[lv[l].yTab], &tabCoef[lv[l].yTab]);
    }

    std::vector<uint4> pyCol, pyRow;
    std::vector<int4> pyBand;
    for (int l = 1; l < nl; ++l) {
        const size_t c0 = pyCol.size(), r0 = pyRow.size();
        lv[l].pyCol = (int)c0;
        lv[l].pyRow = (int)r0;
        lv[l].pyFast = pyramid_level_plan(lv[l - 1], lv[l], &tabOfs[lv[l].xTab], &tabCoef[lv[l].xTab], &tabOfs[lv[l].yTab],
                                          &tabCoef[lv[l].yTab], pyCol, pyRow) ? 1 : 0;
        if (!lv[l].pyFast) { pyCol.resize(c0); pyRow.resize(r0); }
        lv[l].pyBand = (int)pyBand.size();
        lv[l].pyBufBytes = round_up(pyramid_band_plan(lv[l - 1], lv[l], &tabOfs[lv[l].yTab], pyBand), 128);
        lv[l].pyBulkCtas = lv[l].pyFast ? pyramid_bulk_ctas(lv[l].pitch / 4, lv[l].pyBufBytes) : 0;
        lv[l].pyBulk = lv[l].pyBulkCtas > 0;
    }

    // arena
    const long long F = std::max(nFrames, e->maxBatch);
    ORB_CHECK(e->pyr.reserve((size_t)(pyrOff * F) + 256));
    ORB_CHECK(e->blur.reserve((size_t)(blurOff * F) + 256));
    ORB_CHECK(e->slots.reserve((size_t)(slotOff * F) * 4 + 256));
    ORB_CHECK(e->cellCount.reserve((size_t)cells.size() * F * 4 + 256));
    ORB_CHECK(e->sel.reserve((size_t)selOff * F * sizeof(SelKey) + 256));
    ORB_CHECK(e->selCount.reserve((size_t)nl * F * 4 + 256));
    ORB_CHECK(e->keyWs.reserve((size_t)(keyOff * F) * 4 + 256));
    ORB_CHECK(e->dCells.reserve(cells.size() * sizeof(Cell) + 16));
    ORB_CHECK(e->dTiles.reserve(tiles.size() * sizeof(BlurTile) + 16));
    ORB_CHECK(e->dTabOfs.reserve(tabOfs.size() * 4 + 16));
    ORB_CHECK(e->dTabCoef.reserve(tabCoef.size() * 4 + 16));
    ORB_CUDA(cudaMemcpyAsync(e->dCells.p, cells.data(), cells.size() * sizeof(Cell), cudaMemcpyHostToDevice, e->stream));
    ORB_CUDA(cudaMemcpyAsync(e->dTiles.p, tiles.data(), tiles.size() * sizeof(BlurTile), cudaMemcpyHostToDevice, e->stream));
    ORB_CUDA(cudaMemcpyAsync(e->dTabOfs.p, tabOfs.data(), tabOfs.size() * 4, cudaMemcpyHostToDevice, e->stream));
    ORB_CUDA(cudaMemcpyAsync(e->dTabCoef.p, tabCoef.data(), tabCoef.size() * 4, cudaMemcpyHostToDevice, e->stream));
    ORB_CHECK(e->dPyCol.reserve(pyCol.size() * 16 + 16));
    ORB_CHECK(e->dPyRow.reserve(pyRow.size() * 16 + 16));
    if (!pyCol.empty()) ORB_CUDA(cudaMemcpyAsync(e->dPyCol.p, pyCol.data(), pyCol.size() * 16, cudaMemcpyHostToDevice, e->stream));
    if (!pyRow.empty()) ORB_CUDA(cudaMemcpyAsync(e->dPyRow.p, pyRow.data(), pyRow.size() * 16, cudaMemcpyHostToDevice, e->stream));
    ORB_CHECK(e->dPyBand.reserve(pyBand.size() * 16 + 16));
    if (!pyBand.empty()) ORB_CUDA(cudaMemcpyAsync(e->dPyBand.p, pyBand.data(), pyBand.size() * 16, cudaMemcpyHostToDevice, e->stream));
    ORB_CUDA(cudaStreamSynchronize(e->stream));   // the host vectors go out of scope

    std::memset(&P, 0, sizeof P);
    P.nLevels = nl;
    P.iniTh = e->iniTh;
    P.minTh = std::min(e->minTh, e->iniTh);   // a second pass at a threshold >= the first cannot find anything new
    P.nCellsTotal = (int)cells.size();
    P.maxCellW = 1;
    P.maxCellH = 1;
    for (const Cell& c : cells) { P.maxCellW = std::max(P.maxCellW, (int)c.cw); P.maxCellH = std::max(P.maxCellH, (int)c.ch); }
    {
        int cellW[kMaxLevels] = {0}, cellH[kMaxLevels] = {0}, slotCapMax = 1;
        for (const Cell& c : cells) {
            cellW[c.level] = std::max(cellW[c.level], (int)c.cw);
            cellH[c.level] = std::max(cellH[c.level], (int)c.ch);
        }
        for (int l = 0; l < nl; ++l) slotCapMax = std::max(slotCapMax, lv[l].slotCap);
        ORB_CHECK(fast_warp_plan(nl, cellW, cellH, slotCapMax, &P.fw));
    }
    P.selPerFrame = selOff;
    P.pyrFrameBytes = pyrOff;
    P.blurFrameBytes = blurOff;
    P.slotFrameEntries = slotOff;
    P.keyWsFrameEntries = keyOff;
    P.pyr = e->pyr.as<unsigned char>();
    P.blur = e->blur.as<unsigned char>();
    P.slots = e->slots.as<unsigned int>();
    P.cellCount = e->cellCount.as<int>();
    P.sel = e->sel.as<SelKey>();
    P.selCount = e->selCount.as<int>();
    P.keyWs = e->keyWs.as<unsigned int>();
    P.cells = e->dCells.as<Cell>();
    P.tabOfs = e->dTabOfs.as<int>();
    P.tabCoef = e->dTabCoef.as<short2>();
    P.pyColTab = e->dPyCol.as<uint4>();
    P.pyRowTab = e->dPyRow.as<uint4>();
    P.pyBandTab = e->dPyBand.as<int4>();
    ORB_CHECK(e->dPyBarrier.reserve(sizeof(unsigned int) * kMaxLevels));
    P.pyBarrier = e->dPyBarrier.as<unsigned int>();
    {
        const char* v = getenv("ORBB_PYR_BULK_MIN");      // tuning aid: smallest batch that takes the staged resize kernel
        P.pyBulkMinFrames = v ? atoi(v) : 8;
    }
    for (int l = 0; l < nl; ++l) lv[l].blCtas = blur_staged_ctas(lv[l]);
    for (int l = 0; l < nl; ++l) P.lv[l] = lv[l];
    {   // TMA descriptors of the padded pyramid levels: [arena frame][row][pitch], box = one FAST tile
        unsigned char hostMaps[128 * kMaxLevels];
        ORB_CHECK(e->dFastMaps.reserve(sizeof hostMaps));
        ORB_CHECK(fast_warp_encode_maps(P, (int)F, hostMaps));
        ORB_CUDA(cudaMemcpyAsync(e->dFastMaps.p, hostMaps, sizeof hostMaps, cudaMemcpyHostToDevice, e->stream));
        ORB_CUDA(cudaStreamSynchronize(e->stream));
        P.fw.maps = e->dFastMaps.p;
        ORB_CHECK(e->dBriefMaps.reserve(sizeof hostMaps));
        ORB_CHECK(brief_encode_maps(P, (int)F, hostMaps));
        ORB_CUDA(cudaMemcpyAsync(e->dBriefMaps.p, hostMaps, sizeof hostMaps, cudaMemcpyHostToDevice, e->stream));
        ORB_CUDA(cudaStreamSynchronize(e->stream));
        P.brMaps = e->dBriefMaps.p;
        ORB_CHECK(fast_warp_max_warps(P.fw, &P.fw.maxWarps));
        // two sets (scratch queues, ticket counters): two chunks of the host pipeline may run their FAST kernels at the same time
        ORB_CHECK(e->dFastScratch.reserve((size_t)P.fw.maxWarps * P.fw.scratchCap * 2 * 2));
        P.fw.scratch = e->dFastScratch.as<unsigned short>();
        ORB_CHECK(e->dFastCounters.reserve(512));
        ORB_CUDA(cudaMemsetAsync(e->dFastCounters.p, 0, 512, e->stream));
        ORB_CUDA(cudaStreamSynchronize(e->stream));
        P.fw.counters = e->dFastCounters.as<unsigned int>();
    }
    e->cells.swap(cells);
    e->tiles.swap(tiles);
    e->kpCapacity = selOff;
    e->curW = w;
    e->curH = h;
    e->lastFrames = (int)F;
    return ORB_OK;
}

// enqueue one batch (device pointers), no synchronisation
int enqueue(orbx_extractor* e, const uint8_t* dImages, int nFrames, int w, int h, int stride, size_t frameStride,
            orb_keypoint* dKps, uint8_t* dDesc, int capacity, int* dCount, cudaStream_t st, bool timed, int frameBase = 0,
            int set = 0) {
    ORB_CHECK(configure(e, w, h, frameBase + nFrames));
    ExtractParams P = e->P;
    P.nFrames = nFrames;
    P.outCapacity = capacity;
    P.fw.frameBase = frameBase;
    if (set) {         // the second set of the FAST kernel's scratch queues and ticket counters (256 bytes apart)
        P.fw.scratch += (size_t)P.fw.maxWarps * P.fw.scratchCap;
        P.fw.counters += 64;
    }
    if (frameBase) {   // this call works in arena slots [frameBase, frameBase + nFrames)
        P.pyr += (size_t)frameBase * P.pyrFrameBytes;
        P.blur += (size_t)frameBase * P.blurFrameBytes;
        P.slots += (size_t)frameBase * P.slotFrameEntries;
        P.cellCount += (size_t)frameBase * P.nCellsTotal;
        P.sel += (size_t)frameBase * P.selPerFrame;
        P.selCount += (size_t)frameBase * P.nLevels;
        P.keyWs += (size_t)frameBase * P.keyWsFrameEntries;
    }
    e->lastCapacity = capacity;
    e->residentFrames = frameBase + nFrames;   // extract_batch's chunks fill the arena front to back
    e->timed = timed;
    cudaEvent_t* pe = nullptr;
    if (e->profiling && e->profCalls < 4096) {
        if ((int)e->profEvents.size() < 6 * (e->profCalls + 1)) {
            for (int i = 0; i < 6; ++i) {
                cudaEvent_t ev;
                ORB_CUDA(cudaEventCreate(&ev));
                e->profEvents.push_back(ev);
            }
        }
        pe = &e->profEvents[6 * e->profCalls++];
    }
    if (timed) ORB_CUDA(cudaEventRecord(e->ev[0], st));
    if (pe) ORB_CUDA(cudaEventRecord(pe[0], st));
    ORB_CHECK(launch_pyramid(P, dImages, w, h, stride, frameStride, st, &e->launches));
    if (timed) ORB_CUDA(cudaEventRecord(e->ev[1], st));
    if (pe) ORB_CUDA(cudaEventRecord(pe[1], st));
    // a small call leaves most SMs idle during FAST + quadtree: its blur (which needs only the pyramid) runs beside them
    static const bool noFork = getenv("ORBB_NO_FORK") != nullptr;
    const bool fork = !noFork && !pe && nFrames < P.pyBulkMinFrames && e->streamSide && st == e->stream;
    if (fork) {
        ORB_CUDA(cudaEventRecord(e->evFork, st));
        ORB_CUDA(cudaStreamWaitEvent(e->streamSide, e->evFork, 0));
        ORB_CHECK(launch_blur(P, e->dTiles.as<BlurTile>(), (int)e->tiles.size(), e->streamSide, &e->launches));
        ORB_CUDA(cudaEventRecord(e->evJoin, e->streamSide));
    }
    ORB_CHECK(launch_fast_warp(P, st, &e->launches));
    if (pe) ORB_CUDA(cudaEventRecord(pe[2], st));
    ORB_CHECK(launch_octree(P, e->otSmem, e->otKeyCap, e->otNodeCap, e->otCellCap, st, &e->launches));
    if (timed) ORB_CUDA(cudaEventRecord(e->ev[2], st));
    if (pe) ORB_CUDA(cudaEventRecord(pe[3], st));
    if (fork) ORB_CUDA(cudaStreamWaitEvent(st, e->evJoin, 0));
    else ORB_CHECK(launch_blur(P, e->dTiles.as<BlurTile>(), (int)e->tiles.size(), st, &e->launches));
    if (pe) ORB_CUDA(cudaEventRecord(pe[4], st));
    ORB_CHECK(launch_brief(P, std::min(capacity, e->kpCapacity), dKps, dDesc, dCount, st, &e->launches));
    if (timed) ORB_CUDA(cudaEventRecord(e->ev[3], st));
    if (pe) ORB_CUDA(cudaEventRecord(pe[5], st));
    return ORB_OK;
}

}  // namespace

#define ORBX_ENTER(h)                                                                                  \
    if (!(h)) return fail(ORB_ERR_INVALID, "%s: null extractor handle", __func__);                     \
    DeviceGuard guard__((h)->device);                                                                  \
    if (!guard__.ok) return fail(ORB_ERR_CUDA, "%s: cannot select device %d", __func__, (h)->device);

extern "C" {

int orbx_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, int maxW, int maxH, int maxBatch,
                int device, orbx_handle* out) {
    if (!out) return fail(ORB_ERR_INVALID, "orbx_create: null out");
    *out = nullptr;
    if (nfeatures < 1 || nlevels < 1 || nlevels > kMaxLevels || !(scaleFactor > 1.0f) || iniTh < 1 || minTh < 1 ||
        iniTh > 255 || minTh > iniTh || maxBatch < 1)
        return fail(ORB_ERR_INVALID, "orbx_create: need nfeatures>=1, 1<=nlevels<=%d, scaleFactor>1, 1<=minThFAST<=iniThFAST<=255, max_batch>=1",
                    kMaxLevels);
    int n = 0;
    ORB_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(ORB_ERR_INVALID, "orbx_create: device %d of %d", device, n);
    DeviceGuard g(device);
    if (!g.ok) return fail(ORB_ERR_CUDA, "orbx_create: cannot select device %d", device);

    orbx_extractor* e = new orbx_extractor;
    e->device = device;
    e->nfeatures = nfeatures; e->nlevels = nlevels; e->iniTh = iniTh; e->minTh = minTh;
    e->scaleFactor = (double)scaleFactor;
    e->maxW = maxW; e->maxH = maxH; e->maxBatch = maxBatch;
    // scale tables (:417-434): float * double member, stored as float
    e->scale.assign(nlevels, 1.f);
    e->sigma2.assign(nlevels, 1.f);
    for (int i = 1; i < nlevels; ++i) {
        e->scale[i] = (float)((double)e->scale[i - 1] * e->scaleFactor);
        e->sigma2[i] = e->scale[i] * e->scale[i];
    }
    e->invScale.resize(nlevels);
    e->invSigma2.resize(nlevels);
    for (int i = 0; i < nlevels; ++i) {
        e->invScale[i] = 1.0f / e->scale[i];
        e->invSigma2[i] = 1.0f / e->sigma2[i];
    }
    // features per level (:436-448)
    e->perLevel.assign(nlevels, 0);
    const float factor = (float)(1.0 / e->scaleFactor);
    float desired = (float)nfeatures * (1.f - factor) / (1.f - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        e->perLevel[l] = cv_round_f(desired);
        sum += e->perLevel[l];
        desired *= factor;
    }
    e->perLevel[nlevels - 1] = std::max(nfeatures - sum, 0);
    // circular patch extents (:456-471)
    {
        int* um = e->umax;
        std::memset(um, 0, sizeof e->umax);
        const int vmax = (int)floorf(kHalfPatch * sqrtf(2.f) / 2 + 1);
        const int vmin = (int)ceilf(kHalfPatch * sqrtf(2.f) / 2);
        const double hp2 = kHalfPatch * kHalfPatch;
        for (int v = 0; v <= vmax; ++v) um[v] = (int)lrint(sqrt(hp2 - v * v));
        for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
            while (um[v0] == um[v0 + 1]) ++v0;
            um[v] = v0;
            ++v0;
        }
    }
    cudaError_t ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->streamIn, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->streamOut, cudaStreamNonBlocking);
    for (int i = 0; i < 4 && ce == cudaSuccess; ++i) ce = cudaEventCreate(&e->ev[i]);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->streamSide, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e->evFork, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e->evJoin, cudaEventDisableTiming);
    if (ce != cudaSuccess) {
        delete e;
        return fail(ORB_ERR_CUDA, "orbx_create: %s", cudaGetErrorString(ce));
    }
    int st = upload_brief_pattern();
    if (st == ORB_OK) st = upload_orientation_table(e->umax);
    if (st == ORB_OK && maxW > 0 && maxH > 0) st = configure(e, maxW, maxH, maxBatch);
    if (st != ORB_OK) {
        orbx_destroy(e);
        return st;
    }
    *out = e;
    return ORB_OK;
}

int orbx_destroy(orbx_handle e) {
    if (!e) return ORB_OK;
    DeviceGuard g(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    DevBuf* bufs[] = {&e->pyr, &e->blur, &e->slots, &e->cellCount, &e->sel, &e->selCount, &e->keyWs, &e->dCells,
                      &e->dTiles, &e->dTabOfs, &e->dTabCoef, &e->dPyCol, &e->dPyRow, &e->dPyBand, &e->dPyBarrier, &e->dFastMaps, &e->dBriefMaps, &e->dFastScratch, &e->dFastCounters, &e->dImages, &e->dKps, &e->dDesc, &e->dCount,
                      &e->stKeysL, &e->stDescL, &e->stKeysR, &e->stDescR, &e->stOut};
    for (DevBuf* b : bufs) b->release();
    if (e->graphExec) cudaGraphExecDestroy(e->graphExec);
    for (int i = 0; i < 4; ++i)
        if (e->ev[i]) cudaEventDestroy(e->ev[i]);
    for (cudaEvent_t ev : e->profEvents) cudaEventDestroy(ev);
    for (cudaEvent_t ev : e->pipeEvents) cudaEventDestroy(ev);
    if (e->evFork) cudaEventDestroy(e->evFork);
    if (e->evJoin) cudaEventDestroy(e->evJoin);
    e->pinIn.release();
    e->pinOut.release();
    if (e->streamSide) cudaStreamDestroy(e->streamSide);
    if (e->stream2) cudaStreamDestroy(e->stream2);
    if (e->streamIn) cudaStreamDestroy(e->streamIn);
    if (e->streamOut) cudaStreamDestroy(e->streamOut);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return ORB_OK;
}

int orbx_keypoint_capacity(orbx_handle e, int* cap) {
    if (!e || !cap) return fail(ORB_ERR_INVALID, "orbx_keypoint_capacity: null argument");
    // bound for any image size: sum over levels of max(N_l + 3, 4*nIni) + 1, nIni <= round(4096/62) levels aside
    if (e->curW > 0) { *cap = e->kpCapacity; return ORB_OK; }
    int c = 0;
    for (int l = 0; l < e->nlevels; ++l) c += std::max(e->perLevel[l] + 3, 4 * 8) + 1;
    *cap = c;
    return ORB_OK;
}

int orbx_extract_batch_device(orbx_handle e, const uint8_t* dImages, int nFrames, int w, int h, int stride,
                              size_t frameStride, orb_keypoint* dKps, uint8_t* dDesc, int capacity, int* dCount,
                              void* stream) {
    ORBX_ENTER(e);
    if (nFrames < 0) return fail(ORB_ERR_INVALID, "orbx_extract_batch_device: negative frame count");
    if (nFrames == 0) return ORB_OK;
    if (nFrames > 65535) return fail(ORB_ERR_INVALID, "orbx_extract_batch_device: at most 65535 frames per call");
    if (!dImages || !dKps || !dDesc || !dCount || w <= 0 || h <= 0 || stride < w || capacity < 1)
        return fail(ORB_ERR_INVALID, "orbx_extract_batch_device: bad arguments");
    e->launches = 0;
    cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
    e->lastWasHost = false;
    return enqueue(e, dImages, nFrames, w, h, stride, frameStride, dKps, dDesc, capacity, dCount, st, stream == nullptr);
}

int orbx_extract_batch(orbx_handle e, const uint8_t* images, int nFrames, int w, int h, int stride, size_t frameStride,
                       orb_keypoint* kps, uint8_t* desc, int capacity, int* nOut) {
    ORBX_ENTER(e);
    if (nFrames < 0) return fail(ORB_ERR_INVALID, "orbx_extract_batch: negative frame count");
    if (nFrames == 0) return ORB_OK;
    if (!nOut) return fail(ORB_ERR_INVALID, "orbx_extract_batch: null n_out");
    if (!images || w <= 0 || h <= 0) {   // empty image: outputs untouched (:1048-1049)
        for (int f = 0; f < nFrames; ++f) nOut[f] = 0;
        return ORB_OK;
    }
    if (!kps || !desc || stride < w || capacity < 1) return fail(ORB_ERR_INVALID, "orbx_extract_batch: bad arguments");
    e->launches = 0;
    e->lastWasHost = true;
    // The arena holds `super` frames; inside it the batch is cut into pipeline chunks so that the H2D copy of chunk k+1
    // and the D2H copy of chunk k-1 overlap the kernels of chunk k (three streams, events between them).
    const int super = std::min(e->maxBatch, 65535);
    const size_t imgBytes = (size_t)stride * h;
    ORB_CHECK(configure(e, w, h, std::min(nFrames, super)));
    ORB_CHECK(e->dImages.reserve((size_t)super * imgBytes));
    ORB_CHECK(e->dKps.reserve((size_t)super * capacity * sizeof(orb_keypoint)));
    ORB_CHECK(e->dDesc.reserve((size_t)super * capacity * 32));
    ORB_CHECK(e->dCount.reserve((size_t)super * 4));
    // ---- small calls (the live loop: one frame, or a stereo pair): one stream, no pipeline; from the second call of a
    //      size on, the 12 launches are one CUDA graph launch
    static const bool noGraph = getenv("ORBB_NO_GRAPH") != nullptr;
    if (nFrames <= 4 && nFrames <= super) {
        cudaStream_t st = e->stream;
        uint8_t* dImg = e->dImages.as<uint8_t>();
        // A live loop hands over pageable memory (a cv::Mat's data, std::vector outputs), for which every cudaMemcpyAsync is a
        // synchronous staged copy inside the driver.  Staging through the handle's own pinned buffers -- one memcpy in, three
        // truly asynchronous copies out into pinned memory, then memcpy of the n valid entries -- takes ~30 us off a frame.
        static const bool noStage = getenv("ORBB_NO_STAGE") != nullptr;
        bool stage = !noStage;
        if (stage) {
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, images) == cudaSuccess && at.type != cudaMemoryTypeUnregistered) stage = false;   // already pinned
            else cudaGetLastError();
        }
        orb_keypoint* dK = e->dKps.as<orb_keypoint>();
        uint8_t* dD = e->dDesc.as<uint8_t>();
        int* dN = e->dCount.as<int>();
        int* pn = nullptr;
        orb_keypoint* pk = nullptr;
        uint8_t* pd = nullptr;
        if (stage) {
            ORB_CHECK(e->pinIn.reserve((size_t)nFrames * imgBytes));
            ORB_CHECK(e->pinOut.reserve((size_t)nFrames * ((size_t)capacity * 60 + 64)));
            for (int f = 0; f < nFrames; ++f) std::memcpy(e->pinIn.as<uint8_t>() + (size_t)f * imgBytes, images + (size_t)f * frameStride, imgBytes);
            uint8_t* po = e->pinOut.as<uint8_t>();
            pn = reinterpret_cast<int*>(po);
            pk = reinterpret_cast<orb_keypoint*>(po + 64);
            pd = po + 64 + (size_t)nFrames * capacity * sizeof(orb_keypoint);
        }
        // the copies of a call: into the arena, and (staged) out into the pinned buffer -- issued eagerly or captured
        auto copy_in = [&]() -> int {
            if (stage) ORB_CUDA(cudaMemcpyAsync(dImg, e->pinIn.p, (size_t)nFrames * imgBytes, cudaMemcpyHostToDevice, st));
            else if (frameStride == imgBytes) ORB_CUDA(cudaMemcpyAsync(dImg, images, (size_t)nFrames * imgBytes, cudaMemcpyHostToDevice, st));
            else ORB_CUDA(cudaMemcpy2DAsync(dImg, imgBytes, images, frameStride, imgBytes, nFrames, cudaMemcpyHostToDevice, st));
            return ORB_OK;
        };
        auto copy_out_staged = [&]() -> int {
            ORB_CUDA(cudaMemcpyAsync(pn, dN, (size_t)nFrames * 4, cudaMemcpyDeviceToHost, st));
            ORB_CUDA(cudaMemcpyAsync(pk, dK, (size_t)nFrames * capacity * sizeof(orb_keypoint), cudaMemcpyDeviceToHost, st));
            ORB_CUDA(cudaMemcpyAsync(pd, dD, (size_t)nFrames * capacity * 32, cudaMemcpyDeviceToHost, st));
            return ORB_OK;
        };
        orbx_extractor::GraphKey key;
        key.w = w; key.h = h; key.stride = stride; key.nFrames = nFrames; key.capacity = capacity;
        key.img = dImg; key.kps = dK; key.desc = dD; key.cnt = dN; key.pyr = e->P.pyr;
        key.pinIn = stage ? e->pinIn.p : nullptr;
        key.pinOut = stage ? e->pinOut.p : nullptr;
        bool timedCall = false;
        if (!noGraph && !e->profiling && e->graphExec && key == e->graphKey) {
            if (!stage) ORB_CHECK(copy_in());
            ORB_CUDA(cudaGraphLaunch(e->graphExec, st));     // staged: copy in, kernels and copies out are ONE launch
            e->launches = e->graphLaunches;
            e->residentFrames = nFrames;
            e->lastCapacity = capacity;
        } else if (!noGraph && !e->profiling && key == e->graphSeen) {
            // second call with these buffers: capture (attributes and occupancy caches were set by the first, eager call)
            if (e->graphExec) { cudaGraphExecDestroy(e->graphExec); e->graphExec = nullptr; }
            cudaGraph_t graph = nullptr;
            if (!stage) ORB_CHECK(copy_in());
            ORB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
            int stEnq = stage ? copy_in() : ORB_OK;
            if (stEnq == ORB_OK) stEnq = enqueue(e, dImg, nFrames, w, h, stride, imgBytes, dK, dD, capacity, dN, st, false, 0);
            if (stEnq == ORB_OK && stage) stEnq = copy_out_staged();
            const cudaError_t ce = cudaStreamEndCapture(st, &graph);
            if (stEnq != ORB_OK) { if (graph) cudaGraphDestroy(graph); return stEnq; }
            if (ce != cudaSuccess) return fail(ORB_ERR_CUDA, "orbx_extract_batch: stream capture failed: %s", cudaGetErrorString(ce));
            const cudaError_t ci = cudaGraphInstantiate(&e->graphExec, graph, 0);
            cudaGraphDestroy(graph);
            if (ci != cudaSuccess) { e->graphExec = nullptr; return fail(ORB_ERR_CUDA, "orbx_extract_batch: cudaGraphInstantiate: %s", cudaGetErrorString(ci)); }
            e->graphKey = key;
            e->graphLaunches = e->launches;
            ORB_CUDA(cudaGraphLaunch(e->graphExec, st));
        } else {
            e->graphSeen = key;
            timedCall = true;
            ORB_CHECK(copy_in());
            ORB_CHECK(enqueue(e, dImg, nFrames, w, h, stride, imgBytes, dK, dD, capacity, dN, st, true, 0));
            if (stage) ORB_CHECK(copy_out_staged());
        }
        if (stage) {
            ORB_CUDA(cudaStreamSynchronize(st));
            for (int f = 0; f < nFrames; ++f) {
                nOut[f] = pn[f];
                const size_t n = (size_t)std::min(std::max(pn[f], 0), capacity);
                std::memcpy(kps + (size_t)f * capacity, pk + (size_t)f * capacity, n * sizeof(orb_keypoint));
                std::memcpy(desc + (size_t)f * capacity * 32, pd + (size_t)f * capacity * 32, n * 32);
            }
        } else {
            ORB_CUDA(cudaMemcpyAsync(nOut, dN, (size_t)nFrames * 4, cudaMemcpyDeviceToHost, st));
            ORB_CUDA(cudaMemcpyAsync(kps, dK, (size_t)nFrames * capacity * sizeof(orb_keypoint), cudaMemcpyDeviceToHost, st));
            ORB_CUDA(cudaMemcpyAsync(desc, dD, (size_t)nFrames * capacity * 32, cudaMemcpyDeviceToHost, st));
            ORB_CUDA(cudaStreamSynchronize(st));
        }
        if (timedCall) {
            float ms;
            for (int i = 0; i < 3; ++i)
                if (cudaEventElapsedTime(&ms, e->ev[i], e->ev[i + 1]) == cudaSuccess) e->stageMs[i] = ms;
        }
        int status = ORB_OK;
        for (int f = 0; f < nFrames; ++f)
            if (nOut[f] > capacity) {
                status = fail(ORB_ERR_CAPACITY, "frame %d has %d keypoints, capacity %d (see orbx_keypoint_capacity)", f, nOut[f], capacity);
                nOut[f] = capacity;
            }
        return status;
    }
    // Pipeline chunk schedule.  A chunk's kernels can start only when its copy has landed, and the link delivers frames
    // about as fast as the kernels consume them, so every GROWTH in chunk size leaves the GPU idle for the difference;
    // small chunks, on the other hand, pay the per-launch-set overhead (~0.15 ms) more often.  Start small (the pipeline
    // fills quickly), grow by 10-20 % per chunk up to a cap: 31.5 ms instead of 33.2 ms for 4096 EuRoC frames.
    const int nPipe = std::min(nFrames, super);
    int chunk0 = std::max(16, std::min(96, (int)(2.0 * std::sqrt((double)nPipe))));
    if (const char* envChunk = getenv("ORBB_PIPE_CHUNK")) chunk0 = std::max(1, std::min(atoi(envChunk), super));   // tuning knob
    const double chunkGrowth = nPipe >= 2048 ? 1.1 : 1.2;
    const int chunkCap = std::min(768, 8 * chunk0);
    int nStreams = 2, tailMin = 128;
    if (const char* v = getenv("ORBB_PIPE_STREAMS")) nStreams = atoi(v) == 1 ? 1 : 2;       // tuning knobs
    if (const char* v = getenv("ORBB_PIPE_TAIL")) tailMin = std::max(1, atoi(v));
    int tailDiv = 6;
    if (const char* v = getenv("ORBB_PIPE_TAILDIV")) tailDiv = std::max(2, atoi(v));
    if (nStreams == 1) tailMin = 1 << 30;
    auto pipe_event = [&](int i) -> cudaEvent_t {
        while ((int)e->pipeEvents.size() <= i) {
            cudaEvent_t ev;
            if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return nullptr;
            e->pipeEvents.push_back(ev);
        }
        return e->pipeEvents[i];
    };
    const int maxChunks = nPipe / std::max(1, std::min(chunk0, tailMin)) + 24;
    int status = ORB_OK;
    static const bool trace = getenv("ORBB_PIPE_TRACE") != nullptr;   // debugging aid: per-chunk timeline on stderr
    std::vector<cudaEvent_t> tev;
    if (trace) {
        tev.resize(3 * maxChunks + 1);
        for (auto& ev : tev) cudaEventCreate(&ev);
        cudaEventRecord(tev[3 * maxChunks], e->streamIn);
    }
    for (int s0 = 0; s0 < nFrames; s0 += super) {
        const int ns = std::min(super, nFrames - s0);
        int k = 0;
        double want = chunk0;
        for (int f0 = 0, nf = 0; f0 < ns; f0 += nf, ++k) {
            const int remaining = ns - f0;
            nf = std::min(std::max(1, (int)want), remaining);
            // what is left after the last copy -- the last chunks' kernels and their copies back -- is pure latency: a chunk
            // takes at most a sixth of what remains, down to tailMin frames (measured: 1/2 137.3 k, 1/3 140.0 k, 1/4 141.3 k,
            // 1/6 144.1 k, 1/8 143.4 k frames/s end to end on 4096 EuRoC frames)
            if (nStreams == 2) nf = std::min(nf, std::max(tailMin, remaining / tailDiv));
            if (remaining - nf > 0 && remaining - nf < std::min(nf / 3, tailMin)) nf = remaining;    // no sliver at the end
            want = std::min(want * chunkGrowth, (double)chunkCap);
            if (!pipe_event(2 * k + 1)) return fail(ORB_ERR_CUDA, "orbx_extract_batch: cannot create pipeline events");
            uint8_t* dImg = e->dImages.as<uint8_t>() + (size_t)f0 * imgBytes;
            const uint8_t* src = images + (size_t)(s0 + f0) * frameStride;
            if (frameStride == imgBytes) {
                ORB_CUDA(cudaMemcpyAsync(dImg, src, (size_t)nf * imgBytes, cudaMemcpyHostToDevice, e->streamIn));
            } else {
                ORB_CUDA(cudaMemcpy2DAsync(dImg, imgBytes, src, frameStride, imgBytes, nf, cudaMemcpyHostToDevice, e->streamIn));
            }
            ORB_CUDA(cudaEventRecord(e->pipeEvents[2 * k], e->streamIn));
            if (trace) cudaEventRecord(tev[3 * k], e->streamIn);
            // Two compute streams, alternating: sharing the SMs costs the kernels ~8 %, but since the kernels outrun the
            // link the stream that is ahead only ever waits for its copy, while a chunk's launch gaps and kernel tails
            // (~0.2 ms per chunk) hide under the other chunk's kernels -- which is what lets the chunks shrink towards
            // the end of the batch (below) without the kernels falling behind the copies.
            cudaStream_t cs = (nStreams == 2 && (k & 1)) ? e->stream2 : e->stream;
            ORB_CUDA(cudaStreamWaitEvent(cs, e->pipeEvents[2 * k], 0));
            orb_keypoint* dK = e->dKps.as<orb_keypoint>() + (size_t)f0 * capacity;
            uint8_t* dD = e->dDesc.as<uint8_t>() + (size_t)f0 * capacity * 32;
            int* dN = e->dCount.as<int>() + f0;
            ORB_CHECK(enqueue(e, dImg, nf, w, h, stride, imgBytes, dK, dD, capacity, dN, cs, s0 == 0 && f0 == 0, f0,
                              cs == e->stream2 ? 1 : 0));
            ORB_CUDA(cudaEventRecord(e->pipeEvents[2 * k + 1], cs));
            if (trace) cudaEventRecord(tev[3 * k + 1], cs);
            ORB_CUDA(cudaStreamWaitEvent(e->streamOut, e->pipeEvents[2 * k + 1], 0));
            const size_t o = (size_t)(s0 + f0);
            ORB_CUDA(cudaMemcpyAsync(nOut + o, dN, (size_t)nf * 4, cudaMemcpyDeviceToHost, e->streamOut));
            ORB_CUDA(cudaMemcpyAsync(kps + o * capacity, dK, (size_t)nf * capacity * sizeof(orb_keypoint), cudaMemcpyDeviceToHost, e->streamOut));
            ORB_CUDA(cudaMemcpyAsync(desc + o * capacity * 32, dD, (size_t)nf * capacity * 32, cudaMemcpyDeviceToHost, e->streamOut));
            if (trace) cudaEventRecord(tev[3 * k + 2], e->streamOut);
        }
        // the next super-chunk reuses the arena and the staging buffers: drain everything first
        ORB_CUDA(cudaStreamSynchronize(e->streamOut));
        ORB_CUDA(cudaStreamSynchronize(e->stream));
        ORB_CUDA(cudaStreamSynchronize(e->stream2));
        if (trace && s0 == 0) {
            for (int c = 0; c < k; ++c) {
                float a = 0, b = 0, d = 0;
                cudaEventElapsedTime(&a, tev[3 * maxChunks], tev[3 * c]);
                cudaEventElapsedTime(&b, tev[3 * maxChunks], tev[3 * c + 1]);
                cudaEventElapsedTime(&d, tev[3 * maxChunks], tev[3 * c + 2]);
                fprintf(stderr, "chunk %2d  h2d done %7.3f  compute done %7.3f  d2h done %7.3f ms\n", c, a, b, d);
            }
        }
        if (s0 == 0) {
            float ms;
            for (int i = 0; i < 3; ++i)
                if (cudaEventElapsedTime(&ms, e->ev[i], e->ev[i + 1]) == cudaSuccess) e->stageMs[i] = ms;
        }
        for (int f = 0; f < ns; ++f)
            if (nOut[s0 + f] > capacity) {
                status = fail(ORB_ERR_CAPACITY, "frame %d has %d keypoints, capacity %d (see orbx_keypoint_capacity)", s0 + f,
                              nOut[s0 + f], capacity);
                nOut[s0 + f] = capacity;
            }
    }
    for (auto& ev : tev) cudaEventDestroy(ev);
    return status;
}

int orbx_extract(orbx_handle e, const uint8_t* image, int w, int h, int stride, orb_keypoint* kps, uint8_t* desc,
                 int capacity, int* nOut) {
    return orbx_extract_batch(e, image, 1, w, h, stride, (size_t)stride * (h > 0 ? h : 0), kps, desc, capacity, nOut);
}

// Frame::ComputeStereoMatches (Frame.cc:810-984): the pyramids are read where the two extractors' last calls left them
static int stereo_params(orbx_extractor* l, int frameL, orbx_extractor* r, int frameR, float mb, float mbf, StereoParams* S,
                         const char* who) {
    if (l->device != r->device) return fail(ORB_ERR_INVALID, "%s: the two extractors live on different devices", who);
    if (l->curW <= 0 || r->curW <= 0) return fail(ORB_ERR_INVALID, "%s: extract first (no pyramid yet)", who);
    if (l->curW != r->curW || l->curH != r->curH || l->nlevels != r->nlevels || l->scaleFactor != r->scaleFactor)
        return fail(ORB_ERR_INVALID, "%s: left and right extractor differ in image size, levels or scale factor", who);
    // only the frames of the last extraction call are resident (of a call larger than max_batch: its last max_batch-sized part)
    if (frameL < 0 || frameL >= l->residentFrames || frameR < 0 || frameR >= r->residentFrames)
        return fail(ORB_ERR_INVALID, "%s: frame index outside the last batch (%d / %d frames resident)", who, l->residentFrames,
                    r->residentFrames);
    if (!(mb > 0.0f) || !(mbf > 0.0f)) return fail(ORB_ERR_INVALID, "%s: baseline mb and mbf must be positive", who);
    const ExtractParams& P = l->P;
    S->pyrL = P.pyr + (size_t)frameL * P.pyrFrameBytes;
    S->pyrR = r->P.pyr + (size_t)frameR * r->P.pyrFrameBytes;
    S->nLevels = l->nlevels;
    S->nRows = P.lv[0].h;
    S->maxD = mbf / mb;          // Frame.cc:841-843 (minZ = mb)
    S->mbf = mbf;
    for (int i = 0; i < l->nlevels; ++i) {
        S->lvOff[i] = P.lv[i].pyrOff;
        S->pitch[i] = P.lv[i].pitch;
        S->cols[i] = P.lv[i].w;
        S->rows[i] = P.lv[i].h;
        S->scale[i] = l->scale[i];
        S->invScale[i] = l->invScale[i];
    }
    return ORB_OK;
}

int orbx_compute_stereo_matches_device(orbx_handle left, int frameL, orbx_handle right, int frameR,
                                       const orb_keypoint* dKeysL, const uint8_t* dDescL, const int* dNL, int capL,
                                       const orb_keypoint* dKeysR, const uint8_t* dDescR, const int* dNR, int capR,
                                       float mb, float mbf, float* dURight, float* dDepth, int* dSad, int* dKept,
                                       void* stream) {
    ORBX_ENTER(left);
    if (!right) return fail(ORB_ERR_INVALID, "orbx_compute_stereo_matches_device: null right extractor");
    if (!dKeysL || !dDescL || !dKeysR || !dDescR || !dURight || !dDepth || !dSad || !dKept || capL < 1 || capR < 1)
        return fail(ORB_ERR_INVALID, "orbx_compute_stereo_matches_device: bad arguments");
    if (capR > 65535) return fail(ORB_ERR_INVALID, "orbx_compute_stereo_matches_device: at most 65535 right keypoints");
    StereoParams S;
    ORB_CHECK(stereo_params(left, frameL, right, frameR, mb, mbf, &S, "orbx_compute_stereo_matches_device"));
    left->launches = 0;
    cudaStream_t st = stream ? (cudaStream_t)stream : left->stream;
    return launch_stereo(S, dKeysL, dDescL, capL, dNL, dKeysR, dDescR, capR, dNR, dURight, dDepth, dSad, dKept, st,
                         &left->launches);
}

int orbx_compute_stereo_matches(orbx_handle left, int frameL, orbx_handle right, int frameR, const orb_keypoint* keysL,
                                const uint8_t* descL, int nL, const orb_keypoint* keysR, const uint8_t* descR, int nR,
                                float mb, float mbf, float* uRight, float* depth, int* nMatches) {
    ORBX_ENTER(left);
    if (!right) return fail(ORB_ERR_INVALID, "orbx_compute_stereo_matches: null right extractor");
    if (nL < 0 || nR < 0 || (nL > 0 && (!keysL || !descL || !uRight || !depth)) || (nR > 0 && (!keysR || !descR)))
        return fail(ORB_ERR_INVALID, "orbx_compute_stereo_matches: bad arguments");
    if (nR > 65535) return fail(ORB_ERR_INVALID, "orbx_compute_stereo_matches: at most 65535 right keypoints");
    StereoParams S;
    ORB_CHECK(stereo_params(left, frameL, right, frameR, mb, mbf, &S, "orbx_compute_stereo_matches"));
    if (nMatches) *nMatches = 0;
    if (nL == 0) return ORB_OK;
    orbx_extractor* e = left;
    e->launches = 0;
    cudaStream_t st = e->stream;
    ORB_CUDA(cudaStreamSynchronize(right->stream));     // the right pyramid must be complete
    ORB_CHECK(e->stKeysL.reserve((size_t)nL * sizeof(orb_keypoint)));
    ORB_CHECK(e->stDescL.reserve((size_t)nL * 32));
    ORB_CHECK(e->stKeysR.reserve((size_t)(nR + 1) * sizeof(orb_keypoint)));
    ORB_CHECK(e->stDescR.reserve((size_t)(nR + 1) * 32));
    ORB_CHECK(e->stOut.reserve((size_t)nL * 12 + 16));
    ORB_CUDA(cudaMemcpyAsync(e->stKeysL.p, keysL, (size_t)nL * sizeof(orb_keypoint), cudaMemcpyHostToDevice, st));
    ORB_CUDA(cudaMemcpyAsync(e->stDescL.p, descL, (size_t)nL * 32, cudaMemcpyHostToDevice, st));
    if (nR > 0) {
        ORB_CUDA(cudaMemcpyAsync(e->stKeysR.p, keysR, (size_t)nR * sizeof(orb_keypoint), cudaMemcpyHostToDevice, st));
        ORB_CUDA(cudaMemcpyAsync(e->stDescR.p, descR, (size_t)nR * 32, cudaMemcpyHostToDevice, st));
    }
    float* dU = e->stOut.as<float>();
    float* dD = dU + nL;
    int* dSad = (int*)(dD + nL);
    int* dKept = dSad + nL;
    ORB_CHECK(launch_stereo(S, e->stKeysL.as<orb_keypoint>(), e->stDescL.as<unsigned char>(), nL, nullptr,
                            e->stKeysR.as<orb_keypoint>(), e->stDescR.as<unsigned char>(), nR, nullptr, dU, dD, dSad, dKept,
                            st, &e->launches));
    int kept = 0;
    ORB_CUDA(cudaMemcpyAsync(uRight, dU, (size_t)nL * 4, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaMemcpyAsync(depth, dD, (size_t)nL * 4, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaMemcpyAsync(&kept, dKept, 4, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    if (nMatches) *nMatches = kept;
    return ORB_OK;
}

int orbx_synchronize(orbx_handle e) {
    ORBX_ENTER(e);
    ORB_CUDA(cudaStreamSynchronize(e->stream));
    return ORB_OK;
}

int orbx_get_levels(orbx_handle e, int* nlevels) {
    if (!e || !nlevels) return fail(ORB_ERR_INVALID, "orbx_get_levels: null argument");
    *nlevels = e->nlevels;
    return ORB_OK;
}

int orbx_get_scale_tables(orbx_handle e, float* sf, float* inv, float* s2, float* is2, int* perLevel, int* umax) {
    if (!e) return fail(ORB_ERR_INVALID, "orbx_get_scale_tables: null handle");
    for (int i = 0; i < e->nlevels; ++i) {
        if (sf) sf[i] = e->scale[i];
        if (inv) inv[i] = e->invScale[i];
        if (s2) s2[i] = e->sigma2[i];
        if (is2) is2[i] = e->invSigma2[i];
        if (perLevel) perLevel[i] = e->perLevel[i];
    }
    if (umax) std::memcpy(umax, e->umax, sizeof e->umax);
    return ORB_OK;
}

int orbx_get_level(orbx_handle e, int frame, int level, uint8_t* out, int* w, int* hgt) {
    ORBX_ENTER(e);
    if (e->curW < 0) return fail(ORB_ERR_INVALID, "orbx_get_level: no image processed yet");
    if (level < 0 || level >= e->nlevels || frame < 0 || frame >= std::max(e->residentFrames, 1))
        return fail(ORB_ERR_INVALID, "orbx_get_level: bad frame/level (%d frames resident)", e->residentFrames);
    const LevelGeom& L = e->P.lv[level];
    if (w) *w = L.w;
    if (hgt) *hgt = L.h;
    if (!out) return ORB_OK;
    if (frame >= e->residentFrames) return fail(ORB_ERR_INVALID, "orbx_get_level: frame %d is not resident (%d are)", frame, e->residentFrames);
    ORB_CUDA(cudaStreamSynchronize(e->stream));
    const unsigned char* src = e->P.pyr + (size_t)frame * e->P.pyrFrameBytes + L.pyrOff + (kPadLeft - kEdge);
    ORB_CUDA(cudaMemcpy2D(out, L.w + 2 * kEdge, src, L.pitch, L.w + 2 * kEdge, L.h + 2 * kEdge, cudaMemcpyDeviceToHost));
    return ORB_OK;
}

int orbx_stage_times(orbx_handle e, double* ms3) {
    if (!e || !ms3) return fail(ORB_ERR_INVALID, "orbx_stage_times: null argument");
    for (int i = 0; i < 3; ++i) ms3[i] = e->stageMs[i];
    return ORB_OK;
}

int orbx_debug_candidates(orbx_handle e, int frame, int level, orb_keypoint* out, int cap, int* n) {
    ORBX_ENTER(e);
    if (e->curW < 0 || level < 0 || level >= e->nlevels || frame < 0 || frame >= e->residentFrames || !n)
        return fail(ORB_ERR_INVALID, "orbx_debug_candidates: bad arguments");
    const LevelGeom& L = e->P.lv[level];
    ORB_CUDA(cudaStreamSynchronize(e->stream));
    std::vector<int> counts(L.nCells);
    std::vector<unsigned int> slots((size_t)L.nCells * L.slotCap);
    if (L.nCells) {
        ORB_CUDA(cudaMemcpy(counts.data(), e->P.cellCount + (size_t)frame * e->P.nCellsTotal + L.cellBase, (size_t)L.nCells * 4,
                            cudaMemcpyDeviceToHost));
        ORB_CUDA(cudaMemcpy(slots.data(), e->P.slots + (size_t)frame * e->P.slotFrameEntries + L.slotBase, slots.size() * 4,
                            cudaMemcpyDeviceToHost));
    }
    int k = 0;
    for (int c = 0; c < L.nCells; ++c)
        for (int i = 0; i < counts[c]; ++i, ++k) {
            if (k >= cap) continue;
            const unsigned int v = slots[(size_t)c * L.slotCap + i];
            orb_keypoint o;
            o.x = (float)(v >> 20); o.y = (float)((v >> 8) & 0xfff); o.size = 7.f; o.angle = -1.f;
            o.response = (float)(v & 0xff); o.octave = 0; o.class_id = -1;
            out[k] = o;
        }
    *n = k;
    return ORB_OK;
}

int orbx_debug_blurred(orbx_handle e, int frame, int level, uint8_t* out) {
    ORBX_ENTER(e);
    if (e->curW < 0 || level < 0 || level >= e->nlevels || frame < 0 || frame >= e->residentFrames || !out)
        return fail(ORB_ERR_INVALID, "orbx_debug_blurred: bad arguments");
    const LevelGeom& L = e->P.lv[level];
    ORB_CUDA(cudaStreamSynchronize(e->stream));
    ORB_CUDA(cudaMemcpy2D(out, L.w, e->P.blur + (size_t)frame * e->P.blurFrameBytes + L.blurOff, L.bpitch, L.w, L.h,
                          cudaMemcpyDeviceToHost));
    return ORB_OK;
}

int orbx_set_profiling(orbx_handle e, int on) {
    if (!e) return fail(ORB_ERR_INVALID, "orbx_set_profiling: null handle");
    e->profiling = on != 0;
    e->profCalls = 0;
    return ORB_OK;
}

int orbx_kernel_times(orbx_handle e, double* ms5, int* nCalls) {
    ORBX_ENTER(e);
    if (!ms5 || !nCalls) return fail(ORB_ERR_INVALID, "orbx_kernel_times: null argument");
    for (int k = 0; k < 5; ++k) ms5[k] = 0;
    for (int c = 0; c < e->profCalls; ++c) {
        cudaEvent_t* pe = &e->profEvents[6 * c];
        ORB_CUDA(cudaEventSynchronize(pe[5]));
        for (int k = 0; k < 5; ++k) {
            float ms = 0;
            ORB_CUDA(cudaEventElapsedTime(&ms, pe[k], pe[k + 1]));
            ms5[k] += ms;
        }
    }
    *nCalls = e->profCalls;
    e->profCalls = 0;
    return ORB_OK;
}

int orbx_last_device_outputs(orbx_handle e, int frame, const orb_keypoint** dKeys, const uint8_t** dDesc, const int** dCount,
                             int* capacity, void** stream) {
    if (!e || !dKeys || !dDesc || !dCount || !capacity) return fail(ORB_ERR_INVALID, "orbx_last_device_outputs: null argument");
    if (!e->lastWasHost) return fail(ORB_ERR_INVALID, "orbx_last_device_outputs: the last call on this handle was not a host extract call");
    if (frame < 0 || frame >= e->residentFrames)
        return fail(ORB_ERR_INVALID, "orbx_last_device_outputs: frame %d is not resident (%d are)", frame, e->residentFrames);
    *dKeys = e->dKps.as<orb_keypoint>() + (size_t)frame * e->lastCapacity;
    *dDesc = e->dDesc.as<uint8_t>() + (size_t)frame * e->lastCapacity * 32;
    *dCount = e->dCount.as<int>() + frame;
    *capacity = e->lastCapacity;
    if (stream) *stream = e->stream;
    return ORB_OK;
}

int orbx_last_launch_count(orbx_handle e, int* n) {
    if (!e || !n) return fail(ORB_ERR_INVALID, "orbx_last_launch_count: null argument");
    *n = e->launches;
    return ORB_OK;
}

}  // extern "C"
