// TEST INFRASTRUCTURE ONLY -- driver for oracle/_ref/orb_ref: the reference's OWN src/ORBextractor.cc, compiled in
// place from /root/reference against the OpenCV stand-in in include/cvshim.hpp.  Used (a) to pin oracle/orb_oracle.cpp
// against the real reference code and (b) as the "reference" CPU baseline of bench.py.
//
// Address rule: ORBextractor.cc:686 sorts pair<int, ExtractorNode*>, so equal-size nodes are ordered by heap address.
// Every allocation of this process comes from a bump arena (monotonically increasing addresses, free is a no-op),
// which makes "address order" == "creation order", the rule oracle/orb_oracle.cpp and the CUDA path implement.
//
// File protocol (little endian int32 / float32), see tests/ref_runner.py:
//   request : magic 0x0RB1, mode, w, h, nframes, nfeatures, nlevels, iniTh, minTh, dumpLevels, scale(float), frames...
//   mode 0 reply per frame: n, n*28 B keypoints, n*32 B descriptors, [nlevels x (w, h, (w+38)*(h+38) B)]
//   mode 1 request tail   : minX, maxX, minY, maxY, N, nkeys, keys ; reply: n, keys
#include <sys/mman.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <vector>

#include "ORBextractor.h"

// ---------------------------------------------------------------------------------------------- bump arena
static char* g_base = nullptr;
static size_t g_off = 0, g_cap = 0;
static void* bump(size_t n) {
    if (!g_base) {
        g_cap = (size_t)8 << 30;
        g_base = (char*)mmap(nullptr, g_cap, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (g_base == (char*)MAP_FAILED) { std::fprintf(stderr, "arena mmap failed\n"); std::abort(); }
    }
    n = (n + 15) & ~(size_t)15;
    if (g_off + n > g_cap) { std::fprintf(stderr, "arena exhausted\n"); std::abort(); }
    void* p = g_base + g_off;
    g_off += n;
    return p;
}
void* operator new(size_t n) { return bump(n); }
void* operator new[](size_t n) { return bump(n); }
void operator delete(void*) noexcept {}
void operator delete[](void*) noexcept {}
void operator delete(void*, size_t) noexcept {}
void operator delete[](void*, size_t) noexcept {}

namespace {
struct Probe : ORB_SLAM2::ORBextractor {
    using ORB_SLAM2::ORBextractor::ORBextractor;
    std::vector<cv::KeyPoint> distribute(const std::vector<cv::KeyPoint>& k, int minX, int maxX, int minY, int maxY, int N) {
        return DistributeOctTree(k, minX, maxX, minY, maxY, N, 0);
    }
};
struct Req { int magic, mode, w, h, nframes, nfeatures, nlevels, iniTh, minTh, dumpLevels; float scale; };

bool readAll(FILE* f, void* p, size_t n) { return std::fread(p, 1, n, f) == n; }
}  // namespace

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: orb_ref <request> <reply> [bench_iters]\n"); return 2; }
    FILE* fi = std::fopen(argv[1], "rb");
    if (!fi) { std::perror("request"); return 2; }
    Req rq;
    if (!readAll(fi, &rq, sizeof rq) || rq.magic != 0x0B1) { std::fprintf(stderr, "bad request\n"); return 2; }
    const int iters = argc > 3 ? std::atoi(argv[3]) : 0;

    if (rq.mode == 1) {
        int p[6];
        readAll(fi, p, sizeof p);
        cv::KeyPoint* keys = (cv::KeyPoint*)std::malloc(sizeof(cv::KeyPoint) * (size_t)(p[5] > 0 ? p[5] : 1));
        readAll(fi, keys, sizeof(cv::KeyPoint) * (size_t)p[5]);
        std::fclose(fi);
        Probe ex(rq.nfeatures, rq.scale, rq.nlevels, rq.iniTh, rq.minTh);
        std::vector<cv::KeyPoint> in(keys, keys + p[5]);
        std::vector<cv::KeyPoint> out = ex.distribute(in, p[0], p[1], p[2], p[3], p[4]);
        FILE* fo = std::fopen(argv[2], "wb");
        int n = (int)out.size();
        std::fwrite(&n, 4, 1, fo);
        std::fwrite(out.data(), sizeof(cv::KeyPoint), out.size(), fo);
        std::fclose(fo);
        return 0;
    }

    const size_t fsz = (size_t)rq.w * rq.h;
    uchar* frames = (uchar*)std::malloc(fsz * rq.nframes);
    if (!readAll(fi, frames, fsz * rq.nframes)) { std::fprintf(stderr, "short request\n"); return 2; }
    std::fclose(fi);
    const size_t mark = g_off;

    if (iters > 0) {   // timing mode: the reference's own three stage timers + wall time per frame
        double total = 0, st[3] = {0, 0, 0};
        long nkp = 0, count = 0;
        for (int it = 0; it < iters; ++it)
            for (int f = 0; f < rq.nframes; ++f) {
                g_off = mark;
                cv::Mat img(rq.h, rq.w, CV_8UC1);
                std::memcpy(img.data, frames + fsz * f, fsz);
                Probe ex(rq.nfeatures, rq.scale, rq.nlevels, rq.iniTh, rq.minTh);
                std::vector<cv::KeyPoint> kps;
                cv::Mat desc;
                const auto t0 = std::chrono::steady_clock::now();
                ex(img, cv::Mat(), kps, desc);
                const auto t1 = std::chrono::steady_clock::now();
                total += std::chrono::duration<double, std::milli>(t1 - t0).count();
                st[0] += ex.GetTimeOfComputePyramid();
                st[1] += ex.GetTimeOfComputeKeyPointsOctTree();
                st[2] += ex.GetTImeOfComputeDescriptor();
                nkp += (long)kps.size();
                ++count;
            }
        std::printf("{\"frames\": %ld, \"ms_per_frame\": %.6f, \"ms_pyramid\": %.6f, \"ms_keypoints\": %.6f, "
                    "\"ms_descriptors\": %.6f, \"kp_per_frame\": %.2f}\n",
                    count, total / count, st[0] / count, st[1] / count, st[2] / count, (double)nkp / count);
        return 0;
    }

    FILE* fo = std::fopen(argv[2], "wb");
    if (!fo) { std::perror("reply"); return 2; }
    for (int f = 0; f < rq.nframes; ++f) {
        g_off = mark;
        cv::Mat img(rq.h, rq.w, CV_8UC1);
        std::memcpy(img.data, frames + fsz * f, fsz);
        Probe ex(rq.nfeatures, rq.scale, rq.nlevels, rq.iniTh, rq.minTh);
        std::vector<cv::KeyPoint> kps;
        cv::Mat desc;
        ex(img, cv::Mat(), kps, desc);
        int n = (int)kps.size();
        std::fwrite(&n, 4, 1, fo);
        std::fwrite(kps.data(), sizeof(cv::KeyPoint), kps.size(), fo);
        for (int i = 0; i < n; ++i) std::fwrite(desc.ptr(i), 1, 32, fo);
        if (rq.dumpLevels)
            for (int l = 0; l < rq.nlevels; ++l) {
                const cv::Mat& L = ex.mvImagePyramid[l];
                int wh[2] = {L.cols, L.rows};
                std::fwrite(wh, 4, 2, fo);
                const uchar* top = L.data - 19 * (size_t)L.step - 19;   // the 19-px frame around the ROI (:1135-1136)
                for (int y = 0; y < L.rows + 38; ++y) std::fwrite(top + (size_t)y * L.step, 1, L.cols + 38, fo);
            }
    }
    std::fclose(fo);
    return 0;
}
