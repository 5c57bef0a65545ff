// Compiles the C++ adapter (reference class names and method names) against liborbb200.so and runs it once.
// argv[1] = raw 8-bit image file, argv[2] = width, argv[3] = height. Prints "n_keypoints checksum n_matches".
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ORBextractor.h"
#include "ORBmatcher.h"
#include "Frame.h"
#include "ORBVocabulary.h"

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const int w = atoi(argv[2]), h = atoi(argv[3]);
    std::vector<unsigned char> img((size_t)w * h);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(img.data(), 1, img.size(), f) != img.size()) return 2;
    fclose(f);
    try {
        ORB_SLAM2::ORBextractor extractor(1000, 1.2f, 8, 20, 7);
        std::vector<orb_keypoint> keys;
        std::vector<unsigned char> desc;
        extractor(img.data(), w, h, w, keys, desc);
        unsigned long long sum = 0;
        for (size_t i = 0; i < desc.size(); ++i) sum = sum * 131 + desc[i];
        ORB_SLAM2::ORBmatcher matcher(0.9f, true);
        ORB_SLAM2::FrameView F1(matcher.handle(), keys.data(), desc.data(), (int)keys.size(), 0, 0, (float)w, (float)h);
        ORB_SLAM2::FrameView F2(matcher.handle(), keys.data(), desc.data(), (int)keys.size(), 0, 0, (float)w, (float)h);
        std::vector<float> prev(keys.size() * 2);
        for (size_t i = 0; i < keys.size(); ++i) { prev[2 * i] = keys[i].x; prev[2 * i + 1] = keys[i].y; }
        std::vector<int> m12;
        const int n = matcher.SearchForInitialization(F1, F2, prev, m12, 100);
        printf("%zu %llu %d %d %.3f\n", keys.size(), sum, n, extractor.GetLevels(), extractor.GetScaleFactors()[7]);
        const int d = matcher.DescriptorDistance(desc.data(), desc.data() + 32);
        printf("%d\n", d);
        // Frame post-processing mirrors (adapter/Frame.h): third line = "kept minX maxY sum(mvKeysUn.x) best0 best1"
        namespace fo = ORB_SLAM2::frame_ops;
        ORB_SLAM2::ORBextractor right(1000, 1.2f, 8, 20, 7);
        std::vector<orb_keypoint> keysR;
        std::vector<unsigned char> descR;
        right(img.data(), w, h, w, keysR, descR);          // right view = left view: zero disparity
        std::vector<float> uRight, depth;
        const int kept = fo::ComputeStereoMatches(extractor, right, keys, desc, keysR, descR, 0.11f, 47.9f, uRight, depth);
        const orb_camera cam = fo::MakeCamera(458.654f, 457.296f, 367.215f, 248.375f, -0.28340811f, 0.07395907f, 0.00019359f, 1.76187114e-05f);
        float minX, maxX, minY, maxY;
        fo::ComputeImageBounds(matcher.handle(), cam, w, h, minX, maxX, minY, maxY);
        std::vector<orb_keypoint> keysUn;
        fo::UndistortKeyPoints(matcher.handle(), cam, keys, keysUn);
        double sx = 0;
        for (size_t i = 0; i < keysUn.size(); ++i) sx += keysUn[i].x;
        std::vector<int> start(3), best;
        start[0] = 0; start[1] = 5; start[2] = 12;
        std::vector<unsigned char> obs(desc.begin(), desc.begin() + 12 * 32);
        fo::ComputeDistinctiveDescriptors(matcher.handle(), obs, start, best);
        printf("%d %.9g %.9g %.17g %d %d\n", kept, minX, maxY, sx, best[0], best[1]);
        // ComputeBoW on a two-level tree whose 4 + 16 node descriptors are the first extracted descriptors;
        // fourth line = "n_words n_nodes sum(bowValue) fvIdx[0]"
        const int nNodes = 21;
        std::vector<unsigned char> ndesc(nNodes * 32, 0);
        for (int i = 1; i < nNodes; ++i) for (int b = 0; b < 32; ++b) ndesc[i * 32 + b] = desc[(size_t)(i * 7) * 32 + b];
        std::vector<int> cstart(nNodes + 1, 0), kids, wid(nNodes, 0);
        std::vector<double> wgt(nNodes, 0.0);
        for (int i = 1; i <= 4; ++i) kids.push_back(i);
        cstart[1] = 4;
        for (int p = 1; p <= 4; ++p) { for (int c = 0; c < 4; ++c) kids.push_back(5 + (p - 1) * 4 + c); cstart[p + 1] = (int)kids.size(); }
        for (int i = 5; i < nNodes; ++i) { cstart[i + 1] = (int)kids.size(); wid[i] = i - 5; wgt[i] = 1.0 + 0.25 * (i - 5); }
        orbm_vocabulary voc = nullptr;
        fo::check(orbm_vocabulary_create(matcher.handle(), nNodes, 2, ndesc.data(), cstart.data(), kids.data(), wid.data(), wgt.data(), &voc));
        fo::BowResult bow;
        fo::ComputeBoW(matcher.handle(), voc, desc, bow, 1);
        double sv = 0;
        for (size_t i = 0; i < bow.bowValue.size(); ++i) sv += bow.bowValue[i];
        printf("%zu %zu %.17g %d\n", bow.bowWord.size(), bow.fvNode.size(), sv, bow.fvIdx.empty() ? -1 : bow.fvIdx[0]);
        orbm_vocabulary_destroy(voc);
        // ORBVocabulary: argv[4] = vocabulary text file -> load, upload, transform(levelsup 2);
        // fifth line = "nodes words n_bow n_fv sum(bowValue) fvIdx[0]"
        if (argc > 4) {
            ORB_SLAM2::ORBVocabulary orbvoc;
            if (!orbvoc.loadFromTextFile(argv[4])) return 4;
            orbvoc.upload(matcher.handle());
            std::vector<int> bw, fn, fs, fi;
            std::vector<double> bv;
            orbvoc.transform(desc, bw, bv, fn, fs, fi, 2);
            double s2 = 0;
            for (size_t i = 0; i < bv.size(); ++i) s2 += bv[i];
            printf("%d %u %zu %zu %.17g %d\n", orbvoc.nodes(), orbvoc.size(), bw.size(), fn.size(), s2, fi.empty() ? -1 : fi[0]);
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
