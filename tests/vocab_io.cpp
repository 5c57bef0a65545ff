// CPU-only driver of adapter/ORBVocabulary.h's file readers and writers (no device calls):
//   vocab_io <in> <text|binary> <dump> <out.txt> <out.bin>
// loads <in>, dumps the flat arrays to <dump> (little-endian: n, k, L, scoring, weighting, n_children, then desc, parent,
// childStart, children, wordId, isWord (int32 each), weight (float64)), and writes the tree back in both formats.
// Exit code 3 = the reader refused the file.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <map>

#include "ORBVocabulary.h"

// the map-container overload of transform() must instantiate with DBoW2-shaped containers (it needs a device to run)
template void ORB_SLAM2::ORBVocabulary::transform(const std::vector<unsigned char>&, std::map<unsigned int, double>&,
                                                  std::map<unsigned int, std::vector<unsigned int> >&, int) const;

int main(int argc, char** argv) {
    if (argc < 6) return 2;
    ORB_SLAM2::ORBVocabulary voc;
    const bool ok = std::strcmp(argv[2], "binary") == 0 ? voc.loadFromBinaryFile(argv[1]) : voc.loadFromTextFile(argv[1]);
    if (!ok) return 3;
    FILE* f = fopen(argv[3], "wb");
    if (!f) return 2;
    const int n = voc.nodes();
    const int hdr[6] = {n, voc.m_k, voc.m_L, voc.m_scoring, voc.m_weighting, (int)voc.children.size()};
    fwrite(hdr, 4, 6, f);
    fwrite(voc.desc.data(), 1, voc.desc.size(), f);
    fwrite(voc.parent.data(), 4, n, f);
    fwrite(voc.childStart.data(), 4, n + 1, f);
    fwrite(voc.children.data(), 4, voc.children.size(), f);
    fwrite(voc.wordId.data(), 4, n, f);
    std::vector<int> w(voc.isWord.begin(), voc.isWord.end());
    fwrite(w.data(), 4, n, f);
    fwrite(voc.weight.data(), 8, n, f);
    fclose(f);
    voc.saveToTextFile(argv[4]);
    voc.saveToBinaryFile(argv[5]);
    printf("%d %u\n", n, voc.size());
    return 0;
}
