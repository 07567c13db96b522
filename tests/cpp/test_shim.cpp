// C++ host-side check of the drop-in shims: ORB_SLAM::ORBextractor / ORBmatcher over the C ABI.
// Reads a raw 8-bit frame, extracts, prints a checksum the python test compares with the oracle.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../include/ORBextractor.h"
#include "../../include/ORBmatcher.h"
#include "../../include/ORBVocabulary.h"

int main(int argc, char** argv)
{
    if (argc < 6) { std::fprintf(stderr, "usage: %s frame.raw w h nfeatures out.bin [vocabulary.txt]\n", argv[0]); return 2; }
    const int w = std::atoi(argv[2]), h = std::atoi(argv[3]), nf = std::atoi(argv[4]);
    std::vector<unsigned char> img((size_t)w * h);
    FILE* f = std::fopen(argv[1], "rb");
    if (!f || std::fread(img.data(), 1, img.size(), f) != img.size()) { std::fprintf(stderr, "cannot read frame\n"); return 2; }
    std::fclose(f);
    try {
        ORB_SLAM::ORBextractor extractor(nf, 1.2f, 8, ORB_SLAM::ORBextractor::FAST_SCORE, 20, 0, w, h, 1);
        std::vector<orb_keypoint> kps;
        std::vector<unsigned char> desc;
        extractor(img.data(), w, h, w, kps, desc);
        // match the frame against itself: every descriptor finds itself at distance 0
        ORB_SLAM::ORBmatcher matcher(extractor.context(), 0.9f, true);
        std::vector<int32_t> match;
        int n = matcher.MatchBruteForce(desc.data(), (int)kps.size(), desc.data(), (long long)kps.size(), 50, match);
        FILE* o = std::fopen(argv[5], "wb");
        int cnt = (int)kps.size();
        std::fwrite(&cnt, 4, 1, o); std::fwrite(&n, 4, 1, o);
        std::fwrite(kps.data(), sizeof(orb_keypoint), kps.size(), o);
        std::fwrite(desc.data(), 1, desc.size(), o);
        std::fwrite(match.data(), 4, match.size(), o);
        if (argc > 6) {                                   // vocabulary text file: Frame::ComputeBoW through the shim
            ORB_SLAM::ORBVocabulary voc(extractor.context());
            if (!voc.loadFromTextFile(argv[6])) { std::fprintf(stderr, "cannot load vocabulary\n"); return 1; }
            DBoW2::BowVector bow;
            DBoW2::FeatureVector fv;
            voc.transform(desc.data(), cnt, bow, fv, 1);
            int nb = (int)bow.size(), nfv = (int)fv.size();
            float self = (float)voc.score(bow, bow);
            std::fwrite(&nb, 4, 1, o); std::fwrite(&nfv, 4, 1, o); std::fwrite(&self, 4, 1, o);
            for (DBoW2::BowVector::const_iterator it = bow.begin(); it != bow.end(); ++it) { std::fwrite(&it->first, 4, 1, o); std::fwrite(&it->second, 8, 1, o); }
            for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
                int m = (int)it->second.size();
                std::fwrite(&it->first, 4, 1, o); std::fwrite(&m, 4, 1, o); std::fwrite(it->second.data(), 4, it->second.size(), o);
            }
        }
        std::fclose(o);
        std::printf("levels=%d scale=%.3f keypoints=%d selfmatches=%d\n", extractor.GetLevels(), extractor.GetScaleFactor(), cnt, n);
    } catch (const std::exception& e) { std::fprintf(stderr, "error: %s\n", e.what()); return 1; }
    return 0;
}
