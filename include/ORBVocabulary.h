/*
 * ORBVocabulary.h — header-only C++ shim for ORB_SLAM::ORBVocabulary (caomw/ORBSLAM_jpMiniPC include/ORBVocabulary.h:31-32 =
 * DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB>) on top of the C ABI of liborb_b200.so, for the calls the
 * accelerated path makes: loadFromTextFile (src/main.cc:85-97), transform(features, BowVector, FeatureVector, levelsup)
 * (src/Frame.cc:285, src/KeyFrame.cc:63) and score(v1, v2) (src/KeyFrameDatabase.cc:140,247, src/LoopClosing.cc:139).
 * BowVector / FeatureVector keep DBoW2's std::map types so call sites compile unchanged; descriptors are passed as the
 * N x 32 byte matrix Frame::mDescriptors already is (the reference first splits it into rows, src/Converter.cc:28-37).
 */
#ifndef ORBVOCABULARY_H
#define ORBVOCABULARY_H

#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>
#include "orb_b200.h"

#ifdef ORB_B200_WITH_REFERENCE_TYPES
// the reference's own containers (Frame.h / KeyFrame.h include them too), so that mBowVec / mFeatVec keep their types
#include <opencv2/core/core.hpp>
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"
// The reference's headers name vector / list / set / pair unqualified (include/Frame.h:111, include/KeyFrame.h:139,
// include/KeyFrameDatabase.h:66): they rely on the using-directive that the reference's ORBVocabulary.h leaks through
// Thirdparty/DBoW2/DBoW2/FORB.h and TemplatedVocabulary.h.  A header that replaces it has to keep that (unfortunate) contract.
#include <list>
#include <set>
#include <vector>
using namespace std;
#else
namespace DBoW2
{
typedef unsigned int WordId;
typedef double WordValue;
typedef unsigned int NodeId;
class BowVector : public std::map<WordId, WordValue> {};                         // Thirdparty/DBoW2/DBoW2/BowVector.h:56
class FeatureVector : public std::map<NodeId, std::vector<unsigned int> > {};   // Thirdparty/DBoW2/DBoW2/FeatureVector.h:21
}
#endif

namespace ORB_SLAM
{

class ORBVocabulary
{
public:
    explicit ORBVocabulary(orb_ctx* context) : ctx(context), voc(nullptr) {}
    // the reference's default constructor (src/main.cc:85: `ORBVocabulary Vocabulary;`): runs on the process-wide default context
    ORBVocabulary() : ctx(orb_default_context()), voc(nullptr)
    {
        if (!ctx) throw std::runtime_error(std::string("ORBVocabulary: no context (") + orb_last_cuda_error() + ")");
    }
    ~ORBVocabulary() { orb_vocab_destroy(voc); }
    ORBVocabulary(const ORBVocabulary&) = delete;
    ORBVocabulary& operator=(const ORBVocabulary&) = delete;

    bool loadFromTextFile(const std::string& filename)
    {
        orb_vocab_destroy(voc);
        voc = nullptr;
        return orb_vocab_load_text(ctx, filename.c_str(), &voc) == ORB_OK;
    }
    bool empty() const { int nw = 0; return !voc || orb_vocab_info(voc, nullptr, nullptr, nullptr, &nw) != ORB_OK || nw == 0; }
    unsigned int size() const { int nw = 0; if (voc) orb_vocab_info(voc, nullptr, nullptr, nullptr, &nw); return (unsigned)nw; }

    // transform(features, v, fv, levelsup): descriptors = N x 32 bytes
    void transform(const unsigned char* descriptors, int n, DBoW2::BowVector& v, DBoW2::FeatureVector& fv, int levelsup) const
    {
        v.clear();
        fv.clear();
        if (empty() || n <= 0) return;
        const int cap = n;
        std::vector<int32_t> bw(cap), fn(cap), fs(cap + 1), fi(cap);
        std::vector<double> bv(cap);
        int32_t counts = n, nb = 0, nf = 0;
        check(orb_vocab_transform_batch(ctx, voc, descriptors, n, &counts, 1, levelsup, cap, bw.data(), bv.data(), &nb, fn.data(),
                                        fs.data(), fi.data(), &nf));
        for (int i = 0; i < nb; i++) v.insert(v.end(), std::make_pair((DBoW2::WordId)bw[i], bv[i]));
        for (int j = 0; j < nf; j++)
            fv.insert(fv.end(), std::make_pair((DBoW2::NodeId)fn[j], std::vector<unsigned int>(fi.begin() + fs[j], fi.begin() + fs[j + 1])));
    }

#ifdef ORB_B200_WITH_REFERENCE_TYPES
    // the reference's call (src/Frame.cc:284-285, src/KeyFrame.cc:62-63): one 1 x 32 CV_8U row per feature
    void transform(const std::vector<cv::Mat>& features, DBoW2::BowVector& v, DBoW2::FeatureVector& fv, int levelsup) const
    {
        std::vector<unsigned char> rows(features.size() * 32 + 32);
        for (size_t i = 0; i < features.size(); i++) std::memcpy(&rows[i * 32], features[i].ptr<unsigned char>(), 32);
        transform(rows.data(), (int)features.size(), v, fv, levelsup);
    }
    // transform(feature) -> word id, Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1197-1205
    DBoW2::WordId transform(const cv::Mat& feature) const
    {
        int32_t w = 0, node = 0;
        double weight = 0;
        if (empty()) return 0;
        check(orb_vocab_transform_features(ctx, voc, feature.ptr<unsigned char>(), 1, 0, &w, &weight, &node));
        return (DBoW2::WordId)w;
    }
#endif

    // score(v1, v2), rounded to float as every caller stores it (float si = mpVoc->score(...))
    double score(const DBoW2::BowVector& v1, const DBoW2::BowVector& v2) const
    {
        std::vector<int32_t> w1, w2, start(2, 0);
        std::vector<double> x1, x2;
        for (DBoW2::BowVector::const_iterator it = v1.begin(); it != v1.end(); ++it) { w1.push_back((int32_t)it->first); x1.push_back(it->second); }
        for (DBoW2::BowVector::const_iterator it = v2.begin(); it != v2.end(); ++it) { w2.push_back((int32_t)it->first); x2.push_back(it->second); }
        start[1] = (int32_t)w2.size();
        int32_t common = 0; float s = 0.f; int mx = 0;
        check(orb_bow_score_db(ctx, voc, w1.data(), x1.data(), (int)w1.size(), 1, start.data(), w2.data(), x2.data(), 1, &common, &s, &mx));
        return s;
    }

    orb_vocab* handle() const { return voc; }

private:
    static void check(int status)
    {
        if (status != ORB_OK) throw std::runtime_error(std::string("ORBVocabulary: ") + orb_error_string(status) + " (" + orb_last_cuda_error() + ")");
    }
    orb_ctx* ctx;
    orb_vocab* voc;
};

} // namespace ORB_SLAM

#endif
