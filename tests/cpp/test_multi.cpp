// C++ host of the multi-GPU half of the C ABI (include/orb_b200.h): ONE process drives every visible GPU, as a C++ host of the
// reference would (src/main.cc:165-212 is a single process).
//   1. database file (the fork's descriptor dump, orb_db_read_descriptors) -> orb_comm_db_upload -> orb_knn2_sharded,
//      compared bit for bit with ONE scan of the whole database on device 0 (orb_hamming_knn2);
//   2. frames.raw -> orb_extract_batch_multi, compared bit for bit with orb_extract_batch on device 0.
// Prints "PASS ranks=<n> transport=<t>" and writes the merged kNN result for the python test to compare with the oracle.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../include/orb_b200.h"

#define CHECK(x) do { int rc__ = (x); if (rc__ != ORB_OK) { std::fprintf(stderr, "%s -> %s [%s]\n", #x, orb_error_string(rc__), orb_last_cuda_error()); return 1; } } while (0)

int main(int argc, char** argv)
{
    if (argc < 10) { std::fprintf(stderr, "usage: %s ngpus db.bin queries.raw nq frames.raw w h nframes out.bin\n", argv[0]); return 2; }
    const int ngpus = std::atoi(argv[1]), nq = std::atoi(argv[4]), w = std::atoi(argv[6]), h = std::atoi(argv[7]), nfr = std::atoi(argv[8]);
    orb_comm* comm = orb_comm_init(ngpus);
    if (!comm) { std::fprintf(stderr, "orb_comm_init: %s\n", orb_last_cuda_error()); return 1; }
    const int ranks = orb_comm_size(comm);

    // ---- 1. sharded kNN-2 against one scan
    int64_t nrows = 0; int32_t nrec = 0;
    int rc = orb_db_read_descriptors(argv[2], nullptr, 0, nullptr, 0, &nrows, &nrec);
    if (rc != ORB_OK && rc != ORB_ERR_CAPACITY) { std::fprintf(stderr, "cannot size %s\n", argv[2]); return 1; }
    std::vector<uint8_t> db((size_t)nrows * 32 + 32);
    CHECK(orb_db_read_descriptors(argv[2], db.data(), nrows, nullptr, 0, &nrows, &nrec));
    std::vector<uint8_t> q((size_t)nq * 32);
    FILE* f = std::fopen(argv[3], "rb");
    if (!f || std::fread(q.data(), 32, nq, f) != (size_t)nq) { std::fprintf(stderr, "cannot read queries\n"); return 1; }
    std::fclose(f);
    CHECK(orb_comm_db_upload(comm, db.data(), nrows));
    std::vector<int32_t> mi(nq), m1(nq), m2(nq), si(nq), s1(nq), s2(nq);
    CHECK(orb_knn2_sharded(comm, q.data(), nq, mi.data(), m1.data(), m2.data()));
    CHECK(orb_hamming_knn2(orb_comm_context(comm, 0), q.data(), nq, db.data(), nrows, si.data(), s1.data(), s2.data()));
    if (mi != si || m1 != s1 || m2 != s2) { std::fprintf(stderr, "FAIL: sharded kNN differs from the single scan\n"); return 1; }

    // ---- 2. frame-sharded extraction against one device
    std::vector<uint8_t> frames((size_t)nfr * w * h);
    f = std::fopen(argv[5], "rb");
    if (!f || std::fread(frames.data(), (size_t)w * h, nfr, f) != (size_t)nfr) { std::fprintf(stderr, "cannot read frames\n"); return 1; }
    std::fclose(f);
    CHECK(orb_comm_set_extractor(comm, 500, 1.2f, 8, ORB_FAST_SCORE, 20, w, h, 4));
    const int cap = orb_keypoint_capacity(orb_comm_context(comm, 0));
    std::vector<orb_keypoint> ka((size_t)nfr * cap), kb((size_t)nfr * cap);
    std::vector<uint8_t> da((size_t)nfr * cap * 32), dbb((size_t)nfr * cap * 32);
    std::vector<int32_t> ca(nfr), cb(nfr);
    std::memset(ka.data(), 0, ka.size() * sizeof(orb_keypoint)); std::memset(kb.data(), 0, kb.size() * sizeof(orb_keypoint));
    CHECK(orb_extract_batch_multi(comm, frames.data(), nfr, w, h, w, (size_t)w * h, ka.data(), da.data(), cap, ca.data()));
    CHECK(orb_extract_batch(orb_comm_context(comm, 0), frames.data(), nfr, w, h, w, (size_t)w * h, kb.data(), dbb.data(), cap, cb.data()));
    if (ca != cb) { std::fprintf(stderr, "FAIL: frame-sharded keypoint counts differ\n"); return 1; }
    for (int i = 0; i < nfr; i++)
        if (std::memcmp(&ka[(size_t)i * cap], &kb[(size_t)i * cap], (size_t)ca[i] * sizeof(orb_keypoint)) ||
            std::memcmp(&da[(size_t)i * cap * 32], &dbb[(size_t)i * cap * 32], (size_t)ca[i] * 32)) {
            std::fprintf(stderr, "FAIL: frame %d differs between the sharded and the single-device run\n", i); return 1;
        }
    FILE* o = std::fopen(argv[9], "wb");
    std::fwrite(mi.data(), 4, nq, o); std::fwrite(m1.data(), 4, nq, o); std::fwrite(m2.data(), 4, nq, o);
    std::fwrite(ca.data(), 4, nfr, o);
    std::fclose(o);
    std::printf("PASS ranks=%d transport=%s rows=%lld queries=%d frames=%d\n", ranks, orb_comm_transport(comm), (long long)nrows, nq, nfr);
    orb_comm_destroy(comm);
    return 0;
}
