// Latency of the reference's own call pattern from C++ (no python in the way): ORB_SLAM::ORBextractor::operator() of the drop-in shim on one
// frame, std::vector outputs (pageable memory), and the same through the C ABI with pinned buffers.
//   g++ -std=c++17 -O2 -I include -o /tmp/orb_latency tools/cpp/latency.cpp -Lorbslam_jpminipc_b200 -lorb_b200 -Wl,-rpath,$PWD/orbslam_jpminipc_b200
//   /tmp/orb_latency [width height nfeatures [frame.raw]]      (frame.raw: width*height bytes, e.g. python -c "from orbslam_jpminipc_b200.synth import synth_frame; synth_frame(480, 640, 1000).tofile('/tmp/f.raw')")
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "ORBextractor.h"

static void report(const char* what, int w, int h, int nf, size_t n, std::vector<double>& v)
{
    std::sort(v.begin(), v.end());
    std::printf("%dx%d, %d features, %zu keypoints; %s: median %.1f us (p10 %.1f, p90 %.1f)\n", w, h, nf, n, what, v[v.size() / 2], v[v.size() / 10], v[v.size() * 9 / 10]);
}

int main(int argc, char** argv)
{
    const int w = argc > 1 ? atoi(argv[1]) : 640, h = argc > 2 ? atoi(argv[2]) : 480, nf = argc > 3 ? atoi(argv[3]) : 1000;
    // textured synthetic frame: smoothed noise (enough corners for the full feature quota)
    std::vector<unsigned char> img((size_t)w * h);
    std::mt19937 rng(7);
    std::vector<float> a((size_t)w * h);
    for (auto& x : a) x = (float)(rng() & 255);
    for (int y = 1; y + 1 < h; y++)
        for (int x = 1; x + 1 < w; x++) {
            float s = 0;
            for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) s += a[(size_t)(y + dy) * w + x + dx];
            img[(size_t)y * w + x] = (unsigned char)(s / 9.f);
        }
    if (argc > 4) {
        FILE* fp = std::fopen(argv[4], "rb");
        if (!fp || std::fread(img.data(), 1, img.size(), fp) != img.size()) { std::printf("cannot read %s\n", argv[4]); return 1; }
        std::fclose(fp);
    }
    ORB_SLAM::ORBextractor ex(nf, 1.2f, 8, ORB_SLAM::ORBextractor::FAST_SCORE, 20, 0, w, h, 1);
    std::vector<orb_keypoint> kps;
    std::vector<unsigned char> desc;
    for (int i = 0; i < 30; i++) ex(img.data(), w, h, w, kps, desc);
    std::vector<double> t;
    for (int i = 0; i < 500; i++) {
        const auto t0 = std::chrono::steady_clock::now();
        ex(img.data(), w, h, w, kps, desc);
        t.push_back(std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
    }
    report("ORBextractor::operator() (std::vector in / out)", w, h, nf, kps.size(), t);
    // C ABI, pinned buffers
    orb_ctx* c = orb_create(0, nf, 1.2f, 8, 1, 20, w, h, 1);
    const int cap = orb_keypoint_capacity(c);
    unsigned char* pin_img = (unsigned char*)orb_host_alloc((size_t)w * h);
    orb_keypoint* pin_k = (orb_keypoint*)orb_host_alloc((size_t)cap * sizeof(orb_keypoint));
    unsigned char* pin_d = (unsigned char*)orb_host_alloc((size_t)cap * 32);
    if (!c || !pin_img || !pin_k || !pin_d) { std::printf("allocation failed\n"); return 1; }
    std::copy(img.begin(), img.end(), pin_img);
    int n = 0;
    for (int i = 0; i < 30; i++) orb_extract(c, pin_img, w, h, w, pin_k, pin_d, cap, &n);
    t.clear();
    for (int i = 0; i < 500; i++) {
        const auto t0 = std::chrono::steady_clock::now();
        orb_extract(c, pin_img, w, h, w, pin_k, pin_d, cap, &n);
        t.push_back(std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
    }
    report("orb_extract (pinned buffers)", w, h, nf, (size_t)n, t);
    orb_host_free(pin_img); orb_host_free(pin_k); orb_host_free(pin_d);
    orb_destroy(c);
    return 0;
}
