// Host check: orbsel::nth_element must replay libstdc++'s std::nth_element move for move
// (same resulting permutation) on tie-heavy inputs, including the heap-select fallback.
#include <algorithm>
#include <cstdio>
#include <cstdint>
#include <random>
#include <vector>
#define ORBSEL_TRACE_HEAP 1
#include "../../orbslam_jpminipc_b200/csrc/introselect.h"

struct E { float response; int id; };

static int run_case(const std::vector<int>& keys, int nth, const char* what)
{
    const int n = (int)keys.size();
    std::vector<E> ref(n);
    std::vector<uint64_t> mine64(n);
    std::vector<uint32_t> mine32(n);
    for (int i = 0; i < n; i++) {
        ref[i].response = (float)keys[i]; ref[i].id = i;
        mine64[i] = ((uint64_t)keys[i] << 32) | (uint32_t)i;
        mine32[i] = ((uint32_t)keys[i] << 24) | (uint32_t)(i & 0xffffff);
    }
    std::nth_element(ref.begin(), ref.begin() + nth, ref.end(),
                     [](const E& a, const E& b) { return a.response > b.response; });
    orbsel::nth_element(mine64.data(), n, nth, orbsel::KeyGreater<uint64_t, 32>());
    bool small = true;                       // the 32-bit record has an 8-bit key
    for (int v : keys) small = small && v >= 0 && v <= 255;
    orbsel::nth_element(mine32.data(), n, nth, orbsel::KeyGreater<uint32_t, 24>());
    // the data-parallel description of the partition (what k_select's warps execute) must give the same permutation
    for (int serial_below : {4, 48}) {
        std::vector<uint64_t> par(n);
        std::vector<int> scratch(n + 1);
        for (int i = 0; i < n; i++) par[i] = ((uint64_t)keys[i] << 32) | (uint32_t)i;
        orbsel::nth_element_model(par.data(), n, nth, orbsel::KeyGreater<uint64_t, 32>(), scratch.data(), serial_below);
        if (par != mine64) { std::printf("MODEL MISMATCH %s n=%d nth=%d serial_below=%d\n", what, n, nth, serial_below); return 1; }
    }
    for (int i = 0; i < n; i++)
        if ((int)(mine64[i] & 0xffffffffu) != ref[i].id || (small && (int)(mine32[i] & 0xffffff) != ref[i].id)) {
            std::printf("MISMATCH %s n=%d nth=%d at %d\n", what, n, nth, i);
            return 1;
        }
    return 0;
}

int main()
{
    std::mt19937 rng(12345);
    int bad = 0, cases = 0;
    for (int t = 0; t < 20000; t++) {
        int n = 1 + rng() % (t % 7 == 0 ? 3000 : 200);
        int spread = 1 + rng() % (t % 3 == 0 ? 4 : (t % 3 == 1 ? 40 : 250));
        std::vector<int> k(n);
        for (int& v : k) v = 1 + rng() % spread;
        int nth = rng() % n;
        bad += run_case(k, nth, "random"); cases++;
    }
    // sorted / reverse / organ-pipe / constant inputs
    for (int n : {1, 2, 3, 4, 5, 17, 64, 255, 1000, 4097}) {
        std::vector<int> a(n), b(n), c(n), d(n, 7);
        for (int i = 0; i < n; i++) { a[i] = i % 250 + 1; b[i] = (n - i) % 250 + 1; c[i] = std::min(i, n - i) % 250 + 1; }
        for (int nth : {0, n / 3, n / 2, n - 1}) {
            bad += run_case(a, nth, "asc"); bad += run_case(b, nth, "desc");
            bad += run_case(c, nth, "pipe"); bad += run_case(d, nth, "const"); cases += 4;
        }
    }
    // median-of-3 killer (Musser) to force the depth limit and the heap-select branch
    for (int n : {64, 200, 1024, 3000}) {
        std::vector<int> k(n);
        int half = n / 2;
        for (int i = 0; i < half; i++) { if (i % 2 == 0) k[i] = i + 1; else k[i] = half + i + (half % 2 == 0 ? 0 : 1); k[half + i] = 2 * (i + 1); }
        for (int& v : k) v = n + 1 - v;          // descending comparator
        for (int nth : {0, n / 2, n - 2}) { bad += run_case(k, nth, "killer"); cases++; }
        for (int& v : k) v = v % 250 + 1;
        for (int nth : {1, n / 2}) { bad += run_case(k, nth, "killer-mod"); cases++; }
    }
    if (orbsel_heap_hits == 0) { std::printf("heap-select branch never taken\n"); bad++; }
    std::printf("%s cases=%d bad=%d heap_hits=%d\n", bad ? "FAIL" : "PASS", cases, bad, orbsel_heap_hits);
    return bad ? 1 : 0;
}
