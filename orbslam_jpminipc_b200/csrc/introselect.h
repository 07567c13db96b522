// introselect.h — nth_element with the exact comparison / move sequence of libstdc++ 13
// (std::nth_element -> __introselect: median-of-3 pivot moved to the front, unguarded Hoare
// partition, heap-select once the 2*floor(log2 n) depth budget is spent, insertion sort for
// ranges of <= 3).  KeyPointsFilter::retainBest (called at reference src/ORBextractor.cc:683,
// :699) runs std::nth_element over FAST scores, which are small integers with many ties; which
// tied keypoints survive and in what order is defined by this algorithm, so the device code
// must replay it step for step (SURVEY.md §7 hard part 1).  Usable on host (tests compare it
// with std::nth_element) and device (one thread per list).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ORB_HD __host__ __device__ __forceinline__
#else
#define ORB_HD inline
#endif

#ifdef ORBSEL_TRACE_HEAP
static int orbsel_heap_hits = 0;
#endif

namespace orbsel {

// Order: "a before b" when key(a) > key(b)  (KeypointResponseGreater)
template <typename T, int SHIFT>
struct KeyGreater {
    ORB_HD bool operator()(T a, T b) const { return (a >> SHIFT) > (b >> SHIFT); }
};

// HARRIS_SCORE: the key in the high 32 bits is an IEEE float (Harris responses may be negative)
struct FloatKeyGreater64 {
    ORB_HD static float key(unsigned long long a)
    {
        union { uint32_t u; float f; } c;
        c.u = (uint32_t)(a >> 32);
        return c.f;
    }
    ORB_HD bool operator()(unsigned long long a, unsigned long long b) const { return key(a) > key(b); }
};

template <typename T>
ORB_HD void swp(T* v, int a, int b) { T t = v[a]; v[a] = v[b]; v[b] = t; }

template <typename T, typename C>
ORB_HD void push_heap_(T* v, int hole, int top, T val, C lt)
{
    int parent = (hole - 1) / 2;
    while (hole > top && lt(v[parent], val)) {
        v[hole] = v[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    v[hole] = val;
}

template <typename T, typename C>
ORB_HD void adjust_heap_(T* v, int hole, int len, T val, C lt)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (lt(v[child], v[child - 1])) child--;
        v[hole] = v[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        v[hole] = v[child - 1];
        hole = child - 1;
    }
    push_heap_(v, hole, top, val, lt);
}

template <typename T, typename C>
ORB_HD void heap_select_(T* v, int first, int middle, int last, C lt)
{
    T* b = v + first;
    const int len = middle - first;
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            T val = b[parent];
            adjust_heap_(b, parent, len, val, lt);
            if (parent == 0) break;
            parent--;
        }
    }
    for (int i = middle; i < last; ++i)
        if (lt(v[i], v[first])) {
            T val = v[i];
            v[i] = v[first];
            adjust_heap_(b, 0, len, val, lt);
        }
}

template <typename T, typename C>
ORB_HD void insertion_sort_(T* v, int first, int last, C lt)
{
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        T val = v[i];
        if (lt(val, v[first])) {
            for (int k = i; k > first; --k) v[k] = v[k - 1];
            v[first] = val;
        } else {
            int j = i;
            while (lt(val, v[j - 1])) { v[j] = v[j - 1]; --j; }
            v[j] = val;
        }
    }
}

// floor(log2 n) * 2: the depth budget of std::__introselect
ORB_HD int depth_limit(int n)
{
    int depth = 0;
    for (int t = n; t > 1; t >>= 1) depth++;
    return depth * 2;
}

// the loop of std::__introselect from a given state (range [first, last), remaining depth budget)
template <typename T, typename C>
ORB_HD void nth_element_from(T* v, int first, int last, int nth, int depth, C lt)
{
    while (last - first > 3) {
        if (depth == 0) {
#ifdef ORBSEL_TRACE_HEAP
            ++orbsel_heap_hits;     // host tests only: proves the fallback branch is exercised
#endif
            heap_select_(v, first, nth + 1, last, lt);
            swp(v, first, nth);
            return;
        }
        --depth;
        // median of (first+1, mid, last-1) -> first
        const int mid = first + (last - first) / 2;
        const int a = first + 1, b = mid, c = last - 1;
        if (lt(v[a], v[b])) {
            if (lt(v[b], v[c])) swp(v, first, b);
            else if (lt(v[a], v[c])) swp(v, first, c);
            else swp(v, first, a);
        } else if (lt(v[a], v[c])) swp(v, first, a);
        else if (lt(v[b], v[c])) swp(v, first, c);
        else swp(v, first, b);
        // unguarded partition of [first+1, last) around the pivot at first
        int lo = first + 1, hi = last;
        const T pivot = v[first];
        while (true) {
            while (lt(v[lo], pivot)) ++lo;
            --hi;
            while (lt(pivot, v[hi])) --hi;
            if (!(lo < hi)) break;
            swp(v, lo, hi);
            ++lo;
        }
        if (lo <= nth) first = lo; else last = lo;
    }
    insertion_sort_(v, first, last, lt);
}

// nth_element(v, v+nth, v+n) under the strict weak order lt
template <typename T, typename C>
ORB_HD void nth_element(T* v, int n, int nth, C lt)
{
    if (n <= 0 || nth >= n) return;
    nth_element_from(v, 0, n, nth, depth_limit(n), lt);
}

// ---------------------------------------------------------------------------------------------------------------------
// Data-parallel form of the unguarded partition, with the same result as the sequential scan.  Call A the elements of
// [first+1, last) that stop the left scan (!lt(v[i], pivot)) and B those that stop the right scan (!lt(pivot, v[i])); a_k is the
// k-th A from the left, b_k the k-th B from the right, both in the ORIGINAL array.  The sequential loop swaps exactly the pairs
// (a_k, b_k), k = 1..m, where m is the largest k with a_k < b_k (until they cross, the left scan only meets untouched elements and
// elements it has put there itself, likewise the right scan), and returns min(a_{m+1}, b_m) (a_1 if m = 0): after the last swap
// the left scan runs into the next original stopper or into the element it just moved to b_m, whichever comes first.
// The model below executes that description with plain loops; k_select's warp version (orb_extract.cu, warp_partition) executes
// it with ballots.  tests/cpp/test_introselect.cpp checks the model against std::nth_element.
template <typename T, typename C>
inline int partition_model(T* v, int first, int last, C lt, int* scratch /* >= last - first entries */)
{
    const T pivot = v[first];
    int nB = 0;
    for (int i = last - 1; i > first; --i) if (!lt(pivot, v[i])) scratch[nB++] = i;          // b_1, b_2, ... (from the right)
    int k = 0, cut = -1, m = 0;
    for (int i = first + 1; i < last; ++i) {
        if (lt(v[i], pivot)) continue;                                                       // not a left stopper
        const int partner = k < nB ? scratch[k] : -1;
        if (partner > i) { swp(v, i, partner); m = ++k; }                                    // a_k < b_k: swap, as the scan would
        else { cut = i; break; }                                                             // a_{m+1} (or a moved element at b_j >= b_m)
    }
    const int bm = m > 0 ? scratch[m - 1] : last;                                            // last: larger than any index (m = 0 -> a_1)
    if (cut < 0 || bm < cut) cut = bm;
    return cut;
}

template <typename T, typename C>
inline void nth_element_model(T* v, int n, int nth, C lt, int* scratch, int serial_below = 8)
{
    if (n <= 0 || nth >= n) return;
    int first = 0, last = n, depth = depth_limit(n);
    while (last - first > 3) {
        if (depth == 0 || last - first < serial_below) { nth_element_from(v, first, last, nth, depth, lt); return; }
        --depth;
        const int mid = first + (last - first) / 2;
        const int a = first + 1, b = mid, c = last - 1;
        if (lt(v[a], v[b])) {
            if (lt(v[b], v[c])) swp(v, first, b);
            else if (lt(v[a], v[c])) swp(v, first, c);
            else swp(v, first, a);
        } else if (lt(v[a], v[c])) swp(v, first, a);
        else if (lt(v[b], v[c])) swp(v, first, c);
        else swp(v, first, b);
        const int cut = partition_model(v, first, last, lt, scratch);
        if (cut <= nth) first = cut; else last = cut;
    }
    insertion_sort_(v, first, last, lt);
}

} // namespace orbsel
