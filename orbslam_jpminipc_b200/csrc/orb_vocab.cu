// orb_vocab.cu — DBoW2 vocabulary tree on the GPU.
//
// Replaces, for ORB descriptors (FORB, 32 bytes, Hamming distance):
//   TemplatedVocabulary::transform(feature, word, weight, nid, levelsup)   Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1218-1260
//   TemplatedVocabulary::transform(features, BowVector&, FeatureVector&, levelsup)              :1127-1193
//   BowVector::addWeight / addIfNotExist / normalize                        Thirdparty/DBoW2/DBoW2/BowVector.cpp:33-95
//   FeatureVector::addFeature                                               Thirdparty/DBoW2/DBoW2/FeatureVector.cpp:31-45
//   L1Scoring::score                                                        Thirdparty/DBoW2/DBoW2/ScoringObject.cpp:22-64
//   the shared-word count + score loop of KeyFrameDatabase::DetectRelocalisationCandidates  src/KeyFrameDatabase.cc:198-252
// called from Frame::ComputeBoW (src/Frame.cc:279-287) and KeyFrame::ComputeBoW (src/KeyFrame.cc:56-65) with levelsup = 4.
//
// Layout in HBM: nodes are renumbered breadth first so that the children of a node are consecutive ("internal ids"); per internal
// id one int2 (first child, child count), 32 descriptor bytes, the file-order node id, the word id (-1 for inner nodes) and the
// weight.  A k=10, L=6 tree (1 111 111 nodes) takes 35.6 MB of descriptors + 26.7 MB of tables and stays resident in the 126 MB L2.
// Values of a BowVector are doubles and every sum keeps the reference's order (feature order inside a word, ascending word id in
// the norm), so results are bit-identical to the std::map based code.
#include "orb_internal.h"
#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct orb_vocab {
    int device = 0, k = 0, L = 0, scoring = 0, weighting = 0, nnodes = 0, nwords = 0;
    int2* d_kids = nullptr;
    uint4* d_desc = nullptr;
    int32_t* d_orig = nullptr;
    int32_t* d_word = nullptr;
    double* d_weight = nullptr;
};

namespace {

struct VocabDev {
    const int2* kids; const uint4* desc; const int32_t* orig; const int32_t* word; const double* weight;
    int L;
};

constexpr int BOW_MAX_FEATURES = 8192;         // features per frame the in-shared-memory sort handles
constexpr int BOW_THREADS = 1024;

__device__ __forceinline__ int hamming256(const uint32_t (&q)[8], const uint4 a, const uint4 b)
{
    return __popc(q[0] ^ a.x) + __popc(q[1] ^ a.y) + __popc(q[2] ^ a.z) + __popc(q[3] ^ a.w) +
           __popc(q[4] ^ b.x) + __popc(q[5] ^ b.y) + __popc(q[6] ^ b.z) + __popc(q[7] ^ b.w);
}

// Tree descent, G lanes per feature (lane c scores child c, strict '<' in file order == lowest child index among the minima, :1238-1247).
// Features live in per-frame slots of slot_rows rows; counts == nullptr means one flat array of nslots features.
template <int G>
__global__ void __launch_bounds__(256)
k_vocab_descend(VocabDev V, const uint8_t* __restrict__ desc, const int32_t* __restrict__ counts, int slot_rows, long long nslots,
                int levelsup, int32_t* __restrict__ word, double* __restrict__ weight, int32_t* __restrict__ node)
{
    const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int sub = threadIdx.x % G, lane = threadIdx.x & 31;
    if (gid >= nslots) return;
    if (counts && (int)(gid % slot_rows) >= counts[gid / slot_rows]) return;
    uint32_t mask = 0xffffffffu;
    if constexpr (G < 32) mask = ((1u << G) - 1u) << (lane & ~(G - 1));
    uint32_t q[8];
    {
        const uint4* qp = reinterpret_cast<const uint4*>(desc + (size_t)gid * 32);
        const uint4 a = __ldg(qp), b = __ldg(qp + 1);
        q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
    }
    const int nid_level = V.L - levelsup;
    int nid = nid_level <= 0 ? 0 : -1, cur = 0, level = 0;
    int2 kd = __ldg(V.kids);
    while (kd.y > 0) {
        ++level;
        uint32_t best = 0xffffffffu;
        for (int c = sub; c < kd.y; c += G) {
            const uint4* p = V.desc + 2 * (size_t)(kd.x + c);
            best = min(best, ((uint32_t)hamming256(q, __ldg(p), __ldg(p + 1)) << 8) | (uint32_t)c);
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(mask, best, o, G));
        cur = kd.x + (int)(best & 255u);
        if (level == nid_level) nid = cur;
        kd = __ldg(V.kids + cur);
    }
    if (sub == 0) {
        const bool empty_tree = level == 0;
        word[gid] = empty_tree ? -1 : V.word[cur];
        weight[gid] = empty_tree ? 0.0 : V.weight[cur];
        node[gid] = V.orig[nid >= 0 ? nid : cur];            // reference leaves *nid unset for a leaf above nid_level; we report the leaf
    }
}

__device__ void bitonic_sort_u64(unsigned long long* keys, int P)
{
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < P; t += blockDim.x) {
                const int ixj = t ^ j;
                if (ixj > t) {
                    const unsigned long long a = keys[t], b = keys[ixj];
                    const bool up = (t & k) == 0;
                    if ((a > b) == up) { keys[t] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
}

// exclusive scan of 0/1 flags derived from sorted keys: head[i] = i < m && (i == 0 || hi(keys[i]) != hi(keys[i-1])).
// Returns this thread's first position and fills pos for its contiguous chunk via the callback.
template <typename F>
__device__ int scan_heads(const unsigned long long* keys, int P, int m, int* s_warp, F&& emit)
{
    const int per = P / (int)blockDim.x > 0 ? P / (int)blockDim.x : 1;
    const int b = threadIdx.x * per, e = min(b + per, P);
    int cnt = 0;
    for (int i = b; i < e && i < m; i++) cnt += (i == 0 || (keys[i] >> 32) != (keys[i - 1] >> 32));
    // block exclusive scan of cnt
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
        s_warp[lane] = w;                                   // inclusive totals per warp
    }
    __syncthreads();
    int pos = inc - cnt + (warp > 0 ? s_warp[warp - 1] : 0);
    const int total = s_warp[(blockDim.x >> 5) - 1];
    for (int i = b; i < e && i < m; i++)
        if (i == 0 || (keys[i] >> 32) != (keys[i - 1] >> 32)) emit(i, pos++);
    __syncthreads();
    return total;
}

// One CTA per frame: FeatureVector (CSR) and BowVector from the per-feature (word, weight, node) triples.
__global__ void __launch_bounds__(BOW_THREADS)
k_bow_build(const int32_t* __restrict__ word, const double* __restrict__ weight, const int32_t* __restrict__ node,
            const int32_t* __restrict__ counts, int n_flat, int slot_rows, int cap, int scoring, int weighting,
            int32_t* __restrict__ bow_word, double* __restrict__ bow_val, int32_t* __restrict__ nbow,
            int32_t* __restrict__ fv_node, int32_t* __restrict__ fv_start, int32_t* __restrict__ fv_items, int32_t* __restrict__ nfv)
{
    extern __shared__ unsigned long long keys[];
    __shared__ int s_warp[32];
    __shared__ int s_m;
    const int f = blockIdx.x, tid = threadIdx.x;
    const int n = min(counts ? counts[f] : n_flat, slot_rows);
    const size_t in0 = (size_t)f * slot_rows;
    word += in0; weight += in0; node += in0;
    bow_word += (size_t)f * cap; bow_val += (size_t)f * cap;
    fv_node += (size_t)f * cap; fv_items += (size_t)f * cap; fv_start += (size_t)f * (cap + 1);
    int P = 32;
    while (P < n) P <<= 1;
    if (tid == 0) s_m = 0;
    __syncthreads();
    // ---- FeatureVector: stable order by node id, then feature index (map<NodeId, vector<unsigned>>, ascending)
    int valid = 0;
    for (int i = tid; i < P; i += blockDim.x) {
        const bool ok = i < n && weight[i] > 0;                // "not stopped" (:1157)
        keys[i] = ok ? ((unsigned long long)(uint32_t)node[i] << 32) | (uint32_t)i : ~0ull;
        valid += ok;
    }
    atomicAdd(&s_m, valid);
    __syncthreads();
    const int m = s_m;
    bitonic_sort_u64(keys, P);
    for (int i = tid; i < m; i += blockDim.x) fv_items[i] = (int32_t)(keys[i] & 0xffffffffu);
    const int nn = scan_heads(keys, P, m, s_warp, [&](int i, int pos) { fv_node[pos] = (int32_t)(keys[i] >> 32); fv_start[pos] = i; });
    if (tid == 0) { fv_start[nn] = m; nfv[f] = nn; }
    __syncthreads();
    // ---- BowVector: words ascending; a word's value accumulates its features in feature order (addWeight) or keeps the first (addIfNotExist)
    for (int i = tid; i < P; i += blockDim.x) {
        const bool ok = i < n && weight[i] > 0;
        keys[i] = ok ? ((unsigned long long)(uint32_t)word[i] << 32) | (uint32_t)i : ~0ull;
    }
    __syncthreads();
    bitonic_sort_u64(keys, P);
    const bool tf = weighting == 0 || weighting == 1;          // TF_IDF, TF (BowVector.h:36-42)
    const int nw = scan_heads(keys, P, m, s_warp, [&](int i, int pos) {
        const unsigned long long w = keys[i] >> 32;
        double v = weight[(int)(keys[i] & 0xffffffffu)];
        if (tf)
            for (int j = i + 1; j < m && (keys[j] >> 32) == w; j++) v = __dadd_rn(v, weight[(int)(keys[j] & 0xffffffffu)]);
        bow_word[pos] = (int32_t)w;
        bow_val[pos] = v;
    });
    __syncthreads();
    const bool must = scoring != 5;                             // DotProductScoring does not normalise (ScoringObject.h:72-87)
    const bool l2 = scoring == 1;
    __shared__ double s_norm;
    if (tf && nw > 0 && !must) {                                // :1165-1171
        const double nd = (double)nw;
        for (int i = tid; i < nw; i += blockDim.x) bow_val[i] = __ddiv_rn(bow_val[i], nd);
    }
    if (must) {                                                 // BowVector::normalize, in ascending word order
        if (tid == 0) {
            double norm = 0.0;
            if (!l2) for (int i = 0; i < nw; i++) norm = __dadd_rn(norm, fabs(bow_val[i]));
            else { for (int i = 0; i < nw; i++) norm = __dadd_rn(norm, __dmul_rn(bow_val[i], bow_val[i])); norm = sqrt(norm); }
            s_norm = norm;
        }
        __syncthreads();
        const double norm = s_norm;
        if (norm > 0.0) for (int i = tid; i < nw; i += blockDim.x) bow_val[i] = __ddiv_rn(bow_val[i], norm);
    }
    if (tid == 0) nbow[f] = nw;
}

// ---- retrieval scoring: one warp per keyframe BowVector
__device__ __forceinline__ int find_word(const int32_t* __restrict__ qw, int nq, int w)
{
    int lo = 0, hi = nq;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (qw[mid] < w) lo = mid + 1; else hi = mid; }
    return lo < nq && qw[lo] == w ? lo : -1;
}

// excluded (optional): keyframes connected to the query keyframe never enter DetectLoopCandidates' list (src/KeyFrameDatabase.cc:95):
// they count as sharing nothing.  first_pos (optional): position in the query BowVector of the first shared word — the query words are
// walked in ascending order and every inverted-file list in keyframe order (:85-104), so (first_pos, k) is the order of
// lKFsSharingWords.
__global__ void __launch_bounds__(256)
k_bow_common(const int32_t* __restrict__ qw, int nq, int nkf, const int32_t* __restrict__ kf_start, const int32_t* __restrict__ kf_word,
             const uint8_t* __restrict__ excluded, int32_t* __restrict__ common, int* __restrict__ max_common, int32_t* __restrict__ first_pos)
{
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (k >= nkf) return;
    const int s = kf_start[k], e = kf_start[k + 1];
    int c = 0, fp = INT_MAX;
    if (!excluded || !excluded[k])
        for (int j = s + lane; j < e; j += 32) { const int p = find_word(qw, nq, kf_word[j]); if (p >= 0) { c++; fp = min(fp, p); } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { c += __shfl_xor_sync(0xffffffffu, c, o); fp = min(fp, __shfl_xor_sync(0xffffffffu, fp, o)); }
    if (lane == 0) { common[k] = c; atomicMax(max_common, c); if (first_pos) first_pos[k] = fp; }
}

// Covisibility accumulation of DetectRelocalisationCandidates (:262-290) / DetectLoopCandidates (:141-168): one thread per keyframe of
// lScoreAndMatch walks its (at most ten) best covisible keyframes in their order; float sums in that order.  kf_score is the
// mRelocScore / mLoopScore member of every keyframe: k_bow_score has just rewritten it for the keyframes scored by THIS query, the
// others keep what an earlier query left there, exactly as the members do (relocalisation adds such stale scores of neighbours that
// share a word without having been scored, :278-281).  acc[k] < 0: keyframe k is not in lScoreAndMatch.
__global__ void __launch_bounds__(256)
k_bow_accumulate(int nkf, const int32_t* __restrict__ common, const int* __restrict__ max_common, const float* __restrict__ kf_score,
                 const int32_t* __restrict__ cov_start, const int32_t* __restrict__ cov_idx, int loop, float min_score,
                 float* __restrict__ acc, int32_t* __restrict__ best_kf, int* __restrict__ best_acc_bits, unsigned long long* __restrict__ min_key)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkf) return;
    min_key[k] = ~0ull;
    const int min_common = (int)__fmul_rn((float)*max_common, 0.8f);
    const int cm = common[k];
    float a = -1.f;
    int bk = k;
    if (cm > 0 && cm > min_common && (!loop || kf_score[k] >= min_score)) {
        float best = kf_score[k];
        a = best;
        if (cov_start) {
            const int e = min(cov_start[k + 1], cov_start[k] + 10);
            for (int j = cov_start[k]; j < e; j++) {
                const int nb = cov_idx[j];
                if (nb < 0 || nb >= nkf || common[nb] <= 0 || (loop && common[nb] <= min_common)) continue;
                const float sn = kf_score[nb];
                a = __fadd_rn(a, sn);
                if (sn > best) { bk = nb; best = sn; }
            }
        }
        atomicMax(best_acc_bits, __float_as_int(a));         // scores are >= 0: the bit patterns order like the values
    }
    acc[k] = a; best_kf[k] = bk;
}

// keyframes above 0.75 * best accumulated score hand their best keyframe to the result (:170-190, :292-306); a keyframe named several
// times is listed where its FIRST nominator stands in lScoreAndMatch: smallest (first_pos, k) key of its nominators
__global__ void __launch_bounds__(256)
k_bow_mark(int nkf, const float* __restrict__ acc, const int32_t* __restrict__ best_kf, const int32_t* __restrict__ first_pos,
           const int* __restrict__ best_acc_bits, int loop, float min_score, unsigned long long* __restrict__ min_key)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkf || acc[k] < 0.f) return;
    const float best = loop ? fmaxf(min_score, __int_as_float(*best_acc_bits)) : __int_as_float(*best_acc_bits);     // bestAccScore starts at minScore (:139) / 0 (:260)
    if (acc[k] > __fmul_rn(0.75f, best))
        atomicMin(&min_key[best_kf[k]], ((unsigned long long)(uint32_t)first_pos[k] << 32) | (uint32_t)k);
}

// rank by counting: one warp per listed keyframe counts the smaller keys (the list is a handful of keyframes)
__global__ void __launch_bounds__(256)
k_bow_rank(int nkf, const unsigned long long* __restrict__ min_key, int32_t* __restrict__ cand, int* __restrict__ ncand)
{
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= nkf) return;
    const unsigned long long key = min_key[b];
    if (key == ~0ull) return;
    int r = 0;
    for (int j = lane; j < nkf; j += 32) r += min_key[j] < key;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (lane == 0) { cand[r] = b; atomicAdd(ncand, 1); }
}

__global__ void __launch_bounds__(256)
k_bow_score(const int32_t* __restrict__ qw, const double* __restrict__ qv, int nq, int nkf, const int32_t* __restrict__ kf_start,
            const int32_t* __restrict__ kf_word, const double* __restrict__ kf_val, const int32_t* __restrict__ common,
            const int* __restrict__ max_common, int score_all, float* __restrict__ score, int keep_unscored = 0)
{
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (k >= nkf) return;
    const int min_common = (int)__fmul_rn((float)*max_common, 0.8f);     // src/KeyFrameDatabase.cc:233
    const int cm = common[k];
    if (cm <= 0 || (!score_all && cm <= min_common)) { if (lane == 0 && !keep_unscored) score[k] = 0.f; return; }
    const int s = kf_start[k], e = kf_start[k + 1];
    double acc = 0.0;                                                    // lane 0 carries the running sum, in word order
    for (int j0 = s; j0 < e; j0 += 32) {
        const int j = j0 + lane;
        double term = 0.0;
        bool hit = false;
        if (j < e) {
            const int p = find_word(qw, nq, kf_word[j]);
            if (p >= 0) {
                const double vi = qv[p], wi = kf_val[j];
                term = __dsub_rn(__dsub_rn(fabs(__dsub_rn(vi, wi)), fabs(vi)), fabs(wi));     // ScoringObject.cpp:40
                hit = true;
            }
        }
        uint32_t m = __ballot_sync(0xffffffffu, hit);
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const double t = __shfl_sync(0xffffffffu, term, b);
            acc = __dadd_rn(acc, t);
        }
    }
    if (lane == 0) score[k] = __double2float_rn(__ddiv_rn(-acc, 2.0));
}

inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

bool on_device(const void* p)
{
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int launch_descend(const orb_vocab* v, const uint8_t* d_desc, const int32_t* d_counts, int slot_rows, long long nslots, int levelsup,
                   int32_t* d_word, double* d_weight, int32_t* d_node, cudaStream_t s)
{
    if (nslots <= 0) return ORB_OK;
    VocabDev V = { v->d_kids, v->d_desc, v->d_orig, v->d_word, v->d_weight, v->L };
    if (v->k <= 16) {
        const long long threads = nslots * 16;
        k_vocab_descend<16><<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(V, d_desc, d_counts, slot_rows, nslots, levelsup, d_word, d_weight, d_node);
    } else {
        const long long threads = nslots * 32;
        k_vocab_descend<32><<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(V, d_desc, d_counts, slot_rows, nslots, levelsup, d_word, d_weight, d_node);
    }
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

} // namespace

extern "C" {

int orb_vocab_create(orb_ctx* c, int k, int L, int scoring, int weighting, int nnodes, const int32_t* parent, const uint8_t* desc,
                     const double* weight, orb_vocab** out)
{
    if (!c || !out || nnodes < 1 || !parent || !desc || !weight || k < 0 || L < 1 || scoring < 0 || scoring > 5 || weighting < 0 || weighting > 3)
        return ORB_ERR_INVALID;
    *out = nullptr;
    // children lists in file order (m_nodes[pid].children.push_back(nid), TemplatedVocabulary.h:1389)
    std::vector<int32_t> cnt(nnodes, 0), first(nnodes + 1, 0);
    for (int i = 1; i < nnodes; i++) {
        if (parent[i] < 0 || parent[i] >= i) return ORB_ERR_INVALID;     // a parent precedes its children in the file
        cnt[parent[i]]++;
    }
    for (int i = 0; i < nnodes; i++) { if (cnt[i] > 255) return ORB_ERR_CAPACITY; first[i + 1] = first[i] + cnt[i]; }
    std::vector<int32_t> kids_flat(std::max(nnodes - 1, 1)), fill(first.begin(), first.end() - 1);
    for (int i = 1; i < nnodes; i++) kids_flat[fill[parent[i]]++] = i;
    // breadth-first renumbering: children of one node become consecutive
    std::vector<int32_t> order; order.reserve(nnodes);
    order.push_back(0);
    std::vector<int2> kids(nnodes);
    for (size_t h = 0; h < order.size(); h++) {
        const int nd = order[h];
        kids[h] = make_int2((int)order.size(), cnt[nd]);
        for (int j = first[nd]; j < first[nd + 1]; j++) order.push_back(kids_flat[j]);
    }
    if ((int)order.size() != nnodes) return ORB_ERR_INVALID;
    std::vector<int32_t> word_of(nnodes, -1);
    int nwords = 0;
    for (int i = 1; i < nnodes; i++) if (cnt[i] == 0) word_of[i] = nwords++;          // :1408-1414
    std::vector<uint8_t> h_desc((size_t)nnodes * 32);
    std::vector<int32_t> h_word(nnodes);
    std::vector<double> h_weight(nnodes);
    for (int h = 0; h < nnodes; h++) {
        const int nd = order[h];
        memcpy(&h_desc[(size_t)h * 32], desc + (size_t)nd * 32, 32);
        h_word[h] = word_of[nd];
        h_weight[h] = weight[nd];
    }
    ORB_CUDA(cudaSetDevice(c->device));
    orb_vocab* v = new orb_vocab;
    v->device = c->device; v->k = 0; v->L = L; v->scoring = scoring; v->weighting = weighting; v->nnodes = nnodes; v->nwords = nwords;
    for (int i = 0; i < nnodes; i++) v->k = std::max(v->k, (int)cnt[i]);
    v->k = std::max(v->k, k);
    auto fail = [&](int rc) { orb_vocab_destroy(v); return rc; };
    if (cudaMalloc(&v->d_kids, (size_t)nnodes * sizeof(int2)) != cudaSuccess || cudaMalloc(&v->d_desc, (size_t)nnodes * 32) != cudaSuccess ||
        cudaMalloc(&v->d_orig, (size_t)nnodes * 4) != cudaSuccess || cudaMalloc(&v->d_word, (size_t)nnodes * 4) != cudaSuccess ||
        cudaMalloc(&v->d_weight, (size_t)nnodes * 8) != cudaSuccess) return fail(ORB_ERR_CUDA);
    if (cudaMemcpy(v->d_kids, kids.data(), (size_t)nnodes * sizeof(int2), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(v->d_desc, h_desc.data(), (size_t)nnodes * 32, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(v->d_orig, order.data(), (size_t)nnodes * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(v->d_word, h_word.data(), (size_t)nnodes * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(v->d_weight, h_weight.data(), (size_t)nnodes * 8, cudaMemcpyHostToDevice) != cudaSuccess) return fail(ORB_ERR_CUDA);
    *out = v;
    return ORB_OK;
}

// text format of TemplatedVocabulary::loadFromTextFile (:1338-1425): "k L scoring weighting", then per node "parent isLeaf d0 .. d31 weight"
int orb_vocab_load_text(orb_ctx* c, const char* path, orb_vocab** out)
{
    if (!c || !path || !out) return ORB_ERR_INVALID;
    *out = nullptr;
    FILE* f = fopen(path, "r");
    if (!f) return ORB_ERR_INVALID;
    std::vector<char> line(1 << 16);
    if (!fgets(line.data(), (int)line.size(), f)) { fclose(f); return ORB_ERR_INVALID; }
    int k = -1, L = -1, n1 = -1, n2 = -1;
    sscanf(line.data(), "%d %d %d %d", &k, &L, &n1, &n2);
    if (k < 0 || k > 20 || L < 1 || L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) { fclose(f); return ORB_ERR_INVALID; }     // :1358-1362
    std::vector<int32_t> parent(1, 0);
    std::vector<uint8_t> desc(32, 0);
    std::vector<double> weight(1, 0.0);
    bool bad = false;
    while (!bad && fgets(line.data(), (int)line.size(), f)) {
        const char* p = line.data();
        char* end = nullptr;
        const long pid = strtol(p, &end, 10);
        if (end == p) continue;                         // blank line
        p = end;
        strtol(p, &end, 10);                            // nIsLeaf: redundant, leaves are the nodes without children (:328)
        if (end == p) { bad = true; break; }
        p = end;
        parent.push_back((int32_t)pid);
        for (int i = 0; i < 32 && !bad; i++) {
            const long b = strtol(p, &end, 10);
            if (end == p) bad = true;
            desc.push_back((uint8_t)b);
            p = end;
        }
        const double w = strtod(p, &end);
        if (end == p) bad = true;
        weight.push_back(w);
    }
    fclose(f);
    if (bad) return ORB_ERR_INVALID;
    return orb_vocab_create(c, k, L, n1, n2, (int)parent.size(), parent.data(), desc.data(), weight.data(), out);
}

void orb_vocab_destroy(orb_vocab* v)
{
    if (!v) return;
    cudaSetDevice(v->device);
    cudaFree(v->d_kids); cudaFree(v->d_desc); cudaFree(v->d_orig); cudaFree(v->d_word); cudaFree(v->d_weight);
    delete v;
}

int orb_vocab_info(const orb_vocab* v, int* k, int* L, int* nnodes, int* nwords)
{
    if (!v) return ORB_ERR_INVALID;
    if (k) *k = v->k;
    if (L) *L = v->L;
    if (nnodes) *nnodes = v->nnodes;
    if (nwords) *nwords = v->nwords;
    return ORB_OK;
}

int orb_vocab_transform_features(orb_ctx* c, orb_vocab* v, const uint8_t* desc, int n, int levelsup, int32_t* word, double* weight,
                                 int32_t* node)
{
    if (!c || !v || n < 0 || !word || !weight || !node) return ORB_ERR_INVALID;
    if (n == 0) return ORB_OK;
    if (!desc) return ORB_ERR_INVALID;
    ORB_CUDA(cudaSetDevice(c->device));
    LaneGuard lg(c);                              // scratch and stream of this call only: transforms of different threads run side by side
    if (!lg.lane) return ORB_ERR_CUDA;
    cudaStream_t s = lg.lane->stream;
    const bool dev = on_device(desc);
    if (on_device(word) != dev || on_device(weight) != dev || on_device(node) != dev) return ORB_ERR_INVALID;
    if (dev) {
        int rc = launch_descend(v, desc, nullptr, 1, n, levelsup, word, weight, node, s);
        if (rc) return rc;
        ORB_CUDA(cudaStreamSynchronize(s));
        return ORB_OK;
    }
    const size_t N = (size_t)n;
    int rc = orb_lane_scratch(lg.lane, al256(N * 32) + al256(N * 4) * 2 + al256(N * 8));
    if (rc) return rc;
    uint8_t* p = (uint8_t*)lg.lane->d_scratch;
    uint8_t* d_desc = p; p += al256(N * 32);
    int32_t* d_word = (int32_t*)p; p += al256(N * 4);
    int32_t* d_node = (int32_t*)p; p += al256(N * 4);
    double* d_weight = (double*)p;
    ORB_CUDA(cudaMemcpyAsync(d_desc, desc, N * 32, cudaMemcpyHostToDevice, s));
    if ((rc = launch_descend(v, d_desc, nullptr, 1, n, levelsup, d_word, d_weight, d_node, s))) return rc;
    ORB_CUDA(cudaMemcpyAsync(word, d_word, N * 4, cudaMemcpyDeviceToHost, s));
    ORB_CUDA(cudaMemcpyAsync(node, d_node, N * 4, cudaMemcpyDeviceToHost, s));
    ORB_CUDA(cudaMemcpyAsync(weight, d_weight, N * 8, cudaMemcpyDeviceToHost, s));
    ORB_CUDA(cudaStreamSynchronize(s));
    return ORB_OK;
}

int orb_vocab_transform_batch(orb_ctx* c, orb_vocab* v, const uint8_t* desc, int slot_rows, const int32_t* counts, int nframes, int levelsup,
                              int cap, int32_t* bow_word, double* bow_val, int32_t* nbow, int32_t* fv_node, int32_t* fv_start,
                              int32_t* fv_items, int32_t* nfv)
{
    if (!c || !v || nframes < 0 || slot_rows < 0 || cap < 0 || !counts || !nbow || !nfv) return ORB_ERR_INVALID;
    if (nframes == 0) return ORB_OK;
    if (cap > BOW_MAX_FEATURES || slot_rows > cap) return ORB_ERR_CAPACITY;
    if (cap > 0 && (!desc || !bow_word || !bow_val || !fv_node || !fv_start || !fv_items)) return ORB_ERR_INVALID;
    ORB_CUDA(cudaSetDevice(c->device));
    LaneGuard lg(c);                              // scratch and stream of this call only: transforms of different threads run side by side
    if (!lg.lane) return ORB_ERR_CUDA;
    cudaStream_t s = lg.lane->stream;
    const bool dev = on_device(counts);
    const void* ptrs[] = { desc, bow_word, bow_val, nbow, fv_node, fv_start, fv_items, nfv };
    for (const void* p : ptrs) if (p && on_device(p) != dev) return ORB_ERR_INVALID;
    const size_t F = (size_t)nframes, slots = F * (size_t)slot_rows, outs = F * (size_t)cap;
    size_t need = al256(slots * 4) * 2 + al256(slots * 8);
    if (!dev) need += al256(slots * 32) + al256(F * 4) * 3 + al256(outs * 4) * 3 + al256(outs * 8) + al256(F * (size_t)(cap + 1) * 4);
    int rc = orb_lane_scratch(lg.lane, need + 4096);
    if (rc) return rc;
    uint8_t* p = (uint8_t*)lg.lane->d_scratch;
    auto take = [&](size_t bytes) { uint8_t* r = p; p += al256(bytes); return r; };
    int32_t* d_word = (int32_t*)take(slots * 4);
    int32_t* d_node = (int32_t*)take(slots * 4);
    double* d_weight = (double*)take(slots * 8);
    const uint8_t* d_desc = desc; const int32_t* d_counts = counts;
    int32_t *o_bw = bow_word, *o_nb = nbow, *o_fn = fv_node, *o_fs = fv_start, *o_fi = fv_items, *o_nf = nfv;
    double* o_bv = bow_val;
    if (!dev) {
        for (int f = 0; f < nframes; f++) if (counts[f] < 0 || counts[f] > slot_rows) return ORB_ERR_INVALID;
        uint8_t* dd = take(slots * 32); int32_t* dc = (int32_t*)take(F * 4);
        o_nb = (int32_t*)take(F * 4); o_nf = (int32_t*)take(F * 4);
        o_bw = (int32_t*)take(outs * 4); o_fn = (int32_t*)take(outs * 4); o_fi = (int32_t*)take(outs * 4);
        o_bv = (double*)take(outs * 8); o_fs = (int32_t*)take(F * (size_t)(cap + 1) * 4);
        if (slots) ORB_CUDA(cudaMemcpyAsync(dd, desc, slots * 32, cudaMemcpyHostToDevice, s));
        ORB_CUDA(cudaMemcpyAsync(dc, counts, F * 4, cudaMemcpyHostToDevice, s));
        d_desc = dd; d_counts = dc;
    }
    if ((rc = launch_descend(v, d_desc, d_counts, std::max(slot_rows, 1), (long long)slots, levelsup, d_word, d_weight, d_node, s))) return rc;
    int P = 32;
    while (P < std::max(slot_rows, 1)) P <<= 1;
    const size_t smem = (size_t)P * 8;
    // per device and cheap: no caching, a process may drive several GPUs
    ORB_CUDA(cudaFuncSetAttribute(k_bow_build, cudaFuncAttributeMaxDynamicSharedMemorySize, BOW_MAX_FEATURES * 8));
    k_bow_build<<<nframes, BOW_THREADS, smem, s>>>(d_word, d_weight, d_node, d_counts, 0, slot_rows, cap, v->scoring, v->weighting,
                                                   o_bw, o_bv, o_nb, o_fn, o_fs, o_fi, o_nf);
    ORB_CUDA(cudaGetLastError());
    if (!dev) {
        ORB_CUDA(cudaMemcpyAsync(nbow, o_nb, F * 4, cudaMemcpyDeviceToHost, s));
        ORB_CUDA(cudaMemcpyAsync(nfv, o_nf, F * 4, cudaMemcpyDeviceToHost, s));
        if (outs) {
            ORB_CUDA(cudaMemcpyAsync(bow_word, o_bw, outs * 4, cudaMemcpyDeviceToHost, s));
            ORB_CUDA(cudaMemcpyAsync(bow_val, o_bv, outs * 8, cudaMemcpyDeviceToHost, s));
            ORB_CUDA(cudaMemcpyAsync(fv_node, o_fn, outs * 4, cudaMemcpyDeviceToHost, s));
            ORB_CUDA(cudaMemcpyAsync(fv_items, o_fi, outs * 4, cudaMemcpyDeviceToHost, s));
        }
        ORB_CUDA(cudaMemcpyAsync(fv_start, o_fs, F * (size_t)(cap + 1) * 4, cudaMemcpyDeviceToHost, s));
    }
    ORB_CUDA(cudaStreamSynchronize(s));
    return ORB_OK;
}

int orb_bow_score_db(orb_ctx* c, orb_vocab* v, const int32_t* qw, const double* qv, int nq, int nkf, const int32_t* kf_start,
                     const int32_t* kf_word, const double* kf_val, int score_all, int32_t* common, float* score, int* max_common)
{
    if (!c || !v || nq < 0 || nkf < 0 || !max_common) return ORB_ERR_INVALID;
    *max_common = 0;
    if (nkf == 0) return ORB_OK;
    if (!kf_start || !common || !score || (nq > 0 && (!qw || !qv))) return ORB_ERR_INVALID;
    if (v->scoring != 0) return ORB_ERR_UNSUPPORTED;                     // only L1_NORM (what the reference's vocabulary uses)
    ORB_CUDA(cudaSetDevice(c->device));
    LaneGuard lg(c);                              // scratch and stream of this call only: transforms of different threads run side by side
    if (!lg.lane) return ORB_ERR_CUDA;
    cudaStream_t s = lg.lane->stream;
    const bool dev = on_device(kf_start);
    if (on_device(common) != dev || on_device(score) != dev || (nq > 0 && on_device(qw) != dev)) return ORB_ERR_INVALID;
    int total = 0;
    if (dev) ORB_CUDA(cudaMemcpy(&total, kf_start + nkf, 4, cudaMemcpyDeviceToHost)); else total = kf_start[nkf];
    if (total < 0 || (total > 0 && (!kf_word || !kf_val))) return ORB_ERR_INVALID;
    const size_t K = (size_t)nkf, T = (size_t)total, Q = (size_t)nq;
    size_t need = 256;
    if (!dev) need += al256(Q * 4) + al256(Q * 8) + al256((K + 1) * 4) + al256(T * 4) + al256(T * 8) + al256(K * 4) * 2;
    int rc = orb_lane_scratch(lg.lane, need + 1024);
    if (rc) return rc;
    uint8_t* p = (uint8_t*)lg.lane->d_scratch;
    auto take = [&](size_t bytes) { uint8_t* r = p; p += al256(std::max<size_t>(bytes, 1)); return r; };
    int* d_max = (int*)take(4);
    int32_t* d_common = common; float* d_score = score;
    if (!dev) {
        int32_t* a = (int32_t*)take(Q * 4); double* b = (double*)take(Q * 8); int32_t* st = (int32_t*)take((K + 1) * 4);
        int32_t* w = (int32_t*)take(T * 4); double* val = (double*)take(T * 8);
        d_common = (int32_t*)take(K * 4); d_score = (float*)take(K * 4);
        if (Q) { ORB_CUDA(cudaMemcpyAsync(a, qw, Q * 4, cudaMemcpyHostToDevice, s)); ORB_CUDA(cudaMemcpyAsync(b, qv, Q * 8, cudaMemcpyHostToDevice, s)); }
        ORB_CUDA(cudaMemcpyAsync(st, kf_start, (K + 1) * 4, cudaMemcpyHostToDevice, s));
        if (T) { ORB_CUDA(cudaMemcpyAsync(w, kf_word, T * 4, cudaMemcpyHostToDevice, s)); ORB_CUDA(cudaMemcpyAsync(val, kf_val, T * 8, cudaMemcpyHostToDevice, s)); }
        qw = a; qv = b; kf_start = st; kf_word = w; kf_val = val;
    }
    ORB_CUDA(cudaMemsetAsync(d_max, 0, 4, s));
    const unsigned blocks = (unsigned)((K * 32 + 255) / 256);
    k_bow_common<<<blocks, 256, 0, s>>>(qw, nq, nkf, kf_start, kf_word, nullptr, d_common, d_max, nullptr);
    k_bow_score<<<blocks, 256, 0, s>>>(qw, qv, nq, nkf, kf_start, kf_word, kf_val, d_common, d_max, score_all, d_score);
    ORB_CUDA(cudaGetLastError());
    if (!dev) {
        ORB_CUDA(cudaMemcpyAsync(common, d_common, K * 4, cudaMemcpyDeviceToHost, s));
        ORB_CUDA(cudaMemcpyAsync(score, d_score, K * 4, cudaMemcpyDeviceToHost, s));
    }
    ORB_CUDA(cudaMemcpyAsync(max_common, d_max, 4, cudaMemcpyDeviceToHost, s));
    ORB_CUDA(cudaStreamSynchronize(s));
    return ORB_OK;
}

/* KeyFrameDatabase::DetectRelocalisationCandidates (src/KeyFrameDatabase.cc:198-308) and DetectLoopCandidates (:75-196) complete, on
 * flat arrays: shared words, scores, covisibility accumulation, the 0.75 * best cut and the reference's result order. */
int orb_bow_detect_candidates(orb_ctx* c, orb_vocab* v, const int32_t* qw, const double* qv, int nq, int nkf, const int32_t* kf_start,
                              const int32_t* kf_word, const double* kf_val, const uint8_t* excluded, int loop, float min_score,
                              const int32_t* cov_start, const int32_t* cov_idx, float* kf_score, int32_t* common, int32_t* cand, int* ncand)
{
    if (!c || !v || nq < 0 || nkf < 0 || !ncand) return ORB_ERR_INVALID;
    *ncand = 0;
    if (nkf == 0) return ORB_OK;
    if (!kf_start || !kf_score || !common || !cand || (nq > 0 && (!qw || !qv))) return ORB_ERR_INVALID;
    if (v->scoring != 0) return ORB_ERR_UNSUPPORTED;
    ORB_CUDA(cudaSetDevice(c->device));
    LaneGuard lg(c);
    if (!lg.lane) return ORB_ERR_CUDA;
    cudaStream_t s = lg.lane->stream;
    const bool dev = on_device(kf_start);
    if (on_device(common) != dev || on_device(kf_score) != dev || on_device(cand) != dev || (nq > 0 && on_device(qw) != dev) ||
        (excluded && on_device(excluded) != dev) || (cov_start && on_device(cov_start) != dev)) return ORB_ERR_INVALID;
    int total = 0, ctotal = 0;
    if (dev) {
        ORB_CUDA(cudaMemcpy(&total, kf_start + nkf, 4, cudaMemcpyDeviceToHost));
        if (cov_start) ORB_CUDA(cudaMemcpy(&ctotal, cov_start + nkf, 4, cudaMemcpyDeviceToHost));
    } else { total = kf_start[nkf]; if (cov_start) ctotal = cov_start[nkf]; }
    if (total < 0 || ctotal < 0 || (total > 0 && (!kf_word || !kf_val)) || (ctotal > 0 && !cov_idx)) return ORB_ERR_INVALID;
    const size_t K = (size_t)nkf, T = (size_t)total, Q = (size_t)nq, CT = (size_t)ctotal;
    size_t need = 1024 + al256(K * 4) * 3 + al256(K * 8);                  // first_pos, acc, best_kf, min_key
    if (!dev) need += al256(Q * 4) + al256(Q * 8) + al256((K + 1) * 4) * 2 + al256(T * 4) + al256(T * 8) + al256(K * 4) * 3 + al256(K) + al256(CT * 4);
    int rc = orb_lane_scratch(lg.lane, need + 1024);
    if (rc) return rc;
    uint8_t* p = (uint8_t*)lg.lane->d_scratch;
    auto take = [&](size_t bytes) { uint8_t* r = p; p += al256(std::max<size_t>(bytes, 1)); return r; };
    int* d_int = (int*)take(16);                   // [0] max common, [1] best accumulated score (bits), [2] ncand
    int32_t* d_first = (int32_t*)take(K * 4); float* d_acc = (float*)take(K * 4); int32_t* d_best = (int32_t*)take(K * 4);
    unsigned long long* d_key = (unsigned long long*)take(K * 8);
    int32_t* d_common = common; float* d_score = kf_score; int32_t* d_cand = cand;
    if (!dev) {
        int32_t* a = (int32_t*)take(Q * 4); double* b = (double*)take(Q * 8); int32_t* st = (int32_t*)take((K + 1) * 4);
        int32_t* w = (int32_t*)take(T * 4); double* val = (double*)take(T * 8);
        d_common = (int32_t*)take(K * 4); d_score = (float*)take(K * 4); d_cand = (int32_t*)take(K * 4);
        if (Q) { ORB_CUDA(cudaMemcpyAsync(a, qw, Q * 4, cudaMemcpyHostToDevice, s)); ORB_CUDA(cudaMemcpyAsync(b, qv, Q * 8, cudaMemcpyHostToDevice, s)); }
        ORB_CUDA(cudaMemcpyAsync(st, kf_start, (K + 1) * 4, cudaMemcpyHostToDevice, s));
        if (T) { ORB_CUDA(cudaMemcpyAsync(w, kf_word, T * 4, cudaMemcpyHostToDevice, s)); ORB_CUDA(cudaMemcpyAsync(val, kf_val, T * 8, cudaMemcpyHostToDevice, s)); }
        ORB_CUDA(cudaMemcpyAsync(d_score, kf_score, K * 4, cudaMemcpyHostToDevice, s));
        qw = a; qv = b; kf_start = st; kf_word = w; kf_val = val;
        if (excluded) { uint8_t* x = take(K); ORB_CUDA(cudaMemcpyAsync(x, excluded, K, cudaMemcpyHostToDevice, s)); excluded = x; }
        if (cov_start) {
            int32_t* cs = (int32_t*)take((K + 1) * 4); int32_t* ci = (int32_t*)take(CT * 4);
            ORB_CUDA(cudaMemcpyAsync(cs, cov_start, (K + 1) * 4, cudaMemcpyHostToDevice, s));
            if (CT) ORB_CUDA(cudaMemcpyAsync(ci, cov_idx, CT * 4, cudaMemcpyHostToDevice, s));
            cov_start = cs; cov_idx = ci;
        }
    }
    ORB_CUDA(cudaMemsetAsync(d_int, 0, 16, s));
    const unsigned wblocks = (unsigned)((K * 32 + 255) / 256), tblocks = (unsigned)((K + 255) / 256);
    k_bow_common<<<wblocks, 256, 0, s>>>(qw, nq, nkf, kf_start, kf_word, excluded, d_common, d_int, d_first);
    k_bow_score<<<wblocks, 256, 0, s>>>(qw, qv, nq, nkf, kf_start, kf_word, kf_val, d_common, d_int, 0, d_score, 1);
    k_bow_accumulate<<<tblocks, 256, 0, s>>>(nkf, d_common, d_int, d_score, cov_start, cov_idx, loop, min_score, d_acc, d_best, d_int + 1, d_key);
    k_bow_mark<<<tblocks, 256, 0, s>>>(nkf, d_acc, d_best, d_first, d_int + 1, loop, min_score, d_key);
    k_bow_rank<<<wblocks, 256, 0, s>>>(nkf, d_key, d_cand, d_int + 2);
    ORB_CUDA(cudaGetLastError());
    if (!dev) {
        ORB_CUDA(cudaMemcpyAsync(common, d_common, K * 4, cudaMemcpyDeviceToHost, s));
        ORB_CUDA(cudaMemcpyAsync(kf_score, d_score, K * 4, cudaMemcpyDeviceToHost, s));
        ORB_CUDA(cudaMemcpyAsync(cand, d_cand, K * 4, cudaMemcpyDeviceToHost, s));
    }
    ORB_CUDA(cudaMemcpyAsync(ncand, d_int + 2, 4, cudaMemcpyDeviceToHost, s));
    ORB_CUDA(cudaStreamSynchronize(s));
    return ORB_OK;
}

} // extern "C"
