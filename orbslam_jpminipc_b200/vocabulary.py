"""Host-side mirror of ORB_SLAM::ORBVocabulary = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>
(reference include/ORBVocabulary.h, Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h) for the calls on the path:
loadFromTextFile (:1338-1425), transform (:1127-1193, :1218-1260) and score (:1197-1201, L1Scoring).
The tree lives in HBM; every method runs CUDA kernels of csrc/orb_vocab.cu through the C ABI."""
import ctypes as C

import numpy as np

from ._lib import check, lib, ptr

L1_NORM, L2_NORM, CHI_SQUARE, KL, BHATTACHARYYA, DOT_PRODUCT = range(6)    # DBoW2::ScoringType
TF_IDF, TF, IDF, BINARY = range(4)                                          # DBoW2::WeightingType


class ORBVocabulary:
    def __init__(self, ctx):
        self._ctx = ctx
        self._h = ctx._h if hasattr(ctx, "_h") else ctx
        self._v = C.c_void_p(None)

    # ---- construction
    def loadFromTextFile(self, filename):
        """TemplatedVocabulary::loadFromTextFile; returns False on a malformed file like the reference."""
        self._release()
        return lib().orb_vocab_load_text(self._h, str(filename).encode(), C.byref(self._v)) == 0

    def create(self, k, L, parent, desc, weight, scoring=L1_NORM, weighting=TF_IDF):
        """Tree from arrays in text-file node order (node 0 = root)."""
        self._release()
        parent = np.ascontiguousarray(parent, np.int32)
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        weight = np.ascontiguousarray(weight, np.float64)
        if not (len(parent) == len(desc) == len(weight)):
            raise ValueError("parent / desc / weight must have one row per node")
        check(lib().orb_vocab_create(self._h, k, L, scoring, weighting, len(parent), ptr(parent), ptr(desc), ptr(weight),
                                     C.byref(self._v)), "orb_vocab_create")
        return self

    def _release(self):
        if self._v:
            lib().orb_vocab_destroy(self._v)
            self._v = C.c_void_p(None)

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def info(self):
        v = [C.c_int(0) for _ in range(4)]
        check(lib().orb_vocab_info(self._v, *[C.byref(x) for x in v]), "orb_vocab_info")
        return dict(k=v[0].value, L=v[1].value, nnodes=v[2].value, nwords=v[3].value)

    def empty(self):
        return not self._v or self.info()["nwords"] == 0

    def size(self):
        return self.info()["nwords"]

    # ---- transform
    def transform_features(self, desc, levelsup=0):
        """transform(feature, word_id, weight, nid, levelsup) for every row: (word, weight, node) arrays."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        word = np.zeros(n, np.int32); weight = np.zeros(n, np.float64); node = np.zeros(n, np.int32)
        check(lib().orb_vocab_transform_features(self._h, self._v, ptr(desc), n, levelsup, ptr(word), ptr(weight), ptr(node)),
              "orb_vocab_transform_features")
        return word, weight, node

    def transform(self, desc, levelsup=4):
        """transform(features, BowVector, FeatureVector, levelsup) for one frame.
        Returns ((bow_word, bow_val), (fv_node, fv_start, fv_items))."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        bows, fvs = self.transform_batch(desc[None] if len(desc) else np.zeros((1, 0, 32), np.uint8), [len(desc)], levelsup)
        return bows[0], fvs[0]

    def transform_batch(self, desc, counts, levelsup=4):
        """desc: (nframes, slot_rows, 32) uint8, counts[f] valid rows per frame (the layout ORBextractor.extract_batch returns)."""
        desc = np.ascontiguousarray(desc, np.uint8)
        nframes, slot_rows = desc.shape[0], desc.shape[1]
        counts = np.ascontiguousarray(counts, np.int32)
        cap = max(slot_rows, 1)
        bw = np.zeros((nframes, cap), np.int32); bv = np.zeros((nframes, cap), np.float64); nb = np.zeros(nframes, np.int32)
        fn = np.zeros((nframes, cap), np.int32); fs = np.zeros((nframes, cap + 1), np.int32); fi = np.zeros((nframes, cap), np.int32)
        nf = np.zeros(nframes, np.int32)
        check(lib().orb_vocab_transform_batch(self._h, self._v, ptr(desc), slot_rows, ptr(counts), nframes, levelsup, cap,
                                              ptr(bw), ptr(bv), ptr(nb), ptr(fn), ptr(fs), ptr(fi), ptr(nf)), "orb_vocab_transform_batch")
        bows = [(bw[f, :nb[f]].copy(), bv[f, :nb[f]].copy()) for f in range(nframes)]
        fvs = [(fn[f, :nf[f]].copy(), fs[f, :nf[f] + 1].copy(), fi[f, :fs[f, nf[f]]].copy()) for f in range(nframes)]
        return bows, fvs

    # ---- scoring
    def score_db(self, query_bow, kf_bows, score_all=False):
        """Shared words and L1 scores of one BowVector against many (KeyFrameDatabase::DetectRelocalisationCandidates,
        src/KeyFrameDatabase.cc:198-252).  Returns (common, score, max_common)."""
        qw = np.ascontiguousarray(query_bow[0], np.int32); qv = np.ascontiguousarray(query_bow[1], np.float64)
        start = np.zeros(len(kf_bows) + 1, np.int32)
        for i, b in enumerate(kf_bows):
            start[i + 1] = start[i] + len(b[0])
        words = np.concatenate([np.asarray(b[0], np.int32) for b in kf_bows]) if len(kf_bows) else np.zeros(0, np.int32)
        vals = np.concatenate([np.asarray(b[1], np.float64) for b in kf_bows]) if len(kf_bows) else np.zeros(0, np.float64)
        words = np.ascontiguousarray(words, np.int32); vals = np.ascontiguousarray(vals, np.float64)
        common = np.zeros(len(kf_bows), np.int32); score = np.zeros(len(kf_bows), np.float32)
        mx = C.c_int(0)
        check(lib().orb_bow_score_db(self._h, self._v, ptr(qw), ptr(qv), len(qw), len(kf_bows), ptr(start), ptr(words), ptr(vals),
                                     int(score_all), ptr(common), ptr(score), C.byref(mx)), "orb_bow_score_db")
        return common, score, mx.value

    def detect_candidates(self, query_bow, kf_bows, kf_score, covis=None, excluded=None, loop=False, min_score=0.0):
        """KeyFrameDatabase::DetectRelocalisationCandidates (src/KeyFrameDatabase.cc:198-308) / DetectLoopCandidates (:75-196, loop=True)
        over keyframes in add() order.  covis[k] = GetBestCovisibilityKeyFrames(10) of keyframe k as indices; kf_score = the
        mRelocScore / mLoopScore members (float32 array, updated in place).  Returns (candidate indices in the reference's order, common)."""
        qw = np.ascontiguousarray(query_bow[0], np.int32); qv = np.ascontiguousarray(query_bow[1], np.float64)
        n = len(kf_bows)
        start = np.zeros(n + 1, np.int32)
        for i, b in enumerate(kf_bows):
            start[i + 1] = start[i] + len(b[0])
        words = np.ascontiguousarray(np.concatenate([np.asarray(b[0], np.int32) for b in kf_bows]) if n else np.zeros(0, np.int32), np.int32)
        vals = np.ascontiguousarray(np.concatenate([np.asarray(b[1], np.float64) for b in kf_bows]) if n else np.zeros(0, np.float64), np.float64)
        cs = ci = None
        if covis is not None:
            cs = np.zeros(n + 1, np.int32)
            for i, c_ in enumerate(covis):
                cs[i + 1] = cs[i] + len(c_)
            ci = np.ascontiguousarray(np.concatenate([np.asarray(c_, np.int32) for c_ in covis]) if cs[n] else np.zeros(1, np.int32), np.int32)
        ex = None if excluded is None else np.ascontiguousarray(excluded, np.uint8)
        assert kf_score.dtype == np.float32 and kf_score.flags.c_contiguous and len(kf_score) == n
        common = np.zeros(n, np.int32); cand = np.zeros(max(n, 1), np.int32)
        nc = C.c_int(0)
        check(lib().orb_bow_detect_candidates(self._h, self._v, ptr(qw), ptr(qv), len(qw), n, ptr(start), ptr(words), ptr(vals),
                                              ptr(ex) if ex is not None else None, int(loop), float(min_score),
                                              ptr(cs) if cs is not None else None, ptr(ci) if ci is not None else None,
                                              ptr(kf_score), ptr(common), ptr(cand), C.byref(nc)), "orb_bow_detect_candidates")
        return cand[:nc.value].copy(), common

    def score(self, v1, v2):
        """TemplatedVocabulary::score(v1, v2) rounded to float as every caller in the reference does (float si = ...)."""
        _, s, _ = self.score_db(v1, [v2], score_all=True)
        return float(s[0])
