"""GPU parity tests of the vocabulary path (csrc/orb_vocab.cu) against the oracle's DBoW2 restatement: tree descent, BowVector
(doubles, bit-exact), FeatureVector, text loading, retrieval scoring, and the chain extract -> transform -> SearchByBoW."""
import numpy as np
import pytest

from orbslam_jpminipc_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import orbslam_jpminipc_b200 as p
    return p


@pytest.fixture(scope="module")
def po():
    from oracle import pyoracle
    return pyoracle


@pytest.fixture(scope="module")
def ctx(pkg):
    return pkg.ORBmatcher(0.75, True)


def _features(desc, parent, n, seed, flip=0.04):
    rng = np.random.default_rng(seed)
    has_child = np.zeros(len(parent), bool)
    has_child[parent[1:]] = True
    leaves = np.flatnonzero(~has_child)
    leaves = leaves[leaves > 0]
    src = rng.choice(leaves, n)
    bits = np.unpackbits(desc[src], axis=1)
    return np.packbits(bits ^ (rng.random(bits.shape) < flip), axis=1)


def _same_bow_fv(a, b):
    (aw, av), (an, as_, ai) = a
    (bw, bv), (bn, bs, bi) = b
    return (np.array_equal(aw, bw) and np.array_equal(av.view(np.uint64), bv.view(np.uint64)) and np.array_equal(an, bn) and
            np.array_equal(as_, bs) and np.array_equal(ai, bi))


@pytest.mark.parametrize("k,L,levelsup,scoring,weighting,prune,order,n", [
    (10, 3, 2, 0, 0, 0.0, "bfs", 2000), (4, 3, 1, 0, 0, 0.0, "bfs", 500), (5, 4, 4, 0, 0, 0.2, "dfs", 1000),
    (3, 5, 2, 1, 1, 0.15, "bfs", 777), (6, 3, 0, 5, 1, 0.0, "dfs", 300), (4, 3, 1, 0, 2, 0.1, "bfs", 64), (4, 3, 5, 2, 3, 0.0, "bfs", 1),
    (18, 2, 1, 0, 0, 0.0, "bfs", 900), (20, 2, 0, 0, 0, 0.05, "dfs", 4000)])
def test_transform_vs_oracle(pkg, po, ctx, k, L, levelsup, scoring, weighting, prune, order, n):
    parent, desc, weight = synth.synth_vocabulary(k, L, seed=k * 10 + L, stop_frac=0.05, prune_frac=prune, order=order)
    feats = _features(desc, parent, n, 99)
    ov = po.OracleVocabulary(k, L, parent, desc, weight, scoring, weighting)
    gv = pkg.ORBVocabulary(ctx).create(k, L, parent, desc, weight, scoring, weighting)
    assert gv.info()["nnodes"] == ov.nnodes and gv.size() == ov.nwords
    w, wt, nd = gv.transform_features(feats, levelsup)
    rw, rwt, rnd = ov.transform_features(feats, levelsup)
    assert np.array_equal(w, rw) and np.array_equal(wt, rwt) and np.array_equal(nd, rnd)
    assert _same_bow_fv(gv.transform(feats, levelsup), ov.transform(feats, levelsup))


def test_transform_ties_pick_first_child(pkg, po, ctx):
    """Children with identical descriptors: the strict '<' scan keeps the first (TemplatedVocabulary.h:1238-1247)."""
    k, L = 6, 2
    parent, desc, weight = synth.synth_vocabulary(k, L, seed=1, stop_frac=0.0)
    desc[1:] = desc[1]                                     # every node equal -> every distance ties
    feats = np.random.default_rng(0).integers(0, 256, (100, 32), dtype=np.uint8)
    ov = po.OracleVocabulary(k, L, parent, desc, weight)
    gv = pkg.ORBVocabulary(ctx).create(k, L, parent, desc, weight)
    a, b = gv.transform_features(feats, 1), ov.transform_features(feats, 1)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert (a[0] == 0).all()                               # first leaf of the first child


def test_text_loading(pkg, po, ctx, tmp_path):
    parent, desc, weight = synth.synth_vocabulary(5, 3, seed=3, stop_frac=0.05, prune_frac=0.1)
    path = tmp_path / "voc.txt"
    synth.write_vocabulary_text(path, 5, 3, parent, desc, weight)
    gv = pkg.ORBVocabulary(ctx)
    assert gv.loadFromTextFile(path)
    ov = po.OracleVocabulary(path=path)
    feats = _features(desc, parent, 600, 5)
    assert _same_bow_fv(gv.transform(feats, 1), ov.transform(feats, 1))
    bad = tmp_path / "bad.txt"
    bad.write_text("99 3 0 0\n")
    assert not pkg.ORBVocabulary(ctx).loadFromTextFile(bad)
    assert not pkg.ORBVocabulary(ctx).loadFromTextFile(tmp_path / "missing.txt")


def test_batch_ragged_and_empty(pkg, po, ctx):
    k, L = 8, 3
    parent, desc, weight = synth.synth_vocabulary(k, L, seed=11, stop_frac=0.05)
    ov = po.OracleVocabulary(k, L, parent, desc, weight)
    gv = pkg.ORBVocabulary(ctx).create(k, L, parent, desc, weight)
    counts = [700, 0, 1, 33, 1024, 512]
    slot = 1024
    batch = np.zeros((len(counts), slot, 32), np.uint8)
    for f, c in enumerate(counts):
        batch[f, :c] = _features(desc, parent, c, 100 + f) if c else 0
    bows, fvs = gv.transform_batch(batch, counts, 2)
    for f, c in enumerate(counts):
        assert _same_bow_fv((bows[f], fvs[f]), ov.transform(batch[f, :c], 2))
    assert len(bows[1][0]) == 0 and len(fvs[1][0]) == 0
    e = gv.transform(np.zeros((0, 32), np.uint8), 2)
    assert len(e[0][0]) == 0 and len(e[1][0]) == 0


def test_score_db_vs_oracle(pkg, po, ctx):
    k, L = 10, 3
    parent, desc, weight = synth.synth_vocabulary(k, L, seed=21, stop_frac=0.02)
    gv = pkg.ORBVocabulary(ctx).create(k, L, parent, desc, weight)
    nkf = 300
    batch = np.stack([_features(desc, parent, 800, 500 + (f % 40), flip=0.02 + 0.01 * (f % 3)) for f in range(nkf)])
    bows, _ = gv.transform_batch(batch, [800] * nkf, 1)
    q = bows[7]
    common, score, mx = gv.score_db(q, bows)
    rcommon, rscore, rmx = po.bow_score_db(q, bows)
    assert mx == rmx == len(q[0])
    assert np.array_equal(common, rcommon) and np.array_equal(score.view(np.uint32), rscore.view(np.uint32))
    assert score[7] == 1.0 and (score > 0).sum() >= 2
    _, sall, _ = gv.score_db(q, bows, score_all=True)
    for i in (0, 5, 100, 299):
        assert sall[i] == np.float32(po.bow_score_l1(q, bows[i]))
    assert gv.score(q, bows[9]) == np.float32(po.bow_score_l1(q, bows[9]))


@pytest.mark.parametrize("loop,min_score,seed,device_arrays", [(False, 0.0, 1, False), (False, 0.0, 2, True), (True, 0.05, 3, False),
                                                              (True, 0.3, 4, True), (True, 1.5, 5, False)])
def test_detect_candidates_vs_oracle(pkg, po, ctx, loop, min_score, seed, device_arrays):
    """orb_bow_detect_candidates = KeyFrameDatabase::DetectRelocalisationCandidates / DetectLoopCandidates complete (covisibility
    accumulation with stale member scores, 0.75 * best cut, result order), host and device arrays, against the oracle (which
    tests/test_ref_build.py pins to the reference's own KeyFrameDatabase.cc)."""
    import ctypes as C
    k, L = 10, 3
    parent, desc, weight = synth.synth_vocabulary(k, L, seed=21, stop_frac=0.02)
    gv = pkg.ORBVocabulary(ctx).create(k, L, parent, desc, weight)
    rng = np.random.default_rng(100 + seed)
    nkf = 400
    # 40 "places", ten views each with a varying share of the place's features: the query is one more view of place 7
    batch = np.stack([_features(desc, parent, 800, 500 + (f % 40), flip=0.02 + 0.01 * (f % 3)) for f in range(nkf)])
    for f in range(nkf):
        cut = int(rng.integers(100, 800))
        batch[f, cut:] = _features(desc, parent, 800 - cut, 9000 + f)
    bows, _ = gv.transform_batch(batch, [800] * nkf, 1)
    q = gv.transform_batch(_features(desc, parent, 800, 507, flip=0.03)[None], [800], 1)[0][0]
    covis = [list(rng.choice(nkf, int(rng.integers(0, 15)), replace=False)) for _ in range(nkf)]
    covis = [[int(j) for j in c if j != i] for i, c in enumerate(covis)]
    excluded = (rng.random(nkf) < 0.1).astype(np.uint8) if loop else None
    state = (rng.random(nkf) * 0.2).astype(np.float32)
    os_ = state.copy()
    ocand, ocommon = po.bow_detect_candidates(q, bows, os_, covis=covis, excluded=excluded, loop=loop, min_score=min_score)
    gs = state.copy()
    if not device_arrays:
        gcand, gcommon = gv.detect_candidates(q, bows, gs, covis=covis, excluded=excluded, loop=loop, min_score=min_score)
    else:
        import torch
        from orbslam_jpminipc_b200._lib import check, lib, ptr
        dev = torch.device("cuda", 0)
        start = np.zeros(nkf + 1, np.int32); start[1:] = np.cumsum([len(b[0]) for b in bows])
        cs = np.zeros(nkf + 1, np.int32); cs[1:] = np.cumsum([len(c) for c in covis])
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        d = dict(qw=t(q[0].astype(np.int32)), qv=t(q[1]), st=t(start), w=t(np.concatenate([b[0] for b in bows]).astype(np.int32)),
                 v=t(np.concatenate([b[1] for b in bows])), cs=t(cs), ci=t(np.concatenate([np.asarray(c, np.int32) for c in covis])),
                 sc=t(gs), cm=torch.zeros(nkf, dtype=torch.int32, device=dev), cand=torch.zeros(nkf, dtype=torch.int32, device=dev))
        ex = t(excluded) if excluded is not None else None
        nc = C.c_int(0)
        check(lib().orb_bow_detect_candidates(gv._h, gv._v, ptr(d["qw"]), ptr(d["qv"]), len(q[0]), nkf, ptr(d["st"]), ptr(d["w"]), ptr(d["v"]),
                                              ptr(ex) if ex is not None else None, int(loop), float(min_score), ptr(d["cs"]), ptr(d["ci"]),
                                              ptr(d["sc"]), ptr(d["cm"]), ptr(d["cand"]), C.byref(nc)), "orb_bow_detect_candidates")
        gcand, gcommon, gs = d["cand"].cpu().numpy()[:nc.value], d["cm"].cpu().numpy(), d["sc"].cpu().numpy()
    assert np.array_equal(gcommon, ocommon)
    assert np.array_equal(gs.view(np.uint32), os_.view(np.uint32))
    assert list(gcand) == list(ocand)
    if min_score < 1.0:
        assert len(ocand) >= 1 and (os_ != state).sum() >= 3
    else:
        assert len(ocand) == 0                               # an L1 score never exceeds 1: nothing reaches minScore, the reference returns an empty vector (:137)
    # an empty database and a query without words
    e, _ = gv.detect_candidates(q, [], np.zeros(0, np.float32))
    assert len(e) == 0
    e, c0 = gv.detect_candidates((np.zeros(0, np.int32), np.zeros(0, np.float64)), bows[:5], np.zeros(5, np.float32), covis=covis[:5] and [[] for _ in range(5)])
    assert len(e) == 0 and not c0.any()


def test_extract_transform_search_by_bow_chain(pkg, po):
    """Frame::ComputeBoW feeding SearchByBoW (src/Tracking.cc:907-927): both sides of the chain against the oracle."""
    ext = pkg.ORBextractor(1000, 1.2, 8, 1, 20, max_width=640, max_height=480)
    oext = po.OracleExtractor(1000, 1.2, 8, 1, 20)
    img_a = synth.synth_frame(480, 640, 31)
    img_b = synth.shifted_frame(img_a, 3, 2, 32)
    ka, da = ext(img_a)
    kb, db = ext(img_b)
    oka, oda = oext(img_a)
    assert np.array_equal(da, oda)
    k, L = 10, 4
    parent, desc, weight = synth.synth_vocabulary_fast(k, L, seed=2, flip_bits=40, stop_frac=0.01)
    m = pkg.ORBmatcher(0.75, True)
    gv = pkg.ORBVocabulary(m).create(k, L, parent, desc, weight)
    ov = po.OracleVocabulary(k, L, parent, desc, weight)
    (bwa, fva), (bwb, fvb) = gv.transform(da, 2), gv.transform(db, 2)
    assert _same_bow_fv((bwa, fva), ov.transform(da, 2)) and _same_bow_fv((bwb, fvb), ov.transform(db, 2))
    valid = np.ones(len(ka), np.uint8)
    n, match = m.SearchByBoW(fva, da, ka, valid, fvb, db, kb)
    rn, rmatch = po.search_by_bow(fva, da, ka, valid, fvb, db, kb, 0.75, True)
    assert n == rn and np.array_equal(match, rmatch) and n > 50


def test_full_size_tree_k10_L6(pkg, po, ctx):
    """The reference's vocabulary shape (k=10, L=6: 1 111 111 nodes, 10^6 words), levelsup=4 as in Frame::ComputeBoW."""
    k, L = 10, 6
    parent, desc, weight = synth.synth_vocabulary_fast(k, L, seed=7)
    gv = pkg.ORBVocabulary(ctx).create(k, L, parent, desc, weight)
    ov = po.OracleVocabulary(k, L, parent, desc, weight)
    info = gv.info()
    assert info["nnodes"] == 1111111 and info["nwords"] == 1000000
    feats = _features(desc, parent, 2000, 77, flip=0.03)
    assert _same_bow_fv(gv.transform(feats, 4), ov.transform(feats, 4))
    (bw, bv), (fn, fs, fi) = gv.transform(feats, 4)
    assert abs(bv.sum() - 1.0) < 1e-12 and len(fn) <= 100 and (np.diff(fn) > 0).all()
    # batch of 64 frames == frame by frame
    batch = np.stack([_features(desc, parent, 1000, 900 + f) for f in range(64)])
    bows, fvs = gv.transform_batch(batch, [1000] * 64, 4)
    for f in (0, 17, 63):
        assert _same_bow_fv((bows[f], fvs[f]), ov.transform(batch[f], 4))
