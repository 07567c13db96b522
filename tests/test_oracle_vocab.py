"""CPU tests of the vocabulary oracle (oracle/orb_oracle.cpp, DBoW2 section).

The reference carries no tests or golden vectors for DBoW2 and its sources need OpenCV's C++ headers (absent here), so this
part of the oracle is NOT pinned against the reference binary ("parity unpinned").  What is checked instead: an
independent restatement written directly from TemplatedVocabulary.h / BowVector.cpp / ScoringObject.cpp in plain Python
(dict + float, same summation order), the text-format round trip, and the closed form of the L1 score."""
import numpy as np
import pytest

from orbslam_jpminipc_b200 import synth


@pytest.fixture(scope="module")
def po():
    from oracle import pyoracle
    return pyoracle


def _py_transform(k, L, parent, desc, weight, feats, levelsup, scoring=0, weighting=0):
    n = len(parent)
    children = [[] for _ in range(n)]
    for i in range(1, n):
        children[parent[i]].append(i)
    word_id = {}
    for i in range(1, n):
        if not children[i]:
            word_id[i] = len(word_id)
    bits = np.unpackbits(desc, axis=1)
    fbits = np.unpackbits(feats, axis=1)
    bow, fv = {}, {}
    per = []
    for i in range(len(feats)):
        nid_level = L - levelsup
        nid = 0 if nid_level <= 0 else None
        cur, lvl = 0, 0
        while children[cur]:
            lvl += 1
            ch = children[cur]
            d = [(int((fbits[i] != bits[c]).sum())) for c in ch]
            cur = ch[int(np.argmin(d))]          # argmin returns the first minimum == strict '<' scan
            if lvl == nid_level:
                nid = cur
        if nid is None:
            nid = cur
        w = float(weight[cur])
        per.append((word_id[cur], w, nid))
        if w > 0:
            wid = word_id[cur]
            if wid not in bow:
                bow[wid] = w
            elif weighting in (0, 1):
                bow[wid] += w
            fv.setdefault(nid, []).append(i)
    words = sorted(bow)
    vals = [bow[w] for w in words]
    must = scoring != 5
    if weighting in (0, 1) and vals and not must:
        vals = [v / float(len(vals)) for v in vals]
    if must:
        norm = 0.0
        if scoring == 1:
            for v in vals:
                norm += v * v
            norm = float(np.sqrt(norm))
        else:
            for v in vals:
                norm += abs(v)
        if norm > 0:
            vals = [v / norm for v in vals]
    nodes = sorted(fv)
    return per, (np.asarray(words, np.int32), np.asarray(vals, np.float64)), (nodes, [fv[nd] for nd in nodes])


def _features(desc, parent, n, seed):
    rng = np.random.default_rng(seed)
    has_child = np.zeros(len(parent), bool)
    has_child[parent[1:]] = True
    leaves = np.flatnonzero(~has_child)[1:] if not has_child[0] else np.flatnonzero(~has_child)
    src = rng.choice(leaves, n)
    bits = np.unpackbits(desc[src], axis=1)
    flip = rng.random(bits.shape) < 0.04
    return np.packbits(bits ^ flip, axis=1)


@pytest.mark.parametrize("k,L,levelsup,scoring,weighting,prune,order", [
    (4, 3, 1, 0, 0, 0.0, "bfs"), (10, 3, 2, 0, 0, 0.0, "bfs"), (5, 4, 4, 0, 0, 0.2, "dfs"), (3, 5, 2, 1, 1, 0.15, "bfs"),
    (6, 3, 0, 5, 1, 0.0, "dfs"), (4, 3, 1, 0, 2, 0.1, "bfs"), (4, 3, 5, 2, 3, 0.0, "bfs")])
def test_vocab_oracle_vs_python_restatement(po, k, L, levelsup, scoring, weighting, prune, order):
    parent, desc, weight = synth.synth_vocabulary(k, L, seed=k * 10 + L, stop_frac=0.05, prune_frac=prune, order=order)
    feats = _features(desc, parent, 300, 99)
    voc = po.OracleVocabulary(k, L, parent, desc, weight, scoring, weighting)
    per, (pw, pv), (pn, pitems) = _py_transform(k, L, parent, desc, weight, feats, levelsup, scoring, weighting)
    word, wt, node = voc.transform_features(feats, levelsup)
    assert [int(x) for x in word] == [p[0] for p in per]
    assert np.array_equal(wt, np.asarray([p[1] for p in per]))
    assert [int(x) for x in node] == [p[2] for p in per]
    (bw, bv), (fn, fs, fi) = voc.transform(feats, levelsup)
    assert np.array_equal(bw, pw) and np.array_equal(bv, pv)            # doubles bit-exact
    assert list(fn) == pn
    for j in range(len(pn)):
        assert list(fi[fs[j]:fs[j + 1]]) == pitems[j]
    if scoring == 0 and len(bv):
        assert abs(bv.sum() - 1.0) < 1e-12


def test_vocab_text_round_trip(po, tmp_path):
    parent, desc, weight = synth.synth_vocabulary(5, 3, seed=3, stop_frac=0.05, prune_frac=0.1)
    path = tmp_path / "voc.txt"
    synth.write_vocabulary_text(path, 5, 3, parent, desc, weight)
    a = po.OracleVocabulary(5, 3, parent, desc, weight)
    b = po.OracleVocabulary(path=path)
    assert a.nnodes == b.nnodes and a.nwords == b.nwords
    feats = _features(desc, parent, 200, 5)
    ra, rb = a.transform(feats, 1), b.transform(feats, 1)
    for x, y in zip(ra[0] + ra[1], rb[0] + rb[1]):
        assert np.array_equal(x, y)


def test_l1_score_closed_form(po):
    rng = np.random.default_rng(8)
    for _ in range(20):
        def bow(n):
            w = np.sort(rng.choice(500, n, replace=False)).astype(np.int32)
            v = rng.random(n); v /= v.sum()
            return w, v
        a, b = bow(int(rng.integers(1, 200))), bow(int(rng.integers(1, 200)))
        dense_a = np.zeros(500); dense_a[a[0]] = a[1]
        dense_b = np.zeros(500); dense_b[b[0]] = b[1]
        expect = 1.0 - 0.5 * np.abs(dense_a - dense_b).sum()
        assert abs(po.bow_score_l1(a, b) - expect) < 1e-12
    assert po.bow_score_l1(a, a) == pytest.approx(1.0, abs=1e-15)


def test_score_db_oracle(po):
    rng = np.random.default_rng(9)

    def bow(n):
        w = np.sort(rng.choice(300, n, replace=False)).astype(np.int32)
        v = rng.random(n); v /= v.sum()
        return w, v
    q = bow(120)
    kfs = [bow(int(rng.integers(1, 200))) for _ in range(40)]
    common, score, mx = po.bow_score_db(q, kfs)
    exp_common = np.asarray([len(np.intersect1d(q[0], kf[0])) for kf in kfs])
    assert np.array_equal(common, exp_common) and mx == exp_common.max()
    thr = int(np.float32(mx) * np.float32(0.8))
    for i, kf in enumerate(kfs):
        if exp_common[i] > thr:
            assert score[i] == np.float32(po.bow_score_l1(q, kf))
        else:
            assert score[i] == 0
