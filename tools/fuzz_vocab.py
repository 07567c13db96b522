"""Fuzz sweep of the vocabulary path against the oracle's DBoW2 restatement (branching factor, depth, pruning, node order, scoring /
weighting type, levelsup, feature counts, ragged batches, retrieval queries with covisibility):  gpurun -- 'python tools/fuzz_vocab.py 0 200'"""
import sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import orbslam_jpminipc_b200 as pkg
from oracle import pyoracle as po
from orbslam_jpminipc_b200 import synth
from test_gpu_vocab import _features, _same_bow_fv

bad = 0; n = 0; t0 = time.time()
ctx = pkg.ORBmatcher(0.75, True)
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(30000 + seed)
    k, L = int(rng.integers(2, 21)), int(rng.integers(1, 6))
    while k ** L > 60000: L -= 1
    prune = float(rng.choice([0.0, 0.0, 0.05, 0.2])); order = str(rng.choice(["bfs", "dfs"]))
    scoring = int(rng.choice([0, 0, 0, 1, 2, 5])); weighting = int(rng.integers(0, 4)); levelsup = int(rng.integers(0, L + 2))
    parent, desc, weight = synth.synth_vocabulary(k, L, seed=seed, stop_frac=float(rng.choice([0.0, 0.05])), prune_frac=prune, order=order)
    ov = po.OracleVocabulary(k, L, parent, desc, weight, scoring, weighting)
    gv = pkg.ORBVocabulary(ctx).create(k, L, parent, desc, weight, scoring, weighting)
    nfeat = int(rng.choice([0, 1, 31, 257, int(rng.integers(1, 3000))]))
    feats = _features(desc, parent, nfeat, seed, flip=float(rng.choice([0.0, 0.03, 0.2]))) if nfeat else np.zeros((0, 32), np.uint8)
    what = (seed, k, L, prune, order, scoring, weighting, levelsup, nfeat)
    ok = _same_bow_fv(gv.transform(feats, levelsup), ov.transform(feats, levelsup))
    if nfeat:
        a, b = gv.transform_features(feats, levelsup), ov.transform_features(feats, levelsup)
        ok = ok and all(np.array_equal(x, y) for x, y in zip(a, b))
    n += 1
    if not ok: bad += 1; print("MISMATCH transform", what)
    if scoring != 0: continue
    # retrieval: a small database, both queries
    nkf = int(rng.integers(1, 120)); cnt = [int(rng.integers(0, 600)) for _ in range(nkf)]
    slot = max(max(cnt), 1)
    batch = np.zeros((nkf, slot, 32), np.uint8)
    for f, c in enumerate(cnt):
        if c: batch[f, :c] = _features(desc, parent, c, 1000 + seed * 7 + (f % 9))
    bows, _ = gv.transform_batch(batch, cnt, levelsup)
    q = bows[int(rng.integers(0, nkf))]
    covis = [[int(j) for j in rng.choice(nkf, int(rng.integers(0, min(nkf, 14))), replace=False) if j != i] for i in range(nkf)]
    loop = bool(rng.integers(0, 2)); ms = float(rng.choice([0.0, 0.05, 0.3]))
    excl = (rng.random(nkf) < 0.15).astype(np.uint8) if loop else None
    st = (rng.random(nkf) * 0.3).astype(np.float32)
    os_, gs = st.copy(), st.copy()
    oc, ocm = po.bow_detect_candidates(q, bows, os_, covis=covis, excluded=excl, loop=loop, min_score=ms)
    gc, gcm = gv.detect_candidates(q, bows, gs, covis=covis, excluded=excl, loop=loop, min_score=ms)
    n += 1
    if not (list(oc) == list(gc) and np.array_equal(ocm, gcm) and np.array_equal(os_.view(np.uint32), gs.view(np.uint32))):
        bad += 1; print("MISMATCH candidates", what, nkf, loop, ms, list(oc), list(gc))
print("checks", n, "bad", bad, "%.1f s" % (time.time() - t0))
