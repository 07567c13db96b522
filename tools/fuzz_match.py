"""Fuzz sweep of the matcher entry points against the CPU oracle (frame shapes, feature counts, thresholds, ratios, orientation check,
poses, pre-claimed matches, level limits):  gpurun -- 'python tools/fuzz_match.py 0 150'"""
import sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import orbslam_jpminipc_b200 as pkg
from oracle import pyoracle as po
from orbslam_jpminipc_b200.synth import synth_frame, shifted_frame, synth_descriptors

bad = 0; n = 0; t0 = time.time()
def check(ok, what):
    global bad
    if not ok:
        bad += 1; print("MISMATCH", what)

for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(50000 + seed)
    # ---- kNN-2: sizes around the tile boundaries, duplicates, both engines
    nq, ndb = int(rng.integers(0, 700)), int(rng.choice([0, 1, 2, 31, 32, 33, 255, 256, 257, int(rng.integers(1, 9000))]))
    db, q = synth_descriptors(ndb, nq, seed_db=seed, seed_q=seed + 1)
    if ndb > 4: db[rng.integers(0, ndb, 3)] = db[rng.integers(0, ndb, 3)]
    m0 = pkg.ORBmatcher(0.6, True)
    r = po.knn2(q, db)
    g = m0.knn2(q, db)
    check(all(np.array_equal(a, b) for a, b in zip(g, r)), ("knn2", nq, ndb))
    n += 1
    # ---- frame pair
    h, w = int(rng.integers(120, 500)), int(rng.integers(160, 800))
    nf = int(rng.integers(100, 2000))
    ori = bool(rng.integers(0, 2)); ratio = float(rng.choice([0.6, 0.75, 0.8, 0.9, 1.0]))
    a = synth_frame(h, w, 7000 + seed, quadrants=bool(rng.integers(0, 2)))
    b = shifted_frame(a, int(rng.integers(-6, 7)), int(rng.integers(-6, 7)), seed + 1)
    try:
        orc = po.OracleExtractor(nf, 1.2, 8, 1, 20)
        (ka, da), (kb, db_) = orc(a), orc(b)
    except RuntimeError:
        continue
    if len(ka) < 5 or len(kb) < 5: continue
    fx = fy = float(rng.uniform(300, 700)); cx, cy = w / 2.0 + rng.uniform(-5, 5), h / 2.0 + rng.uniform(-5, 5)
    z = rng.uniform(1, 12, len(ka)).astype(np.float32)
    xyz = np.stack([(ka["x"] - cx) / fx * z, (ka["y"] - cy) / fy * z, z], 1).astype(np.float32)
    if rng.random() < 0.2: xyz[rng.integers(0, len(ka), 5), 2] *= -1          # points behind the camera
    T = np.eye(4, dtype=np.float32); T[:3, 3] = rng.uniform(-0.08, 0.08, 3)
    has = (rng.random(len(ka)) < rng.uniform(0.3, 1.0)).astype(np.uint8)
    outl = (rng.random(len(ka)) < 0.05).astype(np.uint8)
    m = pkg.ORBmatcher(ratio, ori)
    gcur, glast = pkg.Frame(m, kb, db_, w, h, fx, fy, cx, cy), pkg.Frame(m, ka, da, w, h, fx, fy, cx, cy)
    ocur, olast = po.OracleFrame(kb, db_, w, h, fx, fy, cx, cy), po.OracleFrame(ka, da, w, h, fx, fy, cx, cy)
    what = (seed, h, w, nf, ori, ratio)
    th = float(rng.choice([3.0, 7.0, 15.0, 30.0]))
    pre = np.full(gcur.N, -1, np.int32); pre[::int(rng.integers(3, 15))] = 4242
    g = m.SearchByProjection(gcur, glast, th, has, outl, xyz, T, match_cur=pre.copy()); r = po.search_by_projection(ocur, olast, has, outl, xyz, T, th, ori, match_cur=pre.copy())
    check(g[0] == r[0] and np.array_equal(g[1], r[1]), ("SearchByProjection",) + what + (th,))
    nn = glast.N
    px = (glast.kps["x"] + rng.normal(0, 2.0, nn)).astype(np.float32); py = (glast.kps["y"] + rng.normal(0, 2.0, nn)).astype(np.float32)
    level = np.clip(glast.kps["octave"] + rng.integers(-1, 2, nn), 0, 7).astype(np.int32)
    vc = rng.uniform(0.9, 1.0, nn).astype(np.float32); inv = (rng.random(nn) < 0.85).astype(np.uint8)
    th2 = float(rng.choice([1.0, 3.0, 5.0]))
    g = m.SearchByProjectionMapPoints(gcur, inv, px, py, level, vc, glast.desc, th2, match_f=pre.copy()); r = po.search_by_projection_mappoints(ocur, inv, px, py, level, vc, glast.desc, th2, ratio, match_f=pre.copy())
    check(g[0] == r[0] and np.array_equal(g[1], r[1]), ("SearchByProjectionMapPoints",) + what + (th2,))
    win = int(rng.choice([5, 15, 30, 100])); minl = int(rng.choice([-1, 0, 2])); maxl = int(rng.choice([2 ** 31 - 1, 5, 3]))
    g = m.WindowSearch(glast, gcur, win, has, minScaleLevel=minl, maxScaleLevel=maxl); r = po.window_search(olast, ocur, has, win, ratio, ori, min_level=minl, max_level=maxl)
    check(g[0] == r[0] and np.array_equal(g[1], r[1]), ("WindowSearch",) + what + (win, minl, maxl))
    g = m.SearchByProjectionWindow(glast, gcur, win, has, xyz, T, pre.copy()); r = po.search_by_projection_window(olast, ocur, has, xyz, T, win, ratio, pre.copy())
    check(g[0] == r[0] and np.array_equal(g[1], r[1]), ("SearchByProjectionWindow",) + what + (win,))
    prev = np.stack([glast.kps["x"], glast.kps["y"]], 1).astype(np.float32)
    g = m.SearchForInitialization(glast, gcur, prev, win); r = po.search_for_initialization(olast, ocur, prev, win, ratio, ori)
    check(g[0] == r[0] and np.array_equal(g[1], r[1]) and np.array_equal(g[2], r[2]), ("SearchForInitialization",) + what + (win,))
    pl = np.clip(glast.kps["octave"] + rng.integers(-1, 2, nn), 0, 7).astype(np.int32)
    dist = int(rng.choice([50, 64, 100]))
    g = m.SearchByProjectionKeyFrame(gcur, has, xyz, T, pl, glast.desc, glast.kps["angle"], th, dist, match_cur=pre.copy()); r = po.search_by_projection_kf(ocur, has, xyz, T, pl, glast.desc, glast.kps["angle"], th, dist, ori, match_cur=pre.copy())
    check(g[0] == r[0] and np.array_equal(g[1], r[1]), ("SearchByProjectionKeyFrame",) + what + (th, dist))
    n += 6
print("checks", n, "bad", bad, "%.1f s" % (time.time() - t0))
