"""The oracle against the REFERENCE's own code.

oracle/_ref/libref_orbslam.so is the reference's src/ORBextractor.cc, ORBmatcher.cc, Frame.cc, KeyFrame.cc, MapPoint.cc, Map.cc,
KeyFrameDatabase.cc and Thirdparty/DBoW2 compiled from /root/reference against oracle/refshim/ (containers re-implemented,
OpenCV arithmetic primitives forwarded to the oracle's cv2-pinned restatements).  These tests pin the oracle's restatement of the
reference's own control flow — grid / quota / fallback / order in the extractor, every ORBmatcher search, the Frame grid, DBoW2's
tree descent, BowVector / FeatureVector construction and L1 score — by running the reference's member functions on the same inputs.
They run wherever the built library is present (the build container; it also travels to the GPU box) and skip otherwise."""
import numpy as np
import pytest

from oracle import pyref

pytestmark = pytest.mark.skipif(not pyref.available(), reason="oracle/_ref not built (needs /root/reference at build time)")


@pytest.fixture(scope="module")
def po():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="module")
def pkg():
    import orbslam_jpminipc_b200 as p
    return p


def _same_keypoints(a, b):
    assert len(a) == len(b)
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(a[f], b[f]), f
    assert np.array_equal(a["angle"].view(np.uint32), b["angle"].view(np.uint32))      # bit pattern of the float


# --------------------------------------------------------------------------------------------- ORBextractor::operator()
@pytest.mark.parametrize("h,w,nf,score,fast_th,nlevels,scale", [
    (480, 640, 1000, 1, 20, 8, 1.2),        # BASELINE config 0
    (480, 752, 1000, 1, 20, 8, 1.2),        # config 1
    (376, 1241, 2000, 1, 20, 8, 1.2),       # config 2
    (480, 640, 2000, 1, 20, 8, 1.2),        # the initialisation extractor, src/Tracking.cc:126
    (480, 640, 1000, 0, 20, 8, 1.2),        # HARRIS_SCORE
    (240, 320, 300, 1, 20, 8, 1.2),
    (200, 300, 500, 1, 5, 4, 1.5),          # fastTh below the fallback threshold, other pyramid
    (120, 160, 100, 1, 40, 3, 2.0),
])
def test_extractor_equals_reference(po, h, w, nf, score, fast_th, nlevels, scale):
    from orbslam_jpminipc_b200.synth import synth_frame
    ref = pyref.RefExtractor(nf, scale, nlevels, score, fast_th)
    orc = po.OracleExtractor(nf, scale, nlevels, score, fast_th)
    total = 0
    for seed in (1000, 1001, 1002):
        for quadrants in (True, False):
            img = synth_frame(h, w, seed, quadrants=quadrants)
            rk, rd = ref(img)
            ok, od = orc(img)
            _same_keypoints(rk, ok)
            assert np.array_equal(rd, od)
            total += len(rk)
    assert total > 100


def test_extractor_degenerate_frames(po):
    ref = pyref.RefExtractor(500, 1.2, 8, 1, 20)
    orc = po.OracleExtractor(500, 1.2, 8, 1, 20)
    rng = np.random.default_rng(5)
    flat = np.full((240, 320), 128, np.uint8)
    noise = rng.integers(0, 256, (240, 320), dtype=np.uint8)
    low = (128 + rng.integers(-9, 10, (240, 320))).astype(np.uint8)         # only the th=7 fallback fires
    half = flat.copy(); half[:, 160:] = noise[:, 160:]                        # empty cells -> quota redistribution
    for img in (flat, noise, low, half):
        rk, rd = ref(img)
        ok, od = orc(img)
        _same_keypoints(rk, ok)
        assert np.array_equal(rd, od)
    assert len(ref(flat)[0]) == 0 and len(ref(noise)[0]) > 400


@pytest.mark.parametrize("h,w,nf,scale,nlevels,fast_th,throws", [
    (261, 623, 1436, 1.5, 6, 9, False),     # level 5 is 82x34: 4 cell rows of height 1 over a 2-row region, the third detects at y = 18 = h - 16
    (673, 239, 1436, 1.25, 1, 9, False),    # 28 columns of width 8 over 207: columns 26 reaches 8 px past w - 16, column 27 is skipped
    (206, 796, 1486, 1.3, 6, 40, False),    # keypoints closer than 16 px to the ROI edge: IC_Angle / rBRIEF read 15 px of the frame
    (495, 298, 1092, 1.44, 1, 12, False),   # 209 cells on one level
    (231, 698, 2278, 1.1, 1, 0, False),     # fastTh = 0
    (360, 915, 2216, 1.7, 8, 1, True),      # a level smaller than its margins: an INNER cell row has negative height -> Mat::rowRange throws
    (234, 523, 1597, 1.5, 8, 20, True),
])
def test_extractor_degenerate_grids(po, h, w, nf, scale, nlevels, fast_th, throws):
    """Grids where cellW = ceil(W / cols) makes inner cells reach past [16, size - 16) or where a level is smaller than its margins
    (found by the GPU fuzz sweep of round 2: the CUDA path cut detection at size - 16 and skipped what the reference throws on).  The
    reference's own ORBextractor.cc decides; the oracle must agree, keypoint for keypoint or throw for throw."""
    from orbslam_jpminipc_b200.synth import synth_frame
    ref = pyref.RefExtractor(nf, scale, nlevels, 1, fast_th)
    orc = po.OracleExtractor(nf, scale, nlevels, 1, fast_th)
    rng = np.random.default_rng(h * w)
    for img in (synth_frame(h, w, 31, quadrants=False), rng.integers(0, 256, (h, w), dtype=np.uint8)):
        if throws:
            with pytest.raises(RuntimeError):
                ref(img)
            with pytest.raises(RuntimeError):
                orc(img)
            continue
        rk, rd = ref(img)
        ok, od = orc(img)
        _same_keypoints(rk, ok)
        assert np.array_equal(rd, od) and len(rk) > 50


# --------------------------------------------------------------------------------------------- Frame: undistortion, bounds, grid
@pytest.mark.parametrize("dist", [(0, 0, 0, 0), (-0.28, 0.07, 0.0002, 0.00002), (0.1, -0.05, 0.001, -0.002)])
def test_frame_constructor_equals_oracle(po, dist):
    from orbslam_jpminipc_b200.synth import synth_frame
    h, w, fx, fy, cx, cy = 480, 752, 458.654, 457.296, 367.215, 248.375          # EuRoC cam0
    img = synth_frame(h, w, 1234)
    ex = pyref.RefExtractor(1000, 1.2, 8, 1, 20)
    rf = pyref.RefFrame.from_image(ex, img, fx, fy, cx, cy, dist)
    keys, keys_un, desc, bounds = rf.get()
    ok, od = po.OracleExtractor(1000, 1.2, 8, 1, 20)(img)
    _same_keypoints(keys, ok)
    assert np.array_equal(desc, od)
    K = (fx, fy, cx, cy)
    d = np.array(dist, np.float32)
    oun = po.undistort_keypoints(ok, K, d)
    assert np.array_equal(keys_un.view(np.uint8), np.ascontiguousarray(oun).view(np.uint8))
    ob = po.image_bounds(w, h, K, d)                                          # minX maxX minY maxY
    assert list(bounds) == list(ob)
    of = po.OracleFrame(oun, od, w, h, fx, fy, cx, cy, bounds=ob)
    start, items = rf.grid()
    assert np.array_equal(start, of.cell_start) and np.array_equal(items, of.cell_items[:start[-1]])
    rng = np.random.default_rng(1)
    for _ in range(300):
        x, y, r = rng.uniform(-20, w + 20), rng.uniform(-20, h + 20), rng.uniform(1, 80)
        lo = int(rng.integers(-1, 7)); hi = int(rng.integers(lo, 8)) if lo >= 0 else -1
        assert np.array_equal(rf.features_in_area(x, y, r, lo, hi), of.features_in_area(x, y, r, lo, hi))


def test_descriptor_distance(po):
    rng = np.random.default_rng(3)
    z, o = np.zeros(32, np.uint8), np.full(32, 255, np.uint8)
    assert pyref.descriptor_distance(z, o) == 256 and pyref.descriptor_distance(z, z) == 0
    for _ in range(500):
        a, b = rng.integers(0, 256, (2, 32), dtype=np.uint8)
        assert pyref.descriptor_distance(a, b) == po.descriptor_distance(a, b) == int(np.unpackbits(a ^ b).sum())


# --------------------------------------------------------------------------------------------- ORBmatcher searches
def _pair(po, h, w, nf, seed):
    from orbslam_jpminipc_b200.synth import synth_frame, shifted_frame
    a = synth_frame(h, w, seed, quadrants=False)
    b = shifted_frame(a, 3, 2, seed + 1)
    orc = po.OracleExtractor(nf, 1.2, 8, 1, 20)
    return orc(a), orc(b)


def _scene(po, h, w, nf, seed):
    (ka, da), (kb, db) = _pair(po, h, w, nf, seed)
    rng = np.random.default_rng(seed)
    fx = fy = 500.0
    cx, cy = w / 2.0, h / 2.0
    z = rng.uniform(2, 10, len(ka)).astype(np.float32)
    xyz = np.stack([(ka["x"] - cx) / fx * z, (ka["y"] - cy) / fy * z, z], 1).astype(np.float32)
    T = np.eye(4, dtype=np.float32)
    T[:3, 3] = [0.03, 0.02, 0.01]
    c, s = np.cos(0.01), np.sin(0.01)
    T[:3, :3] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], np.float32)        # a small roll so Rcw is not the identity
    has = (rng.random(len(ka)) < 0.9).astype(np.uint8)
    outl = (rng.random(len(ka)) < 0.05).astype(np.uint8)
    cam = (w, h, fx, fy, cx, cy)
    return (ka, da), (kb, db), cam, has, outl, xyz, T


@pytest.mark.parametrize("shape,nf,th,ori", [((240, 320), 500, 15.0, True), ((480, 752), 1000, 15.0, True), ((376, 1241), 2000, 15.0, True),
                                             ((240, 320), 500, 7.0, False), ((240, 320), 500, 40.0, True)])
def test_search_by_projection_frame_frame(po, shape, nf, th, ori):
    """src/ORBmatcher.cc:1507-1620 (config 3's matcher), incl. the pose product Rcw*x3Dw+tcw and pre-claimed keypoints."""
    (ka, da), (kb, db), cam, has, outl, xyz, T = _scene(po, shape[0], shape[1], nf, 7000 + nf)
    pre = np.full(len(kb), -1, np.int32)
    free = np.nonzero(has)[0]
    pre[::13] = free[:len(pre[::13])]                      # keypoints that already hold one of the last frame's points
    rcur = pyref.RefFrame(kb, db, *cam).set_pose(T)
    rlast = pyref.RefFrame(ka, da, *cam).set_mappoints(has, xyz, outl)
    n, match = pyref.search_by_projection(rcur, rlast, th, 0.9, ori, pre.copy())
    ocur, olast = po.OracleFrame(kb, db, *cam), po.OracleFrame(ka, da, *cam)
    rn, rmatch = po.search_by_projection(ocur, olast, has, outl, xyz, T, th, ori, pre.copy())
    assert n > 20
    assert n == rn and np.array_equal(match, rmatch)


@pytest.mark.parametrize("shape,nf,th", [((240, 320), 500, 3.0), ((480, 752), 1000, 1.0), ((376, 1241), 2000, 5.0)])
def test_search_by_projection_mappoints(po, shape, nf, th):
    """src/ORBmatcher.cc:49-125"""
    (ka, da), (kb, db), cam, has, outl, xyz, T = _scene(po, shape[0], shape[1], nf, 8000 + nf)
    rng = np.random.default_rng(nf)
    n = len(ka)
    px = (ka["x"] + 3 + rng.normal(0, 1.0, n)).astype(np.float32)
    py = (ka["y"] + 2 + rng.normal(0, 1.0, n)).astype(np.float32)
    level = np.clip(ka["octave"] + rng.integers(0, 2, n), 0, 7).astype(np.int32)
    vcos = rng.uniform(0.99, 1.0, n).astype(np.float32)
    vcos[::4] = rng.uniform(0.9, 0.999, len(vcos[::4])).astype(np.float32)       # both branches of RadiusByViewingCos
    inv = (rng.random(n) < 0.85).astype(np.uint8)
    pre = np.full(len(kb), -1, np.int32)
    pre[::11] = 7
    nm, match = pyref.search_by_projection_mappoints(pyref.RefFrame(kb, db, *cam), inv, px, py, level, vcos, da, th, 0.8, pre.copy())
    rn, rmatch = po.search_by_projection_mappoints(po.OracleFrame(kb, db, *cam), inv, px, py, level, vcos, da, th, 0.8, match_f=pre.copy())
    assert nm > 20
    assert nm == rn and np.array_equal(match, rmatch)


@pytest.mark.parametrize("shape,nf,win,ori,minl", [((240, 320), 500, 20, True, -1), ((480, 752), 1000, 50, True, -1), ((376, 1241), 2000, 100, False, 2)])
def test_window_search(po, shape, nf, win, ori, minl):
    """src/ORBmatcher.cc:409-516"""
    (ka, da), (kb, db), cam, has, outl, xyz, T = _scene(po, shape[0], shape[1], nf, 8100 + nf)
    r1 = pyref.RefFrame(ka, da, *cam).set_mappoints(has, xyz)
    nm, match = pyref.window_search(r1, pyref.RefFrame(kb, db, *cam), win, 0.9, ori, min_level=minl)
    rn, rmatch = po.window_search(po.OracleFrame(ka, da, *cam), po.OracleFrame(kb, db, *cam), has, win, 0.9, ori, min_level=minl)
    assert nm > 20
    assert nm == rn and np.array_equal(match, rmatch)


@pytest.mark.parametrize("shape,nf,win", [((240, 320), 500, 15), ((480, 752), 1000, 15), ((376, 1241), 2000, 30)])
def test_search_by_projection_window(po, shape, nf, win):
    """src/ORBmatcher.cc:519-594"""
    (ka, da), (kb, db), cam, has, outl, xyz, T = _scene(po, shape[0], shape[1], nf, 8200 + nf)
    pre = np.full(len(kb), -1, np.int32)
    pre[::5] = 100000
    r1 = pyref.RefFrame(ka, da, *cam).set_mappoints(has, xyz)
    r2 = pyref.RefFrame(kb, db, *cam).set_pose(T)
    nm, match = pyref.search_by_projection_window(r1, r2, win, 0.9, pre.copy())
    rn, rmatch = po.search_by_projection_window(po.OracleFrame(ka, da, *cam), po.OracleFrame(kb, db, *cam), has, xyz, T, win, 0.9, pre.copy())
    assert nm > 10
    assert nm == rn and np.array_equal(match, rmatch)


@pytest.mark.parametrize("shape,nf,win,ori", [((240, 320), 500, 30, True), ((480, 752), 1000, 100, True), ((376, 1241), 2000, 100, False)])
def test_search_for_initialization(po, shape, nf, win, ori):
    """src/ORBmatcher.cc:598-713, two consecutive calls as the initialiser makes them (src/Tracking.cc:393-401)"""
    (ka, da), (kb, db), cam, has, outl, xyz, T = _scene(po, shape[0], shape[1], nf, 8400 + nf)
    r1, r2 = pyref.RefFrame(ka, da, *cam), pyref.RefFrame(kb, db, *cam)
    o1, o2 = po.OracleFrame(ka, da, *cam), po.OracleFrame(kb, db, *cam)
    prev = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
    nm, m12, pnew = pyref.search_for_initialization(r1, r2, prev, win, 0.9, ori)
    rn, rm12, rprev = po.search_for_initialization(o1, o2, prev, win, 0.9, ori)
    assert nm > 10
    assert nm == rn and np.array_equal(m12, rm12) and np.array_equal(pnew, rprev)
    nm2, m12b, pnew2 = pyref.search_for_initialization(r1, r2, pnew, win, 0.9, ori)
    rn2, rm12b, rprev2 = po.search_for_initialization(o1, o2, rprev, win, 0.9, ori)
    assert nm2 == rn2 and np.array_equal(m12b, rm12b) and np.array_equal(pnew2, rprev2)


def test_search_for_initialization_steal_chains(po, pkg):
    from test_gpu_match import test_search_for_initialization_steal_chains as _  # noqa: F401  (documented twin of the GPU case)
    rng = np.random.default_rng(2)
    n1, n2 = 1500, 120
    k2 = np.zeros(n2, pyref.KP_DTYPE)
    k2["x"] = rng.uniform(100, 200, n2).astype(np.float32); k2["y"] = rng.uniform(100, 200, n2).astype(np.float32)
    k2["angle"] = rng.uniform(0, 360, n2).astype(np.float32); k2["size"] = 31; k2["octave"] = rng.integers(0, 2, n2)
    d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    src = rng.integers(0, n2, n1)
    bits = np.unpackbits(d2[src], axis=1)
    for i in range(n1):
        bits[i, rng.choice(256, int(rng.integers(0, 60)), replace=False)] ^= 1
    d1 = np.packbits(bits, axis=1)
    k1 = np.zeros(n1, pyref.KP_DTYPE)
    k1["x"] = rng.uniform(100, 200, n1).astype(np.float32); k1["y"] = rng.uniform(100, 200, n1).astype(np.float32)
    k1["angle"] = (k2["angle"][src] + rng.choice([0.0] * 6 + [45.0, 90.0, 135.0, 180.0, 270.0], n1)).astype(np.float32) % np.float32(360)
    k1["size"] = 31; k1["octave"] = (rng.random(n1) < 0.1).astype(np.int32)
    cam = (320, 240, 300.0, 300.0, 160.0, 120.0)
    prev = np.stack([k1["x"], k1["y"]], 1)
    nm, m12, pnew = pyref.search_for_initialization(pyref.RefFrame(k1, d1, *cam), pyref.RefFrame(k2, d2, *cam), prev, 120, 0.9, True)
    rn, rm12, rprev = po.search_for_initialization(po.OracleFrame(k1, d1, *cam), po.OracleFrame(k2, d2, *cam), prev, 120, 0.9, True)
    assert nm > 5
    assert nm == rn and np.array_equal(m12, rm12) and np.array_equal(pnew, rprev)


def _bow(pkg, po, n1, n2, nnodes, seed, flip=0.06):
    from test_gpu_match import _bow_case
    return _bow_case(po, pkg, n1, n2, nnodes, seed=seed, flip=flip)


def _ref_kf(kps, desc, fv, valid, cam=(640, 480, 500.0, 500.0, 320.0, 240.0)):
    r = pyref.RefFrame(kps, desc, *cam).set_featvec(*fv)
    if valid is not None:
        r.set_mappoints(valid)
    return r


@pytest.mark.parametrize("n_kf,n_f,nnodes,ori", [(2000, 2000, 100, True), (500, 700, 10, True), (300, 200, 1, False), (64, 64, 40, True)])
def test_search_by_bow_keyframe_frame(po, pkg, n_kf, n_f, nnodes, ori):
    """src/ORBmatcher.cc:155-284"""
    fv1, d1, k1, valid, fv2, d2, k2 = _bow(pkg, po, n_kf, n_f, nnodes, n_kf + nnodes)
    n, match = pyref.search_by_bow(_ref_kf(k1, d1, fv1, valid), _ref_kf(k2, d2, fv2, None), 0.75, ori)
    rn, rmatch = po.search_by_bow(fv1, d1, k1, valid, fv2, d2, k2, 0.75, ori)
    assert n > 5
    assert n == rn and np.array_equal(match, rmatch)


@pytest.mark.parametrize("n1,n2,nnodes,ori", [(2000, 2000, 100, True), (600, 500, 12, True), (300, 300, 3, False)])
def test_search_by_bow_keyframe_keyframe(po, pkg, n1, n2, nnodes, ori):
    """src/ORBmatcher.cc:715-850"""
    fv1, d1, k1, v1, fv2, d2, k2 = _bow(pkg, po, n1, n2, nnodes, n1 + 7 * nnodes)
    v2 = (np.random.default_rng(n2).random(n2) < 0.85).astype(np.uint8)
    n, match = pyref.search_by_bow_kf(_ref_kf(k1, d1, fv1, v1), _ref_kf(k2, d2, fv2, v2), 0.75, ori)
    rn, rmatch = po.search_by_bow_kf(fv1, d1, k1, v1, fv2, d2, k2, v2, 0.75, ori)
    assert n > 5
    assert n == rn and np.array_equal(match, rmatch)


@pytest.mark.parametrize("n1,n2,nnodes,ori,seed", [(2000, 2000, 100, True, 1), (700, 500, 12, True, 2), (300, 300, 2, False, 3), (40, 60, 30, True, 4)])
def test_search_for_triangulation(po, pkg, n1, n2, nnodes, ori, seed):
    """src/ORBmatcher.cc:852-1014 with CheckDistEpipolarLine (:136-153)"""
    fv1, d1, k1, v1, fv2, d2, k2 = _bow(pkg, po, n1, n2, nnodes, seed + 40, flip=0.05)
    rng = np.random.default_rng(seed)
    k1["x"] = rng.uniform(20, 600, n1).astype(np.float32); k1["y"] = rng.uniform(20, 440, n1).astype(np.float32)
    k1["octave"] = rng.integers(0, 8, n1)
    twin = rng.integers(0, n1, n2)
    lam = rng.uniform(0.5, 3.0, n2)
    k2["x"] = (k1["x"][twin] + 6 * lam + rng.normal(0, 1.0, n2)).astype(np.float32)
    k2["y"] = (k1["y"][twin] + 4 * lam + rng.normal(0, 1.0, n2)).astype(np.float32)
    k2["octave"] = rng.integers(0, 8, n2)
    d2 = d1[twin] ^ np.packbits((rng.random((n2, 256)) < 0.04).astype(np.uint8), axis=1)
    t = np.array([6.0, 4.0, 0.0])
    F12 = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]], np.float32)
    sg = np.ones(8, np.float32)
    for i in range(1, 8):
        sg[i] = np.float32(np.float32(1.2) ** i) ** 2
    has1 = (rng.random(n1) < 0.3).astype(np.uint8); has2 = (rng.random(n2) < 0.3).astype(np.uint8)
    node1 = np.zeros(n1, np.int64)
    ids, start, items = fv1
    for j, nid in enumerate(ids):
        node1[items[start[j]:start[j + 1]]] = nid
    node2 = np.where(rng.random(n2) < 0.9, node1[twin], rng.choice(ids, n2))
    ids2 = np.unique(node2)
    st2, it2 = [0], []
    for nid in ids2:
        wv = np.nonzero(node2 == nid)[0]
        it2 += list(wv); st2.append(len(it2))
    fv2 = (ids2.astype(np.int32), np.array(st2, np.int32), np.array(it2, np.int32))
    rk1, rk2 = _ref_kf(k1, d1, fv1, has1), _ref_kf(k2, d2, fv2, has2)
    # the frames' own level sigmas must be the ones handed to the oracle
    n, m12, npairs = pyref.search_for_triangulation(rk1, rk2, F12, 0.6, ori)
    rn, rm12 = po.search_for_triangulation(fv1, d1, k1, has1, fv2, d2, k2, has2, F12, sg, ori)
    assert n > 3 and n == rn and npairs == n and np.array_equal(m12, rm12)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_distinctive_descriptor(po, seed):
    """src/MapPoint.cc:185-250.  The reference walks std::map<KeyFrame*, size_t>, i.e. its observations in POINTER order, so among
    observations tied on the least median it may pick any; the least median itself, and the choice when it is unique, are defined."""
    rng = np.random.default_rng(seed)
    for nobs in (1, 2, 3, 7, 20, 61):
        base = rng.integers(0, 256, 32, dtype=np.uint8)
        bits = np.unpackbits(np.repeat(base[None], nobs, 0), axis=1)
        bits ^= rng.random(bits.shape) < 0.12
        desc = np.packbits(bits, axis=1)
        got = pyref.distinctive_descriptor(desc)
        bi, bm = po.distinctive_descriptors(desc, np.array([0, nobs], np.int32))
        D = np.unpackbits(desc[:, None, :] ^ desc[None, :, :], axis=2).sum(2)
        med = np.sort(D, axis=1)[:, int(0.5 * (nobs - 1))]
        assert med.min() == bm[0] and med[bi[0]] == bm[0]
        cand = np.nonzero(med == med.min())[0]
        assert any(np.array_equal(got, desc[c]) for c in cand)
        if len(cand) == 1:
            assert np.array_equal(got, desc[bi[0]])


# --------------------------------------------------------------------------------------------- DBoW2 vocabulary
@pytest.mark.parametrize("k,L,levelsup,prune,order", [(10, 3, 1, 0.0, "bfs"), (5, 4, 2, 0.0, "bfs"), (4, 5, 4, 0.15, "dfs"), (3, 6, 4, 0.1, "bfs"), (10, 4, 0, 0.05, "dfs")])
def test_vocabulary_transform_and_score(po, tmp_path, k, L, levelsup, prune, order):
    """TemplatedVocabulary.h loadFromTextFile :1338-1425, transform :1127-1193 / :1218-1260, BowVector.cpp, FeatureVector.cpp,
    L1Scoring::score ScoringObject.cpp:22-64 — DBoW2's own code on a synthetic tree in its text format."""
    from orbslam_jpminipc_b200 import synth
    parent, desc, weight = synth.synth_vocabulary(k, L, seed=k * 10 + L, prune_frac=prune, order=order)
    path = str(tmp_path / "voc.txt")
    synth.write_vocabulary_text(path, k, L, parent, desc, weight, trailing_newline=False)
    rv = pyref.RefVocabulary(path)
    ov = po.OracleVocabulary(path=path)
    assert rv.nwords == ov.nwords
    rng = np.random.default_rng(L)
    leaves = np.nonzero(~np.isin(np.arange(len(parent)), parent))[0]
    bows = []
    for n in (0, 1, 37, 1000):
        src = rng.choice(leaves, n)
        feats = desc[src] ^ np.packbits((rng.random((n, 256)) < 0.05).astype(np.uint8), axis=1) if n else np.zeros((0, 32), np.uint8)
        (rw, rvv), (rn_, rs, ri) = rv.transform(feats, levelsup)
        (ow, ovv), (on_, os_, oi) = ov.transform(feats, levelsup)
        assert np.array_equal(rw, ow) and np.array_equal(rvv.view(np.uint64), ovv.view(np.uint64))     # doubles bit for bit
        assert np.array_equal(rn_, on_) and np.array_equal(rs, os_) and np.array_equal(ri, oi)
        if n:
            w, _, _ = ov.transform_features(feats[:50], 0)
            assert [rv.word(f) for f in feats[:50]] == list(w)
        bows.append((ow, ovv))
    for a in bows[1:]:
        for b in bows[1:]:
            s_ref, s_orc = rv.score(a, b), po.bow_score_l1(a, b)
            assert np.float64(s_ref).view(np.uint64) == np.float64(s_orc).view(np.uint64)


def test_relocalisation_candidate_scoring(po, tmp_path):
    """KeyFrameDatabase::add + DetectRelocalisationCandidates (src/KeyFrameDatabase.cc:198-308) on BowVectors made by DBoW2's own
    transform: shared-word counts, the 0.8*max gate and the L1 scores against the oracle's orc_bow_score_db."""
    from orbslam_jpminipc_b200 import synth
    k, L = 6, 4
    parent, desc, weight = synth.synth_vocabulary(k, L, seed=11)
    path = str(tmp_path / "voc.txt")
    synth.write_vocabulary_text(path, k, L, parent, desc, weight, trailing_newline=False)
    rv = pyref.RefVocabulary(path)
    rng = np.random.default_rng(4)
    leaves = np.nonzero(~np.isin(np.arange(len(parent)), parent))[0]
    scene = rng.choice(leaves, 400)                                    # the query sees these words

    def bow_of(words):
        feats = desc[words] ^ np.packbits((rng.random((len(words), 256)) < 0.03).astype(np.uint8), axis=1)
        return rv.transform(feats, 4)[0]
    qbow = bow_of(scene)
    kf_bows = []
    for i in range(40):                                                 # keyframes sharing 0 .. 100 % of the scene
        share = int(len(scene) * (i % 10) / 9.0)
        words = np.concatenate([rng.choice(scene, share), rng.choice(leaves, 300 - min(share, 299))]) if share else rng.choice(leaves, 300)
        kf_bows.append(bow_of(words))
    kp = np.zeros(1, pyref.KP_DTYPE); kp["x"] = 10; kp["y"] = 10
    d1 = np.zeros((1, 32), np.uint8)
    cam = (640, 480, 500.0, 500.0, 320.0, 240.0)
    kfs = [pyref.RefFrame(kp, d1, *cam).set_bowvec(*b) for b in kf_bows]
    q = pyref.RefFrame(kp, d1, *cam).set_bowvec(*qbow)
    common, score, cand = pyref.detect_relocalisation_candidates(rv, q, kfs)
    ocommon, oscore, omax = po.bow_score_db(qbow, kf_bows)
    assert omax > 50 and common.max() == omax
    assert np.array_equal(common, ocommon)
    assert np.array_equal(score.view(np.uint32), oscore.view(np.uint32))
    assert cand.any() and (score[cand] > 0).all()
    # no covisibility edges here, so the returned set is { score > 0.75 * best score } (:262-306)
    assert np.array_equal(cand, score > np.float32(0.75) * score.max())


def _retrieval_scene(tmp_path, seed, nkf=60):
    """a synthetic vocabulary, one query BowVector, nkf keyframe BowVectors sharing 0 .. 100 % of its words, and a covisibility graph
    with distinct weights (KeyFrame::UpdateBestCovisibles sorts (weight, pointer) pairs: equal weights would order by address)"""
    from orbslam_jpminipc_b200 import synth
    k, L = 6, 4
    parent, desc, weight = synth.synth_vocabulary(k, L, seed=11)
    path = str(tmp_path / "voc.txt")
    synth.write_vocabulary_text(path, k, L, parent, desc, weight, trailing_newline=False)
    rv = pyref.RefVocabulary(path)
    rng = np.random.default_rng(seed)
    leaves = np.nonzero(~np.isin(np.arange(len(parent)), parent))[0]
    scene = rng.choice(leaves, 400)

    def bow_of(words):
        feats = desc[words] ^ np.packbits((rng.random((len(words), 256)) < 0.03).astype(np.uint8), axis=1)
        return rv.transform(feats, 4)[0]
    qbow = bow_of(scene)
    kf_bows = []
    for i in range(nkf):
        share = int(len(scene) * rng.choice([0.0, 0.02, 0.3, 0.6, 0.8, 0.9, 0.95, 1.0]))
        words = np.concatenate([rng.choice(scene, share), rng.choice(leaves, 300 - min(share, 299))]) if share else rng.choice(leaves, 300)
        kf_bows.append(bow_of(words))
    edges, wgt = [], 1000
    for a in range(nkf):                                                # up to 14 neighbours: more than GetBestCovisibilityKeyFrames(10) returns
        for b in rng.choice(nkf, int(rng.integers(0, 8)), replace=False):
            if a != b and not any((e[0], e[1]) in ((a, b), (b, a)) for e in edges):
                edges.append((a, int(b), wgt)); wgt -= 1
    kp = np.zeros(1, pyref.KP_DTYPE); kp["x"] = 10; kp["y"] = 10
    d1 = np.zeros((1, 32), np.uint8)
    cam = (640, 480, 500.0, 500.0, 320.0, 240.0)
    kfs = [pyref.RefFrame(kp, d1, *cam).set_bowvec(*b) for b in kf_bows]
    q = pyref.RefFrame(kp, d1, *cam).set_bowvec(*qbow)
    return rv, q, qbow, kfs, kf_bows, edges, rng


@pytest.mark.parametrize("seed", [21, 22, 23])
def test_relocalisation_candidates_with_covisibility(po, tmp_path, seed):
    """DetectRelocalisationCandidates complete (src/KeyFrameDatabase.cc:198-308): covisibility accumulation over
    GetBestCovisibilityKeyFrames(10) including the STALE mRelocScore of neighbours that share a word without being scored by this
    query (:278-281), the 0.75 * best cut and the order / de-duplication of the returned list, against the oracle's flat-array form."""
    rv, q, qbow, kfs, kf_bows, edges, rng = _retrieval_scene(tmp_path, seed)
    stale = (rng.random(len(kfs)) * 0.3).astype(np.float32)                     # what earlier queries left in mRelocScore
    rs = stale.copy()
    rcand, rcommon, best10 = pyref.detect_candidates(rv, q, kfs, edges, rs)
    assert max(len(b) for b in best10) == 10 and len(rcand) >= 1
    os_ = stale.copy()
    ocand, ocommon = po.bow_detect_candidates(qbow, kf_bows, os_, covis=best10)
    assert np.array_equal(rcommon, ocommon)
    assert np.array_equal(rs.view(np.uint32), os_.view(np.uint32))
    assert list(rcand) == list(ocand)
    # the accumulation matters in this scene: without the graph the answer differs
    ncand, _ = po.bow_detect_candidates(qbow, kf_bows, stale.copy(), covis=None)
    assert list(ncand) != list(ocand) or seed != 21


@pytest.mark.parametrize("seed,min_score", [(31, 0.05), (32, 0.2), (33, 0.6)])
def test_loop_candidates(po, tmp_path, seed, min_score):
    """DetectLoopCandidates (src/KeyFrameDatabase.cc:75-196): keyframes connected to the query keyframe never enter the list, minScore
    gates the nominators and seeds the best accumulated score, neighbours need more than minCommonWords shared words."""
    rv, q, qbow, kfs, kf_bows, edges, rng = _retrieval_scene(tmp_path, seed)
    connected = np.zeros(len(kfs), np.uint8)
    connected[rng.choice(len(kfs), 12, replace=False)] = 1
    qedges = [(-1, int(k), 5000 + int(k)) for k in np.nonzero(connected)[0]]
    rs = np.zeros(len(kfs), np.float32)
    rcand, rcommon, best10 = pyref.detect_candidates(rv, q, kfs, edges + qedges, rs, loop=True, min_score=min_score)
    os_ = np.zeros(len(kfs), np.float32)
    ocand, ocommon = po.bow_detect_candidates(qbow, kf_bows, os_, covis=best10, excluded=connected, loop=True, min_score=min_score)
    assert np.array_equal(rcommon, ocommon)
    assert np.array_equal(rs.view(np.uint32), os_.view(np.uint32))
    assert list(rcand) == list(ocand)
    assert not connected[ocand].any()
    if min_score < 0.5:
        assert len(ocand) >= 1


@pytest.mark.parametrize("shape,nf,th,dist,ori", [((240, 320), 500, 10.0, 100, True), ((480, 752), 1000, 3.0, 64, True), ((376, 1241), 2000, 10.0, 100, False)])
def test_search_by_projection_keyframe(po, shape, nf, th, dist, ori):
    """src/ORBmatcher.cc:1622-1746 (relocalisation).  The level prediction (:1662-1669) is the caller's job at the C ABI, so the
    reference's own levels (from MapPoint::UpdateNormalAndDepth + GetMinDistanceInvariance) are handed to the oracle."""
    (ka, da), (kb, db), cam, has, outl, xyz, T = _scene(po, shape[0], shape[1], nf, 8300 + nf)
    rng = np.random.default_rng(nf + 1)
    found = ((rng.random(len(ka)) < 0.1) & (has > 0)).astype(np.uint8)
    pre = np.full(len(kb), -1, np.int32)
    pre[::9] = 4242
    rkf = pyref.RefFrame(ka, da, *cam).set_mappoints(has, xyz).update_points()        # keyframe at the world origin
    rcur = pyref.RefFrame(kb, db, *cam).set_pose(T)
    nm, match, pred = pyref.search_by_projection_kf(rcur, rkf, found, th, dist, 0.9, ori, pre.copy())
    assert (pred[has > 0] >= 0).all() and len(np.unique(pred[has > 0])) > 2
    active = ((has > 0) & (found == 0)).astype(np.uint8)
    rn, rmatch = po.search_by_projection_kf(po.OracleFrame(kb, db, *cam), active, xyz, T, np.maximum(pred, 0), da, ka["angle"], th, dist, ori,
                                            match_cur=pre.copy())
    assert nm > 10
    assert nm == rn and np.array_equal(match, rmatch)


@pytest.mark.skipif(not pyref.fma_available(), reason="oracle/_ref/libref_extractor_fma.so not built")
def test_extractor_fma_contracted_build(po):
    """The reference's flags (-O3 -march=native, CMakeLists.txt:12-13) let GCC contract the descriptor rotation x*b + y*a on an FMA host.
    oracle/_ref/libref_extractor_fma.so is src/ORBextractor.cc built that way; the oracle's desc_fma variant must equal it bit for bit,
    and the two variants differ in (only) a handful of descriptor bits."""
    from orbslam_jpminipc_b200.synth import synth_frame
    nrows = nbits = total = 0
    for (h, w, nf) in [(480, 640, 1000), (480, 752, 1000), (376, 1241, 2000)]:
        ref_fma = pyref.RefExtractor(nf, 1.2, 8, 1, 20, fma=True)
        orc_fma = po.OracleExtractor(nf, 1.2, 8, 1, 20, desc_fma=True)
        orc = po.OracleExtractor(nf, 1.2, 8, 1, 20)
        for seed in range(1000, 1008):
            img = synth_frame(h, w, seed)
            rk, rd = ref_fma(img)
            fk, fd = orc_fma(img)
            _same_keypoints(rk, fk)
            assert np.array_equal(rd, fd)
            pk, pd = orc(img)
            _same_keypoints(pk, fk)                          # keypoints and angles do not depend on the variant
            nrows += int((pd != fd).any(1).sum()); nbits += int(np.unpackbits(pd ^ fd).sum()); total += len(pk)
    assert total > 30000
    assert 0 < nbits <= 20 and nrows <= nbits, (nrows, nbits, total)


def _backend_scene(po, shape, nf, seed, scale=1.0):
    """A source keyframe at the world origin whose features carry map points, and a target keyframe that sees them from a slightly
    different pose (optionally through a similarity of scale `scale`)."""
    (ka, da), (kb, db), cam, has, outl, xyz, T = _scene(po, shape[0], shape[1], nf, seed)
    src = pyref.RefFrame(ka, da, *cam).set_mappoints(has, xyz).update_points()
    S = T.copy()
    S[:3, :] *= np.float32(scale)
    return (ka, da), (kb, db), cam, has, src, T, S


@pytest.mark.parametrize("shape,nf,th,scale", [((240, 320), 500, 10, 1.0), ((480, 752), 1000, 10, 1.3), ((376, 1241), 2000, 4, 0.8)])
def test_search_by_projection_sim3(po, shape, nf, th, scale):
    """src/ORBmatcher.cc:286-407 (loop closing).  The Sim3 decomposition, projection, depth / viewing-angle gates and level prediction are
    the caller's at the C ABI: the reference's own values (restated in the glue with the same member calls) feed the oracle, and the radius
    search, level filter, best-descriptor choice, TH_LOW test and vpMatched bookkeeping must agree with the reference's function."""
    (ka, da), (kb, db), cam, has, src, T, S = _backend_scene(po, shape, nf, 8600 + nf, scale)
    kf = pyref.RefFrame(kb, db, *cam).set_pose(T)
    pre = np.full(len(kb), -1, np.int32)
    pre[::7] = 555555
    n, matched, (active, u, v, level) = pyref.search_by_projection_sim3(kf, src, S, th, pre.copy())
    assert active.sum() > 50 and len(np.unique(level[active > 0])) > 2
    rn, rmatched = po.search_by_projection_sim3(po.OracleFrame(kb, db, *cam), active, u, v, level, da, th, pre.copy())
    assert n > 10
    assert n == rn and np.array_equal(matched, rmatched)


@pytest.mark.parametrize("shape,nf,th", [((240, 320), 500, 2.5), ((480, 752), 1000, 2.5), ((376, 1241), 2000, 7.5)])
def test_fuse(po, shape, nf, th):
    """src/ORBmatcher.cc:1016-1134 (local mapping), one candidate per call so that the keypoint every point was fused with can be read
    back from the reference's Replace / AddObservation bookkeeping; half of the target's keypoints already carry a map point."""
    (ka, da), (kb, db), cam, has, src, T, S = _backend_scene(po, shape, nf, 8700 + nf)
    rng = np.random.default_rng(nf)
    occupied = (rng.random(len(kb)) < 0.5).astype(np.uint8)
    kf = pyref.RefFrame(kb, db, *cam).set_pose(T).set_mappoints(occupied).update_points()
    n, fused, (active, u, v, level) = pyref.fuse(kf, src, th)
    of = po.OracleFrame(kb, db, *cam)
    sf = np.ones(8, np.float32)
    for i in range(1, 8):
        sf[i] = np.float32(sf[i - 1] * np.float32(1.2))
    radius = (np.float32(th) * sf[level]).astype(np.float32)
    bi, bd = po.window_best(of, active, u, v, radius, level, da)
    want = np.where((active > 0) & (bi >= 0) & (bd <= 50), bi, -1)
    assert n > 10 and n == int((want >= 0).sum())
    # the keypoint can be read back from the reference's bookkeeping unless an earlier candidate already went to the same keypoint
    # (MapPoint::Replace then only erases, src/MapPoint.cc:118-150); those few are covered by the count above
    first = np.ones(len(want), bool)
    seen = set()
    for i, w in enumerate(want):
        if w >= 0:
            first[i] = w not in seen
            seen.add(w)
    assert first.sum() > 0.95 * len(want)
    assert np.array_equal(fused[first], want[first])
    assert occupied[fused[fused >= 0]].any() and not occupied[fused[fused >= 0]].all()       # both the Replace and the AddObservation branch ran


@pytest.mark.parametrize("shape,nf,th,s12", [((240, 320), 500, 7.5, 1.0), ((480, 752), 1000, 7.5, 1.1), ((376, 1241), 2000, 5.0, 0.9)])
def test_search_by_sim3(po, shape, nf, th, s12):
    """src/ORBmatcher.cc:1267-1505 (loop closing): both projection directions, TH_HIGH, and the mutual-agreement test, against
    window_best of the oracle on the reference's own projections."""
    (ka, da), (kb, db), cam, has, outl, xyz, T = _scene(po, shape[0], shape[1], nf, 8800 + nf)
    rng = np.random.default_rng(nf + 3)
    # KF1 at the world origin with its points; KF2 at pose T holding points for ITS features: the same 3-D structure seen from T
    k1 = pyref.RefFrame(ka, da, *cam).set_mappoints(has, xyz).update_points()
    fx, fy, cx, cy = cam[2:]
    z2 = rng.uniform(2, 10, len(kb)).astype(np.float32)
    pc2 = np.stack([(kb["x"] - cx) / fx * z2, (kb["y"] - cy) / fy * z2, z2], 1).astype(np.float64)
    R, t = T[:3, :3].astype(np.float64), T[:3, 3].astype(np.float64)
    xyz2 = ((pc2 - t) @ R).astype(np.float32)                         # world = R^T (pc - t)
    has2 = (rng.random(len(kb)) < 0.9).astype(np.uint8)
    k2 = pyref.RefFrame(kb, db, *cam).set_pose(T).set_mappoints(has2, xyz2).update_points()
    # Sim3 camera1 <- camera2: x1 = s12 R12 x2 + t12 with R12 = R^T, t12 = -s12 R^T t (poses: camera1 = world, camera2 = T)
    R12 = R.T.astype(np.float32)
    t12 = (-s12 * (R.T @ t)).astype(np.float32)
    pre = np.full(len(ka), -1, np.int32)
    n, m12, (a12, u12, v12, l12), (a21, u21, v21, l21) = pyref.search_by_sim3(k1, k2, s12, R12, t12, th, pre.copy())
    assert a12.sum() > 50 and a21.sum() > 50
    o1, o2 = po.OracleFrame(ka, da, *cam), po.OracleFrame(kb, db, *cam)
    sf = np.ones(8, np.float32)
    for i in range(1, 8):
        sf[i] = np.float32(sf[i - 1] * np.float32(1.2))
    b1, d1 = po.window_best(o2, a12, u12, v12, (np.float32(th) * sf[l12]).astype(np.float32), l12, da)
    b2, d2 = po.window_best(o1, a21, u21, v21, (np.float32(th) * sf[l21]).astype(np.float32), l21, db)
    m1 = np.where((a12 > 0) & (b1 >= 0) & (d1 <= 100), b1, -1)
    m2 = np.where((a21 > 0) & (b2 >= 0) & (d2 <= 100), b2, -1)
    want = np.full(len(ka), -1, np.int32)
    for i1 in np.flatnonzero(m1 >= 0):
        if m2[m1[i1]] == i1:
            want[i1] = m1[i1]
    assert n > 10
    assert n == int((want >= 0).sum()) and np.array_equal(m12, want)


@pytest.mark.parametrize("shape,nf,th,scale", [((240, 320), 500, 2.5, 1.0), ((480, 752), 1000, 4.0, 1.25)])
def test_fuse_sim3(po, shape, nf, th, scale):
    """src/ORBmatcher.cc:1136-1265 (loop correction): Fuse through a similarity; here every candidate's keypoint can be read back."""
    (ka, da), (kb, db), cam, has, src, T, S = _backend_scene(po, shape, nf, 8900 + nf, scale)
    rng = np.random.default_rng(nf)
    occupied = (rng.random(len(kb)) < 0.5).astype(np.uint8)
    kf = pyref.RefFrame(kb, db, *cam).set_pose(T).set_mappoints(occupied).update_points()
    n, fused, (active, u, v, level) = pyref.fuse_sim3(kf, src, S, th)
    sf = np.ones(8, np.float32)
    for i in range(1, 8):
        sf[i] = np.float32(sf[i - 1] * np.float32(1.2))
    bi, bd = po.window_best(po.OracleFrame(kb, db, *cam), active, u, v, (np.float32(th) * sf[level]).astype(np.float32), level, da)
    want = np.where((active > 0) & (bi >= 0) & (bd <= 50), bi, -1)
    assert n > 10 and n == int((want >= 0).sum())
    assert np.array_equal(fused, want)
