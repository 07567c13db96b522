"""GPU parity tests for the matcher kernels through the C ABI, against the CPU oracle."""
import numpy as np
import pytest

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
def matcher(pkg):
    return pkg.ORBmatcher(0.6, True)


def test_descriptor_distance_kats(pkg, po):
    z, o = np.zeros(32, np.uint8), np.full(32, 255, np.uint8)
    assert pkg.ORBmatcher.DescriptorDistance(z, z) == 0
    assert pkg.ORBmatcher.DescriptorDistance(z, o) == 256
    for byte in (0, 5, 31):
        for bit in (0, 7):
            a = z.copy(); a[byte] = 1 << bit
            assert pkg.ORBmatcher.DescriptorDistance(a, z) == 1
    rng = np.random.default_rng(3)
    for _ in range(200):
        a, b = rng.integers(0, 256, (2, 32), dtype=np.uint8)
        assert pkg.ORBmatcher.DescriptorDistance(a, b) == po.descriptor_distance(a, b)


@pytest.mark.parametrize("nq,ndb", [(2000, 2000), (300, 5000), (1, 1), (257, 255), (2000, 70001), (5, 3), (1000, 0)])
def test_knn2_bit_exact(matcher, po, nq, ndb):
    from orbslam_jpminipc_b200.synth import synth_descriptors
    db, q = synth_descriptors(ndb, nq, seed_db=42 + ndb, seed_q=43 + nq)
    i1, d1, d2 = matcher.knn2(q, db)
    r1, rd1, rd2 = po.knn2(q, db)
    assert np.array_equal(i1, r1) and np.array_equal(d1, rd1) and np.array_equal(d2, rd2)


def test_knn2_ties_lowest_index_wins(matcher, po):
    rng = np.random.default_rng(11)
    db = rng.integers(0, 256, (4096, 32), dtype=np.uint8)
    db[3000] = db[17]; db[1000] = db[17]; db[4095] = db[0]
    q = np.stack([db[17], db[0], np.zeros(32, np.uint8)])
    i1, d1, d2 = matcher.knn2(q, db)
    assert list(i1[:2]) == [17, 0] and list(d1[:2]) == [0, 0] and list(d2[:2]) == [0, 0]
    r = po.knn2(q, db)
    assert np.array_equal(i1, r[0]) and np.array_equal(d1, r[1]) and np.array_equal(d2, r[2])


def test_match_ratio(matcher, po):
    from orbslam_jpminipc_b200.synth import synth_descriptors
    db, q = synth_descriptors(3000, 1000)
    i1, d1, d2 = matcher.knn2(q, db)
    m, n = matcher.match_ratio(i1, d1, d2)
    rm, rn = po.match_ratio(i1, d1, d2, 0.6, 50)
    assert n == rn and np.array_equal(m, rm) and 300 < n < 600


def test_frame_grid(pkg, po, matcher):
    rng = np.random.default_rng(5)
    n = 2000
    kps = np.zeros(n, pkg.KP_DTYPE)
    kps["x"] = rng.uniform(0, 752, n).astype(np.float32); kps["y"] = rng.uniform(0, 480, n).astype(np.float32)
    kps["x"][:5] = [751.9, 0.0, 746.2, 5.8, 740.1]           # round() drops x >= (63.5/64)*W
    kps["octave"] = rng.integers(0, 8, n)
    f = pkg.Frame(matcher, kps, np.zeros((n, 32), np.uint8), 752, 480, 500, 500, 376, 240)
    o = po.OracleFrame(kps, np.zeros((n, 32), np.uint8), 752, 480, 500, 500, 376, 240)
    assert np.array_equal(f.cell_start, o.cell_start)
    cnt = f.cell_start[-1]
    assert cnt < n and np.array_equal(f.cell_items[:cnt], o.cell_items[:cnt])
    # crowded cells (more than 32 keypoints in one cell: the one-warp placement of k_grid_build), every keypoint in ONE cell, and
    # cells of exactly 32 / 33 keypoints on either side of the switch
    for kind in ("clustered", "one_cell", "edge32", "edge33"):
        k2 = kps.copy()
        if kind == "clustered":
            k2["x"][:700] = rng.uniform(100, 130, 700).astype(np.float32); k2["y"][:700] = rng.uniform(200, 225, 700).astype(np.float32)
        elif kind == "one_cell":
            k2["x"] = rng.uniform(300, 305, n).astype(np.float32); k2["y"] = rng.uniform(100, 104, n).astype(np.float32)
        else:
            m = 32 if kind == "edge32" else 33
            idx = rng.permutation(n)[:m]
            k2["x"][idx] = np.float32(400.0); k2["y"][idx] = np.float32(300.0)
            far = np.setdiff1d(np.arange(n), idx)
            near = far[(np.abs(k2["x"][far] - 400.0) < 30) & (np.abs(k2["y"][far] - 300.0) < 30)]
            k2["x"][near] = np.float32(20.0)                       # keep the neighbourhood of that cell empty otherwise
        f2 = pkg.Frame(matcher, k2, np.zeros((n, 32), np.uint8), 752, 480, 500, 500, 376, 240)
        o2 = po.OracleFrame(k2, np.zeros((n, 32), np.uint8), 752, 480, 500, 500, 376, 240)
        c2 = o2.cell_start[-1]
        assert np.array_equal(f2.cell_start, o2.cell_start), kind
        assert np.array_equal(f2.cell_items[:c2], o2.cell_items[:c2]), kind


# ---------------------------------------------------------------- SearchByProjection / SearchByBoW
def _frame_pair(po, h=240, w=320, nf=500, seed=7000):
    from orbslam_jpminipc_b200.synth import synth_frame, shifted_frame
    a = synth_frame(h, w, seed, quadrants=False)
    b = shifted_frame(a, 3, 2, seed + 1)
    orc = po.OracleExtractor(nf, 1.2, 8, 1, 20)
    ka, da = orc(a)
    kb, db = orc(b)
    return (ka, da), (kb, db)


def _scene(po, pkg, matcher, h=240, w=320, nf=500, seed=7000, th=15.0):
    (ka, da), (kb, db) = _frame_pair(po, h, w, nf, seed)
    rng = np.random.default_rng(seed)
    fx = fy = 500.0
    cx, cy = w / 2.0, h / 2.0
    z = rng.uniform(2, 10, len(ka)).astype(np.float32)
    xyz = np.stack([(ka["x"] - cx) / fx * z, (ka["y"] - cy) / fy * z, z], 1).astype(np.float32)
    T = np.eye(4, dtype=np.float32)
    T[:3, 3] = [0.03, 0.02, 0.01]
    has = (rng.random(len(ka)) < 0.9).astype(np.uint8)
    outl = (rng.random(len(ka)) < 0.05).astype(np.uint8)
    args = dict(w=w, h=h, fx=fx, fy=fy, cx=cx, cy=cy)
    gcur = pkg.Frame(matcher, kb, db, w, h, fx, fy, cx, cy)
    glast = pkg.Frame(matcher, ka, da, w, h, fx, fy, cx, cy)
    ocur = po.OracleFrame(kb, db, w, h, fx, fy, cx, cy)
    olast = po.OracleFrame(ka, da, w, h, fx, fy, cx, cy)
    return gcur, glast, ocur, olast, has, outl, xyz, T


@pytest.mark.parametrize("shape,nf,th,ori", [((240, 320), 500, 15.0, True), ((480, 752), 1000, 15.0, True),
                                             ((376, 1241), 2000, 15.0, True), ((240, 320), 500, 7.0, False),
                                             ((240, 320), 500, 40.0, True)])
def test_search_by_projection_vs_oracle(pkg, po, shape, nf, th, ori):
    m = pkg.ORBmatcher(0.9, ori)
    gcur, glast, ocur, olast, has, outl, xyz, T = _scene(po, pkg, m, shape[0], shape[1], nf, 7000 + nf, th)
    n, match = m.SearchByProjection(gcur, glast, th, has, outl, xyz, T)
    rn, rmatch = po.search_by_projection(ocur, olast, has, outl, xyz, T, th, ori)
    assert rn > 20, rn
    assert n == rn and np.array_equal(match, rmatch)


def test_search_by_projection_preclaimed_and_empty(pkg, po):
    m = pkg.ORBmatcher(0.9, True)
    gcur, glast, ocur, olast, has, outl, xyz, T = _scene(po, pkg, m)
    pre = np.full(gcur.N, -1, np.int32)
    pre[::3] = 12345                                   # keypoints that already carry a map point are skipped (:1562)
    n, match = m.SearchByProjection(gcur, glast, 15.0, has, outl, xyz, T, match_cur=pre.copy())
    rn, rmatch = po.search_by_projection(ocur, olast, has, outl, xyz, T, 15.0, True, match_cur=pre.copy())
    assert n == rn and np.array_equal(match, rmatch) and np.all(match[::3] == 12345)
    none = np.zeros_like(has)
    n, match = m.SearchByProjection(gcur, glast, 15.0, none, outl, xyz, T)
    assert n == 0 and np.all(match == -1)
    T2 = T.copy(); T2[:3, 3] = [50, 0, 0]             # everything projects outside the image
    n, match = m.SearchByProjection(gcur, glast, 15.0, has, outl, xyz, T2)
    rn, _ = po.search_by_projection(ocur, olast, has, outl, xyz, T2, 15.0, True)
    assert n == rn == 0


def test_search_by_projection_tie_break_is_scan_order(pkg, po):
    """Two current keypoints with identical descriptors at equal distance: the one met first in the
    GetFeaturesInArea scan (ix outer, iy inner; src/Frame.cc:233-259) wins, not the lower index."""
    m = pkg.ORBmatcher(0.9, False)
    w, h = 640, 480
    rng = np.random.default_rng(1)
    d = rng.integers(0, 256, (1, 32), dtype=np.uint8)
    kb = np.zeros(2, pkg.KP_DTYPE)
    kb["x"] = [330.0, 310.0]; kb["y"] = [240.0, 240.0]; kb["octave"] = 0; kb["angle"] = 10
    db = np.repeat(d, 2, 0)
    ka = np.zeros(1, pkg.KP_DTYPE); ka["x"] = 320; ka["y"] = 240; ka["angle"] = 10
    xyz = np.array([[0, 0, 5.0]], np.float32)
    T = np.eye(4, dtype=np.float32)
    g = (pkg.Frame(m, kb, db, w, h, 500, 500, 320, 240), pkg.Frame(m, ka, d, w, h, 500, 500, 320, 240))
    o = (po.OracleFrame(kb, db, w, h, 500, 500, 320, 240), po.OracleFrame(ka, d, w, h, 500, 500, 320, 240))
    one = np.ones(1, np.uint8); zero = np.zeros(1, np.uint8)
    n, match = m.SearchByProjection(g[0], g[1], 15.0, one, zero, xyz, T)
    rn, rmatch = po.search_by_projection(o[0], o[1], one, zero, xyz, T, 15.0, False)
    assert n == rn == 1 and np.array_equal(match, rmatch) and match[1] == 0 and match[0] == -1


def _bow_case(po, pkg, n_kf=2000, n_f=2000, nnodes=100, seed=9, flip=0.06):
    rng = np.random.default_rng(seed)
    kf_desc = rng.integers(0, 256, (n_kf, 32), dtype=np.uint8)
    twin = rng.permutation(n_kf)[:n_f] if n_f <= n_kf else rng.integers(0, n_kf, n_f)
    flips = np.packbits((rng.random((n_f, 256)) < flip).astype(np.uint8), axis=1)
    f_desc = kf_desc[twin] ^ flips
    f_desc[::7] = rng.integers(0, 256, (len(f_desc[::7]), 32), dtype=np.uint8)
    kf_node = rng.integers(0, nnodes, n_kf)
    f_node = np.where(rng.random(n_f) < 0.9, kf_node[twin], rng.integers(0, nnodes, n_f))
    f_node = f_node * 3 + 5                              # sparse node ids, some present on one side only
    kf_node = kf_node * 3 + 5
    kf_node[kf_node == 5 + 3 * 7] = 4                    # a node id that only the KF has
    def csr(node):
        ids = np.unique(node)
        start = [0]; items = []
        for v in ids:
            it = np.nonzero(node == v)[0]
            items += list(it); start.append(len(items))
        return ids.astype(np.int32), np.array(start, np.int32), np.array(items, np.int32)
    kf_kps = np.zeros(n_kf, pkg.KP_DTYPE); kf_kps["angle"] = rng.uniform(0, 360, n_kf).astype(np.float32)
    f_kps = np.zeros(n_f, pkg.KP_DTYPE)
    f_kps["angle"] = np.where(rng.random(n_f) < 0.8, kf_kps["angle"][twin] - 20 + rng.normal(0, 6, n_f),
                              rng.uniform(0, 360, n_f)).astype(np.float32) % 360
    valid = (rng.random(n_kf) < 0.8).astype(np.uint8)
    return csr(kf_node), kf_desc, kf_kps, valid, csr(f_node), f_desc, f_kps


@pytest.mark.parametrize("n_kf,n_f,nnodes,ori", [(2000, 2000, 100, True), (500, 700, 10, True), (300, 200, 1, False), (64, 64, 40, True)])
def test_search_by_bow_vs_oracle(pkg, po, n_kf, n_f, nnodes, ori):
    m = pkg.ORBmatcher(0.75, ori)
    case = _bow_case(po, pkg, n_kf, n_f, nnodes, seed=n_kf + nnodes)
    n, match = m.SearchByBoW(*case)
    rn, rmatch = po.search_by_bow(*case, 0.75, ori)
    assert rn > 5
    assert n == rn and np.array_equal(match, rmatch)


def test_search_by_bow_overlapping_nodes_serial_path(pkg, po):
    """A frame feature listed under two nodes couples the nodes through the claim array: must still equal the sequential scan."""
    m = pkg.ORBmatcher(0.75, True)
    kfv, kd, kk, valid, ffv, fd, fk = _bow_case(po, pkg, 400, 400, 8, seed=3)
    ids, start, items = ffv
    items2 = np.concatenate([items, items[:50]])            # last node additionally lists the first 50 items
    start2 = start.copy(); start2[-1] = len(items2)
    ffv2 = (ids, start2, items2.astype(np.int32))
    n, match = m.SearchByBoW(kfv, kd, kk, valid, ffv2, fd, fk)
    rn, rmatch = po.search_by_bow(kfv, kd, kk, valid, ffv2, fd, fk, 0.75, True)
    assert n == rn and np.array_equal(match, rmatch)


def test_three_maxima_rule(po):
    assert po.three_maxima([0] * 30) == (-1, -1, -1)
    h = [0] * 30; h[3] = 100; h[5] = 9; h[7] = 50
    assert po.three_maxima(h) == (3, 7, -1)                 # max3 < 0.1*max1 dropped
    h[5] = 10
    assert po.three_maxima(h) == (3, 7, 5)


def test_sharded_merge_kernel_equals_single_scan(pkg, po, matcher):
    """k_knn2_merge over emulated shards (one GPU, several row ranges) == one scan over the whole DB."""
    import ctypes as C
    import torch
    from orbslam_jpminipc_b200._lib import check, lib, ptr
    from orbslam_jpminipc_b200.sharding import shard_range
    from orbslam_jpminipc_b200.synth import synth_descriptors
    db, q = synth_descriptors(30011, 500, dup_frac=0.02)
    dev = torch.device("cuda", 0)
    d_db, d_q = torch.from_numpy(db).to(dev), torch.from_numpy(q).to(dev)
    nq, world = len(q), 5
    allp = torch.zeros((world, 3, nq), dtype=torch.int32, device=dev)
    L = lib()
    st = torch.cuda.current_stream().cuda_stream
    for r in range(world):
        lo, hi = shard_range(len(db), r, world)
        p = allp[r]
        check(L.orb_hamming_knn2_device(matcher._h, ptr(d_q), nq, C.c_void_p(d_db.data_ptr() + lo * 32), hi - lo, 1, lo,
                                        C.c_void_p(p.data_ptr()), C.c_void_p(p.data_ptr() + 4 * nq), C.c_void_p(p.data_ptr() + 8 * nq),
                                        C.c_void_p(st)), "knn2 shard")
    out = torch.zeros((3, nq), dtype=torch.int32, device=dev)
    check(L.orb_knn2_merge_device(matcher._h, ptr(allp), world, nq, C.c_void_p(out.data_ptr()), C.c_void_p(out.data_ptr() + 4 * nq),
                                  C.c_void_p(out.data_ptr() + 8 * nq), C.c_void_p(st)), "merge")
    torch.cuda.synchronize()
    ref = po.knn2(q, db)
    got = out.cpu().numpy()
    assert all(np.array_equal(got[k], ref[k]) for k in range(3))
    assert all(np.array_equal(a, b) for a, b in zip(po.merge_best2(allp.cpu().numpy()), ref))


def test_cpp_shim_program(pkg, po, tmp_path):
    """The header-only C++ shims (include/ORBextractor.h, include/ORBmatcher.h) through a compiled host program."""
    import os
    import subprocess
    from orbslam_jpminipc_b200.synth import synth_frame
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "test_shim")
    libdir = os.path.join(root, "orbslam_jpminipc_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(root, "tests", "cpp", "test_shim.cpp"),
                           "-L" + libdir, "-lorb_b200", "-Wl,-rpath," + libdir])
    img = synth_frame(240, 320, 8100)
    raw, out = str(tmp_path / "f.raw"), str(tmp_path / "o.bin")
    img.tofile(raw)
    from orbslam_jpminipc_b200 import synth
    parent, vdesc, weight = synth.synth_vocabulary(6, 3, seed=4, stop_frac=0.03)
    voc_path = str(tmp_path / "voc.txt")
    synth.write_vocabulary_text(voc_path, 6, 3, parent, vdesc, weight)
    subprocess.check_call([exe, raw, "320", "240", "300", out, voc_path])
    buf = open(out, "rb").read()
    n, nmatch = np.frombuffer(buf[:8], np.int32)
    kps = np.frombuffer(buf[8:8 + 28 * n], pkg.KP_DTYPE)
    desc = np.frombuffer(buf[8 + 28 * n:8 + 60 * n], np.uint8).reshape(n, 32)
    match = np.frombuffer(buf[8 + 60 * n:8 + 64 * n], np.int32)
    # vocabulary part: BowVector / FeatureVector as the std::map based shim returned them
    off = 8 + 64 * n
    nb, nfv = np.frombuffer(buf[off:off + 8], np.int32)
    self_score = np.frombuffer(buf[off + 8:off + 12], np.float32)[0]
    off += 12
    bow = np.frombuffer(buf[off:off + 12 * nb], np.dtype([("w", "<u4"), ("v", "<f8")]))
    off += 12 * nb
    (rbw, rbv), (rfn, rfs, rfi) = po.OracleVocabulary(path=voc_path).transform(desc, 1)
    assert nb == len(rbw) and np.array_equal(bow["w"], rbw) and np.array_equal(bow["v"], rbv) and self_score == 1.0
    assert nfv == len(rfn)
    for j in range(nfv):
        node, m = np.frombuffer(buf[off:off + 8], np.int32)
        items = np.frombuffer(buf[off + 8:off + 8 + 4 * m], np.int32)
        off += 8 + 4 * m
        assert node == rfn[j] and np.array_equal(items, rfi[rfs[j]:rfs[j + 1]])
    rk, rd = po.OracleExtractor(300)(img)
    assert n == len(rk) and np.array_equal(kps.view(np.uint8), rk.view(np.uint8)) and np.array_equal(desc, rd)
    i1, d1, d2 = po.knn2(rd, rd)
    rm, rn = po.match_ratio(i1, d1, d2, np.float32(0.9), 50)
    assert nmatch == rn and np.array_equal(match, rm)


# ---------------------------------------------------------------- further windowed searches
@pytest.mark.parametrize("shape,nf,th", [((240, 320), 500, 3.0), ((480, 752), 1000, 1.0), ((376, 1241), 2000, 5.0)])
def test_search_by_projection_mappoints_vs_oracle(pkg, po, shape, nf, th):
    """ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>, th), src/ORBmatcher.cc:49-125 (same-level ratio rule)."""
    m = pkg.ORBmatcher(0.8, True)
    gcur, glast, ocur, olast, has, outl, xyz, T = _scene(po, pkg, m, shape[0], shape[1], nf, 8000 + nf, 15.0)
    rng = np.random.default_rng(nf)
    n = glast.N                                           # "map points" = last-frame features seen slightly displaced
    proj_x = (glast.kps["x"] + 3 + rng.normal(0, 1.0, n)).astype(np.float32)
    proj_y = (glast.kps["y"] + 2 + rng.normal(0, 1.0, n)).astype(np.float32)
    level = np.clip(glast.kps["octave"] + rng.integers(0, 2, n), 0, 7).astype(np.int32)
    view_cos = rng.uniform(0.99, 1.0, n).astype(np.float32)
    in_view = (rng.random(n) < 0.85).astype(np.uint8)
    pre = np.full(gcur.N, -1, np.int32)
    pre[::11] = 7                                         # keypoints that already carry a map point
    nm, match = m.SearchByProjectionMapPoints(gcur, in_view, proj_x, proj_y, level, view_cos, glast.desc, th, match_f=pre.copy())
    rn, rmatch = po.search_by_projection_mappoints(ocur, in_view, proj_x, proj_y, level, view_cos, glast.desc, th, 0.8, match_f=pre.copy())
    assert rn > 20
    assert nm == rn and np.array_equal(match, rmatch)


@pytest.mark.parametrize("shape,nf,win,ori,minl", [((240, 320), 500, 20, True, -1), ((480, 752), 1000, 50, True, -1),
                                                    ((376, 1241), 2000, 100, False, 2)])
def test_window_search_vs_oracle(pkg, po, shape, nf, win, ori, minl):
    """ORBmatcher::WindowSearch, src/ORBmatcher.cc:409-516."""
    m = pkg.ORBmatcher(0.9, ori)
    gcur, glast, ocur, olast, has, outl, xyz, T = _scene(po, pkg, m, shape[0], shape[1], nf, 8100 + nf, 15.0)
    nm, match = m.WindowSearch(glast, gcur, win, has, minScaleLevel=minl)
    rn, rmatch = po.window_search(olast, ocur, has, win, 0.9, ori, min_level=minl)
    assert rn > 20
    assert nm == rn and np.array_equal(match, rmatch)


@pytest.mark.parametrize("shape,nf,win", [((240, 320), 500, 15), ((480, 752), 1000, 15), ((376, 1241), 2000, 30)])
def test_search_by_projection_window_vs_oracle(pkg, po, shape, nf, win):
    """ORBmatcher::SearchByProjection(F1, F2, windowSize, ...), src/ORBmatcher.cc:519-594 (no bounds test, pre-filled matches)."""
    m = pkg.ORBmatcher(0.9, True)
    gcur, glast, ocur, olast, has, outl, xyz, T = _scene(po, pkg, m, shape[0], shape[1], nf, 8200 + nf, 15.0)
    pre = np.full(gcur.N, -1, np.int32)
    pre[::5] = 100000                                     # F2.mvpMapPoints already set
    nm, match = m.SearchByProjectionWindow(glast, gcur, win, has, xyz, T, pre.copy())
    rn, rmatch = po.search_by_projection_window(olast, ocur, has, xyz, T, win, 0.9, pre.copy())
    assert rn > 10
    assert nm == rn and np.array_equal(match, rmatch)


@pytest.mark.parametrize("n1,n2,nnodes,ori", [(2000, 2000, 100, True), (600, 500, 12, True), (300, 300, 3, False)])
def test_search_by_bow_keyframes_vs_oracle(pkg, po, n1, n2, nnodes, ori):
    """ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, ...), src/ORBmatcher.cc:715-850."""
    m = pkg.ORBmatcher(0.75, ori)
    fv1, d1, k1, v1, fv2, d2, k2 = _bow_case(po, pkg, n1, n2, nnodes, seed=n1 + 7 * nnodes)
    v2 = (np.random.default_rng(n2).random(n2) < 0.85).astype(np.uint8)
    n, match = m.SearchByBoWKeyFrames(fv1, d1, k1, v1, fv2, d2, k2, v2)
    rn, rmatch = po.search_by_bow_kf(fv1, d1, k1, v1, fv2, d2, k2, v2, 0.75, ori)
    assert rn > 5
    assert n == rn and np.array_equal(match, rmatch)


@pytest.mark.parametrize("shape,nf,th,dist,ori", [((240, 320), 500, 10.0, 100, True), ((480, 752), 1000, 3.0, 64, True),
                                                   ((376, 1241), 2000, 10.0, 100, False)])
def test_search_by_projection_keyframe_vs_oracle(pkg, po, shape, nf, th, dist, ori):
    """ORBmatcher::SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist), src/ORBmatcher.cc:1622-1746."""
    m = pkg.ORBmatcher(0.9, ori)
    gcur, glast, ocur, olast, has, outl, xyz, T = _scene(po, pkg, m, shape[0], shape[1], nf, 8300 + nf, 15.0)
    rng = np.random.default_rng(nf + 1)
    pred = np.clip(glast.kps["octave"] + rng.integers(-1, 2, glast.N), 0, 7).astype(np.int32)
    pre = np.full(gcur.N, -1, np.int32)
    pre[::9] = 4242
    nm, match = m.SearchByProjectionKeyFrame(gcur, has, xyz, T, pred, glast.desc, glast.kps["angle"], th, dist, match_cur=pre.copy())
    rn, rmatch = po.search_by_projection_kf(ocur, has, xyz, T, pred, glast.desc, glast.kps["angle"], th, dist, ori, match_cur=pre.copy())
    assert rn > 10
    assert nm == rn and np.array_equal(match, rmatch)


@pytest.mark.parametrize("shape,nf,win,ori", [((240, 320), 500, 30, True), ((480, 752), 1000, 100, True), ((376, 1241), 2000, 100, False)])
def test_search_for_initialization_vs_oracle(pkg, po, shape, nf, win, ori):
    """ORBmatcher::SearchForInitialization, src/ORBmatcher.cc:598-713 (call site src/Tracking.cc:393-401: vbPrevMatched
    starts as the first frame's own keypoint positions)."""
    m = pkg.ORBmatcher(0.9, ori)
    gcur, glast, ocur, olast, has, outl, xyz, T = _scene(po, pkg, m, shape[0], shape[1], nf, 8400 + nf, 15.0)
    prev = np.stack([glast.kps["x"], glast.kps["y"]], 1).astype(np.float32)
    nm, m12, pnew = m.SearchForInitialization(glast, gcur, prev, win)
    rn, rm12, rprev = po.search_for_initialization(olast, ocur, prev, win, 0.9, ori)
    assert rn > 10
    assert nm == rn and np.array_equal(m12, rm12) and np.array_equal(pnew, rprev)
    # second call with the updated centres, as the initializer does frame after frame
    nm2, m12b, pnew2 = m.SearchForInitialization(glast, gcur, pnew, win)
    rn2, rm12b, rprev2 = po.search_for_initialization(olast, ocur, rprev, win, 0.9, ori)
    assert nm2 == rn2 and np.array_equal(m12b, rm12b) and np.array_equal(pnew2, rprev2)


@pytest.mark.parametrize("n1,n2,seed", [(600, 40, 1), (2000, 150, 2), (64, 64, 3)])
def test_search_for_initialization_steal_chains(pkg, po, n1, n2, seed):
    """Many F1 features compete for few F2 keypoints with ever closer descriptors, so matches are stolen repeatedly
    (src/ORBmatcher.cc:637,:656-660)."""
    rng = np.random.default_rng(seed)
    m = pkg.ORBmatcher(0.9, True)
    k2 = np.zeros(n2, pkg.KP_DTYPE)
    k2["x"] = rng.uniform(100, 200, n2).astype(np.float32); k2["y"] = rng.uniform(100, 200, n2).astype(np.float32)
    k2["angle"] = rng.uniform(0, 360, n2).astype(np.float32); k2["size"] = 31; k2["octave"] = rng.integers(0, 2, n2)
    d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    src = rng.integers(0, n2, n1)
    d1 = d2[src].copy()
    bits = np.unpackbits(d1, axis=1)
    for i in range(n1):                                    # later queries tend to be closer: long steal chains
        nflip = int(rng.integers(0, 60))
        bits[i, rng.choice(256, nflip, replace=False)] ^= 1
    d1 = np.packbits(bits, axis=1)
    k1 = np.zeros(n1, pkg.KP_DTYPE)
    k1["x"] = rng.uniform(100, 200, n1).astype(np.float32); k1["y"] = rng.uniform(100, 200, n1).astype(np.float32)
    k1["angle"] = (k2["angle"][src] + rng.choice([0.0] * 6 + [45.0, 90.0, 135.0, 180.0, 270.0], n1)).astype(np.float32) % np.float32(360)
    k1["size"] = 31; k1["octave"] = (rng.random(n1) < 0.1).astype(np.int32)
    F1 = pkg.Frame(m, k1, d1, 320, 240, 300.0, 300.0, 160.0, 120.0); F2 = pkg.Frame(m, k2, d2, 320, 240, 300.0, 300.0, 160.0, 120.0)
    O1 = po.OracleFrame(k1, d1, 320, 240, 300.0, 300.0, 160.0, 120.0); O2 = po.OracleFrame(k2, d2, 320, 240, 300.0, 300.0, 160.0, 120.0)
    prev = np.stack([k1["x"], k1["y"]], 1)
    nm, m12, pnew = m.SearchForInitialization(F1, F2, prev, 120)
    rn, rm12, rprev = po.search_for_initialization(O1, O2, prev, 120, 0.9, True)
    assert rn > 5
    assert nm == rn and np.array_equal(m12, rm12) and np.array_equal(pnew, rprev)


def test_search_for_initialization_empty(pkg):
    m = pkg.ORBmatcher(0.9, True)
    k = np.zeros(0, pkg.KP_DTYPE); d = np.zeros((0, 32), np.uint8)
    E = pkg.Frame(m, k, d, 320, 240, 300.0, 300.0, 160.0, 120.0)
    rng = np.random.default_rng(0)
    k1 = np.zeros(10, pkg.KP_DTYPE); k1["x"] = 50; k1["y"] = 60
    F = pkg.Frame(m, k1, rng.integers(0, 256, (10, 32), dtype=np.uint8), 320, 240, 300.0, 300.0, 160.0, 120.0)
    n, m12, _ = m.SearchForInitialization(F, E, np.zeros((10, 2), np.float32), 50)
    assert n == 0 and (m12 == -1).all()
    n, m12, _ = m.SearchForInitialization(E, F, np.zeros((0, 2), np.float32), 50)
    assert n == 0 and len(m12) == 0


@pytest.mark.parametrize("npoints,maxobs,seed", [(2000, 12, 1), (300, 70, 2), (5, 700, 3), (50, 1, 4)])
def test_distinctive_descriptors_vs_oracle(pkg, po, npoints, maxobs, seed):
    """MapPoint::ComputeDistinctiveDescriptors, src/MapPoint.cc:185-250 (batched over map points; groups of 0, 1, 2 and many
    observations, duplicates for ties, one group larger than the shared-memory staging)."""
    rng = np.random.default_rng(seed)
    m = pkg.ORBmatcher(0.6, True)
    sizes = rng.integers(0, maxobs + 1, npoints)
    sizes[:3] = [0, 1, 2][:min(3, npoints)]
    start = np.zeros(npoints + 1, np.int32); start[1:] = np.cumsum(sizes)
    base = rng.integers(0, 256, (npoints, 32), dtype=np.uint8)
    desc = np.zeros((start[-1], 32), np.uint8)
    for p in range(npoints):
        n = sizes[p]
        if n == 0:
            continue
        bits = np.unpackbits(np.repeat(base[p][None], n, 0), axis=1)
        bits ^= rng.random(bits.shape) < rng.uniform(0.0, 0.2)
        d = np.packbits(bits, axis=1)
        if n > 3:
            d[n // 2] = d[0]                                # exact duplicates -> tied medians
        desc[start[p]:start[p + 1]] = d
    bi, bm = m.ComputeDistinctiveDescriptors(desc, start)
    rbi, rbm = po.distinctive_descriptors(desc, start)
    assert np.array_equal(bi, rbi) and np.array_equal(bm, rbm)
    assert bi[0] == -1 and (bi[1:3] == 0).all()


def _projected_points(rng, src, dst_w, dst_h, jitter=3.0):
    """Map points observed at src keypoints, 'projected' into another keyframe near the same place, with a predicted level."""
    n = len(src)
    u = (src["x"] + rng.normal(0, jitter, n)).astype(np.float32)
    v = (src["y"] + rng.normal(0, jitter, n)).astype(np.float32)
    lvl = np.clip(src["octave"] + rng.integers(0, 2, n), 0, 7).astype(np.int32)
    active = ((rng.random(n) < 0.85) & (u >= 0) & (u < dst_w) & (v >= 0) & (v < dst_h)).astype(np.uint8)
    return active, u, v, lvl


@pytest.mark.parametrize("shape,nf,th", [((240, 320), 500, 10), ((480, 752), 1000, 10), ((376, 1241), 2000, 4)])
def test_search_by_projection_sim3_vs_oracle(pkg, po, shape, nf, th):
    """ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th), src/ORBmatcher.cc:286-407 (loop closing, :389)."""
    m = pkg.ORBmatcher(0.75, True)
    gcur, glast, ocur, olast, has, outl, xyz, T = _scene(po, pkg, m, shape[0], shape[1], nf, 8500 + nf, 15.0)
    rng = np.random.default_rng(nf)
    active, u, v, lvl = _projected_points(rng, glast.kps, shape[1], shape[0])
    pre = np.full(gcur.N, -1, np.int32)
    pre[::7] = 1 << 20                                     # vpMatched already holds a point there
    n, match = m.SearchByProjectionSim3(gcur, active, u, v, lvl, glast.desc, th, pre.copy())
    rn, rmatch = po.search_by_projection_sim3(ocur, active, u, v, lvl, glast.desc, th, pre.copy())
    assert rn > 10 and n == rn and np.array_equal(match, rmatch)


@pytest.mark.parametrize("shape,nf,th", [((240, 320), 500, 2.5), ((480, 752), 1000, 2.5), ((376, 1241), 2000, 7.5)])
def test_fuse_and_sim3_candidates_vs_oracle(pkg, po, shape, nf, th):
    """Scoring loops of ORBmatcher::Fuse (src/ORBmatcher.cc:1016-1265) and SearchBySim3 (:1267-1505): independent best per point."""
    m = pkg.ORBmatcher(0.75, True)
    gcur, glast, ocur, olast, has, outl, xyz, T = _scene(po, pkg, m, shape[0], shape[1], nf, 8600 + nf, 15.0)
    rng = np.random.default_rng(nf + 1)
    a1, u1, v1, l1 = _projected_points(rng, glast.kps, shape[1], shape[0])
    a2, u2, v2, l2 = _projected_points(rng, gcur.kps, shape[1], shape[0])
    sf = np.ones(8, np.float32)
    for i in range(1, 8):
        sf[i] = np.float32(sf[i - 1] * np.float32(1.2))
    # Fuse: map points seen in `last` fused into `cur`
    fused = m.FuseCandidates(gcur, a1, u1, v1, l1, glast.desc, th)
    rbi, rbd = po.window_best(ocur, a1, u1, v1, np.float32(th) * sf[l1], l1, glast.desc)
    assert np.array_equal(fused, np.where(rbd <= 50, rbi, -1)) and (fused >= 0).sum() > 10
    bi, bd = m.BestInWindow(gcur, a1, u1, v1, np.float32(th) * sf[l1], l1 - 1, l1, glast.desc)
    assert np.array_equal(bi, rbi) and np.array_equal(bd, rbd)
    # SearchBySim3: both directions + agreement
    nfound, m12 = m.SearchBySim3(glast, gcur, a1, u1, v1, l1, glast.desc, a2, u2, v2, l2, gcur.desc, th)
    b1, d1 = po.window_best(ocur, a1, u1, v1, np.float32(th) * sf[l1], l1, glast.desc)
    b2, d2 = po.window_best(olast, a2, u2, v2, np.float32(th) * sf[l2], l2, gcur.desc)
    exp = np.full(glast.N, -1, np.int32)
    for i1 in range(glast.N):                              # src/ORBmatcher.cc:1478-1493
        idx2 = b1[i1] if d1[i1] <= 100 else -1
        if idx2 >= 0 and (b2[idx2] if d2[idx2] <= 100 else -1) == i1:
            exp[i1] = idx2
    assert nfound == (exp >= 0).sum() > 5 and np.array_equal(m12, exp)


def test_window_best_empty_cases(pkg):
    m = pkg.ORBmatcher(0.75, True)
    k = np.zeros(0, pkg.KP_DTYPE); d = np.zeros((0, 32), np.uint8)
    E = pkg.Frame(m, k, d, 320, 240, 300.0, 300.0, 160.0, 120.0)
    q = np.zeros((4, 32), np.uint8)
    bi, bd = m.BestInWindow(E, np.ones(4, np.uint8), np.full(4, 50.0), np.full(4, 50.0), np.full(4, 10.0), np.zeros(4), np.zeros(4), q)
    assert (bi == -1).all() and (bd == np.iinfo(np.int32).max).all()
    k1 = np.zeros(3, pkg.KP_DTYPE); k1["x"] = [10, 50, 52]; k1["y"] = [10, 50, 50]
    F = pkg.Frame(m, k1, np.zeros((3, 32), np.uint8), 320, 240, 300.0, 300.0, 160.0, 120.0)
    bi, bd = m.BestInWindow(F, np.array([1, 0, 1, 1], np.uint8), np.array([50, 50, 300, 51], np.float32), np.array([50, 50, 200, 50], np.float32),
                            np.full(4, 5.0, np.float32), np.full(4, -1), np.full(4, 0), q)
    assert list(bi) == [1, -1, -1, 1] and list(bd[[0, 3]]) == [0, 0]      # ties keep the first in scan order


@pytest.mark.parametrize("n1,n2,nnodes,ori,seed", [(2000, 2000, 100, True, 1), (700, 500, 12, True, 2), (300, 300, 2, False, 3), (40, 60, 30, True, 4)])
def test_search_for_triangulation_vs_oracle(pkg, po, n1, n2, nnodes, ori, seed):
    """ORBmatcher::SearchForTriangulation, src/ORBmatcher.cc:852-1014 with CheckDistEpipolarLine (:136-153)."""
    m = pkg.ORBmatcher(0.6, ori)
    fv1, d1, k1, v1, fv2, d2, k2 = _bow_case(po, pkg, n1, n2, nnodes, seed=seed + 40, flip=0.05)
    rng = np.random.default_rng(seed)
    # geometry: KF2 sees the scene shifted by about (6, 4) px; F12 of a pure sideways translation in pixel units (K = identity)
    k1["x"] = rng.uniform(20, 600, n1).astype(np.float32); k1["y"] = rng.uniform(20, 440, n1).astype(np.float32)
    k1["octave"] = rng.integers(0, 8, n1)
    twin = rng.integers(0, n1, n2)
    lam = rng.uniform(0.5, 3.0, n2)
    k2["x"] = (k1["x"][twin] + 6 * lam + rng.normal(0, 1.0, n2)).astype(np.float32)
    k2["y"] = (k1["y"][twin] + 4 * lam + rng.normal(0, 1.0, n2)).astype(np.float32)
    k2["octave"] = rng.integers(0, 8, n2)
    # descriptors of the twins so that nodes hold real matches
    d2 = d1[twin] ^ np.packbits((rng.random((n2, 256)) < 0.04).astype(np.uint8), axis=1)
    t = np.array([6.0, 4.0, 0.0])
    F12 = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]], np.float32)      # x1' F12 x2 = 0 for x2 = x1 + lambda t
    sg = np.ones(8, np.float32)
    for i in range(1, 8):
        sg[i] = np.float32(np.float32(1.2) ** i) ** 2
    has1 = (rng.random(n1) < 0.3).astype(np.uint8); has2 = (rng.random(n2) < 0.3).astype(np.uint8)
    # node membership of KF2 follows its twin most of the time
    node1 = np.zeros(n1, np.int64)
    ids, start, items = fv1
    for j, nid in enumerate(ids):
        node1[items[start[j]:start[j + 1]]] = nid
    node2 = np.where(rng.random(n2) < 0.9, node1[twin], rng.choice(ids, n2))
    ids2 = np.unique(node2)
    st2, it2 = [0], []
    for nid in ids2:
        w = np.nonzero(node2 == nid)[0]
        it2 += list(w); st2.append(len(it2))
    fv2 = (ids2.astype(np.int32), np.array(st2, np.int32), np.array(it2, np.int32))
    n, pairs, m12 = m.SearchForTriangulation(fv1, d1, k1, has1, fv2, d2, k2, has2, F12, sg)
    rn, rm12 = po.search_for_triangulation(fv1, d1, k1, has1, fv2, d2, k2, has2, F12, sg, ori)
    assert rn > 3 and n == rn and np.array_equal(m12, rm12)
    assert len(pairs) == n and not has1[pairs[:, 0]].any() and not has2[pairs[:, 1]].any()
    assert len(np.unique(pairs[:, 1])) == len(pairs)       # a KF2 feature is matched once


@pytest.mark.parametrize("nq,ndb", [(2000, 2000), (300, 5000), (1, 1), (257, 255), (2000, 70001), (5, 3), (129, 513), (128, 256), (640, 300000)])
def test_knn2_tensor_core_engine_bit_exact(pkg, po, nq, ndb):
    """ORB_KNN_TENSOR (csrc/orb_match_tc.cu): descriptor bits as +-1 int8 through tcgen05.mma kind::i8, Hamming = (256 - dot) / 2.
    Exact integer arithmetic: (idx1, d1, d2) must equal the oracle's scan (and therefore the POPC engine) bit for bit."""
    from orbslam_jpminipc_b200._lib import check, lib
    from orbslam_jpminipc_b200.synth import synth_descriptors
    m = pkg.ORBmatcher(0.6, True)
    check(lib().orb_set_knn_engine(m._h, 1), "orb_set_knn_engine")
    db, q = synth_descriptors(ndb, nq, dup_frac=0.02)
    got = m.knn2(q, db)
    ref = po.knn2(q, db)
    assert all(np.array_equal(a, b) for a, b in zip(got, ref))
    check(lib().orb_set_knn_engine(m._h, 0), "orb_set_knn_engine")
    assert all(np.array_equal(a, b) for a, b in zip(m.knn2(q, db), ref))


def test_knn2_tensor_core_engine_extremes(pkg, po):
    """all-equal / all-different bits (dot = +256 / -256), ties across tiles and column halves, multi-pair launches"""
    import ctypes as C
    import torch
    from orbslam_jpminipc_b200._lib import check, lib, ptr
    L = lib()
    m = pkg.ORBmatcher(0.6, True)
    check(L.orb_set_knn_engine(m._h, 1), "orb_set_knn_engine")
    rng = np.random.default_rng(3)
    db = rng.integers(0, 256, (1500, 32), dtype=np.uint8)
    db[700] = db[3]; db[1300] = db[3]; db[200] = ~db[5]; db[129] = db[128]
    q = np.stack([db[3], ~db[3], db[5], db[128], np.zeros(32, np.uint8), np.full(32, 255, np.uint8)] + [db[i] for i in range(300, 420)])
    assert all(np.array_equal(a, b) for a, b in zip(m.knn2(q, db), po.knn2(q, db)))
    # several (query, database) blocks in one launch
    npair, nq, nd = 5, 200, 700
    qs = rng.integers(0, 256, (npair, nq, 32), dtype=np.uint8); ds = rng.integers(0, 256, (npair, nd, 32), dtype=np.uint8)
    qs[:, ::2] = ds[:, :nq // 2] ^ np.packbits((rng.random((npair, nq // 2, 256)) < 0.05).astype(np.uint8), axis=2)
    dev = torch.device("cuda", 0)
    dq, dd = torch.from_numpy(qs).to(dev), torch.from_numpy(ds).to(dev)
    o = [torch.zeros(npair * nq, dtype=torch.int32, device=dev) for _ in range(3)]
    check(L.orb_hamming_knn2_device(m._h, ptr(dq), nq, ptr(dd), nd, npair, 0, ptr(o[0]), ptr(o[1]), ptr(o[2]),
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)), "knn2 pairs")
    torch.cuda.synchronize()
    for p in range(npair):
        ref = po.knn2(qs[p], ds[p])
        assert all(np.array_equal(o[k][p * nq:(p + 1) * nq].cpu().numpy(), ref[k]) for k in range(3)), p
