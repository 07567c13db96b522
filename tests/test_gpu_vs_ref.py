"""The CUDA path against the REFERENCE's own code (oracle/_ref/libref_orbslam.so, see tests/test_ref_build.py), through the C ABI.
The library is built where /root/reference exists and travels to the GPU box as a built file; the tests skip when it is absent."""
import numpy as np
import pytest

from oracle import pyref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not pyref.available(), reason="oracle/_ref not built")]


@pytest.fixture(scope="module")
def pkg():
    import orbslam_jpminipc_b200 as p
    return p


@pytest.mark.parametrize("h,w,nf,score", [(480, 640, 1000, 1), (480, 752, 1000, 1), (376, 1241, 2000, 1), (480, 640, 2000, 1), (480, 640, 1000, 0)])
def test_extraction_equals_reference(pkg, h, w, nf, score):
    """ORBextractor::operator() (reference src/ORBextractor.cc:718-779): keypoints, octaves, responses, descriptors bit-exact,
    angles within 1e-4 rad (they are in fact bit-exact too)."""
    from orbslam_jpminipc_b200.synth import synth_frames
    frames = synth_frames(4, h, w, 3000 + w)
    ex = pkg.ORBextractor(nf, 1.2, 8, score, 20, device=0, max_width=w, max_height=h, max_batch=4)
    ref = pyref.RefExtractor(nf, 1.2, 8, score, 20)
    out = ex.extract_batch(frames)
    for img, (k, d) in zip(frames, out):
        rk, rd = ref(img)
        assert len(k) == len(rk) and len(k) > nf // 2
        for f in ("x", "y", "size", "response", "octave", "class_id"):
            assert np.array_equal(k[f], rk[f]), f
        assert np.max(np.abs(k["angle"] - rk["angle"])) * np.pi / 180 < 1e-4
        assert np.array_equal(k["angle"].view(np.uint32), rk["angle"].view(np.uint32))
        assert np.array_equal(d, rd)


@pytest.mark.parametrize("shape,nf,th", [((480, 752), 1000, 15.0), ((376, 1241), 2000, 15.0)])
def test_tracking_step_equals_reference(pkg, shape, nf, th):
    """BASELINE config 3: extract two frames on the GPU, then ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th)
    (reference src/ORBmatcher.cc:1507-1620), against the reference's Frame / MapPoint / ORBmatcher objects fed the
    reference's own extraction of the same images."""
    from orbslam_jpminipc_b200.synth import synth_frame, shifted_frame
    h, w = shape
    a = synth_frame(h, w, 9000 + nf, quadrants=False)
    b = shifted_frame(a, 3, 2, 9001 + nf)
    ex = pkg.ORBextractor(nf, 1.2, 8, 1, 20, device=0, max_width=w, max_height=h, max_batch=2)
    (ka, da), (kb, db) = ex.extract_batch(np.stack([a, b]))
    ref = pyref.RefExtractor(nf, 1.2, 8, 1, 20)
    (rka, rda), (rkb, rdb) = ref(a), ref(b)
    rng = np.random.default_rng(nf)
    fx = fy = 500.0
    cx, cy = w / 2.0, h / 2.0
    z = rng.uniform(2, 10, len(ka)).astype(np.float32)
    xyz = np.stack([(ka["x"] - cx) / fx * z, (ka["y"] - cy) / fy * z, z], 1).astype(np.float32)
    T = np.eye(4, dtype=np.float32)
    T[:3, 3] = [0.03, 0.02, 0.01]
    has = (rng.random(len(ka)) < 0.9).astype(np.uint8)
    outl = (rng.random(len(ka)) < 0.05).astype(np.uint8)
    m = pkg.ORBmatcher(0.9, True, extractor=ex)
    n, match = m.SearchByProjection(pkg.Frame(m, kb, db, w, h, fx, fy, cx, cy), pkg.Frame(m, ka, da, w, h, fx, fy, cx, cy), th, has, outl, xyz, T)
    rcur = pyref.RefFrame(rkb, rdb, w, h, fx, fy, cx, cy).set_pose(T)
    rlast = pyref.RefFrame(rka, rda, w, h, fx, fy, cx, cy).set_mappoints(has, xyz, outl)
    rn, rmatch = pyref.search_by_projection(rcur, rlast, th, 0.9, True)
    assert rn > 100
    assert n == rn and np.array_equal(match, rmatch)


def test_search_by_bow_equals_reference(pkg):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) (reference src/ORBmatcher.cc:155-284)"""
    from oracle import pyoracle as po
    from test_gpu_match import _bow_case
    m = pkg.ORBmatcher(0.75, True)
    fv1, d1, k1, valid, fv2, d2, k2 = _bow_case(po, pkg, 2000, 2000, 100, seed=77)
    n, match = m.SearchByBoW(fv1, d1, k1, valid, fv2, d2, k2)
    cam = (640, 480, 500.0, 500.0, 320.0, 240.0)
    rkf = pyref.RefFrame(k1, d1, *cam).set_featvec(*fv1).set_mappoints(valid)
    rf = pyref.RefFrame(k2, d2, *cam).set_featvec(*fv2)
    rn, rmatch = pyref.search_by_bow(rkf, rf, 0.75, True)
    assert rn > 50
    assert n == rn and np.array_equal(match, rmatch)


def test_vocabulary_transform_equals_reference(pkg, tmp_path):
    """DBoW2 TemplatedVocabulary::transform + L1 score (reference Thirdparty/DBoW2) against the GPU vocabulary kernels."""
    from orbslam_jpminipc_b200 import synth
    k, L = 10, 4
    parent, desc, weight = synth.synth_vocabulary(k, L, seed=3, prune_frac=0.05)
    path = str(tmp_path / "voc.txt")
    synth.write_vocabulary_text(path, k, L, parent, desc, weight, trailing_newline=False)
    rv = pyref.RefVocabulary(path)
    ex = pkg.ORBextractor(500, 1.2, 8, 1, 20, device=0, max_width=320, max_height=240, max_batch=1)
    gv = pkg.ORBVocabulary(ex).create(k, L, parent, desc, weight)
    rng = np.random.default_rng(9)
    leaves = np.nonzero(~np.isin(np.arange(len(parent)), parent))[0]
    for n in (1, 333, 2000):
        feats = desc[rng.choice(leaves, n)] ^ np.packbits((rng.random((n, 256)) < 0.05).astype(np.uint8), axis=1)
        (rw, rvv), (rn_, rs, ri) = rv.transform(feats, 2)
        bow, fv = gv.transform(feats, 2)
        assert np.array_equal(bow[0], rw) and np.array_equal(np.asarray(bow[1]).view(np.uint64), rvv.view(np.uint64))
        assert np.array_equal(fv[0], rn_) and np.array_equal(fv[1], rs) and np.array_equal(fv[2], ri)


def test_backend_searches_equal_reference(pkg):
    """SearchByProjection(KeyFrame,Scw), Fuse and SearchBySim3 (reference src/ORBmatcher.cc:286-407, :1016-1134, :1267-1505) through the
    CUDA window searches, fed the reference's own projections and predicted levels (the adapter's part at the C ABI)."""
    from oracle import pyoracle as po
    import test_ref_build as T
    (ka, da), (kb, db), cam, has, src, Tcw, S = T._backend_scene(po, (480, 752), 1000, 9900, 1.2)
    w, h, fx, fy, cx, cy = cam
    m = pkg.ORBmatcher(0.6, True)
    gkf = pkg.Frame(m, kb, db, w, h, fx, fy, cx, cy)
    # SearchByProjection(KF, Scw)
    kf = pyref.RefFrame(kb, db, *cam).set_pose(Tcw)
    pre = np.full(len(kb), -1, np.int32); pre[::7] = 555555
    rn, rmatched, (act, u, v, lv) = pyref.search_by_projection_sim3(kf, src, S, 10, pre.copy())
    n, matched = m.SearchByProjectionSim3(gkf, act, u, v, lv, da, 10, pre.copy())
    assert rn > 20 and n == rn and np.array_equal(matched, rmatched)
    # Fuse
    rng = np.random.default_rng(1)
    occupied = (rng.random(len(kb)) < 0.5).astype(np.uint8)
    kf2 = pyref.RefFrame(kb, db, *cam).set_pose(Tcw).set_mappoints(occupied).update_points()
    rn, rfused, (act, u, v, lv) = pyref.fuse_sim3(kf2, src, S, 4.0)
    fused = m.FuseCandidates(gkf, act, u, v, lv, da, th=4.0)
    assert rn > 20 and np.array_equal(fused, rfused)


def test_search_by_sim3_equals_reference(pkg):
    from oracle import pyoracle as po
    import test_ref_build as T
    (ka, da), (kb, db), cam, has, outl, xyz, Tcw = T._scene(po, 480, 752, 1000, 9901)
    w, h, fx, fy, cx, cy = cam
    rng = np.random.default_rng(4)
    k1 = pyref.RefFrame(ka, da, *cam).set_mappoints(has, xyz).update_points()
    z2 = rng.uniform(2, 10, len(kb)).astype(np.float32)
    pc2 = np.stack([(kb["x"] - cx) / fx * z2, (kb["y"] - cy) / fy * z2, z2], 1).astype(np.float64)
    R, t = Tcw[:3, :3].astype(np.float64), Tcw[:3, 3].astype(np.float64)
    has2 = (rng.random(len(kb)) < 0.9).astype(np.uint8)
    k2 = pyref.RefFrame(kb, db, *cam).set_pose(Tcw).set_mappoints(has2, ((pc2 - t) @ R).astype(np.float32)).update_points()
    s12 = 1.1
    rn, rm12, (a12, u12, v12, l12), (a21, u21, v21, l21) = pyref.search_by_sim3(k1, k2, s12, R.T.astype(np.float32), (-s12 * (R.T @ t)).astype(np.float32), 7.5)
    m = pkg.ORBmatcher(0.6, True)
    g1, g2 = pkg.Frame(m, ka, da, w, h, fx, fy, cx, cy), pkg.Frame(m, kb, db, w, h, fx, fy, cx, cy)
    n, m12 = m.SearchBySim3(g1, g2, a12, u12, v12, l12, da, a21, u21, v21, l21, db, th=7.5)
    assert rn > 20 and n == rn and np.array_equal(m12, rm12)
