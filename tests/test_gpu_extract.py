"""GPU parity tests (run on the B200 box): the CUDA extractor through the C ABI against the CPU
oracle and against the golden fixtures generated from real OpenCV.  Integer outputs bit-exact;
angles must agree within 1e-4 rad (north_star) — we additionally expect bit equality."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ANGLE_TOL_RAD = 1e-4


def _bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


def _same(kps, desc, rk, rd, what=""):
    assert len(kps) == len(rk), (what, len(kps), len(rk))
    for f in ("x", "y", "size", "response"):
        assert np.array_equal(_bits(kps[f]), _bits(rk[f])), (what, f)
    assert np.array_equal(kps["octave"], rk["octave"]) and np.all(kps["class_id"] == -1), what
    if len(kps):
        dang = np.abs(kps["angle"].astype(np.float64) - rk["angle"].astype(np.float64))
        dang = np.minimum(dang, 360 - dang) * np.pi / 180
        assert dang.max() < ANGLE_TOL_RAD, (what, dang.max())
        assert np.array_equal(_bits(kps["angle"]), _bits(rk["angle"])), (what, "angle bits")
    assert np.array_equal(desc, rd), what


@pytest.fixture(scope="module")
def pkg():
    import orbslam_jpminipc_b200 as p
    return p


@pytest.fixture(scope="module")
def po():
    from oracle import pyoracle
    return pyoracle


def _roi_ring(plane, info, ring=4):
    """the part of a padded pyramid plane the library materialises: the ROI and the ORB_RING = 4 px reflect-101 ring around it (the
    reference writes a 16 px frame, but nothing on the path reads further out than 3 px: csrc/orb_internal.h, ORB_RING)"""
    return plane[16 - ring:16 + info["h"] + ring, 16 - ring:16 + info["w"] + ring]


@pytest.mark.parametrize("name", ["e2e_320x240_n300", "e2e_640x480_n1000", "e2e_620x188_n700"])
def test_golden_frames(pkg, name):
    d = np.load(os.path.join(G, name + ".npz"))
    img = d["img"]
    ex = pkg.ORBextractor(int(d["nfeatures"]), 1.2, 8, 1, 20, max_width=img.shape[1], max_height=img.shape[0], max_batch=2)
    kps, desc = ex(img)
    _same(kps, desc, d["kps"], d["desc"], name)
    for l in (1, 4):
        info = ex.level_info(l)
        assert np.array_equal(_roi_ring(ex.level_plane(l, False), info), _roi_ring(d["L%d_plane" % l], info))
        got = ex.level_plane(l, True)[16:16 + info["h"], 16:16 + info["w"]]
        assert np.array_equal(got, d["L%d_blur" % l][16:-16, 16:-16])
    ex.close()


@pytest.mark.parametrize("shape,nf", [((480, 640), 1000), ((480, 752), 1000), ((376, 1241), 2000), ((480, 640), 2000),
                                      ((300, 301), 500), ((100, 140), 120)])
def test_vs_oracle_shapes(pkg, po, shape, nf):
    from orbslam_jpminipc_b200.synth import synth_frame
    h, w = shape
    ex = pkg.ORBextractor(nf, 1.2, 8, 1, 20, max_width=w, max_height=h, max_batch=4)
    orc = po.OracleExtractor(nf, 1.2, 8, 1, 20)
    for seed in (1000, 1001, 1002):
        img = synth_frame(h, w, seed, quadrants=(seed != 1001))
        kps, desc = ex(img)
        rk, rd = orc(img)
        _same(kps, desc, rk, rd, (shape, nf, seed))
        for l in range(8):
            assert ex.level_info(l)["nKept"] == orc.level_info(l)["nKept"]
            assert np.array_equal(_roi_ring(ex.level_plane(l), ex.level_info(l)), _roi_ring(orc.level_plane(l), ex.level_info(l))), (shape, l)
    ex.close()


@pytest.mark.parametrize("h,w", [(161, 225), (162, 239), (287, 240), (289, 241), (300, 257), (193, 272), (416, 287), (290, 352)])
def test_partly_filled_edge_tiles(pkg, po, h, w):
    """k_fast_nms / k_blur / k_resize walk only the filled part of a tile on the right / bottom edge of a level: widths with
    (w - 32) mod 64 in {1, 15, 16, 17, 33, 48, 63, 0} and heights with (h - 32) mod 128 in {1, 2, 127, 129, ...} at level 0 (the
    other levels add their own residues), dense and sparse frames through one context so that stale tile contents would show."""
    from orbslam_jpminipc_b200.synth import synth_frame
    rng = np.random.default_rng(h * 1000 + w)
    frames = [synth_frame(h, w, 500 + h), rng.integers(0, 256, (h, w), dtype=np.uint8), synth_frame(h, w, 501 + w, quadrants=False)]
    ex = pkg.ORBextractor(400, 1.2, 8, 1, 20, max_width=w, max_height=h, max_batch=3)
    orc = po.OracleExtractor(400, 1.2, 8, 1, 20)
    try:
        ref = [orc(f) for f in frames]
    except RuntimeError:
        with pytest.raises(pkg.OrbError):
            ex(frames[0])
        ex.close()
        return
    for f, (rk, rd) in zip(frames, ref):
        kps, desc = ex(f)
        _same(kps, desc, rk, rd, (h, w))
    out = ex.extract_batch(np.stack(frames))
    for (kps, desc), (rk, rd) in zip(out, ref):
        _same(kps, desc, rk, rd, (h, w, "batch"))
    ex.close()


@pytest.mark.parametrize("h,w,nf", [(97, 203, 150), (120, 160, 100), (90, 300, 200), (64, 64, 60), (200, 200, 50), (150, 170, 130)])
def test_degenerate_geometry_agrees_with_oracle(pkg, po, h, w, nf):
    """Shapes on which the reference itself throws / divides by zero: product and oracle must agree on the verdict."""
    from orbslam_jpminipc_b200.synth import synth_frame
    img = synth_frame(h, w, 77)
    ex = pkg.ORBextractor(nf, 1.2, 8, 1, 20, max_width=w, max_height=h, max_batch=1)
    try:
        rk, rd = po.OracleExtractor(nf)(img)
    except RuntimeError:
        with pytest.raises(pkg.OrbError):
            ex(img)
    else:
        kps, desc = ex(img)
        _same(kps, desc, rk, rd, (h, w, nf))
    ex.close()


def test_batch_equals_single_and_chunks(pkg, po):
    from orbslam_jpminipc_b200.synth import synth_frames
    frames = synth_frames(7, 240, 320, seed0=4000)
    ex = pkg.ORBextractor(400, 1.2, 8, 1, 20, max_width=320, max_height=240, max_batch=3)   # 3 chunks: 3+3+1
    orc = po.OracleExtractor(400, 1.2, 8, 1, 20)
    res = ex.extract_batch(frames)
    for i, (kps, desc) in enumerate(res):
        rk, rd = orc(frames[i])
        _same(kps, desc, rk, rd, ("batch", i))
    ex.close()


def test_full_machine_batch_is_deterministic_and_exact(pkg, po):
    """A batch large enough to keep every SM's resident CTAs busy for many work items (k_fast_nms re-uses ONE raw TMA buffer per CTA,
    k_resize_u / k_blur two, all three draw items from atomic queues): sampled frames equal the oracle and ten repeats of the launch
    give identical bytes, dense, sparse and flat frames mixed so that item costs differ widely."""
    from orbslam_jpminipc_b200.synth import synth_frame
    h, w, n = 480, 640, 192
    rng = np.random.default_rng(77)
    frames = []
    for i in range(n):
        kind = i % 4
        if kind == 0: f = synth_frame(h, w, 6000 + i)
        elif kind == 1: f = synth_frame(h, w, 6000 + i, quadrants=False)
        elif kind == 2: f = rng.integers(0, 256, (h, w), dtype=np.uint8)
        else:
            f = np.full((h, w), 90, np.uint8)
            f[100:380, 150:500] = synth_frame(280, 350, 6000 + i)            # texture island in a flat frame: most tiles reject early
        frames.append(f)
    frames = np.stack(frames)
    ex = pkg.ORBextractor(1000, 1.2, 8, 1, 20, max_width=w, max_height=h, max_batch=n)
    orc = po.OracleExtractor(1000, 1.2, 8, 1, 20)
    first = ex.extract_batch(frames)
    for i in (0, 1, 2, 3, 95, 126, 190, 191):
        rk, rd = orc(frames[i])
        _same(first[i][0], first[i][1], rk, rd, ("full-machine batch", i))
    for rep in range(10):
        again = ex.extract_batch(frames)
        for i in range(n):
            assert np.array_equal(again[i][0].view(np.uint8), first[i][0].view(np.uint8)) and np.array_equal(again[i][1], first[i][1]), (rep, i)
    ex.close()


def test_other_parameters(pkg, po):
    from orbslam_jpminipc_b200.synth import synth_frame
    img = synth_frame(360, 480, 5000)
    for nf, sf, nl, th in [(800, 1.2, 8, 12), (600, 1.5, 5, 20), (500, 1.1, 6, 5), (1200, 1.2, 4, 30), (300, 1.2, 1, 20),
                           (400, 2.0, 3, 20), (400, 1.2, 2, 9), (500, 2.5, 2, 20)]:
        ex = pkg.ORBextractor(nf, sf, nl, 1, th, max_width=480, max_height=360, max_batch=1)
        kps, desc = ex(img)
        rk, rd = po.OracleExtractor(nf, sf, nl, 1, th)(img)
        _same(kps, desc, rk, rd, (nf, sf, nl, th))
        ex.close()


@pytest.mark.parametrize("seed", range(10))
def test_random_shapes_and_parameters(pkg, po, seed):
    """Seeded sweep over image shapes, feature counts, scale factors, level counts and thresholds (six per seed): tile residues of every
    kind at every level for k_resize_u (packed taps, role-swapping row sets), k_fast_nms<true> (112-row tiles, half-lane tile) and
    k_blur; where the reference itself throws on a geometry, product and oracle must agree on that."""
    from orbslam_jpminipc_b200.synth import synth_frame
    rng = np.random.default_rng(9000 + seed)
    for _ in range(6):
        h, w = int(rng.integers(120, 700)), int(rng.integers(160, 900))
        nf = int(rng.integers(100, 1500))
        sf = float(rng.choice([1.1, 1.2, 1.2, 1.2, 1.25, 1.3, 1.44, 1.5, 2.0]))
        nl = int(rng.integers(1, 9))
        th = int(rng.choice([7, 9, 12, 20, 20, 25, 40]))
        kind = int(rng.integers(0, 3))
        if seed >= 6:                                          # wider ranges: very many features on few levels, fastTh 0, tiny levels
            nf, th = int(rng.integers(50, 3000)), int(rng.choice([0, 1, 5, 20, 80]))
        img = (synth_frame(h, w, 9100 + seed) if kind == 0 else synth_frame(h, w, 9200 + seed, quadrants=False) if kind == 1
               else rng.integers(0, 256, (h, w), dtype=np.uint8))
        what = (h, w, nf, sf, nl, th, kind)
        ex = pkg.ORBextractor(nf, sf, nl, 1, th, max_width=w, max_height=h, max_batch=2)
        try:
            rk, rd = po.OracleExtractor(nf, sf, nl, 1, th)(img)
        except RuntimeError:
            with pytest.raises(pkg.OrbError):
                ex(img)
            ex.close()
            continue
        kps, desc = ex(img)
        _same(kps, desc, rk, rd, what)
        (k2, d2), (k3, d3) = ex.extract_batch(np.stack([img, img[::-1].copy()]))
        _same(k2, d2, rk, rd, what + ("batch",))
        rk3, rd3 = po.OracleExtractor(nf, sf, nl, 1, th)(img[::-1].copy())
        _same(k3, d3, rk3, rd3, what + ("flipped",))
        ex.close()


@pytest.mark.parametrize("h,w,nf,scale,nlevels,fast_th", [
    (261, 623, 1436, 1.5, 6, 9), (673, 239, 1436, 1.25, 1, 9), (206, 796, 1486, 1.3, 6, 40), (495, 298, 1092, 1.44, 1, 12),
    (231, 698, 2278, 1.1, 1, 0), (360, 915, 2216, 1.7, 8, 1), (234, 523, 1597, 1.5, 8, 20)])
def test_degenerate_grids(pkg, po, h, w, nf, scale, nlevels, fast_th):
    """The configurations of tests/test_ref_build.py::test_extractor_degenerate_grids (oracle == the reference's own ORBextractor.cc
    there): inner cells reaching past size - 16, keypoints that need the whole 16 px frame, 209 cells on one level, fastTh 0, and two
    geometries the reference throws on."""
    from orbslam_jpminipc_b200.synth import synth_frame
    rng = np.random.default_rng(h * w)
    ex = pkg.ORBextractor(nf, scale, nlevels, 1, fast_th, max_width=w, max_height=h, max_batch=2)
    orc = po.OracleExtractor(nf, scale, nlevels, 1, fast_th)
    for img in (synth_frame(h, w, 31, quadrants=False), rng.integers(0, 256, (h, w), dtype=np.uint8)):
        try:
            rk, rd = orc(img)
        except RuntimeError:
            with pytest.raises(pkg.OrbError):
                ex(img)
            continue
        kps, desc = ex(img)
        _same(kps, desc, rk, rd, (h, w, nf, scale, nlevels, fast_th))
    ex.close()


def test_strided_input_and_edge_cases(pkg, po):
    from orbslam_jpminipc_b200.synth import synth_frame
    big = synth_frame(300, 500, 6000)
    view = big[10:250, 20:340]                       # non-contiguous rows (stride 500)
    ex = pkg.ORBextractor(300, 1.2, 8, 1, 20, max_width=320, max_height=240, max_batch=1)
    kps, desc = ex(view)
    rk, rd = po.OracleExtractor(300, 1.2, 8, 1, 20)(np.ascontiguousarray(view))
    _same(kps, desc, rk, rd, "strided")
    k, d = ex(np.zeros((0, 0), np.uint8))            # empty image: silent, no keypoints
    assert len(k) == 0 and d.shape == (0, 32)
    k, d = ex(np.full((240, 320), 9, np.uint8))      # flat image: no corners
    assert len(k) == 0
    with pytest.raises(pkg.OrbError):                # larger than the context was sized for
        ex(np.zeros((241, 320), np.uint8))
    ex.close()
    tiny = pkg.ORBextractor(10, 1.2, 8, 1, 20, max_width=100, max_height=100, max_batch=1)
    with pytest.raises(pkg.OrbError):                # grid with zero columns: the reference divides by zero
        tiny(np.zeros((100, 100), np.uint8))
    tiny.close()


def test_async_tickets_match_sync_call(pkg):
    """orb_extract_batch_async / orb_wait: several calls in flight (more than the ticket ring holds) give the results of the
    blocking call, chunked over both work sets."""
    import ctypes as C
    from orbslam_jpminipc_b200._lib import check, lib, ptr
    from orbslam_jpminipc_b200.synth import synth_frames
    L = lib()
    h, w, nb = 240, 320, 6
    ex = pkg.ORBextractor(300, 1.2, 8, 1, 20, max_width=w, max_height=h, max_batch=4)
    cap = ex.capacity
    batches = [np.stack(synth_frames(nb, h, w, seed0=4000 + 10 * b)) for b in range(11)]
    ref = []
    for fr in batches:
        k = np.zeros((nb, cap), pkg.KP_DTYPE); d = np.zeros((nb, cap, 32), np.uint8); c = np.zeros(nb, np.int32)
        check(L.orb_extract_batch(ex._h, ptr(fr), nb, w, h, w, w * h, ptr(k), ptr(d), cap, ptr(c)), "sync")
        ref.append((k, d, c))
    outs, tickets = [], []
    for fr in batches:
        k = np.zeros((nb, cap), pkg.KP_DTYPE); d = np.zeros((nb, cap, 32), np.uint8); c = np.zeros(nb, np.int32)
        t = C.c_longlong(-1)
        check(L.orb_extract_batch_async(ex._h, ptr(fr), nb, w, h, w, w * h, ptr(k), ptr(d), cap, ptr(c), C.byref(t)), "async")
        outs.append((k, d, c)); tickets.append(t.value)
    assert tickets == list(range(tickets[0], tickets[0] + len(batches)))
    for t in reversed(tickets):                    # any order, also tickets whose record was recycled
        check(L.orb_wait(ex._h, t), "wait")
    for (k, d, c), (rk, rd, rc) in zip(outs, ref):
        assert np.array_equal(c, rc)
        for i in range(nb):
            assert np.array_equal(k[i, :c[i]].view(np.uint8), rk[i, :c[i]].view(np.uint8)) and np.array_equal(d[i, :c[i]], rd[i, :c[i]])
    assert L.orb_wait(ex._h, tickets[-1] + 5) != 0          # unknown ticket
    # the blocking call still works with nothing in flight
    k = np.zeros((nb, cap), pkg.KP_DTYPE); d = np.zeros((nb, cap, 32), np.uint8); c = np.zeros(nb, np.int32)
    check(L.orb_extract_batch(ex._h, ptr(batches[0]), nb, w, h, w, w * h, ptr(k), ptr(d), cap, ptr(c)), "sync again")
    assert np.array_equal(c, ref[0][2])


def test_threshold_fallback_cells(pkg, po):
    """Cells with <= 3 corners at fastTh are re-detected at threshold 7 (src/ORBextractor.cc:609-614): low-contrast frames where
    most cells fall back, huge cells (tiny nfeatures -> one CTA walks a cell wider than its bitmap), and fastTh on both sides of 7."""
    from orbslam_jpminipc_b200.synth import synth_frame
    rng = np.random.default_rng(77)
    base = synth_frame(480, 640, 7100).astype(np.float32)
    low = np.clip((base - 128.0) * 0.18 + 120.0, 0, 255).astype(np.uint8)            # contrast too low for th=20 almost everywhere
    mixed = low.copy(); mixed[100:300, 200:500] = base[100:300, 200:500].astype(np.uint8)
    sparse = np.full((480, 640), 100, np.uint8)
    for _ in range(40):                                                               # a few isolated bright blobs on a flat frame
        y, x = int(rng.integers(30, 450)), int(rng.integers(30, 610))
        sparse[y:y + 3, x:x + 3] = int(rng.integers(110, 140))
    for img, tag in ((low, "low"), (mixed, "mixed"), (sparse, "sparse")):
        for nf, th in ((1000, 20), (60, 20), (20, 40), (1000, 7), (1000, 5), (500, 8)):
            ex = pkg.ORBextractor(nf, 1.2, 8, 1, th, max_width=640, max_height=480, max_batch=1)
            orc = po.OracleExtractor(nf, 1.2, 8, 1, th)
            try:
                rk, rd = orc(img)
            except Exception:
                ex.close()
                continue                              # geometry the reference cannot process either (covered elsewhere)
            kps, desc = ex(img)
            _same(kps, desc, rk, rd, (tag, nf, th))
            ex.close()


@pytest.mark.parametrize("shape,nf,th", [((240, 320), 300, 20), ((480, 640), 1000, 20), ((480, 752), 1000, 12), ((376, 1241), 2000, 20)])
def test_harris_score_vs_oracle(pkg, po, shape, nf, th):
    """scoreType = HARRIS_SCORE (src/ORBextractor.cc:616-620, HarrisResponses :79-120): float responses drive both retainBest
    passes; keypoints (incl. response bit patterns), order and descriptors must equal the oracle's."""
    from orbslam_jpminipc_b200.synth import synth_frame
    h, w = shape
    ex = pkg.ORBextractor(nf, 1.2, 8, pkg.ORBextractor.HARRIS_SCORE, th, max_width=w, max_height=h, max_batch=2)
    orc = po.OracleExtractor(nf, 1.2, 8, 0, th)
    for seed in (3000, 3001):
        img = synth_frame(h, w, seed, quadrants=(seed == 3000))
        kps, desc = ex(img)
        rk, rd = orc(img)
        _same(kps, desc, rk, rd, ("harris", shape, nf, seed))
        assert len(kps) > 50 and (kps["response"] != np.round(kps["response"])).any()      # Harris values, not FAST scores
    # the FAST_SCORE result on the same frame differs (different selection)
    fk, _ = pkg.ORBextractor(nf, 1.2, 8, 1, th, max_width=w, max_height=h, max_batch=1)(img)
    assert len(fk) != len(kps) or not np.array_equal(fk["x"], kps["x"])
    ex.close()


def test_two_extractors_of_different_shapes_coexist(pkg, po):
    """ORB-SLAM keeps two long-lived extractors (src/Tracking.cc:111,126: nFeatures and 2*nFeatures); a later, smaller context must
    not shrink per-kernel attributes the earlier one relies on."""
    from orbslam_jpminipc_b200.synth import synth_frame
    big = pkg.ORBextractor(2000, 1.2, 8, 1, 20, device=0, max_width=752, max_height=480, max_batch=2)
    a = synth_frame(480, 752, 4100)
    k1, d1 = big(a)
    small = pkg.ORBextractor(300, 1.2, 8, 1, 20, device=0, max_width=320, max_height=240, max_batch=1)
    b = synth_frame(240, 320, 4101)
    ks, ds = small(b)
    k2, d2 = big(a)                                          # the first context still works and gives the same answer
    assert np.array_equal(k1, k2) and np.array_equal(d1, d2)
    rk, rd = po.OracleExtractor(2000, 1.2, 8, 1, 20)(a)
    assert len(k2) == len(rk) and np.array_equal(d2, rd)
    rks, rds = po.OracleExtractor(300, 1.2, 8, 1, 20)(b)
    assert len(ks) == len(rks) and np.array_equal(ds, rds)


def test_repeated_single_frame_calls_replay_a_graph(pkg, po):
    """The tracking thread calls the extractor once per frame with the same buffers: from the third call on the pass is replayed from a
    CUDA graph (orb_api.cu launch_extract).  Results must not depend on that, also when the image content and then the shape change."""
    from orbslam_jpminipc_b200.synth import synth_frame
    ex = pkg.ORBextractor(500, 1.2, 8, 1, 20, device=0, max_width=640, max_height=480, max_batch=1)
    orc = po.OracleExtractor(500, 1.2, 8, 1, 20)
    imgs = [synth_frame(240, 320, 5000 + i) for i in range(3)]
    want = [orc(im) for im in imgs]
    for rep in range(4):
        for im, (rk, rd) in zip(imgs, want):
            k, d = ex(im)
            assert len(k) == len(rk) and np.array_equal(k["response"], rk["response"]) and np.array_equal(d, rd)
    big = synth_frame(480, 640, 5100)                      # new plan: the captured graph must not be replayed
    rk, rd = orc(big)
    for rep in range(3):
        k, d = ex(big)
        assert len(k) == len(rk) and np.array_equal(d, rd)
    k, d = ex(imgs[0])
    assert np.array_equal(d, want[0][1])


def test_descriptor_fma_variant(pkg, po):
    """orb_set_descriptor_fma: the descriptor rotation as GCC contracts it under the reference's own flags; against the oracle's variant
    (which tests/test_ref_build.py pins to the reference built that way) and, where it travelled, that reference build itself."""
    from oracle import pyref
    from orbslam_jpminipc_b200.synth import synth_frames
    frames = synth_frames(8, 480, 752, 6000)
    ex = pkg.ORBextractor(1000, 1.2, 8, 1, 20, device=0, max_width=752, max_height=480, max_batch=8, desc_fma=True)
    orc = po.OracleExtractor(1000, 1.2, 8, 1, 20, desc_fma=True)
    ref = pyref.RefExtractor(1000, 1.2, 8, 1, 20, fma=True) if pyref.fma_available() else None
    for img, (k, d) in zip(frames, ex.extract_batch(frames)):
        rk, rd = orc(img)
        assert len(k) == len(rk) and np.array_equal(k["angle"].view(np.uint32), rk["angle"].view(np.uint32)) and np.array_equal(d, rd)
        if ref is not None:
            assert np.array_equal(d, ref(img)[1])
    # switching the flag back restores the default form
    from orbslam_jpminipc_b200._lib import lib, check
    check(lib().orb_set_descriptor_fma(ex._h, 0), "orb_set_descriptor_fma")
    plain = po.OracleExtractor(1000, 1.2, 8, 1, 20)
    for img, (k, d) in zip(frames, ex.extract_batch(frames)):
        assert np.array_equal(d, plain(img)[1])


@pytest.mark.parametrize("kind,h,w,nf", [("two_level", 240, 320, 300), ("two_level", 480, 752, 1000), ("big_cells", 720, 1280, 200),
                                          ("big_cells", 1080, 1920, 1000), ("blocks", 376, 1241, 2000)])
def test_selection_ties_and_long_lists(pkg, po, kind, h, w, nf):
    """The selection kernel replays std::nth_element with a warp per cell (k_select_fast): tie-heavy scores (images with two or a
    few gray levels), cell lists longer than the 1024-entry staging buffer (few features on a large image: the serial in-place path)
    and level lists far above the quota must all keep the reference's order."""
    rng = np.random.default_rng(h * 7 + nf)
    if kind == "two_level":                               # two gray levels: a handful of distinct FAST scores
        small = rng.random((h // 3 + 1, w // 3 + 1)) < 0.5
        img = np.where(np.kron(small, np.ones((3, 3), bool))[:h, :w], 200, 60).astype(np.uint8)
    elif kind == "blocks":
        lv = rng.integers(0, 4, (h // 5 + 1, w // 5 + 1)) * 60 + 20
        img = np.kron(lv, np.ones((5, 5), np.int64))[:h, :w].astype(np.uint8)
    else:                                                  # dense noise: thousands of candidates per (large) cell
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    ex = pkg.ORBextractor(nf, 1.2, 8, 1, 20, device=0, max_width=w, max_height=h, max_batch=1)
    k, d = ex(img)
    rk, rd = po.OracleExtractor(nf, 1.2, 8, 1, 20)(img)
    assert len(k) == len(rk) and len(k) > nf // 4
    for f in ("x", "y", "response", "octave"):
        assert np.array_equal(k[f], rk[f]), f
    assert np.array_equal(d, rd)


@pytest.mark.parametrize("shape,nf,score", [((480, 640), 1000, 1), ((376, 1241), 2000, 1), ((240, 320), 300, 0)])
def test_small_calls_take_the_latency_paths_and_agree(pkg, po, shape, nf, score):
    """Calls of a few frames (the reference's one frame per Frame::Frame, src/Frame.cc:60) run a latency-tuned form of the pass:
    short resize tiles, CTA-per-cell compaction, 32-warp selection, programmatic dependent launches and, for pageable outputs, one
    staged result block.  Every combination of call size (below / at / above the limit) and output memory kind must equal the oracle,
    and a context with all of it switched off must agree byte for byte."""
    import torch
    from orbslam_jpminipc_b200._lib import check, lib, ptr
    from orbslam_jpminipc_b200.synth import synth_frames
    h, w = shape
    frames = np.stack(synth_frames(14, h, w, seed0=8100 + nf))
    orc = po.OracleExtractor(nf, 1.2, 8, score, 20)
    want = [orc(f) for f in frames]
    off = {"ORB_SMALL_CALL": "0", "ORB_SELECT_WIDE": "0", "ORB_COMPACT_WIDE": "0", "ORB_PDL": "0", "ORB_STAGE_SMALL": "0"}
    L = lib()
    for env in ({}, off):
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            ex = pkg.ORBextractor(nf, 1.2, 8, score, 20, max_width=w, max_height=h, max_batch=14)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        cap = ex.capacity
        for n in (1, 2, 4, 5, 12, 13, 14, 1):          # PDL / staged results up to 4 frames, the kernel forms up to 12
            for pinned in (False, True):
                k = torch.zeros((n, cap, 7), dtype=torch.int32); d = torch.full((n, cap, 32), 0xAB, dtype=torch.uint8)
                c = torch.zeros(n, dtype=torch.int32)
                if pinned:
                    k, d, c = k.pin_memory(), d.pin_memory(), c.pin_memory()
                fr = frames[:n]
                for rep in range(3):                       # repeated identical calls: the captured-graph / PDL paths of later calls
                    check(L.orb_extract_batch(ex._h, ptr(fr), n, w, h, w, w * h, ptr(k), ptr(d), cap, ptr(c)), "orb_extract_batch")
                    kk = k.numpy().view(np.uint8).reshape(n, cap, 28).copy().view(pkg.KP_DTYPE).reshape(n, cap)
                    for i in range(n):
                        rk, rd = want[i]
                        _same(kk[i, :int(c[i])], d.numpy()[i, :int(c[i])], rk, rd, (env != {}, n, pinned, rep, i))
        ex.close()


def test_rotation_factors_equal_libm_on_every_angle():
    """cosf / sinf of the descriptor rotation (reference src/ORBextractor.cc:160): the device function k_describe calls (csrc/orb_trig.h)
    against this host's libm on EVERY float angle in [0, 360] degrees, 1 135 869 953 arguments (tools/cpp/sincos_exhaustive.cu, built by
    __graft_entry__.build(); a few seconds).  Exit code 1 = some angle differs."""
    import subprocess
    exe = os.path.join(os.path.dirname(G), os.pardir, "tools", "cpp", "build", "sincos_ex")
    if not os.path.exists(exe):
        pytest.skip("tools/cpp/build/sincos_ex not built")
    try:
        r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    except (OSError, subprocess.TimeoutExpired) as e:       # a host that cannot run the prebuilt checker (or has a single slow core) is not a parity failure
        pytest.skip("sincos_ex did not run: %r" % (e,))
    if r.returncode == 2:
        pytest.skip("sincos_ex found no usable CUDA device")
    tail = r.stdout.strip().splitlines()[-5:]
    assert r.returncode == 0, tail
    assert any("the function k_describe calls" in l and "differs from the host's cosf / sinf: 0 of" in l for l in tail), tail
