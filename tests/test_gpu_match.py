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
