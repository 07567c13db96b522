"""GPU parity of the frame plumbing (csrc/orb_frame.cu) against the oracle: colour conversion feeding the extractor,
Frame::UndistortKeyPoints, Frame::ComputeImageBounds, and a tracked frame pair on a distorted camera."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
K = (np.float32(517.306408), np.float32(516.469215), np.float32(318.643040), np.float32(255.313989))
DISTS = [np.array([0.262383, -0.953104, -0.005358, 0.002628], np.float32),
         np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314], np.float32),
         np.array([-0.28, 0.07, 0.0002, 0.00002], np.float32),
         np.array([0.1, -0.2, 0.001, -0.002, 0.05, 0.01, -0.02, 0.003], np.float32),
         np.array([5.0, 40.0, 0.3, 0.3], np.float32)]                       # strong enough to hit the icdist < 0 branch


@pytest.fixture(scope="module")
def pkg():
    import orbslam_jpminipc_b200 as p
    return p


@pytest.fixture(scope="module")
def po():
    from oracle import pyoracle
    return pyoracle


@pytest.mark.parametrize("shape", [(97, 131), (480, 640), (33, 7), (1, 1)])
def test_cvt_gray_vs_oracle(pkg, po, shape):
    rng = np.random.default_rng(shape[0])
    ex = pkg.ORBextractor(300, max_width=640, max_height=480, max_batch=2)
    batch = rng.integers(0, 256, (3,) + shape + (3,), dtype=np.uint8)
    for order, o in (("RGB", 0), ("BGR", 1)):
        got = ex.cvt_gray(batch, order)
        for i in range(3):
            assert np.array_equal(got[i], po.cvt_gray(batch[i], o))
    wide = rng.integers(0, 256, (shape[0], shape[1] + 5, 3), dtype=np.uint8)[:, 2:2 + shape[1]]     # unaligned rows, padded stride
    assert np.array_equal(ex.cvt_gray(np.ascontiguousarray(wide), "RGB"), po.cvt_gray(np.ascontiguousarray(wide), 0))


def test_extract_color_equals_extract_of_gray(pkg, po):
    from orbslam_jpminipc_b200.synth import synth_frame
    ex = pkg.ORBextractor(500, max_width=640, max_height=480, max_batch=2)
    r, g, b = (synth_frame(480, 640, s) for s in (61, 62, 63))
    rgb = np.ascontiguousarray(np.stack([r, g, b], -1))
    for order, o in (("RGB", 0), ("BGR", 1)):
        kps, desc = ex.extract_color(rgb, order)
        rk, rd = po.OracleExtractor(500)(po.cvt_gray(rgb, o))
        assert len(kps) == len(rk) > 100
        assert np.array_equal(kps.view(np.uint8), rk.view(np.uint8)) and np.array_equal(desc, rd)


@pytest.mark.parametrize("di", range(len(DISTS)))
def test_undistort_keypoints_and_bounds_vs_oracle(pkg, po, di):
    rng = np.random.default_rng(di)
    ex = pkg.ORBextractor(300, max_width=640, max_height=480, max_batch=1)
    n = 3000
    kps = np.zeros(n, pkg.KP_DTYPE)
    kps["x"] = rng.uniform(0, 640, n).astype(np.float32); kps["y"] = rng.uniform(0, 480, n).astype(np.float32)
    kps["size"] = 31; kps["angle"] = rng.uniform(0, 360, n).astype(np.float32); kps["octave"] = rng.integers(0, 8, n)
    got = ex.undistort_keypoints(kps, K, DISTS[di])
    ref = po.undistort_keypoints(kps, K, DISTS[di])
    assert np.array_equal(got.view(np.uint8), ref.view(np.uint8))
    assert np.array_equal(ex.image_bounds(640, 480, K, DISTS[di]), po.image_bounds(640, 480, K, DISTS[di]))


def test_zero_distortion_and_empty(pkg, po):
    ex = pkg.ORBextractor(300, max_width=640, max_height=480, max_batch=1)
    kps = np.zeros(5, pkg.KP_DTYPE); kps["x"] = np.arange(5); kps["y"] = 9
    z = np.zeros(4, np.float32)
    assert np.array_equal(ex.undistort_keypoints(kps, K, z).view(np.uint8), kps.view(np.uint8))
    assert list(ex.image_bounds(752, 480, K, z)) == [0, 752, 0, 480]
    assert len(ex.undistort_keypoints(kps[:0], K, DISTS[0])) == 0
    with pytest.raises(pkg.OrbError):
        ex.undistort_keypoints(kps, K, np.zeros(3, np.float32) + 1)          # 3 coefficients: not a size cv::undistortPoints accepts


def test_distorted_camera_frame_pair(pkg, po):
    """Frame::Frame on a distorted camera (src/Frame.cc:56-128): undistorted keypoints + undistorted bounds feed the grid and
    SearchByProjection exactly like the oracle's."""
    from orbslam_jpminipc_b200.synth import synth_frame, shifted_frame
    d = DISTS[2]
    ex = pkg.ORBextractor(1000, max_width=640, max_height=480, max_batch=2)
    a = synth_frame(480, 640, 71); b = shifted_frame(a, 3, 2, 72)
    (ka, da), (kb, db) = ex.extract_batch(np.stack([a, b]))
    kua, kub = ex.undistort_keypoints(ka, K, d), ex.undistort_keypoints(kb, K, d)
    bnd = ex.image_bounds(640, 480, K, d)
    m = pkg.ORBmatcher(0.9, True, extractor=ex)
    cur = pkg.Frame(m, kub, db, 640, 480, K[0], K[1], K[2], K[3], bounds=bnd)
    last = pkg.Frame(m, kua, da, 640, 480, K[0], K[1], K[2], K[3], bounds=bnd)
    ocur = po.OracleFrame(po.undistort_keypoints(kb, K, d), db, 640, 480, K[0], K[1], K[2], K[3], bounds=po.image_bounds(640, 480, K, d))
    olast = po.OracleFrame(po.undistort_keypoints(ka, K, d), da, 640, 480, K[0], K[1], K[2], K[3], bounds=po.image_bounds(640, 480, K, d))
    assert np.array_equal(cur.cell_start, ocur.cell_start) and np.array_equal(cur.cell_items[:cur.N], ocur.cell_items[:cur.N])
    rng = np.random.default_rng(5)
    z = rng.uniform(2, 10, len(ka)).astype(np.float32)
    xyz = np.stack([(kua["x"] - K[2]) / K[0] * z, (kua["y"] - K[3]) / K[1] * z, z], 1).astype(np.float32)
    T = np.eye(4, dtype=np.float32); T[:3, 3] = [0.03, 0.02, 0.01]
    has = np.ones(len(ka), np.uint8); outl = np.zeros(len(ka), np.uint8)
    n, match = m.SearchByProjection(cur, last, 15.0, has, outl, xyz, T)
    rn, rmatch = po.search_by_projection(ocur, olast, has, outl, xyz, T, 15.0, True)
    assert n == rn > 50 and np.array_equal(match, rmatch)
