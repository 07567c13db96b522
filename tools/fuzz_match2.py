"""Second fuzz sweep of matcher-side entry points against the CPU oracle: SearchByBoW (keyframe-frame, keyframe-keyframe),
SearchForTriangulation, ComputeDistinctiveDescriptors, the window-best / Sim3 / Fuse candidate searches, colour conversion and
keypoint undistortion:  gpurun -- 'python tools/fuzz_match2.py 0 200'"""
import sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import orbslam_jpminipc_b200 as pkg
from oracle import pyoracle as po
import test_gpu_match as T

bad = 0; n = 0; t0 = time.time()
def check(ok, what):
    global bad, n
    n += 1
    if not ok:
        bad += 1; print("MISMATCH", what)

ex = pkg.ORBextractor(300, max_width=640, max_height=480, max_batch=1)
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(40000 + seed)
    ori = bool(rng.integers(0, 2)); ratio = float(rng.choice([0.6, 0.75, 0.9]))
    m = pkg.ORBmatcher(ratio, ori)
    # ---- SearchByBoW (KF, F) and (KF, KF)
    n_kf, n_f, nnodes = int(rng.integers(1, 2500)), int(rng.integers(1, 2500)), int(rng.choice([1, 2, 7, 40, 100, 400]))
    case = T._bow_case(po, pkg, n_kf, n_f, nnodes, seed=seed, flip=float(rng.choice([0.0, 0.03, 0.08, 0.2])))
    g = m.SearchByBoW(*case); r = po.search_by_bow(*case, ratio, ori)
    check(g[0] == r[0] and np.array_equal(g[1], r[1]), ("SearchByBoW", seed, n_kf, n_f, nnodes, ori, ratio))
    fv1, d1, k1, v1, fv2, d2, k2 = case
    v2 = (rng.random(len(k2)) < 0.7).astype(np.uint8)
    g = m.SearchByBoWKeyFrames(fv1, d1, k1, v1, fv2, d2, k2, v2); r = po.search_by_bow_kf(fv1, d1, k1, v1, fv2, d2, k2, v2, ratio, ori)
    check(g[0] == r[0] and np.array_equal(g[1], r[1]), ("SearchByBoWKeyFrames", seed, n_kf, n_f, nnodes, ori, ratio))
    # ---- SearchForTriangulation on the same node structure with an epipolar geometry
    k1 = k1.copy(); k2 = k2.copy()
    k1["x"] = rng.uniform(20, 600, len(k1)).astype(np.float32); k1["y"] = rng.uniform(20, 440, len(k1)).astype(np.float32)
    k1["octave"] = rng.integers(0, 8, len(k1)); k2["octave"] = rng.integers(0, 8, len(k2))
    tw = rng.integers(0, len(k1), len(k2)); lam = rng.uniform(0.5, 3.0, len(k2))
    k2["x"] = (k1["x"][tw] + 6 * lam + rng.normal(0, 1.0, len(k2))).astype(np.float32)
    k2["y"] = (k1["y"][tw] + 4 * lam + rng.normal(0, 1.0, len(k2))).astype(np.float32)
    F12 = np.array([[0, 0, 4.0], [0, 0, -6.0], [-4.0, 6.0, 0]], np.float32)
    sg = np.ones(8, np.float32)
    for i in range(1, 8): sg[i] = np.float32(np.float32(1.2) ** i) ** 2
    h1 = (rng.random(len(k1)) < 0.3).astype(np.uint8); h2 = (rng.random(len(k2)) < 0.3).astype(np.uint8)
    g = m.SearchForTriangulation(fv1, d1, k1, h1, fv2, d2, k2, h2, F12, sg); r = po.search_for_triangulation(fv1, d1, k1, h1, fv2, d2, k2, h2, F12, sg, ori)
    check(g[0] == r[0] and np.array_equal(g[2], r[1]), ("SearchForTriangulation", seed, n_kf, n_f, nnodes, ori, ratio))
    # ---- ComputeDistinctiveDescriptors
    npts, maxobs = int(rng.integers(1, 1500)), int(rng.choice([1, 2, 5, 12, 70, 300]))
    sizes = rng.integers(0, maxobs + 1, npts)
    start = np.zeros(npts + 1, np.int32); start[1:] = np.cumsum(sizes)
    base = rng.integers(0, 256, (npts, 32), dtype=np.uint8)
    desc = np.zeros((start[-1], 32), np.uint8)
    for p in range(npts):
        if sizes[p]:
            bits = np.unpackbits(np.repeat(base[p][None], sizes[p], 0), axis=1)
            bits ^= rng.random(bits.shape) < rng.uniform(0.0, 0.2)
            d = np.packbits(bits, axis=1)
            if sizes[p] > 3: d[sizes[p] // 2] = d[0]
            desc[start[p]:start[p + 1]] = d
    g = m.ComputeDistinctiveDescriptors(desc, start); r = po.distinctive_descriptors(desc, start)
    check(np.array_equal(g[0], r[0]) and np.array_equal(g[1], r[1]), ("Distinctive", seed, npts, maxobs))
    # ---- window best / Sim3 / Fuse on a frame pair
    h, w, nf = int(rng.integers(140, 480)), int(rng.integers(200, 752)), int(rng.integers(200, 1500))
    try:
        gcur, glast, ocur, olast, has, outl, xyz, Tm = T._scene(po, pkg, m, h, w, nf, 100 + seed, 15.0)
    except RuntimeError:
        continue
    a1, u1, v1_, l1 = T._projected_points(rng, glast.kps, w, h, jitter=float(rng.choice([1.0, 3.0, 8.0])))
    th = float(rng.choice([2.5, 4.0, 7.5, 10.0]))
    pre = np.full(gcur.N, -1, np.int32); pre[::int(rng.integers(3, 12))] = 1 << 20
    g = m.SearchByProjectionSim3(gcur, a1, u1, v1_, l1, glast.desc, th, pre.copy()); r = po.search_by_projection_sim3(ocur, a1, u1, v1_, l1, glast.desc, th, pre.copy())
    check(g[0] == r[0] and np.array_equal(g[1], r[1]), ("SearchByProjectionSim3", seed, h, w, nf, th))
    sf = np.ones(8, np.float32)
    for i in range(1, 8): sf[i] = np.float32(sf[i - 1] * np.float32(1.2))
    fused = m.FuseCandidates(gcur, a1, u1, v1_, l1, glast.desc, th)
    rbi, rbd = po.window_best(ocur, a1, u1, v1_, np.float32(th) * sf[l1], l1, glast.desc)
    check(np.array_equal(fused, np.where(rbd <= 50, rbi, -1)), ("FuseCandidates", seed, h, w, nf, th))
    # ---- frame plumbing
    hh, ww = int(rng.integers(1, 300)), int(rng.integers(1, 400))
    rgb = rng.integers(0, 256, (hh, ww, 3), dtype=np.uint8)
    for order, o in (("RGB", 0), ("BGR", 1)):
        check(np.array_equal(ex.cvt_gray(rgb, order), po.cvt_gray(rgb, o)), ("cvt_gray", seed, hh, ww, order))
    K = (np.float32(rng.uniform(300, 700)), np.float32(rng.uniform(300, 700)), np.float32(rng.uniform(250, 400)), np.float32(rng.uniform(180, 300)))   # fx, fy, cx, cy
    dist = np.array([rng.uniform(-0.4, 0.3), rng.uniform(-0.1, 0.2), rng.uniform(-0.002, 0.002), rng.uniform(-0.002, 0.002)], np.float32)
    if rng.random() < 0.3: dist = np.concatenate([dist, [np.float32(rng.uniform(-0.05, 0.05))]]).astype(np.float32)
    kk = np.zeros(500, pkg.KP_DTYPE); kk["x"] = rng.uniform(0, 640, 500).astype(np.float32); kk["y"] = rng.uniform(0, 480, 500).astype(np.float32)
    check(np.array_equal(ex.undistort_keypoints(kk, K, dist).view(np.uint8), po.undistort_keypoints(kk, K, dist).view(np.uint8)), ("undistort", seed, list(dist)))
    check(np.array_equal(ex.image_bounds(640, 480, K, dist), po.image_bounds(640, 480, K, dist)), ("bounds", seed, list(dist)))
print("checks", n, "bad", bad, "%.1f s" % (time.time() - t0))
