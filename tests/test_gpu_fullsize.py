"""GPU tests at BASELINE.json's full sizes, through size-independent properties (the oracle would take minutes here):
planted-neighbour recovery, recomputed distances, duplicate tie-breaking, shard-merge equivalence, batch == single."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import orbslam_jpminipc_b200 as p
    return p


def _popcount_rows(a, b):
    return np.unpackbits(a ^ b, axis=1).sum(1).astype(np.int32)


def test_knn2_config5_10m_rows_properties(pkg):
    """config 5: 2000 queries vs a 10M-row DB (320 MB) on one GPU, and the same DB as 8 emulated shards + merge."""
    import torch
    from orbslam_jpminipc_b200._lib import check, lib, ptr
    from orbslam_jpminipc_b200.sharding import shard_range
    dev = torch.device("cuda", 0)
    ND, NQ = 10_000_000, 2000
    g = torch.Generator(device=dev)
    g.manual_seed(42)
    d_db = torch.randint(0, 256, (ND, 32), dtype=torch.uint8, device=dev, generator=g)
    rng = np.random.default_rng(43)
    # planted neighbours for even queries, exact duplicates of the planted rows further down the DB (lowest index must win)
    rows = np.sort(rng.choice(ND // 2, NQ // 2, replace=False)).astype(np.int64)
    dup = rows + ND // 2
    d_db[torch.from_numpy(dup).to(dev)] = d_db[torch.from_numpy(rows).to(dev)]
    planted = d_db[torch.from_numpy(rows).to(dev)].cpu().numpy()
    flips = np.packbits((rng.random((NQ // 2, 256)) < 0.08).astype(np.uint8), axis=1)
    q = rng.integers(0, 256, (NQ, 32), dtype=np.uint8)
    q[0::2] = planted ^ flips
    d_q = torch.from_numpy(q).to(dev)
    m = pkg.ORBmatcher(0.6, True)
    L = lib()
    st = torch.cuda.current_stream().cuda_stream
    out = torch.zeros((3, NQ), dtype=torch.int32, device=dev)
    check(L.orb_hamming_knn2_device(m._h, ptr(d_q), NQ, ptr(d_db), ND, 1, 0, C.c_void_p(out.data_ptr()),
                                    C.c_void_p(out.data_ptr() + 4 * NQ), C.c_void_p(out.data_ptr() + 8 * NQ), C.c_void_p(st)), "knn2")
    torch.cuda.synchronize()
    idx1, d1, d2 = out.cpu().numpy()
    # (a) planted rows are recovered at their FIRST occurrence with the planted distance, duplicate => d2 == d1
    nflip = np.unpackbits(flips, axis=1).sum(1)
    assert np.array_equal(idx1[0::2], rows) and np.array_equal(d1[0::2], nflip) and np.array_equal(d2[0::2], nflip)
    # (b) every reported distance equals the distance recomputed from the reported row; d1 <= d2; random queries are far
    best_rows = d_db[torch.from_numpy(idx1.astype(np.int64)).to(dev)].cpu().numpy()
    assert np.array_equal(_popcount_rows(q, best_rows), d1) and np.all(d1 <= d2) and d1[1::2].min() > 50
    # (c) 8 shards + exact merge == one scan
    world = 8
    allp = torch.zeros((world, 3, NQ), dtype=torch.int32, device=dev)
    for r in range(world):
        lo, hi = shard_range(ND, r, world)
        p = allp[r]
        check(L.orb_hamming_knn2_device(m._h, ptr(d_q), NQ, C.c_void_p(d_db.data_ptr() + lo * 32), hi - lo, 1, lo,
                                        C.c_void_p(p.data_ptr()), C.c_void_p(p.data_ptr() + 4 * NQ), C.c_void_p(p.data_ptr() + 8 * NQ),
                                        C.c_void_p(st)), "shard")
    mg = torch.zeros((3, NQ), dtype=torch.int32, device=dev)
    check(L.orb_knn2_merge_device(m._h, ptr(allp), world, NQ, C.c_void_p(mg.data_ptr()), C.c_void_p(mg.data_ptr() + 4 * NQ),
                                  C.c_void_p(mg.data_ptr() + 8 * NQ), C.c_void_p(st)), "merge")
    torch.cuda.synchronize()
    assert torch.equal(mg, out)


def test_extract_config2_batch_properties(pkg):
    """config 2: a 128-frame batch of 752x480 frames: deterministic, batch == single-frame calls, counts and bounds."""
    from orbslam_jpminipc_b200.synth import synth_frames
    base = synth_frames(8, 480, 752, 1000)
    frames = np.concatenate([base] * 16)                     # 128 frames, every frame appears 16 times
    ex = pkg.ORBextractor(1000, 1.2, 8, 1, 20, max_width=752, max_height=480, max_batch=48)     # 3 chunks, ragged last one
    res = ex.extract_batch(frames)
    res2 = ex.extract_batch(frames)
    one = pkg.ORBextractor(1000, 1.2, 8, 1, 20, max_width=752, max_height=480, max_batch=1)
    singles = [one(f) for f in base]
    for i, (k, d) in enumerate(res):
        assert len(k) == 1000 and d.shape == (1000, 32)
        assert np.array_equal(k.view(np.uint8), res2[i][0].view(np.uint8)) and np.array_equal(d, res2[i][1])        # idempotent
        sk, sd = singles[i % 8]
        assert np.array_equal(k.view(np.uint8), sk.view(np.uint8)) and np.array_equal(d, sd)                      # batch == single
        assert k["x"].min() >= 16 and k["x"].max() < 752 and k["y"].min() >= 16 and k["y"].max() < 480
        assert np.array_equal(np.bincount(k["octave"], minlength=8), [217, 181, 151, 126, 105, 87, 73, 60])
        assert np.all(np.diff(k["octave"]) >= 0)                                                                  # level-major order
    ex.close(); one.close()


def test_extract_config3_kitti_shape_and_tracking(pkg):
    """config 3: 1241x376, 2000 kp: extraction + frame-to-frame SearchByProjection; translation-consistent matches."""
    from orbslam_jpminipc_b200.synth import synth_frame, shifted_frame
    h, w = 376, 1241
    a = synth_frame(h, w, 9100, quadrants=False)
    b = shifted_frame(a, 3, 2, 9101)
    ex = pkg.ORBextractor(2000, 1.2, 8, 1, 20, max_width=w, max_height=h, max_batch=2)
    (ka, da), (kb, db) = ex.extract_batch(np.stack([a, b]))
    assert len(ka) == 2000 and len(kb) == 2000
    m = pkg.ORBmatcher(0.9, True, extractor=ex)
    fx = fy = 500.0
    z = np.full(len(ka), 5.0, np.float32)
    xyz = np.stack([(ka["x"] - w / 2) / fx * z, (ka["y"] - h / 2) / fy * z, z], 1).astype(np.float32)
    T = np.eye(4, dtype=np.float32)
    cur, last = pkg.Frame(m, kb, db, w, h, fx, fy, w / 2, h / 2), pkg.Frame(m, ka, da, w, h, fx, fy, w / 2, h / 2)
    n, match = m.SearchByProjection(cur, last, 15.0, np.ones(len(ka), np.uint8), np.zeros(len(ka), np.uint8), xyz, T)
    assert n > 800 and n == int((match >= 0).sum())
    i2 = np.nonzero(match >= 0)[0]
    assert len(np.unique(match[i2])) == len(i2)                              # a map point is assigned at most once
    dx = kb["x"][i2] - ka["x"][match[i2]]
    dy = kb["y"][i2] - ka["y"][match[i2]]
    scale = 1.2 ** kb["octave"][i2]
    assert np.median(np.abs(dx - 3) / scale) < 1.5 and np.median(np.abs(dy - 2) / scale) < 1.5   # the planted (3,2) shift
    ex.close()
