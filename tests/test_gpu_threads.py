"""-m gpu: the concurrency contract of a context (INTEGRATION.md "Threads").

The reference calls ORBmatcher from three threads (src/main.cc:165 Tracking, :182 LocalMapping, :193 LoopClosing), each with a matcher built
on the stack.  Here the three threads SHARE one context (one extractor + every matcher built on it): matcher / vocabulary / frame
calls borrow a lane (own stream + scratch) per call, extraction calls are serialised by the context's mutex.  200 iterations per
thread, every result compared with the serial run of the same call."""
import threading

import numpy as np
import pytest

import test_gpu_match as tm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import orbslam_jpminipc_b200 as p
    return p


@pytest.fixture(scope="module")
def po():
    from oracle import pyoracle
    return pyoracle


def test_three_threads_share_one_context(pkg, po):
    from orbslam_jpminipc_b200.synth import synth_frames
    ITER = 200
    h, w = 240, 320
    ex = pkg.ORBextractor(500, 1.2, 8, 1, 20, max_width=w, max_height=h, max_batch=2)
    frames = synth_frames(4, h, w, 6100)
    # --- tracking thread: extract + SearchByProjection(Frame, Frame)
    m_track = pkg.ORBmatcher(0.9, True, extractor=ex)
    gcur, glast, ocur, olast, has, outl, xyz, T = tm._scene(po, pkg, m_track, h, w, 500, 7000, 15.0)
    # --- local mapping thread: SearchForTriangulation + the Fuse scoring loop
    m_map = pkg.ORBmatcher(0.6, True, extractor=ex)
    fv1, d1, k1, v1, fv2, d2, k2 = tm._bow_case(po, pkg, 700, 500, 12, seed=42, flip=0.05)
    rng = np.random.default_rng(2)
    k1["x"] = rng.uniform(20, 600, len(k1)).astype(np.float32); k1["y"] = rng.uniform(20, 440, len(k1)).astype(np.float32)
    k2["x"] = rng.uniform(20, 600, len(k2)).astype(np.float32); k2["y"] = rng.uniform(20, 440, len(k2)).astype(np.float32)
    F12 = np.array([[0, 0, 4], [0, 0, -6], [-4, 6, 0]], np.float32)
    sg = np.array([np.float32(np.float32(1.2) ** i) ** 2 for i in range(8)], np.float32)
    has1 = (rng.random(len(k1)) < 0.3).astype(np.uint8); has2 = (rng.random(len(k2)) < 0.3).astype(np.uint8)
    a1, u1, vv1, l1 = tm._projected_points(rng, glast.kps, w, h)
    # --- loop closing thread: SearchByBoW(KF, KF) + brute-force kNN
    m_loop = pkg.ORBmatcher(0.75, True, extractor=ex)
    case = tm._bow_case(po, pkg, 600, 500, 12, seed=77)
    kfv, kd, kk, kvalid, ffv, fd, fk = case
    fvalid = (np.random.default_rng(5).random(len(fd)) < 0.8).astype(np.uint8)
    from orbslam_jpminipc_b200.synth import synth_descriptors
    db, q = synth_descriptors(30000, 300)

    def track(i):
        k, d = ex(frames[i % len(frames)])
        n, match = m_track.SearchByProjection(gcur, glast, 15.0, has, outl, xyz, T)
        return (k.tobytes(), d.tobytes(), n, match.tobytes())

    def mapping(i):
        n, pairs, m12 = m_map.SearchForTriangulation(fv1, d1, k1, has1, fv2, d2, k2, has2, F12, sg)
        fused = m_map.FuseCandidates(gcur, a1, u1, vv1, l1, glast.desc, 2.5)
        return (n, m12.tobytes(), fused.tobytes())

    def loop(i):
        n, m12 = m_loop.SearchByBoWKeyFrames(kfv, kd, kk, kvalid, ffv, fd, fk, fvalid)
        i1, dd1, dd2 = m_loop.knn2(q, db)
        return (n, m12.tobytes(), i1.tobytes(), dd1.tobytes(), dd2.tobytes())

    jobs = [track, mapping, loop]
    serial = [[job(i) for i in range(len(frames))] for job in jobs]          # the expected answers, one caller at a time
    # the serial answers are themselves right (oracle)
    rn, rmatch = po.search_by_projection(ocur, olast, has, outl, xyz, T, 15.0, True)
    assert serial[0][0][2] == rn and serial[0][0][3] == np.asarray(rmatch, np.int32).tobytes()
    r = po.knn2(q, db)
    assert serial[2][0][2] == r[0].tobytes() and serial[2][0][3] == r[1].tobytes()
    errors = []
    go = threading.Barrier(3)

    def run(j):
        try:
            go.wait()
            for i in range(ITER):
                if jobs[j](i) != serial[j][i % len(frames)]:
                    errors.append("thread %d iteration %d differs from the serial run" % (j, i))
                    return
        except Exception as e:                                             # noqa: BLE001
            errors.append("thread %d: %r" % (j, e))
    th = [threading.Thread(target=run, args=(j,)) for j in range(3)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errors, errors


def test_two_streams_one_context_are_ordered(pkg, po):
    """two device-pointer calls on different caller streams share the context's work buffers: the library orders them"""
    import ctypes as C
    import torch
    from orbslam_jpminipc_b200._lib import check, lib, ptr
    from orbslam_jpminipc_b200.synth import synth_frames
    L = lib()
    h, w, B = 240, 320, 8
    ex = pkg.ORBextractor(400, 1.2, 8, 1, 20, max_width=w, max_height=h, max_batch=B)
    cap = ex.capacity
    dev = torch.device("cuda", 0)
    fa, fb = synth_frames(B, h, w, 9100), synth_frames(B, h, w, 9200)
    d_a, d_b = torch.from_numpy(fa).to(dev), torch.from_numpy(fb).to(dev)
    outs = [[torch.zeros((B, cap, 7), dtype=torch.int32, device=dev), torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev),
             torch.zeros(B, dtype=torch.int32, device=dev)] for _ in range(2)]
    s = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for rep in range(20):
        for k, src in enumerate((d_a, d_b)):
            check(L.orb_extract_batch_device(ex._h, ptr(src), B, w, h, w, w * h, ptr(outs[k][0]), ptr(outs[k][1]), cap, ptr(outs[k][2]),
                                             C.c_void_p(s[k].cuda_stream)), "orb_extract_batch_device")
    torch.cuda.synchronize()
    orc = po.OracleExtractor(400, 1.2, 8, 1, 20)
    for k, fr in enumerate((fa, fb)):
        cnt = outs[k][2].cpu().numpy()
        kp = outs[k][0].cpu().numpy().view(np.uint8).reshape(B, cap, 28)
        de = outs[k][1].cpu().numpy()
        for i in (0, B - 1):
            rk, rd = orc(fr[i])
            assert cnt[i] == len(rk) and np.array_equal(kp[i, :cnt[i]].reshape(-1), rk.view(np.uint8).reshape(-1)) and np.array_equal(de[i, :cnt[i]], rd)
