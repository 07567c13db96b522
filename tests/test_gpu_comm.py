"""-m gpu: the multi-GPU half of the C ABI (orb_comm_*, orb_knn2_sharded*, orb_extract_batch_multi; csrc/orb_comm.cu).

Every test runs with however many GPUs the box has: with one GPU the 1-rank path of the same entry points is checked; the multi-rank
cases are skipped below 2 devices (`gpurun --gpus 2` runs them).  The database follows SURVEY.md §8d's config-5 recipe
(synth.db_rows_*): planted neighbours whose two exact copies ALWAYS sit in different shards, so "lowest global index wins, d2 == d1"
is decided by the cross-shard merge (SURVEY.md §4 test pyramid item 4: shard-merge equivalence vs single GPU)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def po():
    from oracle import pyoracle
    return pyoracle


def _db_case(ndb=60_011, nq=400, seed=11):
    from orbslam_jpminipc_b200.synth import db_queries, db_rows_np
    db = db_rows_np(np.arange(ndb), ndb, seed)
    q, planted = db_queries(nq, ndb, seed)
    return db, q, planted


@pytest.mark.parametrize("transport", ["nccl", "p2p"])
def test_local_comm_knn2_equals_single_scan(po, transport, monkeypatch):
    from orbslam_jpminipc_b200.sharding import LocalComm
    monkeypatch.setenv("ORB_COMM_TRANSPORT", transport)
    comm = LocalComm(0)                                   # every visible device
    db, q, planted = _db_case()
    comm.db_upload(db)
    i1, d1, d2 = comm.knn2(q)
    r1, rd1, rd2 = po.knn2(q, db)
    assert np.array_equal(i1, r1) and np.array_equal(d1, rd1) and np.array_equal(d2, rd2)
    ev = planted >= 0
    assert np.array_equal(i1[ev], planted[ev]) and np.array_equal(d1[ev], d2[ev])      # the lower copy wins, the upper copy is the second
    if comm.world > 1:
        assert comm.transport == transport
    # empty and ragged: fewer rows than ranks, and no rows at all
    comm.db_upload(db[:1])
    i1, d1, d2 = comm.knn2(q[:7])
    assert np.array_equal(i1, np.zeros(7, np.int32)) and np.array_equal(d2, np.full(7, 2**31 - 1, np.int32))
    comm.db_upload(db[:0])
    i1, d1, d2 = comm.knn2(q[:7])
    assert np.array_equal(i1, np.full(7, -1, np.int32)) and np.array_equal(d1, np.full(7, 2**31 - 1, np.int32))
    comm.close()


def test_local_comm_device_generated_shards(po):
    """shards generated on their own devices (db_rows_torch) and attached; result equals the host-generated database's scan"""
    import torch
    from orbslam_jpminipc_b200.sharding import LocalComm, shard_range
    from orbslam_jpminipc_b200.synth import db_queries, db_rows_np, db_rows_torch
    ndb, seed = 200_003, 5
    comm = LocalComm(0)
    for r in range(comm.world):
        lo, hi = shard_range(ndb, r, comm.world)
        comm.db_attach(r, db_rows_torch(lo, hi, ndb, seed, torch.device("cuda", r)), lo)
    torch.cuda.synchronize()
    q, planted = db_queries(300, ndb, seed)
    got = comm.knn2(q)
    ref = po.knn2(q, db_rows_np(np.arange(ndb), ndb, seed))
    assert all(np.array_equal(a, b) for a, b in zip(got, ref))
    comm.close()


def test_extract_batch_multi_equals_single_device(po):
    import orbslam_jpminipc_b200 as pkg
    from orbslam_jpminipc_b200.sharding import LocalComm
    from orbslam_jpminipc_b200.synth import synth_frames
    frames = synth_frames(11, 240, 320, 4100)             # 11 frames: ragged over 2, 4 or 8 ranks
    comm = LocalComm(0)
    comm.set_extractor(400, 1.2, 8, 1, 20, 320, 240, 3)
    kps, desc, counts = comm.extract_batch(frames)
    ex = po.OracleExtractor(400, 1.2, 8, 1, 20)
    for i, f in enumerate(frames):
        rk, rd = ex(f)
        n = counts[i]
        assert n == len(rk) and np.array_equal(kps[i, :n].view(np.uint8), rk.view(np.uint8)) and np.array_equal(desc[i, :n], rd)
    comm.close()


def _rank_worker(rank, world, port, ndb, seed, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import orbslam_jpminipc_b200 as pkg
    from orbslam_jpminipc_b200.sharding import RankComm, shard_range
    from orbslam_jpminipc_b200.synth import db_queries, db_rows_torch
    dev = torch.device("cuda", rank)
    ex = pkg.ORBextractor(500, device=rank, max_width=64, max_height=64, max_batch=1)
    comm = RankComm(ex)
    lo, hi = shard_range(ndb, rank, world)
    shard = db_rows_torch(lo, hi, ndb, seed, dev)
    q, _ = db_queries(500, ndb, seed)
    d_q = torch.from_numpy(q).to(dev)
    side = torch.cuda.Stream()                            # not torch's current stream: ordering must come from the library alone
    side.wait_stream(torch.cuda.current_stream())
    out = None
    for _ in range(3):                                    # repeated calls reuse the exchange buffers
        out = comm.knn2_sharded(d_q, shard, lo, stream=side.cuda_stream)
    side.synchronize()
    ret[rank] = tuple(o.cpu().numpy() for o in out) + (comm.transport,)
    comm.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_rank_comm_processes_equal_single_scan(po):
    """one process per GPU: NCCL all-gather + merge inside the library, result on EVERY rank equals one scan of the whole database"""
    import torch.multiprocessing as mp
    from orbslam_jpminipc_b200.synth import db_queries, db_rows_np
    world, ndb, seed = min(_ngpu(), 4), 300_007, 21
    ret = mp.Manager().dict()
    mp.spawn(_rank_worker, args=(world, 29700 + os.getpid() % 200, ndb, seed, ret), nprocs=world, join=True)
    q, planted = db_queries(500, ndb, seed)
    ref = po.knn2(q, db_rows_np(np.arange(ndb), ndb, seed))
    for r in range(world):
        assert ret[r][3] == "nccl"
        assert all(np.array_equal(a, b) for a, b in zip(ret[r][:3], ref)), "rank %d" % r


def test_cpp_multi_gpu_host_program(po, tmp_path):
    """tests/cpp/test_multi.cpp: a C++ host reaches every GPU of the box through the C ABI alone (no python, no torch)"""
    from orbslam_jpminipc_b200 import dbio
    from orbslam_jpminipc_b200.synth import synth_frames
    exe = str(tmp_path / "test_multi")
    libdir = os.path.join(ROOT, "orbslam_jpminipc_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_multi.cpp"),
                           "-L" + libdir, "-lorb_b200", "-Wl,-rpath," + libdir])
    db, q, planted = _db_case(40_009, 256, seed=3)
    dbp, qp, fp, out = (str(tmp_path / n) for n in ("db.bin", "q.raw", "f.raw", "o.bin"))
    dbio.write_descriptors(dbp, db, np.array([0, 10_000, 10_000, len(db)], np.int32))       # three keyframe records, one empty
    q.tofile(qp)
    frames = synth_frames(9, 240, 320, 5100)
    frames.tofile(fp)
    txt = subprocess.check_output([exe, "0", dbp, qp, str(len(q)), fp, "320", "240", str(len(frames)), out]).decode()
    assert ("PASS ranks=%d" % _ngpu()) in txt, txt                 # NCCL prints its version banner on stdout first
    buf = np.fromfile(out, np.int32)
    nq = len(q)
    ref = po.knn2(q, db)
    assert all(np.array_equal(buf[k * nq:(k + 1) * nq], ref[k]) for k in range(3))
    ex = po.OracleExtractor(500, 1.2, 8, 1, 20)
    assert buf[3 * nq:].tolist() == [len(ex(f)[0]) for f in frames]
