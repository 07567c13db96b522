"""CPU, world_size 2, gloo: the multi-GPU host logic (shard ranges, partial-result gather layout,
exactness of the best-two merge, frame sharding) with the oracle standing in for the GPU kernels."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as po
    from orbslam_jpminipc_b200.sharding import gather_partials, shard_range
    from orbslam_jpminipc_b200.synth import synth_descriptors, synth_frames
    # ---- DB-sharded kNN: shard, local best-two with global indices, gather, merge
    db, q = synth_descriptors(5001, 300, dup_frac=0.02)
    lo, hi = shard_range(len(db), rank, world)
    i1, d1, d2 = po.knn2(q, db[lo:hi])
    i1 = np.where(i1 >= 0, i1 + lo, -1).astype(np.int32)
    part = torch.from_numpy(np.stack([i1, d1, d2]))
    allp = gather_partials(part)
    assert tuple(allp.shape) == (world, 3, 300)
    merged = po.merge_best2(allp.numpy())
    full = po.knn2(q, db)
    ok_knn = all(np.array_equal(a, b) for a, b in zip(merged, full))
    # ---- frame sharding: every rank extracts its block, counts gathered on every rank
    frames = synth_frames(5, 120, 160, seed0=300)
    flo, fhi = shard_range(len(frames), rank, world)
    ex = po.OracleExtractor(150)
    mine = torch.zeros(len(frames), dtype=torch.int64)
    for f in range(flo, fhi):
        mine[f] = len(ex(frames[f])[0])
    dist.all_reduce(mine)
    single = [len(ex(f)[0]) for f in frames]
    ok_frames = mine.tolist() == single
    ret[rank] = (ok_knn, ok_frames, (lo, hi))
    dist.destroy_process_group()


def test_two_rank_sharding_and_merge():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0][0] and ret[1][0], "merged best-two differs from the single-shard scan"
    assert ret[0][1] and ret[1][1], "frame-sharded extraction differs from the single-process run"
    assert ret[0][2] == (0, 2501) and ret[1][2] == (2501, 5001)


def test_shard_ranges_cover_exactly():
    from orbslam_jpminipc_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 1000, 10_000_000):
        for world in (1, 2, 3, 4, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_merge_tie_rules():
    from oracle import pyoracle as po
    # two shards with the same best distance: the lower shard's (lower global) index wins, second == best
    parts = np.array([[[5], [10], [30]], [[105], [10], [12]]], np.int32)
    i, d1, d2 = po.merge_best2(parts)
    assert (i[0], d1[0], d2[0]) == (5, 10, 10)
    parts = np.array([[[-1], [2**31 - 1], [2**31 - 1]], [[7], [40], [41]]], np.int32)     # empty first shard
    i, d1, d2 = po.merge_best2(parts)
    assert (i[0], d1[0], d2[0]) == (7, 40, 41)
