"""Multi-GPU plumbing (one process per GPU, torch.distributed).

The path shards two ways:
  * extraction / per-frame-pair matching: frames are independent -> contiguous blocks of frames per
    rank, no collective on the data path;
  * relocalisation-sized kNN: the descriptor DB is split by contiguous row ranges, queries are
    replicated, each rank computes (idx1, d1, d2) over its rows with global row indices, ONE
    all-gather of nq*12 bytes per rank follows and the exact merge kernel (k_knn2_merge) runs on
    every rank.  best = lexicographic min of (d1, idx1); second = 2nd smallest of the union.
Device-agnostic helpers (shard ranges, the gather layout) are exercised on CPU with gloo in
tests/test_sharding_gloo.py; the compute calls need the CUDA library.
"""
import ctypes as C

import torch
import torch.distributed as dist

from ._lib import check, lib, ptr


def shard_range(n, rank, world):
    """contiguous [lo, hi) block of `n` items for `rank`; blocks differ by at most one item"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_partials(part, group=None):
    """part: int32 tensor [3, nq] = (idx1, d1, d2) of this rank's shard -> [world, 3, nq], rank-major
    (ascending global row ranges, which is what the merge's first-index-wins rule relies on)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return part.unsqueeze(0).contiguous()
    flat = torch.empty((world * part.shape[0],) + tuple(part.shape[1:]), dtype=part.dtype, device=part.device)
    dist.all_gather_into_tensor(flat, part.contiguous(), group=group)       # concatenation along dim 0, rank-major
    return flat.view((world,) + tuple(part.shape))


def knn2_sharded(extractor, d_q, d_db_shard, row_base, group=None, stream=None):
    """d_q: [nq,32] uint8 CUDA tensor (replicated), d_db_shard: [rows,32] uint8 CUDA tensor (this rank's rows,
    global index = row_base + local).  Returns (idx1, d1, d2) int32 CUDA tensors, identical on every rank."""
    L = lib()
    nq = d_q.shape[0]
    st = torch.cuda.current_stream().cuda_stream if stream is None else stream
    part = torch.empty((3, nq), dtype=torch.int32, device=d_q.device)
    check(L.orb_hamming_knn2_device(extractor._h, ptr(d_q), nq, ptr(d_db_shard), d_db_shard.shape[0], 1, int(row_base),
                                    C.c_void_p(part.data_ptr()), C.c_void_p(part.data_ptr() + 4 * nq),
                                    C.c_void_p(part.data_ptr() + 8 * nq), C.c_void_p(st)), "orb_hamming_knn2_device")
    allp = gather_partials(part, group)
    if allp.shape[0] == 1:
        return part[0], part[1], part[2]
    out = torch.empty((3, nq), dtype=torch.int32, device=d_q.device)
    check(L.orb_knn2_merge_device(extractor._h, ptr(allp), allp.shape[0], nq, C.c_void_p(out.data_ptr()),
                                  C.c_void_p(out.data_ptr() + 4 * nq), C.c_void_p(out.data_ptr() + 8 * nq), C.c_void_p(st)),
          "orb_knn2_merge_device")
    return out[0], out[1], out[2]
