"""Multi-GPU plumbing: thin callers of the C ABI's communicator (include/orb_b200.h, "multi-GPU"; csrc/orb_comm.cu).

The path shards two ways:
  * extraction / per-frame-pair matching: frames are independent -> contiguous blocks of frames per
    rank, no collective on the data path;
  * relocalisation-sized kNN: the descriptor DB is split by contiguous row ranges, queries are
    replicated, each rank computes (idx1, d1, d2) over its rows with global row indices, ONE
    exchange of nq*12 bytes per rank follows (ncclAllGather on the rank's own stream, inside the library) and the exact merge
    kernel (k_knn2_merge) runs on every rank.  best = lexicographic min of (d1, idx1); second = 2nd smallest of the union.

Two owners of the ranks, both in the library:
  RankComm   one process per GPU (torchrun): torch.distributed only carries the 128-byte NCCL unique id to the ranks;
  LocalComm  one process, all devices (what a C++ host of the reference would use): orb_comm_init.
Device-agnostic helpers (shard ranges, the gather layout) are exercised on CPU with gloo in
tests/test_sharding_gloo.py; the compute calls need the CUDA library.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from ._lib import KP_DTYPE, OrbError, check, lib, ptr


def shard_range(n, rank, world):
    """contiguous [lo, hi) block of `n` items for `rank`; blocks differ by at most one item"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_partials(part, group=None):
    """part: int32 tensor [3, nq] = (idx1, d1, d2) of this rank's shard -> [world, 3, nq], rank-major
    (ascending global row ranges, which is what the merge's first-index-wins rule relies on).  Host-logic twin of the
    library's exchange (used by the gloo CPU test)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return part.unsqueeze(0).contiguous()
    flat = torch.empty((world * part.shape[0],) + tuple(part.shape[1:]), dtype=part.dtype, device=part.device)
    dist.all_gather_into_tensor(flat, part.contiguous(), group=group)       # concatenation along dim 0, rank-major
    return flat.view((world,) + tuple(part.shape))


class RankComm:
    """orb_comm_init_rank: this process owns one rank; the unique id is broadcast through torch.distributed (any backend)."""

    def __init__(self, extractor, group=None):
        L = lib()
        self._L, self._ex = L, extractor
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.world > 1:
            if self.rank == 0:
                buf = (C.c_ubyte * 128)()
                check(L.orb_comm_unique_id(buf), "orb_comm_unique_id")
                uid = torch.tensor(list(buf), dtype=torch.uint8)
            if dist.get_backend(group) == "nccl":
                t = uid.to(torch.device("cuda", torch.cuda.current_device()))
                dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
                uid = t.cpu()
            else:
                dist.broadcast(uid, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self._id = np.ascontiguousarray(uid.numpy())
        self._h = L.orb_comm_init_rank(extractor._h, self.world, self.rank, ptr(self._id))
        if not self._h:
            raise OrbError(-4, "orb_comm_init_rank")

    @property
    def transport(self):
        return self._L.orb_comm_transport(self._h).decode()

    def knn2_sharded(self, d_q, d_db_shard, row_base, stream=None):
        """d_q: [nq,32] uint8 CUDA tensor (replicated), d_db_shard: [rows,32] uint8 CUDA tensor (this rank's rows,
        global index = row_base + local).  Returns (idx1, d1, d2) int32 CUDA tensors, identical on every rank.
        kNN, NCCL all-gather and merge are enqueued on one stream (torch's current one by default)."""
        nq = d_q.shape[0]
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream
        out = torch.empty((3, nq), dtype=torch.int32, device=d_q.device)
        check(self._L.orb_knn2_sharded_device(self._h, ptr(d_q), nq, ptr(d_db_shard), d_db_shard.shape[0], int(row_base),
                                              C.c_void_p(out.data_ptr()), C.c_void_p(out.data_ptr() + 4 * nq),
                                              C.c_void_p(out.data_ptr() + 8 * nq), C.c_void_p(st)), "orb_knn2_sharded_device")
        return out[0], out[1], out[2]

    def close(self):
        if self._h:
            torch.cuda.synchronize()
            self._L.orb_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_rank_comms = {}


def knn2_sharded(extractor, d_q, d_db_shard, row_base, group=None, stream=None):
    """functional form: one RankComm per (context, group), created on first use (a collective: every rank must call it)"""
    key = (extractor._h, id(group))
    if key not in _rank_comms:
        _rank_comms[key] = RankComm(extractor, group)
    return _rank_comms[key].knn2_sharded(d_q, d_db_shard, row_base, stream)


class LocalComm:
    """orb_comm_init: one process drives `ngpus` devices (the reference is a single process, src/main.cc:165-212)."""

    def __init__(self, ngpus=0):
        self._L = lib()
        self._h = self._L.orb_comm_init(int(ngpus))
        if not self._h:
            raise OrbError(-4, "orb_comm_init")
        self.world = self._L.orb_comm_size(self._h)

    @property
    def transport(self):
        return self._L.orb_comm_transport(self._h).decode()

    def set_extractor(self, nfeatures=1000, scale=1.2, nlevels=8, score_type=1, fast_th=20, max_width=752, max_height=480, max_batch=64):
        check(self._L.orb_comm_set_extractor(self._h, nfeatures, scale, nlevels, score_type, fast_th, max_width, max_height, max_batch),
              "orb_comm_set_extractor")
        self.capacity = self._L.orb_keypoint_capacity(self._L.orb_comm_context(self._h, 0))

    def db_upload(self, db):
        db = np.ascontiguousarray(db, np.uint8)
        check(self._L.orb_comm_db_upload(self._h, ptr(db), len(db)), "orb_comm_db_upload")

    def db_attach(self, rank, d_rows, row_base):
        self._keep = getattr(self, "_keep", {})
        self._keep[rank] = d_rows
        check(self._L.orb_comm_db_attach(self._h, rank, ptr(d_rows), d_rows.shape[0], int(row_base)), "orb_comm_db_attach")

    def knn2(self, q):
        q = np.ascontiguousarray(q, np.uint8)
        nq = len(q)
        o = [np.empty(nq, np.int32) for _ in range(3)]
        check(self._L.orb_knn2_sharded(self._h, ptr(q), nq, ptr(o[0]), ptr(o[1]), ptr(o[2])), "orb_knn2_sharded")
        return tuple(o)

    def extract_batch(self, imgs, kps=None, desc=None, counts=None):
        """imgs: [n, h, w] uint8 (numpy, or a pinned torch tensor); returns (kps [n, cap], desc [n, cap, 32], counts [n])"""
        n, h, w = imgs.shape
        cap = self.capacity
        kps = np.zeros((n, cap), KP_DTYPE) if kps is None else kps
        desc = np.zeros((n, cap, 32), np.uint8) if desc is None else desc
        counts = np.zeros(n, np.int32) if counts is None else counts
        check(self._L.orb_extract_batch_multi(self._h, ptr(imgs), n, w, h, w, w * h, ptr(kps), ptr(desc), cap, ptr(counts)),
              "orb_extract_batch_multi")
        return kps, desc, counts

    def close(self):
        if self._h:
            self._L.orb_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
