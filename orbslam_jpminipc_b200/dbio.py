"""Keypoint / descriptor dump files of the fork (reference include/SaveLoadWorld.h:1408-1459) as database files for the
sharded Hamming search: thin wrappers over the C ABI readers / writers (host-side file I/O)."""
import ctypes as C

import numpy as np

from ._lib import KP_DTYPE, check, lib, ptr


def _read(fn, path, dtype, shape_tail):
    n, r = C.c_int64(0), C.c_int32(0)
    check(fn(str(path).encode(), None, 0, None, 0, C.byref(n), C.byref(r)), "db read (sizing)")
    rows = np.zeros((n.value,) + shape_tail, dtype)
    start = np.zeros(r.value + 1, np.int32)
    check(fn(str(path).encode(), ptr(rows) if n.value else None, n.value, ptr(start), r.value + 1, C.byref(n), C.byref(r)), "db read")
    return rows, start


def read_descriptors(path):
    """-> (descriptors N x 32 uint8, rec_start: row offsets of the keyframes)"""
    return _read(lib().orb_db_read_descriptors, path, np.uint8, (32,))


def read_keypoints(path):
    return _read(lib().orb_db_read_keypoints, path, KP_DTYPE, ())


def write_descriptors(path, desc, rec_start):
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    start = np.ascontiguousarray(rec_start, np.int32)
    check(lib().orb_db_write_descriptors(str(path).encode(), ptr(desc) if len(desc) else None, ptr(start), len(start) - 1), "db write")


def write_keypoints(path, kps, rec_start):
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    start = np.ascontiguousarray(rec_start, np.int32)
    check(lib().orb_db_write_keypoints(str(path).encode(), ptr(kps) if len(kps) else None, ptr(start), len(start) - 1), "db write")
