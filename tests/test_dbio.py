"""Host-side readers / writers of the fork's dump formats (reference include/SaveLoadWorld.h:1408-1459).  No GPU needed: the
expected bytes are written here with struct, straight from the layout the reference's save code produces."""
import struct

import numpy as np
import pytest


@pytest.fixture(scope="module")
def dbio():
    from orbslam_jpminipc_b200 import dbio
    return dbio


def _ref_desc_file(path, groups):
    with open(path, "wb") as f:
        for g in groups:                                   # saveHeader {0xeb,0x90}, int desnumi, rows of 32 bytes (:1450-1458)
            f.write(bytes([0xEB, 0x90])); f.write(struct.pack("<i", len(g))); f.write(g.tobytes())


def _ref_kp_file(path, groups):
    with open(path, "wb") as f:
        for g in groups:                                   # saveHeader, size_t nKeys, 5 floats + 2 ints per key (:1410-1424)
            f.write(bytes([0xEB, 0x90])); f.write(struct.pack("<Q", len(g)))
            for k in g:
                f.write(struct.pack("<fffffii", k["x"], k["y"], k["size"], k["angle"], k["response"], k["octave"], k["class_id"]))


def test_descriptor_dump_round_trip(dbio, tmp_path):
    rng = np.random.default_rng(1)
    groups = [rng.integers(0, 256, (n, 32), dtype=np.uint8) for n in (5, 0, 1000, 1)]
    p = tmp_path / "des.bin"
    _ref_desc_file(p, groups)
    rows, start = dbio.read_descriptors(p)
    assert list(start) == [0, 5, 5, 1005, 1006] and np.array_equal(rows, np.concatenate(groups))
    q = tmp_path / "des2.bin"
    dbio.write_descriptors(q, rows, start)
    assert open(p, "rb").read() == open(q, "rb").read()


def test_keypoint_dump_round_trip(dbio, tmp_path):
    from orbslam_jpminipc_b200 import KP_DTYPE
    rng = np.random.default_rng(2)
    groups = []
    for n in (3, 700, 0):
        k = np.zeros(n, KP_DTYPE)
        for name in ("x", "y", "size", "angle", "response"):
            k[name] = rng.uniform(0, 500, n).astype(np.float32)
        k["octave"] = rng.integers(0, 8, n); k["class_id"] = -1
        groups.append(k)
    p = tmp_path / "keys.bin"
    _ref_kp_file(p, groups)
    rows, start = dbio.read_keypoints(p)
    assert list(start) == [0, 3, 703, 703] and np.array_equal(rows.view(np.uint8), np.concatenate(groups).view(np.uint8))
    q = tmp_path / "keys2.bin"
    dbio.write_keypoints(q, rows, start)
    assert open(p, "rb").read() == open(q, "rb").read()


def test_malformed_and_missing(dbio, tmp_path):
    from orbslam_jpminipc_b200 import OrbError
    bad = tmp_path / "bad.bin"
    bad.write_bytes(bytes([0xEB, 0x91, 1, 0, 0, 0]) + bytes(32))
    with pytest.raises(OrbError):
        dbio.read_descriptors(bad)
    trunc = tmp_path / "trunc.bin"
    trunc.write_bytes(bytes([0xEB, 0x90]) + struct.pack("<i", 4) + bytes(40))
    with pytest.raises(OrbError):
        dbio.read_descriptors(trunc)
    with pytest.raises(OrbError):
        dbio.read_descriptors(tmp_path / "missing.bin")
    empty = tmp_path / "empty.bin"
    empty.write_bytes(b"")
    rows, start = dbio.read_descriptors(empty)
    assert len(rows) == 0 and list(start) == [0]
