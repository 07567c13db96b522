"""CPU checks of the drop-in boundary: liborb_b200.so loads, exports every symbol that
include/orb_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(ROOT, "orbslam_jpminipc_b200", "liborb_b200.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "orbslam_jpminipc_b200", "csrc"), "-s"])
    import orbslam_jpminipc_b200 as pkg
    return pkg.lib()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "orb_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(orb_[a-z0-9_]+)\s*\(", hdr))
    from orbslam_jpminipc_b200._lib import EXPORTS
    assert names == set(EXPORTS), names ^ set(EXPORTS)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.orb_abi_version() == 1


def test_keypoint_record_is_cv_keypoint_sized():
    from orbslam_jpminipc_b200 import KP_DTYPE
    assert KP_DTYPE.itemsize == 28 and KP_DTYPE.fields["octave"][1] == 20


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = lib.orb_create(0, 1000, ctypes.c_float(1.2), 8, 1, 20, 640, 480, 4)
    assert not h
    assert b"no usable CUDA device" in lib.orb_last_cuda_error()
    import orbslam_jpminipc_b200 as pkg
    with pytest.raises(RuntimeError):
        pkg.ORBextractor(1000)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "orbslam_jpminipc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in src and "orb_oracle" not in src and "liborb_oracle" not in src, f


def test_introselect_matches_libstdcxx(tmp_path):
    exe = str(tmp_path / "t")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_introselect.cpp")])
    out = subprocess.check_output([exe]).decode()
    assert out.startswith("PASS") and "heap_hits=0" not in out, out
