"""CPU tests: the oracle against the golden fixtures generated from the real OpenCV (cv2 4.13)
by tests/golden/make_golden.py.  No cv2, no GPU and no /root/reference needed at test time."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def prim():
    return np.load(os.path.join(G, "primitives.npz"))


def test_resize_bit_exact(prim):
    img = prim["img"]
    for i in range(4):
        ref = prim["resize_%d" % i]
        got = po.resize_linear(img, ref.shape[1], ref.shape[0])
        assert np.array_equal(got, ref)


def test_border_reflect101(prim):
    assert np.array_equal(po.border_reflect101(prim["img"], 16), prim["border16"])


@pytest.mark.parametrize("th", [20, 7])
def test_fast_keypoints_order_and_response(prim, th):
    x, y, s = po.fast9_nms(prim["img"], th)
    ref = prim["fast_%d" % th]
    assert len(ref) > 50
    assert np.array_equal(np.stack([x, y, s], 1), ref)


def test_fast_small_images_yield_nothing():
    img = np.random.default_rng(0).integers(0, 256, (6, 50), dtype=np.uint8)
    assert len(po.fast9_nms(img, 7)[0]) == 0
    assert len(po.fast9_nms(np.ascontiguousarray(img.T), 7)[0]) == 0


def test_fallback_subset_property(prim):
    """th=20 survivors == th=7 survivors with score >= 20 (basis of the one-pass GPU form)."""
    x7, y7, s7 = po.fast9_nms(prim["img"], 7)
    x20, y20, s20 = po.fast9_nms(prim["img"], 20)
    k = s7 >= 20
    assert np.array_equal(np.stack([x7[k], y7[k], s7[k]]), np.stack([x20, y20, s20]))


def test_blur_variants(prim):
    assert [int(v) for v in prim["gauss_taps_bits"]] == [0x3d8fafb1, 0x3e06387e, 0x3e434a39, 0x3e5d4ae0,
                                                        0x3e434a39, 0x3e06387e, 0x3d8fafb1]
    assert np.array_equal(po.gaussian_blur7(prim["border16"], 16, po.BLUR_F32), prim["blur_f32"])
    assert np.array_equal(po.gaussian_blur7(prim["border16"], 16, po.BLUR_FIXED_256), prim["blur_fixed256"])


def test_fast_atan2_bits(prim):
    yx = prim["atan2_yx"]
    got = np.array([po.fast_atan2(float(y), float(x)) for y, x in yx], np.float32)
    assert np.array_equal(_bits(got), prim["atan2_deg_bits"])
    assert po.fast_atan2(0.0, 0.0) == 0.0


def test_cv2_orb_single_level_kat():
    """OpenCV's own C++ retainBest order + IC_Angle + sub-matrix blur + rBRIEF, via cv2.ORB(nlevels=1)."""
    d = np.load(os.path.join(G, "cv2_orb_single_level.npz"))
    for t in range(3):
        img, n = d["img_%d" % t], int(d["n_%d" % t])
        ref, h, w = d["kp_%d" % t], img.shape[0], img.shape[1]
        x, y, s = po.fast9_nms(img, 20)
        keep = (x >= 31) & (x < w - 31) & (y >= 31) & (y < h - 31)
        x, y, s = x[keep], y[keep], s[keep]
        assert len(x) > n
        r, idx = po.nth_element_desc(s.astype(np.float32), n - 1)
        first = idx[:n]
        assert np.array_equal(np.stack([x[first], y[first], s[first]], 1), ref[:n])          # order-exact
        rest = idx[n:][r[n:] >= r[n - 1]]
        assert set(zip(x[rest], y[rest])) == set(zip(ref[n:, 0], ref[n:, 1]))              # tie-inclusive tail
        pl = po.border_reflect101(img, 32)
        bl = pl.copy()
        bl[32:-32, 32:-32] = po.gaussian_blur7(pl, 32, po.BLUR_F32)
        ang = np.array([po.ic_angle(pl, int(px) + 32, int(py) + 32) for px, py, _ in ref], np.float32)
        assert np.array_equal(_bits(ang), d["angle_bits_%d" % t])
        desc = np.stack([po.rbrief(bl, int(px) + 32, int(py) + 32, a) for (px, py, _), a in zip(ref, ang)])
        assert np.array_equal(desc, d["desc_%d" % t])


@pytest.mark.parametrize("name", ["e2e_320x240_n300", "e2e_640x480_n1000", "e2e_620x188_n700"])
def test_extract_end_to_end_vs_cv2_composed(name):
    d = np.load(os.path.join(G, name + ".npz"))
    ex = po.OracleExtractor(int(d["nfeatures"]), 1.2, 8, 1, 20)
    kps, desc = ex(d["img"])
    assert np.array_equal(_bits(ex.scale_factors()), _bits(d["scale"]))
    assert np.array_equal(_bits(ex.inv_scale_factors()), _bits(d["inv_scale"]))
    assert ex.features_per_level() == [int(v) for v in d["per_level"]]
    assert ex.umax() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    ref = d["kps"]
    assert len(kps) == len(ref)
    for f in ("x", "y", "size", "angle", "response"):
        assert np.array_equal(_bits(kps[f]), _bits(ref[f])), f
    assert np.array_equal(kps["octave"], ref["octave"]) and np.all(kps["class_id"] == -1)
    assert np.array_equal(desc, d["desc"])
    fallback_cells = 0
    for l in range(8):
        cell, x, y, s = ex.level_candidates(l)
        assert np.array_equal(np.stack([cell, x, y, s], 1), d["L%d_cand" % l])
        nt, nr = ex.level_quota(l)
        assert np.array_equal(np.stack([nt, nr]), d["L%d_quota" % l])
        for c in np.unique(cell):
            if s[cell == c].min() < 20:
                fallback_cells += 1
    assert fallback_cells > 0          # the th=7 fallback really fires in the fixtures
    for l in (1, 4):
        assert np.array_equal(ex.level_plane(l, False), d["L%d_plane" % l])
        assert np.array_equal(ex.level_plane(l, True), d["L%d_blur" % l])


def test_extract_empty_and_tiny():
    ex = po.OracleExtractor(1000)
    k, d = ex(np.zeros((0, 0), np.uint8))
    assert len(k) == 0 and d.shape == (0, 32)
    with pytest.raises(RuntimeError):        # reference divides by zero / throws on such geometry
        po.OracleExtractor(10)(np.zeros((100, 100), np.uint8))
    k, d = ex(np.full((480, 640), 77, np.uint8))   # flat image: no corners at all
    assert len(k) == 0


def test_gemm_projection_pin(prim):
    """Rcw*x3Dw+tcw (src/ORBmatcher.cc:1530): FP32 sum of three products, translation added in double."""
    T, X, Y = prim["gemm_T"], prim["gemm_X"], prim["gemm_Y_bits"]
    for i in range(len(T)):
        for r in range(3):
            t0 = np.float32(np.float32(T[i, r, 0] * X[i, 0, 0]) + np.float32(T[i, r, 1] * X[i, 1, 0]))
            t0 = np.float32(t0 + np.float32(T[i, r, 2] * X[i, 2, 0]))
            v = np.float32(float(t0) + float(T[i, r, 3]))
            assert v.view(np.uint32) == Y[i, r]


def test_descriptor_trig_equals_host_libm():
    """The oracle restates glibc's cosf / sinf (the reference's calls at src/ORBextractor.cc:160) instead of calling libm, so that the pin does
    not move with the host.  Here the restatement is compared with this host's own cosf / sinf on every 89th float angle of [0, 360] degrees
    (12.8 M angles over all binades) and on two dense runs; tools/cpp/sincos_exhaustive.cu covers all 1 135 869 953 angles on the GPU box."""
    import ctypes as C
    L = po.lib()
    L.orc_trig_mismatches.restype = C.c_longlong
    L.orc_trig_mismatches.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    last = int(np.float32(360.0).view(np.uint32))
    assert L.orc_trig_mismatches(0, last, 89) == 0
    lo = int(np.float32(0.1).view(np.uint32))                       # where glibc's sinf is most often one ulp from the correctly rounded value
    assert L.orc_trig_mismatches(lo, lo + 2_000_000, 1) == 0
    lo = int(np.float32(44.9).view(np.uint32))                      # across the pi/4 switch of the argument reduction
    assert L.orc_trig_mismatches(lo, lo + 2_000_000, 1) == 0


def test_trig_constants_shared_by_kernel_and_oracle():
    """csrc/orb_trig.h (device) and oracle/orb_oracle.cpp (checker) restate the same libm algorithm: their hexadecimal float constants
    (reduction factor, pi/2, polynomial coefficients) and the two argument-range thresholds must be the same set."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dev = open(os.path.join(root, "orbslam_jpminipc_b200", "csrc", "orb_trig.h")).read()
    ora = open(os.path.join(root, "oracle", "orb_oracle.cpp")).read()
    ora = ora[ora.index("static float trig_poly"):ora.index("/* computeOrbDescriptor, src/ORBextractor.cc:155-194. */")]
    hexf = lambda s: sorted(set(float.fromhex(m) for m in re.findall(r"0x1(?:\.[0-9a-fA-F]+)?p[+-]?\d+", s)))
    assert hexf(dev) == hexf(ora) and len(hexf(dev)) == 9      # 2/pi * 2^24, pi/2, four cosine and three sine coefficients (c0 = 1 is spelled as the sign)
    for t in ("0x3f4", "0x398", "0x800000"):
        assert t in dev and t in ora
