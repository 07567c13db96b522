"""ctypes binding of the CPU oracle (oracle/liborb_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liborb_oracle.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

BLUR_F32, BLUR_FIXED_256, BLUR_FIXED_257 = 0, 1, 2


def build(force=False):
    src = os.path.join(_HERE, "orb_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        # only this library: oracle/_ref/ can be rebuilt only where the reference tree exists and must survive on the GPU box
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s", "liborb_oracle.so"])
    return _SO


class _Frame(C.Structure):
    _fields_ = [("n", C.c_int), ("kps", C.c_void_p), ("desc", C.c_void_p),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("min_x", C.c_int), ("max_x", C.c_int), ("min_y", C.c_int), ("max_y", C.c_int),
                ("nlevels", C.c_int), ("scale_factor", C.c_float),
                ("cell_start", C.c_void_p), ("cell_items", C.c_void_p)]


class _FeatVec(C.Structure):
    _fields_ = [("nnodes", C.c_int), ("node_id", C.c_void_p), ("start", C.c_void_p), ("items", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_extractor_create.restype = C.c_void_p
        L.orc_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_extractor_destroy.argtypes = [C.c_void_p]
        L.orc_extractor_set_descriptor_fma.argtypes = [C.c_void_p, C.c_int]
        L.orc_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_int, C.POINTER(C.c_int)]
        L.orc_scale_factor.restype = C.c_float
        L.orc_scale_factor.argtypes = [C.c_void_p, C.c_int]
        L.orc_inv_scale_factor.restype = C.c_float
        L.orc_inv_scale_factor.argtypes = [C.c_void_p, C.c_int]
        L.orc_features_per_level.argtypes = [C.c_void_p, C.c_int]
        L.orc_umax.restype = C.POINTER(C.c_int)
        L.orc_umax.argtypes = [C.c_void_p]
        L.orc_level_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_level_plane.restype = C.c_void_p
        L.orc_level_plane.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_level_candidates.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 4
        L.orc_level_quota.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_resize_linear_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_border_reflect101.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_fast9_nms.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 3
        L.orc_fast_atan2.restype = C.c_float
        L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_gaussian_blur7.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.orc_nth_element_desc.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orc_retain_best.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orc_ic_angle.restype = C.c_float
        L.orc_ic_angle.argtypes = [C.c_void_p, C.c_int]
        L.orc_rbrief.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p]
        L.orc_descriptor_distance.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_knn2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_match_ratio.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p]
        L.orc_frame_grid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_features_in_area.argtypes = [C.POINTER(_Frame), C.c_float, C.c_float, C.c_float, C.c_int, C.c_int,
                                           C.c_void_p, C.c_int]
        L.orc_search_by_projection.argtypes = [C.POINTER(_Frame), C.POINTER(_Frame), C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p]
        L.orc_search_by_bow.argtypes = [C.POINTER(_FeatVec), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                        C.POINTER(_FeatVec), C.c_void_p, C.c_void_p, C.c_int,
                                        C.c_float, C.c_int, C.c_void_p]
        L.orc_search_by_projection_mappoints.argtypes = [C.POINTER(_Frame), C.c_int] + [C.c_void_p] * 6 + [C.c_float, C.c_float, C.c_void_p]
        L.orc_window_search.argtypes = [C.POINTER(_Frame), C.POINTER(_Frame), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p]
        L.orc_search_by_projection_window.argtypes = [C.POINTER(_Frame), C.POINTER(_Frame), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                                      C.c_float, C.c_void_p]
        L.orc_search_by_bow_kf.argtypes = [C.POINTER(_FeatVec), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                           C.POINTER(_FeatVec), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                           C.c_float, C.c_int, C.c_void_p]
        L.orc_search_by_projection_kf.argtypes = [C.POINTER(_Frame), C.c_int] + [C.c_void_p] * 6 + [C.c_float, C.c_int, C.c_int, C.c_void_p]
        L.orc_search_for_initialization.argtypes = [C.POINTER(_Frame), C.POINTER(_Frame), C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p]
        L.orc_vocab_create.restype = C.c_void_p
        L.orc_vocab_create.argtypes = [C.c_int] * 5 + [C.c_void_p] * 3
        L.orc_vocab_load_text.restype = C.c_void_p
        L.orc_vocab_load_text.argtypes = [C.c_char_p]
        L.orc_vocab_destroy.argtypes = [C.c_void_p]
        L.orc_vocab_nnodes.argtypes = [C.c_void_p]
        L.orc_vocab_nwords.argtypes = [C.c_void_p]
        L.orc_vocab_transform_feature.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_vocab_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 7
        L.orc_bow_score_l1.restype = C.c_double
        L.orc_bow_score_l1.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_bow_score_db.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6
        L.orc_cvt_gray.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_undistort_points.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * 4 + [C.c_void_p, C.c_int]
        L.orc_undistort_keypoints.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * 4 + [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_image_bounds.argtypes = [C.c_int, C.c_int] + [C.c_float] * 4 + [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_distinctive_descriptors.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_harris_response.restype = C.c_float
        L.orc_harris_response.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_search_by_projection_sim3.argtypes = [C.POINTER(_Frame), C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p]
        L.orc_window_best.argtypes = [C.POINTER(_Frame), C.c_int] + [C.c_void_p] * 8
        L.orc_search_for_triangulation.argtypes = [C.POINTER(_FeatVec), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(_FeatVec), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_three_maxima.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 3
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleExtractor:
    """ORB_SLAM::ORBextractor restated on the CPU (reference include/ORBextractor.h:32-77)."""
    HARRIS_SCORE, FAST_SCORE = 0, 1

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, scoreType=1, fastTh=20, blur=BLUR_F32, desc_fma=False):
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self._h = lib().orc_extractor_create(nfeatures, scaleFactor, nlevels, scoreType, fastTh, blur)
        if desc_fma:
            lib().orc_extractor_set_descriptor_fma(self._h, 1)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_extractor_destroy(self._h)
            self._h = None

    def __call__(self, image, mask=None):
        image = np.asarray(image)
        if image.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2 and image.strides[1] == 1
        h, w = image.shape
        cap = self.nfeatures + 64
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        rc = lib().orc_extract(self._h, _p(image), w, h, image.strides[0], _p(kps), _p(desc), cap, C.byref(n))
        if rc != 0:
            raise RuntimeError("orc_extract failed: %d" % rc)
        return kps[:n.value].copy(), desc[:n.value].copy()

    def scale_factors(self):
        return np.array([lib().orc_scale_factor(self._h, l) for l in range(self.nlevels)], np.float32)

    def inv_scale_factors(self):
        return np.array([lib().orc_inv_scale_factor(self._h, l) for l in range(self.nlevels)], np.float32)

    def features_per_level(self):
        return [lib().orc_features_per_level(self._h, l) for l in range(self.nlevels)]

    def umax(self):
        p = lib().orc_umax(self._h)
        return [p[i] for i in range(16)]

    def level_info(self, level):
        info = np.zeros(10, np.int32)
        lib().orc_level_info(self._h, level, _p(info))
        return dict(zip(["w", "h", "stride", "nDesired", "cols", "rows", "cellW", "cellH", "nfCell", "nKept"],
                        [int(v) for v in info]))

    def level_plane(self, level, blurred=False):
        i = self.level_info(level)
        ptr = lib().orc_level_plane(self._h, level, int(blurred))
        buf = (C.c_uint8 * (i["stride"] * (i["h"] + 32))).from_address(ptr)
        return np.frombuffer(buf, np.uint8).reshape(i["h"] + 32, i["stride"]).copy()

    def level_candidates(self, level):
        cap = 1 << 20
        a = [np.zeros(cap, np.int32) for _ in range(4)]
        n = lib().orc_level_candidates(self._h, level, cap, *[_p(x) for x in a])
        assert n >= 0
        return [x[:n].copy() for x in a]

    def level_quota(self, level):
        i = self.level_info(level)
        nt = np.zeros(i["cols"] * i["rows"], np.int32)
        nr = np.zeros_like(nt)
        lib().orc_level_quota(self._h, level, _p(nt), _p(nr))
        return nt, nr


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    return dst


def border_reflect101(img, b=16):
    h, w = img.shape
    plane = np.zeros((h + 2 * b, w + 2 * b), np.uint8)
    plane[b:b + h, b:b + w] = img
    lib().orc_border_reflect101(_p(plane), w, h, w + 2 * b, b)
    return plane


def fast9_nms(img, th):
    assert img.dtype == np.uint8 and img.strides[1] == 1
    cap = max(16, img.shape[0] * img.shape[1] // 4 + 16)
    x, y, s = (np.zeros(cap, np.int32) for _ in range(3))
    n = lib().orc_fast9_nms(_p(img), img.shape[1], img.shape[0], img.strides[0], th, cap, _p(x), _p(y), _p(s))
    assert n >= 0
    return x[:n].copy(), y[:n].copy(), s[:n].copy()


def fast_atan2(y, x):
    return lib().orc_fast_atan2(float(y), float(x))


def gaussian_blur7(padded, b, variant=BLUR_F32):
    """padded: (h+2b, w+2b) plane with a reflect-101 border (b>=3); returns the blurred (h, w) ROI."""
    padded = np.ascontiguousarray(padded, np.uint8)
    h, w = padded.shape[0] - 2 * b, padded.shape[1] - 2 * b
    dst = np.zeros((h, w), np.uint8)
    roi = padded.ctypes.data + b * padded.strides[0] + b
    lib().orc_gaussian_blur7(C.c_void_p(roi), w, h, padded.strides[0], _p(dst), w, variant)
    return dst


def nth_element_desc(resp, nth):
    r = np.ascontiguousarray(resp, np.float32).copy()
    idx = np.arange(len(r), dtype=np.int32)
    lib().orc_nth_element_desc(_p(r), _p(idx), len(r), nth)
    return r, idx


def retain_best(resp, n):
    r = np.ascontiguousarray(resp, np.float32).copy()
    idx = np.arange(len(r), dtype=np.int32)
    m = lib().orc_retain_best(_p(r), _p(idx), len(r), n)
    return r[:m], idx[:m]


def ic_angle(padded, x, y):
    padded = np.ascontiguousarray(padded, np.uint8)
    return lib().orc_ic_angle(C.c_void_p(padded.ctypes.data + y * padded.strides[0] + x), padded.strides[0])


def rbrief(padded, x, y, angle_deg):
    padded = np.ascontiguousarray(padded, np.uint8)
    d = np.zeros(32, np.uint8)
    lib().orc_rbrief(C.c_void_p(padded.ctypes.data + y * padded.strides[0] + x), padded.strides[0],
                     float(angle_deg), _p(d))
    return d


def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return lib().orc_descriptor_distance(_p(a), _p(b))


def knn2(q, db, use_popcnt=True):
    q = np.ascontiguousarray(q, np.uint8)
    db = np.ascontiguousarray(db, np.uint8)
    nq = q.shape[0]
    idx1, d1, d2 = (np.zeros(nq, np.int32) for _ in range(3))
    lib().orc_knn2(_p(q), nq, _p(db), db.shape[0], _p(idx1), _p(d1), _p(d2), int(use_popcnt))
    return idx1, d1, d2


def match_ratio(idx1, d1, d2, nnratio, th):
    m = np.zeros(len(idx1), np.int32)
    n = lib().orc_match_ratio(_p(idx1), _p(d1), _p(d2), len(idx1), nnratio, th, _p(m))
    return m, n


class OracleFrame:
    """The slice of ORB_SLAM::Frame the matcher reads (reference src/Frame.cc:56-128)."""

    def __init__(self, kps, desc, w, h, fx, fy, cx, cy, nlevels=8, scale_factor=1.2, bounds=None):
        self.kps = np.ascontiguousarray(kps, KP_DTYPE)
        self.desc = np.ascontiguousarray(desc, np.uint8)
        self.n = len(self.kps)
        self.cell_start = np.zeros(64 * 48 + 1, np.int32)
        self.cell_items = np.zeros(max(self.n, 1), np.int32)
        b = (0, w, 0, h) if bounds is None else tuple(int(v) for v in bounds)
        lib().orc_frame_grid(_p(self.kps), self.n, b[0], b[1], b[2], b[3], _p(self.cell_start), _p(self.cell_items))
        self.c = _Frame(self.n, self.kps.ctypes.data, self.desc.ctypes.data, fx, fy, cx, cy, b[0], b[1], b[2], b[3],
                        nlevels, scale_factor, self.cell_start.ctypes.data, self.cell_items.ctypes.data)

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        out = np.zeros(max(self.n, 1), np.int32)
        n = lib().orc_features_in_area(C.byref(self.c), x, y, r, min_level, max_level, _p(out), len(out))
        return out[:n].copy()


def search_by_projection(cur, last, last_has_mp, last_outlier, last_xyz, Tcw, th, check_ori=True, match_cur=None):
    if match_cur is None:
        match_cur = np.full(cur.n, -1, np.int32)
    has = np.ascontiguousarray(last_has_mp, np.uint8)
    out = np.ascontiguousarray(last_outlier, np.uint8)
    xyz = np.ascontiguousarray(last_xyz, np.float32)
    T = np.ascontiguousarray(Tcw, np.float32).reshape(16)
    n = lib().orc_search_by_projection(C.byref(cur.c), C.byref(last.c), _p(has), _p(out), _p(xyz), _p(T),
                                       th, int(check_ori), _p(match_cur))
    return n, match_cur


def _fv(node_id, start, items):
    node_id = np.ascontiguousarray(node_id, np.int32)
    start = np.ascontiguousarray(start, np.int32)
    items = np.ascontiguousarray(items, np.int32)
    return _FeatVec(len(node_id), node_id.ctypes.data, start.ctypes.data, items.ctypes.data), (node_id, start, items)


def search_by_bow(kf_fv, kf_desc, kf_kps, kf_mp_valid, f_fv, f_desc, f_kps, nnratio, check_ori=True):
    a, keep_a = _fv(*kf_fv)
    b, keep_b = _fv(*f_fv)
    kf_desc = np.ascontiguousarray(kf_desc, np.uint8)
    f_desc = np.ascontiguousarray(f_desc, np.uint8)
    kf_kps = np.ascontiguousarray(kf_kps, KP_DTYPE)
    f_kps = np.ascontiguousarray(f_kps, KP_DTYPE)
    valid = np.ascontiguousarray(kf_mp_valid, np.uint8)
    m = np.full(len(f_kps), -1, np.int32)
    n = lib().orc_search_by_bow(C.byref(a), _p(kf_desc), _p(kf_kps), _p(valid), len(kf_kps),
                                C.byref(b), _p(f_desc), _p(f_kps), len(f_kps), nnratio, int(check_ori), _p(m))
    return n, m


def three_maxima(sizes):
    s = np.ascontiguousarray(sizes, np.int32)
    i1, i2, i3 = C.c_int(), C.c_int(), C.c_int()
    lib().orc_three_maxima(_p(s), len(s), C.byref(i1), C.byref(i2), C.byref(i3))
    return i1.value, i2.value, i3.value


def merge_best2(parts):
    """Exact merge of per-shard (idx1, d1, d2) triples, shards in ascending global-index order
    (SURVEY.md §8e): best = first strict minimum, second = 2nd smallest of the multiset union.
    parts: int array [nshards, 3, nq] -> (idx1, d1, d2)."""
    parts = np.asarray(parts)
    nq = parts.shape[2]
    I, D1, D2 = np.full(nq, -1, np.int32), np.full(nq, 2**31 - 1, np.int32), np.full(nq, 2**31 - 1, np.int32)
    for s in range(parts.shape[0]):
        pi, p1, p2 = parts[s]
        for i in range(nq):
            if pi[i] < 0:
                continue
            if p1[i] < D1[i]:
                D2[i] = D1[i]; D1[i] = p1[i]; I[i] = pi[i]
            elif p1[i] < D2[i]:
                D2[i] = p1[i]
            if p2[i] < D2[i]:
                D2[i] = p2[i]
    return I, D1, D2


def search_by_projection_mappoints(f, in_view, proj_x, proj_y, level, view_cos, mp_desc, th, nnratio, match_f=None):
    if match_f is None:
        match_f = np.full(f.n, -1, np.int32)
    a = [np.ascontiguousarray(in_view, np.uint8), np.ascontiguousarray(proj_x, np.float32), np.ascontiguousarray(proj_y, np.float32),
         np.ascontiguousarray(level, np.int32), np.ascontiguousarray(view_cos, np.float32), np.ascontiguousarray(mp_desc, np.uint8)]
    n = lib().orc_search_by_projection_mappoints(C.byref(f.c), len(a[0]), *[_p(x) for x in a], th, nnratio, _p(match_f))
    return n, match_f


def window_search(f1, f2, f1_has_mp, window, nnratio, check_ori=True, min_level=-1, max_level=2**31 - 1):
    has = np.ascontiguousarray(f1_has_mp, np.uint8)
    m = np.full(f2.n, -1, np.int32)
    n = lib().orc_window_search(C.byref(f1.c), C.byref(f2.c), _p(has), window, min_level, max_level, nnratio, int(check_ori), _p(m))
    return n, m


def search_by_projection_window(f1, f2, f1_active, f1_xyz, Tc2w, window, nnratio, match2):
    act = np.ascontiguousarray(f1_active, np.uint8)
    xyz = np.ascontiguousarray(f1_xyz, np.float32)
    T = np.ascontiguousarray(Tc2w, np.float32).reshape(16)
    n = lib().orc_search_by_projection_window(C.byref(f1.c), C.byref(f2.c), _p(act), _p(xyz), _p(T), window, nnratio, _p(match2))
    return n, match2


def search_by_bow_kf(fv1, desc1, kps1, valid1, fv2, desc2, kps2, valid2, nnratio, check_ori=True):
    a, keep_a = _fv(*fv1)
    b, keep_b = _fv(*fv2)
    desc1 = np.ascontiguousarray(desc1, np.uint8); desc2 = np.ascontiguousarray(desc2, np.uint8)
    kps1 = np.ascontiguousarray(kps1, KP_DTYPE); kps2 = np.ascontiguousarray(kps2, KP_DTYPE)
    v1 = np.ascontiguousarray(valid1, np.uint8); v2 = np.ascontiguousarray(valid2, np.uint8)
    m = np.full(len(kps1), -1, np.int32)
    n = lib().orc_search_by_bow_kf(C.byref(a), _p(desc1), _p(kps1), _p(v1), len(kps1), C.byref(b), _p(desc2), _p(kps2), _p(v2),
                                   len(kps2), nnratio, int(check_ori), _p(m))
    return n, m


def search_by_projection_kf(cur, active, xyz, Tcw, pred_level, mp_desc, kf_angle, th, orb_dist, check_ori=True, match_cur=None):
    if match_cur is None:
        match_cur = np.full(cur.n, -1, np.int32)
    a = [np.ascontiguousarray(active, np.uint8), np.ascontiguousarray(xyz, np.float32), np.ascontiguousarray(Tcw, np.float32).reshape(16),
         np.ascontiguousarray(pred_level, np.int32), np.ascontiguousarray(mp_desc, np.uint8), np.ascontiguousarray(kf_angle, np.float32)]
    n = lib().orc_search_by_projection_kf(C.byref(cur.c), len(a[0]), *[_p(x) for x in a], th, orb_dist, int(check_ori), _p(match_cur))
    return n, match_cur


def search_for_initialization(f1, f2, prev_matched, window, nnratio, check_ori=True):
    prev = np.ascontiguousarray(prev_matched, np.float32).copy()
    m = np.full(f1.n, -1, np.int32)
    n = lib().orc_search_for_initialization(C.byref(f1.c), C.byref(f2.c), _p(prev), window, nnratio, int(check_ori), _p(m))
    return n, m, prev


class OracleVocabulary:
    """DBoW2 TemplatedVocabulary<FORB> restated (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h)."""

    def __init__(self, k=None, L=None, parent=None, desc=None, weight=None, scoring=0, weighting=0, path=None):
        if path is not None:
            self._h = lib().orc_vocab_load_text(str(path).encode())
        else:
            parent = np.ascontiguousarray(parent, np.int32)
            desc = np.ascontiguousarray(desc, np.uint8)
            weight = np.ascontiguousarray(weight, np.float64)
            self._h = lib().orc_vocab_create(k, L, scoring, weighting, len(parent), _p(parent), _p(desc), _p(weight))
        if not self._h:
            raise ValueError("invalid vocabulary")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_vocab_destroy(self._h)
            self._h = None

    @property
    def nnodes(self):
        return lib().orc_vocab_nnodes(self._h)

    @property
    def nwords(self):
        return lib().orc_vocab_nwords(self._h)

    def transform_features(self, desc, levelsup=0):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        word = np.zeros(n, np.int32); weight = np.zeros(n, np.float64); node = np.zeros(n, np.int32)
        w = C.c_int32(0); wt = C.c_double(0); nd = C.c_int32(0)
        for i in range(n):
            lib().orc_vocab_transform_feature(self._h, desc[i].ctypes.data, levelsup, C.addressof(w), C.addressof(wt), C.addressof(nd))
            word[i], weight[i], node[i] = w.value, wt.value, nd.value
        return word, weight, node

    def transform(self, desc, levelsup=4):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        bw = np.zeros(max(n, 1), np.int32); bv = np.zeros(max(n, 1), np.float64)
        fn = np.zeros(max(n, 1), np.int32); fs = np.zeros(n + 1, np.int32); fi = np.zeros(max(n, 1), np.int32)
        nb = C.c_int(0); nf = C.c_int(0)
        lib().orc_vocab_transform(self._h, _p(desc), n, levelsup, _p(bw), _p(bv), C.addressof(nb), _p(fn), _p(fs), _p(fi), C.addressof(nf))
        return (bw[:nb.value].copy(), bv[:nb.value].copy()), (fn[:nf.value].copy(), fs[:nf.value + 1].copy(), fi[:fs[nf.value]].copy())


def bow_score_l1(v1, v2):
    w1 = np.ascontiguousarray(v1[0], np.int32); x1 = np.ascontiguousarray(v1[1], np.float64)
    w2 = np.ascontiguousarray(v2[0], np.int32); x2 = np.ascontiguousarray(v2[1], np.float64)
    return lib().orc_bow_score_l1(_p(w1), _p(x1), len(w1), _p(w2), _p(x2), len(w2))


def bow_score_db(query, kf_bows):
    qw = np.ascontiguousarray(query[0], np.int32); qv = np.ascontiguousarray(query[1], np.float64)
    start = np.zeros(len(kf_bows) + 1, np.int32)
    for i, b in enumerate(kf_bows):
        start[i + 1] = start[i] + len(b[0])
    words = np.ascontiguousarray(np.concatenate([np.asarray(b[0], np.int32) for b in kf_bows]), np.int32)
    vals = np.ascontiguousarray(np.concatenate([np.asarray(b[1], np.float64) for b in kf_bows]), np.float64)
    common = np.zeros(len(kf_bows), np.int32); score = np.zeros(len(kf_bows), np.float32)
    mx = C.c_int(0)
    lib().orc_bow_score_db(_p(qw), _p(qv), len(qw), len(kf_bows), _p(start), _p(words), _p(vals), _p(common), _p(score), C.addressof(mx))
    return common, score, mx.value


def bow_detect_candidates(query, kf_bows, kf_score, covis=None, excluded=None, loop=False, min_score=0.0):
    """src/KeyFrameDatabase.cc:198-308 (loop=False) / :75-196 (loop=True) -> (candidates in the reference's order, common); kf_score in place"""
    qw = np.ascontiguousarray(query[0], np.int32); qv = np.ascontiguousarray(query[1], np.float64)
    n = len(kf_bows)
    start = np.zeros(n + 1, np.int32)
    for i, b in enumerate(kf_bows):
        start[i + 1] = start[i] + len(b[0])
    words = np.ascontiguousarray(np.concatenate([np.asarray(b[0], np.int32) for b in kf_bows]), np.int32)
    vals = np.ascontiguousarray(np.concatenate([np.asarray(b[1], np.float64) for b in kf_bows]), np.float64)
    cs = ci = None
    if covis is not None:
        cs = np.zeros(n + 1, np.int32)
        for i, c_ in enumerate(covis):
            cs[i + 1] = cs[i] + len(c_)
        ci = np.ascontiguousarray(np.concatenate([np.asarray(c_, np.int32) for c_ in covis]) if cs[n] else np.zeros(1, np.int32), np.int32)
    ex = None if excluded is None else np.ascontiguousarray(excluded, np.uint8)
    assert kf_score.dtype == np.float32 and kf_score.flags.c_contiguous
    common = np.zeros(n, np.int32); cand = np.zeros(max(n, 1), np.int32)
    f = lib().orc_bow_detect_candidates
    f.restype = C.c_int
    f.argtypes = [C.c_void_p] * 2 + [C.c_int] * 2 + [C.c_void_p] * 4 + [C.c_int, C.c_float] + [C.c_void_p] * 5
    nc = f(_p(qw), _p(qv), len(qw), n, _p(start), _p(words), _p(vals), _p(ex) if ex is not None else None, int(loop), float(min_score),
           _p(cs) if cs is not None else None, _p(ci) if ci is not None else None, _p(kf_score), _p(common), _p(cand))
    return cand[:nc].copy(), common


def cvt_gray(img, order=0):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape[:2]
    out = np.zeros((h, w), np.uint8)
    lib().orc_cvt_gray(_p(img), w, h, img.strides[0], order, _p(out), w)
    return out


def undistort_points(xy, K, dist):
    xy = np.ascontiguousarray(xy, np.float32).copy()
    d = np.ascontiguousarray(dist, np.float32)
    lib().orc_undistort_points(_p(xy), len(xy), K[0], K[1], K[2], K[3], _p(d), len(d))
    return xy


def undistort_keypoints(kps, K, dist):
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    out = np.zeros_like(kps)
    d = np.ascontiguousarray(dist, np.float32)
    lib().orc_undistort_keypoints(_p(kps), len(kps), K[0], K[1], K[2], K[3], _p(d), len(d), _p(out))
    return out


def image_bounds(w, h, K, dist):
    d = np.ascontiguousarray(dist, np.float32)
    b = np.zeros(4, np.int32)
    lib().orc_image_bounds(w, h, K[0], K[1], K[2], K[3], _p(d), len(d), _p(b))
    return b


def distinctive_descriptors(desc, start):
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    start = np.ascontiguousarray(start, np.int32)
    n = len(start) - 1
    bi = np.zeros(n, np.int32); bm = np.zeros(n, np.int32)
    lib().orc_distinctive_descriptors(_p(desc), _p(start), n, _p(bi), _p(bm))
    return bi, bm


def harris_response(img, x, y):
    img = np.ascontiguousarray(img, np.uint8)
    return lib().orc_harris_response(_p(img), img.strides[0], int(x), int(y))


def search_by_projection_sim3(kf, active, u, v, pred_level, mp_desc, th, matched):
    a = np.ascontiguousarray(active, np.uint8); uu = np.ascontiguousarray(u, np.float32); vv = np.ascontiguousarray(v, np.float32)
    lv = np.ascontiguousarray(pred_level, np.int32); d = np.ascontiguousarray(mp_desc, np.uint8)
    n = lib().orc_search_by_projection_sim3(C.byref(kf.c), len(a), _p(a), _p(uu), _p(vv), _p(lv), _p(d), int(th), _p(matched))
    return n, matched


def window_best(f, active, u, v, radius, pred_level, desc):
    a = np.ascontiguousarray(active, np.uint8); uu = np.ascontiguousarray(u, np.float32); vv = np.ascontiguousarray(v, np.float32)
    r = np.ascontiguousarray(radius, np.float32); lv = np.ascontiguousarray(pred_level, np.int32); d = np.ascontiguousarray(desc, np.uint8)
    bi = np.zeros(len(a), np.int32); bd = np.zeros(len(a), np.int32)
    lib().orc_window_best(C.byref(f.c), len(a), _p(a), _p(uu), _p(vv), _p(r), _p(lv), _p(d), _p(bi), _p(bd))
    return bi, bd


def search_for_triangulation(fv1, desc1, kps1, has_mp1, fv2, desc2, kps2, has_mp2, F12, level_sigma2, check_ori=True):
    a, keep_a = _fv(*fv1)
    b, keep_b = _fv(*fv2)
    desc1 = np.ascontiguousarray(desc1, np.uint8); desc2 = np.ascontiguousarray(desc2, np.uint8)
    kps1 = np.ascontiguousarray(kps1, KP_DTYPE); kps2 = np.ascontiguousarray(kps2, KP_DTYPE)
    h1 = np.ascontiguousarray(has_mp1, np.uint8); h2 = np.ascontiguousarray(has_mp2, np.uint8)
    F = np.ascontiguousarray(F12, np.float32).reshape(9); sg = np.ascontiguousarray(level_sigma2, np.float32)
    m = np.full(len(kps1), -1, np.int32)
    n = lib().orc_search_for_triangulation(C.byref(a), _p(desc1), _p(kps1), _p(h1), len(kps1), C.byref(b), _p(desc2), _p(kps2), _p(h2),
                                           len(kps2), _p(F), _p(sg), int(check_ori), _p(m))
    return n, m
