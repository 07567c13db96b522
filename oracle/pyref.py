"""ctypes binding of oracle/_ref/libref_orbslam.so — the REFERENCE's own ORBextractor / ORBmatcher / Frame / KeyFrame /
MapPoint / DBoW2 sources compiled from /root/reference against oracle/refshim/ (see oracle/Makefile, refshim/minicv.hpp).
TEST INFRASTRUCTURE ONLY: tests/ and bench.py's `--impl reference` / cpu_baseline legs may import it, the product never does.
The library is built in the build container (where /root/reference exists) and travels to the GPU box as a built file;
`available()` is False when it is missing."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libref_orbslam.so")
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
_lib = None


def available():
    return os.path.exists(_SO)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libref_orbslam.so is not built (make -C oracle, needs /root/reference)")
        L = C.CDLL(_SO)
        vp, i, f, d = C.c_void_p, C.c_int, C.c_float, C.c_double
        L.ref_extractor_create.restype = vp
        L.ref_extractor_create.argtypes = [i, f, i, i, i]
        L.ref_extractor_destroy.argtypes = [vp]
        L.ref_extract.argtypes = [vp, vp, i, i, i, vp, vp, i, C.POINTER(i)]
        L.ref_frame_create.restype = vp
        L.ref_frame_create.argtypes = [vp, vp, i, i, i, i, i, f, f, f, f, i, f]
        L.ref_frame_from_image.restype = vp
        L.ref_frame_from_image.argtypes = [vp, vp, vp, i, i, i, f, f, f, f, vp]
        L.ref_frame_destroy.argtypes = [vp]
        L.ref_frame_n.argtypes = [vp]
        L.ref_frame_get.argtypes = [vp, vp, vp, vp, vp]
        L.ref_frame_grid.argtypes = [vp, vp, vp]
        L.ref_features_in_area.argtypes = [vp, f, f, f, i, i, vp, i]
        L.ref_frame_set_pose.argtypes = [vp, vp]
        L.ref_frame_set_featvec.argtypes = [vp, i, vp, vp, vp]
        L.ref_frame_set_mappoints.argtypes = [vp, vp, vp, vp]
        L.ref_frame_set_bowvec.argtypes = [vp, i, vp, vp]
        L.ref_detect_relocalisation_candidates.argtypes = [vp, vp, vp, i, vp, vp, vp]
        if hasattr(L, "ref_detect_candidates"):
            L.ref_detect_candidates.argtypes = [vp, vp, vp, i, i, vp, vp, vp, i, f, vp, vp, vp, vp]
        L.ref_frame_update_points.argtypes = [vp]
        if hasattr(L, "ref_fuse"):                       # not in the drop-in flavour (oracle/pydropin.py): their projection stays with the caller
            L.ref_search_by_projection_kf.argtypes = [vp, vp, vp, f, i, f, i, vp, vp]
            L.ref_search_by_projection_sim3.argtypes = [vp, vp, vp, i, vp, vp, vp, vp, vp]
            L.ref_fuse.argtypes = [vp, vp, f, vp, vp, vp, vp, vp]
            L.ref_search_by_sim3.argtypes = [vp, vp, f, vp, vp, f, vp, vp, vp, vp, vp, vp, vp, vp, vp]
            L.ref_fuse_sim3.argtypes = [vp, vp, vp, f, vp, vp, vp, vp, vp]
        L.ref_track_with_motion_model.argtypes = [vp, vp, vp, vp]
        L.ref_glue_flavour.restype = C.c_char_p
        L.ref_descriptor_distance.argtypes = [vp, vp]
        L.ref_search_by_projection_ff.argtypes = [vp, vp, f, f, i, vp]
        L.ref_search_by_projection_mappoints.argtypes = [vp, i, vp, vp, vp, vp, vp, vp, f, f, vp]
        L.ref_search_by_bow.argtypes = [vp, vp, f, i, vp]
        L.ref_search_by_bow_kf.argtypes = [vp, vp, f, i, vp]
        L.ref_window_search.argtypes = [vp, vp, i, i, i, f, i, vp]
        L.ref_search_by_projection_window.argtypes = [vp, vp, i, f, vp]
        L.ref_search_for_initialization.argtypes = [vp, vp, vp, i, f, i, vp]
        L.ref_search_for_triangulation.argtypes = [vp, vp, vp, f, i, vp, C.POINTER(i)]
        L.ref_distinctive_descriptor.argtypes = [vp, i, vp]
        L.ref_vocab_load_text.restype = vp
        L.ref_vocab_load_text.argtypes = [C.c_char_p]
        L.ref_vocab_destroy.argtypes = [vp]
        L.ref_vocab_nwords.argtypes = [vp]
        L.ref_vocab_transform.argtypes = [vp, vp, i, i, vp, vp, C.POINTER(i), vp, vp, vp, C.POINTER(i)]
        L.ref_vocab_transform_feature.argtypes = [vp, vp, C.POINTER(C.c_int32)]
        L.ref_vocab_score.restype = d
        L.ref_vocab_score.argtypes = [vp, vp, vp, i, vp, vp, i]
        _lib = L
    return _lib


def _ok(rc, what):
    if rc <= -1000:
        raise RuntimeError("reference %s raised" % what)
    return rc


_SO_FMA = os.path.join(_HERE, "_ref", "libref_extractor_fma.so")
_lib_fma = None


def fma_available():
    return os.path.exists(_SO_FMA)


def lib_fma():
    """the reference extractor built with floating-point contraction on (its own -O3 -march=native on an FMA host)"""
    global _lib_fma
    if _lib_fma is None:
        L = C.CDLL(_SO_FMA)
        L.ref_extractor_create.restype = C.c_void_p
        L.ref_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.ref_extractor_destroy.argtypes = [C.c_void_p]
        L.ref_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        _lib_fma = L
    return _lib_fma


class RefExtractor:
    """ORB_SLAM::ORBextractor (reference include/ORBextractor.h:32-77, src/ORBextractor.cc), the real thing.
    fma=True: the build with floating-point contraction on."""

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, scoreType=1, fastTh=20, fma=False):
        self._L = lib_fma() if fma else lib()
        self._h = self._L.ref_extractor_create(nfeatures, scaleFactor, nlevels, scoreType, fastTh)
        self.cap = max(4 * nfeatures, 64)

    def __del__(self):
        if getattr(self, "_h", None) and getattr(self, "_L", None) is not None:
            self._L.ref_extractor_destroy(self._h)
            self._h = None

    def __call__(self, image):
        img = np.ascontiguousarray(image, np.uint8)
        k = np.zeros(self.cap, KP_DTYPE)
        d = np.zeros((self.cap, 32), np.uint8)
        n = C.c_int(0)
        rc = self._L.ref_extract(self._h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(k), _p(d), self.cap, C.byref(n))
        if rc != 0:
            raise RuntimeError("ref_extract failed: %d" % rc)
        return k[:n.value].copy(), d[:n.value].copy()


class RefFrame:
    """ORB_SLAM::Frame (+ the KeyFrame / Map / MapPoints hung on it) built by the reference's own code."""

    def __init__(self, kps=None, desc=None, w=0, h=0, fx=0.0, fy=0.0, cx=0.0, cy=0.0, nlevels=8, scale_factor=1.2, bounds=None, _h=None):
        if _h is not None:
            self._h = _h
        else:
            self.kps = np.ascontiguousarray(kps, KP_DTYPE)
            self.desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
            b = (0, w, 0, h) if bounds is None else tuple(int(v) for v in bounds)
            self._h = lib().ref_frame_create(_p(self.kps), _p(self.desc), len(self.kps), b[0], b[1], b[2], b[3], fx, fy, cx, cy, nlevels, scale_factor)
        self.n = lib().ref_frame_n(self._h)

    @classmethod
    def from_image(cls, extractor, image, fx, fy, cx, cy, dist=(0, 0, 0, 0), vocab=None):
        img = np.ascontiguousarray(image, np.uint8)
        d = np.ascontiguousarray(dist, np.float32)
        h = lib().ref_frame_from_image(extractor._h, vocab._h if vocab is not None else None, _p(img), img.shape[1], img.shape[0],
                                       img.strides[0], fx, fy, cx, cy, _p(d))
        if not h:
            raise RuntimeError("reference Frame::Frame raised")
        return cls(_h=h)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ref_frame_destroy(self._h)
            self._h = None

    def get(self):
        k, ku, d, b = np.zeros(self.n, KP_DTYPE), np.zeros(self.n, KP_DTYPE), np.zeros((self.n, 32), np.uint8), np.zeros(4, np.int32)
        lib().ref_frame_get(self._h, _p(k), _p(ku), _p(d), _p(b))
        return k, ku, d, b

    def grid(self):
        start, items = np.zeros(64 * 48 + 1, np.int32), np.zeros(max(self.n, 1), np.int32)
        lib().ref_frame_grid(self._h, _p(start), _p(items))
        return start, items[:start[-1]].copy()

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        out = np.zeros(max(self.n, 1), np.int32)
        n = lib().ref_features_in_area(self._h, x, y, r, min_level, max_level, _p(out), len(out))
        return out[:n].copy()

    def set_pose(self, Tcw):
        T = np.ascontiguousarray(Tcw, np.float32).reshape(16)
        lib().ref_frame_set_pose(self._h, _p(T))
        return self

    def set_featvec(self, node_id, start, items):
        a, b, c = (np.ascontiguousarray(v, np.int32) for v in (node_id, start, items))
        lib().ref_frame_set_featvec(self._h, len(a), _p(a), _p(b), _p(c))
        return self

    def update_points(self):
        lib().ref_frame_update_points(self._h)
        return self

    def set_bowvec(self, word, val):
        w, v = np.ascontiguousarray(word, np.int32), np.ascontiguousarray(val, np.float64)
        lib().ref_frame_set_bowvec(self._h, len(w), _p(w), _p(v))
        return self

    def set_mappoints(self, has, xyz=None, outlier=None):
        has = np.ascontiguousarray(has, np.uint8)
        xyz = None if xyz is None else np.ascontiguousarray(xyz, np.float32)
        outlier = None if outlier is None else np.ascontiguousarray(outlier, np.uint8)
        lib().ref_frame_set_mappoints(self._h, _p(has), _p(xyz), _p(outlier))
        return self


def descriptor_distance(a, b):
    a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
    return lib().ref_descriptor_distance(_p(a), _p(b))


def search_by_projection(cur, last, th, nnratio=0.9, check_ori=True, match_cur=None):
    m = np.full(cur.n, -1, np.int32) if match_cur is None else np.ascontiguousarray(match_cur, np.int32)
    n = _ok(lib().ref_search_by_projection_ff(cur._h, last._h, th, nnratio, int(check_ori), _p(m)), "SearchByProjection(F,F)")
    return n, m


def track_with_motion_model(cur, last, velocity):
    """the call sequence of Tracking::TrackWithMotionModel (src/Tracking.cc:594-606) -> (nmatches, match_cur)"""
    V = np.ascontiguousarray(velocity, np.float32).reshape(16)
    m = np.full(cur.n, -1, np.int32)
    return _ok(lib().ref_track_with_motion_model(cur._h, last._h, _p(V), _p(m)), "TrackWithMotionModel"), m


def flavour():
    return lib().ref_glue_flavour().decode()


def search_by_projection_mappoints(f, in_view, proj_x, proj_y, level, view_cos, mp_desc, th, nnratio, match_f=None):
    m = np.full(f.n, -1, np.int32) if match_f is None else np.ascontiguousarray(match_f, np.int32)
    iv, px, py = np.ascontiguousarray(in_view, np.uint8), np.ascontiguousarray(proj_x, np.float32), np.ascontiguousarray(proj_y, np.float32)
    lv, vc, dd = np.ascontiguousarray(level, np.int32), np.ascontiguousarray(view_cos, np.float32), np.ascontiguousarray(mp_desc, np.uint8)
    n = _ok(lib().ref_search_by_projection_mappoints(f._h, len(iv), _p(iv), _p(px), _p(py), _p(lv), _p(vc), _p(dd), th, nnratio, _p(m)),
            "SearchByProjection(F,MapPoints)")
    return n, m


def search_by_projection_kf(cur, kf, already_found, th, orb_dist, nnratio=0.9, check_ori=True, match_cur=None):
    m = np.full(cur.n, -1, np.int32) if match_cur is None else np.ascontiguousarray(match_cur, np.int32)
    af = np.ascontiguousarray(already_found, np.uint8)
    pred = np.full(kf.n, -1, np.int32)
    n = _ok(lib().ref_search_by_projection_kf(cur._h, kf._h, _p(af), th, orb_dist, nnratio, int(check_ori), _p(m), _p(pred)), "SearchByProjection(F,KF)")
    return n, m, pred


def search_by_projection_sim3(kf, src, Scw, th, matched=None):
    """-> (nmatches, matched, (active, u, v, level)) : the reference's result and the projections it worked from"""
    S = np.ascontiguousarray(Scw, np.float32).reshape(16)
    m = np.full(kf.n, -1, np.int32) if matched is None else np.ascontiguousarray(matched, np.int32)
    a, u, v, lv = np.zeros(src.n, np.uint8), np.zeros(src.n, np.float32), np.zeros(src.n, np.float32), np.zeros(src.n, np.int32)
    n = _ok(lib().ref_search_by_projection_sim3(kf._h, src._h, _p(S), int(th), _p(m), _p(a), _p(u), _p(v), _p(lv)), "SearchByProjection(KF,Scw)")
    return n, m, (a, u, v, lv)


def search_by_sim3(k1, k2, s12, R12, t12, th=7.5, matches12=None):
    """ORBmatcher::SearchBySim3 -> (nFound, matches12, (act, u, v, level) of KF1's points in KF2, the same of KF2's points in KF1)"""
    R = np.ascontiguousarray(R12, np.float32).reshape(9); t = np.ascontiguousarray(t12, np.float32).reshape(3)
    m = np.full(k1.n, -1, np.int32) if matches12 is None else np.ascontiguousarray(matches12, np.int32)
    p12 = (np.zeros(k1.n, np.uint8), np.zeros(k1.n, np.float32), np.zeros(k1.n, np.float32), np.zeros(k1.n, np.int32))
    p21 = (np.zeros(k2.n, np.uint8), np.zeros(k2.n, np.float32), np.zeros(k2.n, np.float32), np.zeros(k2.n, np.int32))
    n = _ok(lib().ref_search_by_sim3(k1._h, k2._h, s12, _p(R), _p(t), th, _p(m), *[_p(x) for x in p12], *[_p(x) for x in p21]), "SearchBySim3")
    return n, m, p12, p21


def fuse(kf, src, th=2.5):
    """ORBmatcher::Fuse(pKF, points of src, th) -> (nFused, fused keypoint per source point, (active, u, v, level))"""
    fz = np.full(src.n, -1, np.int32)
    a, u, v, lv = np.zeros(src.n, np.uint8), np.zeros(src.n, np.float32), np.zeros(src.n, np.float32), np.zeros(src.n, np.int32)
    n = _ok(lib().ref_fuse(kf._h, src._h, th, _p(fz), _p(a), _p(u), _p(v), _p(lv)), "Fuse")
    return n, fz, (a, u, v, lv)


def fuse_sim3(kf, src, Scw, th=2.5):
    """ORBmatcher::Fuse(pKF, Scw, points of src, th) -> (nFused, fused keypoint per source point, (active, u, v, level))"""
    S = np.ascontiguousarray(Scw, np.float32).reshape(16)
    fz = np.full(src.n, -1, np.int32)
    a, u, v, lv = np.zeros(src.n, np.uint8), np.zeros(src.n, np.float32), np.zeros(src.n, np.float32), np.zeros(src.n, np.int32)
    n = _ok(lib().ref_fuse_sim3(kf._h, src._h, _p(S), th, _p(fz), _p(a), _p(u), _p(v), _p(lv)), "Fuse(Scw)")
    return n, fz, (a, u, v, lv)


def search_by_bow(kf, f, nnratio, check_ori=True):
    m = np.full(f.n, -1, np.int32)
    return _ok(lib().ref_search_by_bow(kf._h, f._h, nnratio, int(check_ori), _p(m)), "SearchByBoW(KF,F)"), m


def search_by_bow_kf(k1, k2, nnratio, check_ori=True):
    m = np.full(k1.n, -1, np.int32)
    return _ok(lib().ref_search_by_bow_kf(k1._h, k2._h, nnratio, int(check_ori), _p(m)), "SearchByBoW(KF,KF)"), m


def window_search(f1, f2, window, nnratio, check_ori=True, min_level=-1, max_level=2**31 - 1):
    m = np.full(f2.n, -1, np.int32)
    return _ok(lib().ref_window_search(f1._h, f2._h, window, min_level, max_level, nnratio, int(check_ori), _p(m)), "WindowSearch"), m


def search_by_projection_window(f1, f2, window, nnratio, match2):
    m = np.ascontiguousarray(match2, np.int32)
    return _ok(lib().ref_search_by_projection_window(f1._h, f2._h, window, nnratio, _p(m)), "SearchByProjection(F1,F2,window)"), m


def search_for_initialization(f1, f2, prev_matched, window, nnratio, check_ori=True):
    prev = np.ascontiguousarray(prev_matched, np.float32).copy()
    m = np.full(f1.n, -1, np.int32)
    n = _ok(lib().ref_search_for_initialization(f1._h, f2._h, _p(prev), window, nnratio, int(check_ori), _p(m)), "SearchForInitialization")
    return n, m, prev


def search_for_triangulation(k1, k2, F12, nnratio=0.6, check_ori=True):
    F = np.ascontiguousarray(F12, np.float32).reshape(9)
    m = np.full(k1.n, -1, np.int32)
    npairs = C.c_int(0)
    n = _ok(lib().ref_search_for_triangulation(k1._h, k2._h, _p(F), nnratio, int(check_ori), _p(m), C.byref(npairs)), "SearchForTriangulation")
    return n, m, npairs.value


def distinctive_descriptor(desc):
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    out = np.zeros(32, np.uint8)
    rc = _ok(lib().ref_distinctive_descriptor(_p(d), len(d), _p(out)), "ComputeDistinctiveDescriptors")
    return out if rc == 0 else None


def detect_relocalisation_candidates(vocab, query, keyframes):
    """KeyFrameDatabase::DetectRelocalisationCandidates over a database holding `keyframes` -> (common words, score, returned?)"""
    n = len(keyframes)
    hs = (C.c_void_p * n)(*[k._h for k in keyframes])
    common, score, cand = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros(n, np.int32)
    _ok(lib().ref_detect_relocalisation_candidates(vocab._h, query._h, hs, n, _p(common), _p(score), _p(cand)), "DetectRelocalisationCandidates")
    return common, score, cand.astype(bool)


def detect_candidates(vocab, query, keyframes, edges, kf_score, loop=False, min_score=0.0):
    """KeyFrameDatabase::DetectRelocalisationCandidates / DetectLoopCandidates (loop=True) with a covisibility graph.
    edges = [(a, b, weight)], a = -1 for the query keyframe.  kf_score (float32) = the score members before the query, updated in place.
    -> (candidates in returned order, common words, GetBestCovisibilityKeyFrames(10) per keyframe as index lists)"""
    n = len(keyframes)
    hs = (C.c_void_p * n)(*[k._h for k in keyframes])
    e = np.ascontiguousarray(np.asarray(edges, np.int32).reshape(-1, 3))
    ea, eb, ew = (np.ascontiguousarray(e[:, j]) for j in range(3))
    common, cand, best = np.zeros(n, np.int32), np.zeros(max(n, 1), np.int32), np.zeros((n, 10), np.int32)
    nc = _ok(lib().ref_detect_candidates(vocab._h, query._h, hs, n, len(e), _p(ea), _p(eb), _p(ew), int(loop), float(min_score),
                                         _p(kf_score), _p(common), _p(cand), _p(best)), "DetectCandidates")
    return cand[:nc].copy(), common, [row[row >= 0].tolist() for row in best]


class RefVocabulary:
    """ORB_SLAM::ORBVocabulary = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> loaded from the text format."""

    def __init__(self, path):
        self._h = lib().ref_vocab_load_text(path.encode())
        if not self._h:
            raise RuntimeError("loadFromTextFile failed: " + path)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ref_vocab_destroy(self._h)
            self._h = None

    @property
    def nwords(self):
        return lib().ref_vocab_nwords(self._h)

    def transform(self, desc, levelsup=4):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(d)
        bw, bv = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.float64)
        fn, fs, fi = np.zeros(max(n, 1), np.int32), np.zeros(n + 2, np.int32), np.zeros(max(n, 1), np.int32)
        nb, nf = C.c_int(0), C.c_int(0)
        _ok(lib().ref_vocab_transform(self._h, _p(d), n, levelsup, _p(bw), _p(bv), C.byref(nb), _p(fn), _p(fs), _p(fi), C.byref(nf)), "transform")
        return (bw[:nb.value].copy(), bv[:nb.value].copy()), (fn[:nf.value].copy(), fs[:nf.value + 1].copy(), fi[:fs[nf.value]].copy())

    def word(self, desc32):
        d = np.ascontiguousarray(desc32, np.uint8)
        w = C.c_int32(0)
        lib().ref_vocab_transform_feature(self._h, _p(d), C.byref(w))
        return w.value

    def score(self, a, b):
        w1, v1 = np.ascontiguousarray(a[0], np.int32), np.ascontiguousarray(a[1], np.float64)
        w2, v2 = np.ascontiguousarray(b[0], np.int32), np.ascontiguousarray(b[1], np.float64)
        return lib().ref_vocab_score(self._h, _p(w1), _p(v1), len(w1), _p(w2), _p(v2), len(w2))
