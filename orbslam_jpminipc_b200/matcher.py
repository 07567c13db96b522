"""ORBmatcher / Frame — python mirrors of ORB_SLAM::ORBmatcher (reference include/ORBmatcher.h:37-107)
and of the slice of ORB_SLAM::Frame the matcher reads (include/Frame.h, src/Frame.cc:56-128), on top
of the C ABI.  All compute happens in liborb_b200.so on the GPU.
"""
import ctypes as C

import numpy as np

from ._lib import (GRID_COLS, GRID_ROWS, KP_DTYPE, FeatVecView, FrameView, WindowQuerySet, check, lib, ptr)
from .extractor import ORBextractor


class Frame:
    """Keypoints + descriptors + the 64x48 lookup grid (src/Frame.cc:109-123), zero distortion
    (mvKeysUn == mvKeys, image bounds = image rectangle, src/Frame.cc:291-295,:342-348)."""

    def __init__(self, ctx, kps, desc, width, height, fx, fy, cx, cy, nlevels=8, scale_factor=1.2, bounds=None):
        self.kps = np.ascontiguousarray(kps, KP_DTYPE)
        self.desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        self.N = len(self.kps)
        self.width, self.height = width, height
        self.fx, self.fy, self.cx, self.cy = fx, fy, cx, cy
        self.nlevels, self.scale_factor = nlevels, scale_factor
        self.cell_start = np.zeros(GRID_COLS * GRID_ROWS + 1, np.int32)
        self.cell_items = np.zeros(max(self.N, 1), np.int32)
        # mnMinX, mnMaxX, mnMinY, mnMaxY: the image rectangle, or Frame::ComputeImageBounds for a distorted camera
        self.bounds = (0, width, 0, height) if bounds is None else tuple(int(v) for v in bounds)
        h = ctx._h if hasattr(ctx, "_h") else ctx
        check(lib().orb_frame_grid_build(h, ptr(self.kps), self.N, self.bounds[0], self.bounds[1], self.bounds[2], self.bounds[3],
                                         ptr(self.cell_start), ptr(self.cell_items)), "orb_frame_grid_build")

    def view(self):
        return FrameView(self.N, self.kps.ctypes.data, self.desc.ctypes.data, self.fx, self.fy, self.cx, self.cy,
                         self.bounds[0], self.bounds[1], self.bounds[2], self.bounds[3], self.nlevels, self.scale_factor,
                         self.cell_start.ctypes.data, self.cell_items.ctypes.data)


class ORBmatcher:
    TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30       # src/ORBmatcher.cc:40-42

    def __init__(self, nnratio=0.6, checkOri=True, extractor=None, device=0):
        self.mfNNratio, self.mbCheckOrientation = float(nnratio), bool(checkOri)
        self._own = None
        if extractor is None:                             # the matcher only needs a context (streams, scratch)
            extractor = self._own = ORBextractor(device=device, max_width=64, max_height=64, max_batch=1)
        self._ex = extractor
        self._h = extractor._h

    @staticmethod
    def DescriptorDistance(a, b):
        a = np.ascontiguousarray(a, np.uint8).reshape(32)
        b = np.ascontiguousarray(b, np.uint8).reshape(32)
        return lib().orb_descriptor_distance(ptr(a), ptr(b))

    def knn2(self, q, db):
        """best / second-best over all db rows (scan semantics of src/ORBmatcher.cc:197-222)."""
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
        db = np.ascontiguousarray(db, np.uint8).reshape(-1, 32)
        nq = len(q)
        idx1, d1, d2 = (np.zeros(nq, np.int32) for _ in range(3))
        check(lib().orb_hamming_knn2(self._h, ptr(q), nq, ptr(db), len(db), ptr(idx1), ptr(d1), ptr(d2)), "orb_hamming_knn2")
        return idx1, d1, d2

    def match_ratio(self, idx1, d1, d2, th=None):
        th = self.TH_LOW if th is None else th
        m = np.zeros(len(idx1), np.int32)
        n = C.c_int(0)
        a, b, c = (np.ascontiguousarray(v, np.int32) for v in (idx1, d1, d2))
        check(lib().orb_match_ratio(self._h, ptr(a), ptr(b), ptr(c), len(a), self.mfNNratio, th, ptr(m), C.byref(n)),
              "orb_match_ratio")
        return m, n.value

    def SearchByProjection(self, CurrentFrame, LastFrame, th, last_has_mp, last_outlier, last_xyz, Tcw, match_cur=None):
        """ORBmatcher::SearchByProjection(Frame&, const Frame&, float) (src/ORBmatcher.cc:1507-1620).
        Returns (nmatches, match_cur) with match_cur[i2] = last-frame feature index or -1."""
        if match_cur is None:
            match_cur = np.full(CurrentFrame.N, -1, np.int32)
        has = np.ascontiguousarray(last_has_mp, np.uint8)
        out = np.ascontiguousarray(last_outlier, np.uint8)
        xyz = np.ascontiguousarray(last_xyz, np.float32)
        T = np.ascontiguousarray(Tcw, np.float32).reshape(16)
        cur, last = CurrentFrame.view(), LastFrame.view()
        n = C.c_int(0)
        check(lib().orb_search_by_projection(self._h, C.byref(cur), C.byref(last), ptr(has), ptr(out), ptr(xyz), ptr(T),
                                             th, int(self.mbCheckOrientation), ptr(match_cur), C.byref(n)),
              "orb_search_by_projection")
        return n.value, match_cur

    # ---- the other windowed searches, all through orb_search_window ----
    ACCEPT_BEST, ACCEPT_RATIO, ACCEPT_LEVEL_RATIO = 0, 1, 2

    def _search_window(self, target, n, active, desc, u, v, xyz, Tcw, check_bounds, radius, radius_const, min_level, max_level,
                       angle, accept, th_dist, check_ori, match):
        keep = []

        def arr(a, dt):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dt)
            keep.append(a)
            return a.ctypes.data
        q = WindowQuerySet(n, arr(active, np.uint8), arr(desc, np.uint8), arr(u, np.float32), arr(v, np.float32),
                           arr(xyz, np.float32), arr(None if Tcw is None else np.asarray(Tcw, np.float32).reshape(16), np.float32),
                           int(check_bounds), arr(radius, np.float32), float(radius_const), arr(min_level, np.int32),
                           arr(max_level, np.int32), arr(angle, np.float32))
        tv = target.view()
        nm = C.c_int(0)
        check(lib().orb_search_window(self._h, C.byref(tv), C.byref(q), accept, self.mfNNratio, th_dist, int(check_ori),
                                      ptr(match), C.byref(nm)), "orb_search_window")
        return nm.value, match

    def SearchByProjectionMapPoints(self, F, in_view, proj_x, proj_y, level, view_cos, mp_desc, th=3.0, match_f=None):
        """ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:49-125).
        Per map point: mbTrackInView && !isBad, mTrackProjX/Y, mnTrackScaleLevel, mTrackViewCos, GetDescriptor()."""
        n = len(in_view)
        if match_f is None:
            match_f = np.full(F.N, -1, np.int32)
        sf = np.ones(F.nlevels, np.float32)
        for i in range(1, F.nlevels):
            sf[i] = np.float32(sf[i - 1] * np.float32(F.scale_factor))
        level = np.asarray(level, np.int32)
        r = np.where(np.asarray(view_cos, np.float32).astype(np.float64) > 0.998, np.float32(2.5), np.float32(4.0)).astype(np.float32)
        if th != 1.0:                                      # bFactor, :53,:69-70
            r = (r * np.float32(th)).astype(np.float32)
        radius = (r * sf[np.clip(level, 0, F.nlevels - 1)]).astype(np.float32)
        return self._search_window(F, n, in_view, mp_desc, proj_x, proj_y, None, None, 0, radius, 0.0, level - 1, level, None,
                                   self.ACCEPT_LEVEL_RATIO, self.TH_HIGH, False, match_f)

    def WindowSearch(self, F1, F2, windowSize, f1_has_mp, minScaleLevel=-1, maxScaleLevel=2**31 - 1):
        """ORBmatcher::WindowSearch (src/ORBmatcher.cc:409-516).  Returns (nmatches, vnMatches21)."""
        lv = F1.kps["octave"].astype(np.int32)
        active = np.asarray(f1_has_mp, np.uint8).copy()
        if minScaleLevel > 0:
            active[lv < minScaleLevel] = 0
        if maxScaleLevel < 2**31 - 1:
            active[lv > maxScaleLevel] = 0
        match2 = np.full(F2.N, -1, np.int32)
        return self._search_window(F2, F1.N, active, F1.desc, F1.kps["x"], F1.kps["y"], None, None, 0, None, float(windowSize),
                                   lv, lv, F1.kps["angle"], self.ACCEPT_RATIO, self.TH_HIGH, self.mbCheckOrientation, match2)

    def SearchByProjectionWindow(self, F1, F2, windowSize, f1_active, f1_xyz, Tc2w, match2):
        """ORBmatcher::SearchByProjection(Frame &F1, Frame &F2, int windowSize, ...) (src/ORBmatcher.cc:519-594).
        f1_active: F1 map point is live and not already among F2's; match2 (in/out) starts as F2.mvpMapPoints (>=0 = set)."""
        lv = F1.kps["octave"].astype(np.int32)
        return self._search_window(F2, F1.N, f1_active, F1.desc, None, None, f1_xyz, Tc2w, 0, None, float(windowSize), lv, lv, None,
                                   self.ACCEPT_RATIO, self.TH_HIGH, False, match2)

    def SearchForInitialization(self, F1, F2, vbPrevMatched, windowSize=10):
        """ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:598-713).  vbPrevMatched: n1 x 2 float32 window centres.
        Returns (nmatches, vnMatches12, updated vbPrevMatched)."""
        prev = np.ascontiguousarray(vbPrevMatched, np.float32).reshape(-1, 2).copy()
        if len(prev) != F1.N:
            raise ValueError("vbPrevMatched must hold one point per F1 keypoint")
        m12 = np.full(F1.N, -1, np.int32)
        n = C.c_int(0)
        v1, v2 = F1.view(), F2.view()
        check(lib().orb_search_for_initialization(self._h, C.byref(v1), C.byref(v2), ptr(prev), int(windowSize), self.mfNNratio,
                                                  int(self.mbCheckOrientation), ptr(m12), C.byref(n)), "orb_search_for_initialization")
        return n.value, m12, prev

    def SearchByProjectionKeyFrame(self, CurrentFrame, active, xyz, Tcw, pred_level, mp_desc, kf_angle, th, ORBdist, match_cur=None):
        """ORBmatcher::SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist) (src/ORBmatcher.cc:1622-1746).
        active: KF map point live and not in sAlreadyFound; pred_level: the level predicted from dist3D/minDistance
        (:1663-1669, host side); kf_angle: pKF->GetKeyPointUn(i).angle."""
        n = len(active)
        if match_cur is None:
            match_cur = np.full(CurrentFrame.N, -1, np.int32)
        sf = np.ones(CurrentFrame.nlevels, np.float32)
        for i in range(1, CurrentFrame.nlevels):
            sf[i] = np.float32(sf[i - 1] * np.float32(CurrentFrame.scale_factor))
        lv = np.asarray(pred_level, np.int32)
        radius = (np.float32(th) * sf[np.clip(lv, 0, CurrentFrame.nlevels - 1)]).astype(np.float32)
        return self._search_window(CurrentFrame, n, active, mp_desc, None, None, xyz, Tcw, 1, radius, 0.0, lv - 1, lv + 1, kf_angle,
                                   self.ACCEPT_BEST, int(ORBdist), self.mbCheckOrientation, match_cur)

    # ---- back-end searches on pre-projected map points (the Sim3 / pose projection and the level prediction of the reference stay
    #      in the adapter, like for SearchByProjectionKeyFrame)
    @staticmethod
    def _scale_factors(F):
        sf = np.ones(F.nlevels, np.float32)
        for i in range(1, F.nlevels):
            sf[i] = np.float32(sf[i - 1] * np.float32(F.scale_factor))
        return sf

    def SearchByProjectionSim3(self, KF, active, u, v, pred_level, mp_desc, th, matched):
        """ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) (src/ORBmatcher.cc:286-407) from the radius search
        on.  matched (in/out, one per KF keypoint): >= 0 where vpMatched is already set; returns (nmatches, matched)."""
        lv = np.asarray(pred_level, np.int32)
        radius = (np.float32(int(th)) * self._scale_factors(KF)[np.clip(lv, 0, KF.nlevels - 1)]).astype(np.float32)
        return self._search_window(KF, len(lv), active, mp_desc, u, v, None, None, 0, radius, 0.0, lv - 1, lv, None,
                                   self.ACCEPT_BEST, self.TH_LOW, False, matched)

    def BestInWindow(self, F, active, u, v, radius, min_level, max_level, desc):
        """orb_search_window_best: most similar keypoint in each query's radius, no claims -> (best_idx, best_dist)."""
        keep = []

        def arr(a, dt):
            a = np.ascontiguousarray(a, dt)
            keep.append(a)
            return a.ctypes.data
        n = len(active)
        q = WindowQuerySet(n, arr(active, np.uint8), arr(desc, np.uint8), arr(u, np.float32), arr(v, np.float32), None, None, 0,
                           arr(radius, np.float32), 0.0, arr(min_level, np.int32), arr(max_level, np.int32), None)
        tv = F.view()
        bi = np.full(n, -1, np.int32); bd = np.zeros(n, np.int32)
        check(lib().orb_search_window_best(self._h, C.byref(tv), C.byref(q), ptr(bi), ptr(bd)), "orb_search_window_best")
        return bi, bd

    def FuseCandidates(self, KF, active, u, v, pred_level, mp_desc, th=2.5):
        """Scoring loop of ORBmatcher::Fuse (src/ORBmatcher.cc:1016-1134; the Scw form :1136-1265 is the same after projection):
        for every projected map point the KF keypoint it fuses with (bestDist <= TH_LOW), or -1.  Replace / AddObservation on the
        result is graph bookkeeping and stays with the caller."""
        lv = np.asarray(pred_level, np.int32)
        radius = (np.float32(th) * self._scale_factors(KF)[np.clip(lv, 0, KF.nlevels - 1)]).astype(np.float32)
        bi, bd = self.BestInWindow(KF, active, u, v, radius, lv - 1, lv, mp_desc)
        return np.where(bd <= self.TH_LOW, bi, -1).astype(np.int32)

    def SearchBySim3(self, KF1, KF2, act1, u12, v12, lvl12, desc1, act2, u21, v21, lvl21, desc2, th=7.5):
        """ORBmatcher::SearchBySim3 (src/ORBmatcher.cc:1267-1505).  act1/u12/v12/lvl12/desc1: map points of KF1 (one per KF1 keypoint)
        projected into KF2 with their predicted level there; act2/... the other way round.  Returns (nFound, matches12) where
        matches12[i1] = KF2 keypoint whose map point matches, or -1 (agreement test :1478-1493)."""
        l12 = np.asarray(lvl12, np.int32); l21 = np.asarray(lvl21, np.int32)
        r12 = (np.float32(th) * self._scale_factors(KF2)[np.clip(l12, 0, KF2.nlevels - 1)]).astype(np.float32)
        r21 = (np.float32(th) * self._scale_factors(KF1)[np.clip(l21, 0, KF1.nlevels - 1)]).astype(np.float32)
        b1, d1 = self.BestInWindow(KF2, act1, u12, v12, r12, l12 - 1, l12, desc1)
        b2, d2 = self.BestInWindow(KF1, act2, u21, v21, r21, l21 - 1, l21, desc2)
        m1 = np.where(d1 <= self.TH_HIGH, b1, -1)
        m2 = np.where(d2 <= self.TH_HIGH, b2, -1)
        out = np.full(len(m1), -1, np.int32)
        ok = m1 >= 0
        idx = np.flatnonzero(ok)
        agree = m2[m1[idx]] == idx
        out[idx[agree]] = m1[idx[agree]]
        return int(agree.sum()), out

    def SearchForTriangulation(self, fv1, desc1, kps1, has_mp1, fv2, desc2, kps2, has_mp2, F12, level_sigma2):
        """ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:852-1014).  fv = (node_id, start, items) CSR; has_mp: the feature
        already has a map point; F12 3x3 float32; level_sigma2 = pKF2's mvLevelSigma2.  Returns (nmatches, vMatchedPairs as N x 2)."""
        keep = []

        def fv(t):
            arrs = [np.ascontiguousarray(a, np.int32) for a in t]
            keep.append(arrs)
            return FeatVecView(len(arrs[0]), arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data)
        a, b = fv(fv1), fv(fv2)
        desc1 = np.ascontiguousarray(desc1, np.uint8); desc2 = np.ascontiguousarray(desc2, np.uint8)
        kps1 = np.ascontiguousarray(kps1, KP_DTYPE); kps2 = np.ascontiguousarray(kps2, KP_DTYPE)
        h1 = np.ascontiguousarray(has_mp1, np.uint8); h2 = np.ascontiguousarray(has_mp2, np.uint8)
        F = np.ascontiguousarray(F12, np.float32).reshape(9); sg = np.ascontiguousarray(level_sigma2, np.float32)
        m12 = np.full(len(kps1), -1, np.int32)
        n = C.c_int(0)
        check(lib().orb_search_for_triangulation(self._h, C.byref(a), ptr(desc1), ptr(kps1), ptr(h1), len(kps1), C.byref(b), ptr(desc2),
                                                 ptr(kps2), ptr(h2), len(kps2), ptr(F), ptr(sg), len(sg), int(self.mbCheckOrientation),
                                                 ptr(m12), C.byref(n)), "orb_search_for_triangulation")
        i1 = np.flatnonzero(m12 >= 0)
        return n.value, np.stack([i1, m12[i1]], 1).astype(np.int32), m12

    def ComputeDistinctiveDescriptors(self, desc, start):
        """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:185-250) for many map points: desc = all observation descriptors,
        start = CSR offsets per point.  Returns (BestIdx within each group or -1, BestMedian)."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        start = np.ascontiguousarray(start, np.int32)
        n = len(start) - 1
        bi = np.zeros(n, np.int32); bm = np.zeros(n, np.int32)
        check(lib().orb_distinctive_descriptors(self._h, ptr(desc), ptr(start), n, ptr(bi), ptr(bm)), "orb_distinctive_descriptors")
        return bi, bm

    def SearchByBoW(self, kf_featvec, kf_desc, kf_kps, kf_mp_valid, f_featvec, f_desc, f_kps):
        """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) scoring (src/ORBmatcher.cc:155-284).
        featvec = (node_id, start, items) CSR arrays.  Returns (nmatches, match_f)."""
        keep = []

        def fv(t):
            arrs = [np.ascontiguousarray(a, np.int32) for a in t]
            keep.append(arrs)
            return FeatVecView(len(arrs[0]), arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data)
        a, b = fv(kf_featvec), fv(f_featvec)
        kf_desc = np.ascontiguousarray(kf_desc, np.uint8)
        f_desc = np.ascontiguousarray(f_desc, np.uint8)
        kf_kps = np.ascontiguousarray(kf_kps, KP_DTYPE)
        f_kps = np.ascontiguousarray(f_kps, KP_DTYPE)
        valid = np.ascontiguousarray(kf_mp_valid, np.uint8)
        m = np.full(len(f_kps), -1, np.int32)
        n = C.c_int(0)
        check(lib().orb_search_by_bow(self._h, C.byref(a), ptr(kf_desc), ptr(kf_kps), ptr(valid), len(kf_kps),
                                      C.byref(b), ptr(f_desc), ptr(f_kps), len(f_kps), self.mfNNratio,
                                      int(self.mbCheckOrientation), ptr(m), C.byref(n)), "orb_search_by_bow")
        return n.value, m

    def SearchByBoWKeyFrames(self, fv1, desc1, kps1, valid1, fv2, desc2, kps2, valid2):
        """ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, ...) (src/ORBmatcher.cc:715-850).  Returns (nmatches, match12)."""
        keep = []

        def fv(t):
            arrs = [np.ascontiguousarray(a, np.int32) for a in t]
            keep.append(arrs)
            return FeatVecView(len(arrs[0]), arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data)
        a, b = fv(fv1), fv(fv2)
        desc1 = np.ascontiguousarray(desc1, np.uint8); desc2 = np.ascontiguousarray(desc2, np.uint8)
        kps1 = np.ascontiguousarray(kps1, KP_DTYPE); kps2 = np.ascontiguousarray(kps2, KP_DTYPE)
        v1 = np.ascontiguousarray(valid1, np.uint8); v2 = np.ascontiguousarray(valid2, np.uint8)
        m = np.full(len(kps1), -1, np.int32)
        n = C.c_int(0)
        check(lib().orb_search_by_bow_kf(self._h, C.byref(a), ptr(desc1), ptr(kps1), ptr(v1), len(kps1),
                                         C.byref(b), ptr(desc2), ptr(kps2), ptr(v2), len(kps2), self.mfNNratio,
                                         int(self.mbCheckOrientation), ptr(m), C.byref(n)), "orb_search_by_bow_kf")
        return n.value, m
