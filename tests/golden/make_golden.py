#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the REAL OpenCV (python cv2 4.13.0).

The reference (caomw/ORBSLAM_jpMiniPC) ships no tests or golden vectors and cannot be
compiled in the build container; its arithmetic lives in OpenCV.  This script therefore
re-composes the control flow of src/ORBextractor.cc in python around *real* cv2 calls
(cv2.resize, cv2.copyMakeBorder, cv2.FastFeatureDetector, cv2.sepFilter2D with the CV_32F
Gaussian taps = the path GaussianBlur takes on a non-isolated sub-matrix, cv2.fastAtan2,
cv2.gemm) and records inputs + outputs.  std::nth_element (libstdc++ 13, part of the spec)
is reached through the oracle's tiny helper.  It also records cv2.ORB(nlevels=1) runs, which
exercise OpenCV's own C++ retainBest / IC_Angle / sub-matrix blur / rBRIEF.

Run in the build container only (needs cv2):  python tests/golden/make_golden.py
The .npz files are committed; tests read them without cv2 and without /root/reference.
"""
import os
import sys
import numpy as np
import cv2

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po                       # noqa: E402  (nth_element helper only)
from orbslam_jpminipc_b200.synth import synth_frame      # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32
cv2.setNumThreads(1)

PATTERN = np.array([int(t) for l in open(os.path.join(ROOT, "oracle", "orb_pattern.inc"))
                    if not l.startswith("//") for t in l.replace(",", " ").split()], np.int32).reshape(512, 2)
UMAX = [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
EDGE = 16


def tables(nfeatures, scale, nlevels):
    sf64 = float(f32(scale))
    mvScale = [f32(1)]
    for i in range(1, nlevels):
        mvScale.append(f32(float(mvScale[-1]) * sf64))
    inv = f32(1.0 / sf64)
    mvInv = [f32(1)]
    for i in range(1, nlevels):
        mvInv.append(f32(mvInv[-1] * inv))
    factor = f32(1.0 / sf64)
    nd = f32(f32(f32(nfeatures) * f32(f32(1) - factor)) / f32(f32(1) - f32(float(factor) ** nlevels)))
    per = []
    for l in range(nlevels - 1):
        per.append(int(np.rint(nd)))
        nd = f32(nd * factor)
    per.append(max(nfeatures - sum(per), 0))
    return mvScale, mvInv, per


def ic_angle(plane, x, y):
    m01 = m10 = 0
    for v in range(-15, 16):
        d = UMAX[abs(v)]
        row = plane[y + v, x - d:x + d + 1].astype(np.int64)
        m10 += int((np.arange(-d, d + 1) * row).sum())
        m01 += v * int(row.sum())
    return cv2.fastAtan2(float(m01), float(m10))


def rbrief(plane, x, y, angle_deg):
    ang = f32(f32(angle_deg) * f32(np.pi / f32(180.0)))
    a = f32(np.cos(np.float64(ang)))
    b = f32(np.sin(np.float64(ang)))
    px = PATTERN[:, 0].astype(f32)
    py = PATTERN[:, 1].astype(f32)
    ry = np.rint(px * b + py * a).astype(np.int64)
    rx = np.rint(px * a - py * b).astype(np.int64)
    v = plane[y + ry, x + rx].astype(np.int32)
    bits = (v[0::2] < v[1::2]).astype(np.uint8).reshape(32, 8)
    return (bits << np.arange(8, dtype=np.uint8)).sum(axis=1).astype(np.uint8)


def retain_first_n(resp, n):
    """first n after std::nth_element(begin, begin+n-1, end, response>)  (OpenCV-4 retainBest + resize(n))"""
    if len(resp) <= n:
        return np.arange(len(resp))
    if n == 0:
        return np.arange(0)
    _, idx = po.nth_element_desc(np.asarray(resp, f32), n - 1)
    return idx[:n]


def cv2_composed_extract(img, nfeatures=1000, scale=1.2, nlevels=8, fast_th=20):
    """src/ORBextractor.cc:718-822 re-composed around real cv2 primitives."""
    mvScale, mvInv, per = tables(nfeatures, scale, nlevels)
    h0, w0 = img.shape
    planes = []
    for l in range(nlevels):
        if l == 0:
            roi = img
        else:
            sz = (int(np.rint(f32(w0) * mvInv[l])), int(np.rint(f32(h0) * mvInv[l])))
            roi = cv2.resize(planes[l - 1][EDGE:-EDGE, EDGE:-EDGE], sz, interpolation=cv2.INTER_LINEAR)
        planes.append(cv2.copyMakeBorder(roi, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_REFLECT_101))
    ratio = f32(w0) / f32(h0)
    det = {th: cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True,
                                              type=cv2.FastFeatureDetector_TYPE_9_16) for th in (fast_th, 7)}
    out_kp, out_desc, stage = [], [], {}
    for l in range(nlevels):
        plane = planes[l]
        roi = plane[EDGE:-EDGE, EDGE:-EDGE]
        h, w = roi.shape
        nDes = per[l]
        cols = int(np.sqrt(f32(nDes) / f32(f32(5) * ratio)))
        rows = int(f32(ratio * f32(cols)))
        W, H = w - 2 * EDGE, h - 2 * EDGE
        cellW = int(np.ceil(f32(W) / f32(cols)))
        cellH = int(np.ceil(f32(H) / f32(rows)))
        nCells = rows * cols
        nfc = int(np.ceil(f32(nDes) / f32(nCells)))
        cells = {}
        nTot = np.zeros((rows, cols), int)
        nRet = np.zeros((rows, cols), int)
        noMore = np.zeros((rows, cols), bool)
        iniX = [0] * cols
        iniY = [0] * rows
        nNoMore = nDist = 0
        hY = cellH + 6
        for i in range(rows):
            iniY[i] = EDGE + i * cellH - 3
            if i == rows - 1:
                hY = h - EDGE + 3 - iniY[i]
                if hY <= 0:
                    continue
            hX = cellW + 6
            for j in range(cols):
                if i == 0:
                    iniX[j] = EDGE + j * cellW - 3
                if j == cols - 1:
                    hX = w - EDGE + 3 - iniX[j]
                    if hX <= 0:
                        continue
                cell = np.ascontiguousarray(roi[iniY[i]:iniY[i] + hY, iniX[j]:iniX[j] + hX])
                k = det[fast_th].detect(cell)
                if len(k) <= 3:
                    k = det[7].detect(cell)
                cells[(i, j)] = [(int(p.pt[0]), int(p.pt[1]), int(p.response)) for p in k]
                nTot[i, j] = len(k)
                if len(k) > nfc:
                    nRet[i, j] = nfc
                else:
                    nRet[i, j] = len(k)
                    nDist += nfc - len(k)
                    noMore[i, j] = True
                    nNoMore += 1
        while nDist > 0 and nNoMore < nCells:
            nNew = nfc + int(np.ceil(f32(nDist) / f32(nCells - nNoMore)))
            nDist = 0
            for i in range(rows):
                for j in range(cols):
                    if not noMore[i, j]:
                        if nTot[i, j] > nNew:
                            nRet[i, j] = nNew
                        else:
                            nRet[i, j] = nTot[i, j]
                            nDist += nNew - nTot[i, j]
                            noMore[i, j] = True
                            nNoMore += 1
        lev = []
        for i in range(rows):
            for j in range(cols):
                c = cells.get((i, j), [])
                keep = retain_first_n([s for _, _, s in c], int(nRet[i, j]))
                lev += [(c[t][0] + iniX[j], c[t][1] + iniY[i], c[t][2]) for t in keep]
        if len(lev) > nDes:
            keep = retain_first_n([s for _, _, s in lev], nDes)
            lev = [lev[t] for t in keep]
        stage["L%d_cand" % l] = np.array([(i * cols + j, x, y, s) for (i, j), c in sorted(cells.items())
                                          for (x, y, s) in c], np.int32).reshape(-1, 4)
        stage["L%d_quota" % l] = np.stack([nTot.ravel(), nRet.ravel()]).astype(np.int32)
        angles = [ic_angle(plane, x + EDGE, y + EDGE) for x, y, _ in lev]
        blurred = plane.copy()
        if lev:
            k = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
            blurred[EDGE:-EDGE, EDGE:-EDGE] = cv2.sepFilter2D(plane, cv2.CV_8U, k, k,
                                                              borderType=cv2.BORDER_REFLECT_101)[EDGE:-EDGE, EDGE:-EDGE]
        stage["L%d_plane" % l] = plane
        stage["L%d_blur" % l] = blurred
        size = float(int(f32(31) * mvScale[l]))
        for (x, y, s), a in zip(lev, angles):
            out_desc.append(rbrief(blurred, x + EDGE, y + EDGE, a))
            fx, fy = f32(x), f32(y)
            if l != 0:
                fx, fy = f32(fx * mvScale[l]), f32(fy * mvScale[l])
            out_kp.append((fx, fy, size, a, float(s), l, -1))
    kps = np.array(out_kp, po.KP_DTYPE)
    desc = np.array(out_desc, np.uint8).reshape(-1, 32)
    return kps, desc, stage, (mvScale, mvInv, per)


def golden_extract():
    cases = [("e2e_320x240_n300", 240, 320, 300, 2000),
             ("e2e_640x480_n1000", 480, 640, 1000, 1000),
             ("e2e_620x188_n700", 188, 620, 700, 2002)]
    for name, h, w, nf, seed in cases:
        img = synth_frame(h, w, seed)
        kps, desc, stage, (sc, inv, per) = cv2_composed_extract(img, nf)
        keep = {k: v for k, v in stage.items() if k.endswith("_quota") or k.endswith("_cand")}
        # planes are bulky: keep two levels' planes as spot checks
        for l in (1, 4):
            keep["L%d_plane" % l] = stage["L%d_plane" % l]
            keep["L%d_blur" % l] = stage["L%d_blur" % l]
        np.savez_compressed(os.path.join(OUT, name + ".npz"), img=img, nfeatures=nf, kps=kps, desc=desc,
                            scale=np.array(sc, f32), inv_scale=np.array(inv, f32), per_level=np.array(per), **keep)
        print(name, "kp", len(kps), "per level", np.bincount(kps["octave"], minlength=8))


def golden_primitives():
    rng = np.random.default_rng(7)
    img = synth_frame(120, 160, 3000)
    d = {"img": img}
    for i, (dw, dh) in enumerate([(133, 100), (111, 83), (160, 120), (77, 59)]):
        d["resize_%d" % i] = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
    d["border16"] = cv2.copyMakeBorder(img, 16, 16, 16, 16, cv2.BORDER_REFLECT_101)
    for th in (20, 7):
        det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True,
                                             type=cv2.FastFeatureDetector_TYPE_9_16)
        k = det.detect(img)
        d["fast_%d" % th] = np.array([(p.pt[0], p.pt[1], p.response) for p in k], np.int32).reshape(-1, 3)
    k = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    d["gauss_taps_bits"] = k.ravel().view(np.uint32)
    d["blur_f32"] = cv2.sepFilter2D(d["border16"], cv2.CV_8U, k, k, borderType=cv2.BORDER_REFLECT_101)[16:-16, 16:-16]
    d["blur_fixed256"] = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    yx = rng.integers(-2900000, 2900000, (4000, 2))
    yx[:40, 0] = 0
    yx[20:60, 1] = 0
    yx[100:140, 0] = yx[100:140, 1]
    d["atan2_yx"] = yx.astype(np.int32)
    d["atan2_deg_bits"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], f32).view(np.uint32)
    # cv::gemm 3x3 * 3x1 + 3x1 on CV_32F (Rcw*x3Dw+tcw, src/ORBmatcher.cc:1530)
    T = np.zeros((64, 4, 4), f32)
    X = (rng.standard_normal((64, 3, 1)) * 5).astype(f32)
    Y = np.zeros((64, 3), f32)
    for i in range(64):
        T[i] = np.eye(4)
        T[i, :3, :3] = rng.standard_normal((3, 3))
        T[i, :3, 3] = rng.standard_normal(3) * 3
        Y[i] = cv2.gemm(T[i, 0:3, 0:3], X[i], 1.0, T[i, 0:3, 3:4], 1.0).ravel()
    d["gemm_T"], d["gemm_X"], d["gemm_Y_bits"] = T, X, Y.view(np.uint32)
    np.savez_compressed(os.path.join(OUT, "primitives.npz"), **d)
    print("primitives ok")


def golden_cv2_orb():
    """cv2.ORB(nlevels=1): OpenCV's own retainBest order, IC_Angle, sub-matrix blur and rBRIEF."""
    d = {}
    for t, (seed, n, h, w) in enumerate([(3100, 120, 160, 200), (3101, 37, 120, 160), (3102, 400, 200, 260)]):
        img = synth_frame(h, w, seed, quadrants=bool(t % 2))
        orb = cv2.ORB_create(nfeatures=n, scaleFactor=1.2, nlevels=1, edgeThreshold=31, firstLevel=0, WTA_K=2,
                             scoreType=cv2.ORB_FAST_SCORE, patchSize=31, fastThreshold=20)
        kps, desc = orb.detectAndCompute(img, None)
        d["img_%d" % t] = img
        d["n_%d" % t] = n
        d["kp_%d" % t] = np.array([(k.pt[0], k.pt[1], k.response) for k in kps], np.int32).reshape(-1, 3)
        d["angle_bits_%d" % t] = np.array([k.angle for k in kps], f32).view(np.uint32)
        d["desc_%d" % t] = desc
    np.savez_compressed(os.path.join(OUT, "cv2_orb_single_level.npz"), **d)
    print("cv2 ORB KAT ok")


if __name__ == "__main__":
    golden_primitives()
    golden_cv2_orb()
    golden_extract()
