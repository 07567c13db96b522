"""Latency of one ORBmatcher::SearchByProjection(Frame&, Frame&, th) call (src/ORBmatcher.cc:1507-1620, the second per-frame call of
Tracking::TrackWithMotionModel) through the C ABI: host buffers, device-resident views, and the 640x480 / 1241x376 shapes.
   gpurun -- 'python tools/latency_sbp.py'            (add `ncu --metrics gpu__time_duration.sum ...` in front for per-kernel times)"""
import sys, time, ctypes as C
import numpy as np
sys.path.insert(0, ".")
import torch
import orbslam_jpminipc_b200 as pkg
from orbslam_jpminipc_b200._lib import check, lib, ptr, FrameView
from orbslam_jpminipc_b200.synth import shifted_frame, synth_frame

L = lib()
REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 200


def med(ts):
    return "%.1f us (p10 %.1f p90 %.1f)" % (np.median(ts) * 1e6, np.percentile(ts, 10) * 1e6, np.percentile(ts, 90) * 1e6)


for (h, w, nf) in [(480, 640, 1000), (376, 1241, 2000)]:
    ex = pkg.ORBextractor(nf, 1.2, 8, 1, 20, max_width=w, max_height=h, max_batch=2)
    fa = synth_frame(h, w, 9000, quadrants=False)
    fb = shifted_frame(fa, 3, 2, 9001)
    (ka, da), (kb, db_) = ex.extract_batch(np.stack([fa, fb]))
    m = pkg.ORBmatcher(0.9, True, extractor=ex)
    fx = fy = 500.0
    rng = np.random.default_rng(9000)
    z = rng.uniform(2, 10, len(ka)).astype(np.float32)
    xyz = np.stack([(ka["x"] - w / 2) / fx * z, (ka["y"] - h / 2) / fy * z, z], 1).astype(np.float32)
    Tcw = np.eye(4, dtype=np.float32); Tcw[:3, 3] = [0.03, 0.02, 0.01]
    has, outl = np.ones(len(ka), np.uint8), np.zeros(len(ka), np.uint8)
    cur = pkg.Frame(m, kb, db_, w, h, fx, fy, w / 2, h / 2)
    last = pkg.Frame(m, ka, da, w, h, fx, fy, w / 2, h / 2)
    nm, mt = m.SearchByProjection(cur, last, 15.0, has, outl, xyz, Tcw)
    ts = []
    for _ in range(REPS):
        t0 = time.perf_counter(); m.SearchByProjection(cur, last, 15.0, has, outl, xyz, Tcw); ts.append(time.perf_counter() - t0)
    print((h, w, nf), "matches", nm, "| host buffers (python wrapper):", med(ts))
    # the same call on device-resident views (what a pipeline that keeps the extractor's outputs in HBM passes)
    keep = []

    def dv(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda(); keep.append(t); return t.data_ptr()

    def view(F):
        return FrameView(F.N, dv(F.kps.view(np.uint8)), dv(F.desc), F.fx, F.fy, F.cx, F.cy, F.bounds[0], F.bounds[1], F.bounds[2], F.bounds[3],
                         F.nlevels, F.scale_factor, dv(F.cell_start), dv(F.cell_items))
    vc, vl = view(cur), view(last)
    d_has, d_out, d_xyz = dv(has), dv(outl), dv(xyz)
    T = np.ascontiguousarray(Tcw, np.float32).reshape(16)
    d_match = torch.full((cur.N,), -1, dtype=torch.int32, device="cuda")
    n = C.c_int(0)

    def call_dev():
        d_match.fill_(-1)
        check(L.orb_search_by_projection(ex._h, C.byref(vc), C.byref(vl), C.c_void_p(d_has), C.c_void_p(d_out), C.c_void_p(d_xyz), ptr(T),
                                         C.c_float(15.0), 1, ptr(d_match), C.byref(n)), "orb_search_by_projection")
    for _ in range(5): call_dev()
    assert n.value == nm and np.array_equal(d_match.cpu().numpy(), mt)
    ts = []
    for _ in range(REPS):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); call_dev(); ts.append(time.perf_counter() - t0)
    print((h, w, nf), "| device views (incl. a torch fill + the result read-back):", med(ts))
    ex.close()

# ---- ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:49-125): the local-map search of
# Tracking::SearchReferencePointsInFrustum, once per tracked frame: ~1500 projected map points against one 640x480 frame
h, w, nf = 480, 640, 1000
ex = pkg.ORBextractor(nf, 1.2, 8, 1, 20, max_width=w, max_height=h, max_batch=2)
fa = synth_frame(h, w, 9000, quadrants=False)
fb = shifted_frame(fa, 3, 2, 9001)
(ka, da), (kb, db_) = ex.extract_batch(np.stack([fa, fb]))
m = pkg.ORBmatcher(0.8, True, extractor=ex)
F = pkg.Frame(m, kb, db_, w, h, 500.0, 500.0, w / 2, h / 2)
rng = np.random.default_rng(1)
NMP = 1500
src = rng.integers(0, len(ka), NMP)
px = (ka["x"][src] + 3 + rng.normal(0, 1.5, NMP)).astype(np.float32); py = (ka["y"][src] + 2 + rng.normal(0, 1.5, NMP)).astype(np.float32)
lvl = ka["octave"][src].astype(np.int32)
in_view = ((px > 0) & (px < w) & (py > 0) & (py < h)).astype(np.uint8)
vcos = rng.uniform(0.9, 1.0, NMP).astype(np.float32)
n1, mt1 = m.SearchByProjectionMapPoints(F, in_view, px, py, lvl, vcos, da[src], 3.0)
ts = []
for _ in range(REPS):
    t0 = time.perf_counter(); m.SearchByProjectionMapPoints(F, in_view, px, py, lvl, vcos, da[src], 3.0); ts.append(time.perf_counter() - t0)
print("SearchByProjection(F, %d map points, th 3) on a 640x480 / %d-keypoint frame: matches %d | host buffers (python wrapper): %s" % (NMP, F.N, n1, med(ts)))
ex.close()
