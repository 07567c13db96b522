"""Small single-frame / small-call workload for compute-sanitizer (memcheck, racecheck, synccheck) over the latency forms of the pass and the
matcher kernels:   gpurun -- 'compute-sanitizer --tool racecheck python tools/sanitize_small.py'"""
import sys
import numpy as np
sys.path.insert(0, ".")
import orbslam_jpminipc_b200 as pkg
from orbslam_jpminipc_b200.synth import shifted_frame, synth_frame

h, w, nf = 240, 320, 300
ex = pkg.ORBextractor(nf, 1.2, 8, 1, 20, max_width=w, max_height=h, max_batch=6)
fa = synth_frame(h, w, 9000, quadrants=False)
fb = shifted_frame(fa, 3, 2, 9001)
for n in (1, 2, 6):                                   # small-call forms (1, 2) and the batch forms (6)
    res = ex.extract_batch(np.stack([fa, fb] * 3)[:n])
(ka, da), (kb, db_) = res[0], res[1]
noise = np.random.default_rng(1).integers(0, 256, (h, w), dtype=np.uint8)
ex2 = pkg.ORBextractor(20, 1.2, 2, 1, 20, max_width=w, max_height=h, max_batch=1)    # few huge cells: the overflow path of k_cell_compact_wide
ex2(noise)
m = pkg.ORBmatcher(0.9, True, extractor=ex)
z = np.random.default_rng(2).uniform(2, 10, len(ka)).astype(np.float32)
xyz = np.stack([(ka["x"] - w / 2) / 500.0 * z, (ka["y"] - h / 2) / 500.0 * z, z], 1).astype(np.float32)
Tcw = np.eye(4, dtype=np.float32); Tcw[:3, 3] = [0.03, 0.02, 0.01]
cur = pkg.Frame(m, kb, db_, w, h, 500.0, 500.0, w / 2, h / 2)
last = pkg.Frame(m, ka, da, w, h, 500.0, 500.0, w / 2, h / 2)
nm, mt = m.SearchByProjection(cur, last, 15.0, np.ones(len(ka), np.uint8), np.zeros(len(ka), np.uint8), xyz, Tcw)
print("keypoints", len(ka), len(kb), "matches", nm)
