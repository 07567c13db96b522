import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import orbslam_jpminipc_b200 as pkg
from oracle import pyoracle as po
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_match as T
m = pkg.ORBmatcher(0.9, True)
gcur, glast, ocur, olast, has, outl, xyz, Tm = T._scene(po, pkg, m, 376, 1241, 2000, 9000, 15.0)
def t(f, n=20):
    f(); t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e3
print("keypoints", gcur.N, glast.N)
print("GPU SearchByProjection (host buffers) ms", t(lambda: m.SearchByProjection(gcur, glast, 15.0, has, outl, xyz, Tm)))
print("CPU oracle SearchByProjection ms", t(lambda: po.search_by_projection(ocur, olast, has, outl, xyz, Tm, 15.0, True)))
case = T._bow_case(po, pkg, 2000, 2000, 100, seed=5)
mb = pkg.ORBmatcher(0.75, True)
print("GPU SearchByBoW ms", t(lambda: mb.SearchByBoW(*case)))
print("CPU oracle SearchByBoW ms", t(lambda: po.search_by_bow(*case, 0.75, True)))
from orbslam_jpminipc_b200.synth import synth_descriptors
db, q = synth_descriptors(2000, 2000)
print("GPU knn2 2000x2000 (host buffers) ms", t(lambda: m.knn2(q, db)))
print("CPU oracle knn2 2000x2000 popcnt ms", t(lambda: po.knn2(q, db), 3))
print("GPU grid build ms", t(lambda: pkg.Frame(m, gcur.kps, gcur.desc, 1241, 376, 500, 500, 620, 188)))
