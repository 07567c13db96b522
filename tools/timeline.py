"""Timeline of ONE single-frame extraction pass in its production form (replayed graph + programmatic dependent launch), from
%globaltimer stamps taken at the top of every kernel (diagnostic build only):
   make -C orbslam_jpminipc_b200/csrc OUT=../liborb_b200_tl.so EXTRA=-DORB_TIMELINE
   gpurun -- 'ORB_B200_LIB=$PWD/orbslam_jpminipc_b200/liborb_b200_tl.so python tools/timeline.py'
per kernel: arrival of its first CTA (scheduled, parked in griddepcontrol.wait) and the time its dependency resolved (work starts)."""
import ctypes as C, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import orbslam_jpminipc_b200 as pkg
from orbslam_jpminipc_b200._lib import check, lib, ptr
from orbslam_jpminipc_b200.synth import synth_frames

L = lib()
L.orb_debug_timeline.argtypes = [C.c_void_p, C.c_int]
H, W, NF = 480, 640, 1000
NAMES = {1: "k_level0", 2: "k_resize (first .. last level)", 3: "k_border", 4: "k_fast_nms", 5: "k_cell_compact", 6: "k_harris", 7: "k_select",
         8: "k_blur", 9: "k_describe"}
frames = torch.from_numpy(np.stack(synth_frames(4, H, W, seed0=1000))).pin_memory()
ex = pkg.ORBextractor(NF, 1.2, 8, 1, 20, max_width=W, max_height=H, max_batch=1)
cap = ex.capacity
pk = torch.zeros((cap, 7), dtype=torch.int32).pin_memory(); pd = torch.zeros((cap, 32), dtype=torch.uint8).pin_memory()
n = C.c_int(0)
for i in range(20):
    check(L.orb_extract(ex._h, ptr(frames[i % 4]), W, H, W, ptr(pk), ptr(pd), cap, C.byref(n)), "orb_extract")
buf = (C.c_ulonglong * 64)()
for rep in range(4):
    L.orb_debug_timeline(None, 1)
    t0 = time.perf_counter()
    check(L.orb_extract(ex._h, ptr(frames[rep % 4]), W, H, W, ptr(pk), ptr(pd), cap, C.byref(n)), "orb_extract")
    wall = (time.perf_counter() - t0) * 1e6
    L.orb_debug_timeline(buf, 0)
    t = np.array(buf[:], dtype=np.uint64).reshape(16, 4)
    base = int(t[1, 0])
    print("pass %d: blocking call %.1f us (pinned buffers); microseconds after the arrival of k_level0's first CTA:" % (rep, wall))
    for i, name in NAMES.items():
        if t[i, 2] == 0: continue
        a0, s0, a1, s1 = [(int(v) - base) / 1e3 for v in t[i]]
        print("   %-32s first CTA arrives %7.1f  starts %7.1f | last CTA arrives %7.1f  starts %7.1f" % (name, a0, s0, a1, s1))
    for i, name in {10: "k_select: quotas done", 11: "k_select: cells done", 12: "k_select: level cut done", 13: "k_select: lists written", 14: "k_describe: last warp done"}.items():
        if t[i, 3]: print("   %-32s %7.1f" % (name, (int(t[i, 3]) - base) / 1e3))
ex.close()
