"""Blocking orb_extract_batch calls of N = 1..32 frames (stereo / multi-camera rigs) with the small-call forms on (limit >= N) and off:
where should `small_call_frames` / `pdl_frames` sit?   gpurun -- 'python tools/latency_small_batch.py'"""
import os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, ".")
import torch
import orbslam_jpminipc_b200 as pkg
from orbslam_jpminipc_b200._lib import check, lib, ptr
from orbslam_jpminipc_b200.synth import synth_frames

L = lib()
H, W, NF = 480, 640, 1000
frames = torch.from_numpy(np.stack(synth_frames(32, H, W, seed0=1000))).pin_memory()


def ctx(env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    ex = pkg.ORBextractor(NF, 1.2, 8, 1, 20, max_width=W, max_height=H, max_batch=4 if "--hd" in sys.argv else 32)
    for k, v in old.items():
        if v is None: del os.environ[k]
        else: os.environ[k] = v
    return ex


variants = {"forms on (limit 64)": {"ORB_SMALL_CALL": "64", "ORB_PDL_FRAMES": "64"},
            "kernels on, PDL off": {"ORB_SMALL_CALL": "64", "ORB_PDL_FRAMES": "0"},
            "kernels off, PDL on": {"ORB_SMALL_CALL": "0", "ORB_PDL_FRAMES": "64", "ORB_SELECT_WIDE": "0"},
            "forms off": {"ORB_SMALL_CALL": "0", "ORB_PDL_FRAMES": "0", "ORB_SELECT_WIDE": "0"}}
if "--each" in sys.argv:        # which of the kernel forms pays at which call size (PDL off throughout)
    base = {"ORB_SMALL_CALL": "64", "ORB_PDL_FRAMES": "0"}
    variants = {"all on": dict(base), "FAST 128 threads": dict(base, ORB_FAST_WIDE="0"), "compaction per warp": dict(base, ORB_COMPACT_WIDE="0"),
                "selection 8 warps": dict(base, ORB_SELECT_WIDE="0"), "border on main stream": dict(base, ORB_SIDE_BORDER="0"),
                "resize 8 rows": dict(base, ORB_RESIZE_ROWS_SMALL="8")}
if "--forms4" in sys.argv:      # FAST CTA size x resize tiling, forced
    base = {"ORB_SMALL_CALL": "64", "ORB_PDL_FRAMES": "2"}
    variants = {"FAST 512 + short tiles": dict(base, ORB_FAST_WIDE="2", ORB_RESIZE_ROWS_SMALL="2"), "FAST 512 + 8-row tiles": dict(base, ORB_FAST_WIDE="2", ORB_RESIZE_ROWS_SMALL="8"),
                "FAST 128 + short tiles": dict(base, ORB_FAST_WIDE="0", ORB_RESIZE_ROWS_SMALL="2"), "FAST 128 + 8-row tiles": dict(base, ORB_FAST_WIDE="0", ORB_RESIZE_ROWS_SMALL="8")}
if "--hd" in sys.argv:
    H, W = 1080, 1920
    frames = torch.from_numpy(np.stack(synth_frames(4, H, W, seed0=1000))).pin_memory()
if "--noise" in sys.argv:       # dense noise: every FAST tile is heavy, thousands of candidates per cell
    frames = torch.from_numpy(np.random.default_rng(3).integers(0, 256, tuple(frames.shape), dtype=np.uint8)).pin_memory()
exs = {k: ctx(v) for k, v in variants.items()}
cap = next(iter(exs.values())).capacity
for n in ((1, 2, 4) if "--hd" in sys.argv else (1, 2, 4, 6, 8, 12, 16, 32)):
    k = torch.zeros((n, cap, 7), dtype=torch.int32).pin_memory(); d = torch.zeros((n, cap, 32), dtype=torch.uint8).pin_memory()
    c = torch.zeros(n, dtype=torch.int32).pin_memory()
    T = {a: [] for a in exs}
    for rnd in range(8 if "--forms4" in sys.argv else 4):
        for a, ex in exs.items():
            for i in range(30):
                t0 = time.perf_counter()
                check(L.orb_extract_batch(ex._h, ptr(frames), n, W, H, W, W * H, ptr(k), ptr(d), cap, ptr(c)), "orb_extract_batch")
                if i >= 5: T[a].append(time.perf_counter() - t0)
    print("frames per call", n, "| " + " | ".join("%s: %.1f us" % (a, np.median(T[a]) * 1e6) for a in exs))
for ex in exs.values(): ex.close()
