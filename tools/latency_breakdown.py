"""Where the single-frame latency of orb_extract goes (the call the reference makes once per camera frame, src/Frame.cc:60):
host API with pageable / pinned buffers, device-pointer API timed with CUDA events (the kernel chain alone), each with the
CUDA-graph replay on / off and the fused pyramid on / off.
   gpurun -- 'python tools/latency_breakdown.py'"""
import os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, ".")
import torch
import orbslam_jpminipc_b200 as pkg
from orbslam_jpminipc_b200._lib import check, lib, ptr
from orbslam_jpminipc_b200.synth import synth_frames

L = lib()
H, W, NF = 480, 640, 1000
frames = np.stack(synth_frames(8, H, W, seed0=1000))


def ctx(env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    ex = pkg.ORBextractor(NF, 1.2, 8, 1, 20, max_width=W, max_height=H, max_batch=1)
    for k, v in old.items():
        if v is None: del os.environ[k]
        else: os.environ[k] = v
    return ex


def med(ts):
    return "%.1f us (p10 %.1f p90 %.1f)" % (np.median(ts) * 1e6, np.percentile(ts, 10) * 1e6, np.percentile(ts, 90) * 1e6)


ENVS = [dict(kv.split("=") for kv in a.split(",") if kv) for a in sys.argv[1:] if not a.startswith("--")] or [{"ORB_GRAPH": "1"}, {"ORB_GRAPH": "0"}]


def stages(env, h=H, w=W, nf=NF):
    """per-stage device time of ONE frame (profiling mode: stage-boundary events, blur not forked)"""
    fr = np.stack(synth_frames(4, h, w, seed0=1000))
    ex = ctx(env) if (h, w, nf) == (H, W, NF) else pkg.ORBextractor(nf, 1.2, 8, 1, 20, max_width=w, max_height=h, max_batch=1)
    cap = ex.capacity
    df = torch.from_numpy(fr).cuda(); dk = torch.zeros((1, cap, 7), dtype=torch.int32, device="cuda")
    dd = torch.zeros((1, cap, 32), dtype=torch.uint8, device="cuda"); dc = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream()
    L.orb_profile_enable(ex._h, 1)
    nstage = 7
    for i in range(10): ex.extract_batch_device(df[i % 4:i % 4 + 1], dk, dd, dc, st.cuda_stream)
    torch.cuda.synchronize()
    check(L.orb_profile_read(ex._h, (C.c_double * nstage)(), C.byref(C.c_int(0))), "orb_profile_read")
    R = 100
    for i in range(R): ex.extract_batch_device(df[i % 4:i % 4 + 1], dk, dd, dc, st.cuda_stream)
    ms = (C.c_double * nstage)(); nc = C.c_int(0)
    check(L.orb_profile_read(ex._h, ms, C.byref(nc)), "orb_profile_read")
    L.orb_profile_enable(ex._h, 0)
    print(env, (h, w, nf), "stage us per frame:", {L.orb_profile_stage_name(i).decode(): round(ms[i] / R * 1e3, 1) for i in range(nstage)},
          "sum %.1f" % (sum(ms) / R * 1e3))
    ex.close()


# host API: every variant gets its context first, then the variants are measured in interleaved rounds (order / warm-up effects of
# the host path are as large as the differences looked for)
exs = [ctx(env) for env in ENVS]
cap = exs[0].capacity
k = np.zeros(cap, pkg.KP_DTYPE); d = np.zeros((cap, 32), np.uint8); n = C.c_int(0)
pf = torch.from_numpy(frames).pin_memory()
pk = torch.zeros((cap, 7), dtype=torch.int32).pin_memory(); pd = torch.zeros((cap, 32), dtype=torch.uint8).pin_memory()
df = pf.cuda(); dk = torch.zeros((1, cap, 7), dtype=torch.int32, device="cuda"); dd = torch.zeros((1, cap, 32), dtype=torch.uint8, device="cuda")
dc = torch.zeros(1, dtype=torch.int32, device="cuda")
st = torch.cuda.Stream()
T = {(i, kind): [] for i in range(len(exs)) for kind in ("pageable", "pinned", "device", "device_wall", "h2d_only", "d2h_only")}
pc = torch.zeros(1, dtype=torch.int32).pin_memory()


def call(ex, img, k, d):
    check(L.orb_extract(ex._h, ptr(img), W, H, W, ptr(k), ptr(d), cap, C.byref(n)), "orb_extract")


for rnd in range(6):
    for i, ex in enumerate(exs):
        for j in range(10): call(ex, frames[j % 8], k, d)
        for j in range(60):
            t0 = time.perf_counter(); call(ex, frames[j % 8], k, d); T[i, "pageable"].append(time.perf_counter() - t0)
        for j in range(10): call(ex, pf[j % 8], pk, pd)
        for j in range(60):
            t0 = time.perf_counter(); call(ex, pf[j % 8], pk, pd); T[i, "pinned"].append(time.perf_counter() - t0)
        # blocking host-API call with only one side on the host: pinned input + device outputs, device input + pinned outputs
        def mixed(img, ok, od, oc):
            check(L.orb_extract_batch(ex._h, ptr(img), 1, W, H, W, W * H, ptr(ok), ptr(od), cap, ptr(oc)), "orb_extract_batch")
        for kind, args in (("h2d_only", lambda j: (pf[j % 8], dk, dd, dc)), ("d2h_only", lambda j: (df[j % 8], pk, pd, pc))):
            for j in range(10): mixed(*args(j))
            for j in range(40):
                t0 = time.perf_counter(); mixed(*args(j)); T[i, kind].append(time.perf_counter() - t0)
        with torch.cuda.stream(st):
            for j in range(10): ex.extract_batch_device(df[j % 8:j % 8 + 1], dk, dd, dc, st.cuda_stream)
            st.synchronize()
            for j in range(40):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                t0 = time.perf_counter()
                e0.record(st); ex.extract_batch_device(df[j % 8:j % 8 + 1], dk, dd, dc, st.cuda_stream); e1.record(st)
                st.synchronize()
                T[i, "device_wall"].append(time.perf_counter() - t0); T[i, "device"].append(e0.elapsed_time(e1) * 1e-3)
for i, env in enumerate(ENVS):
    print(env, "| host API pageable:", med(T[i, "pageable"]), "| pinned:", med(T[i, "pinned"]), "| device API events:", med(T[i, "device"]),
          "| wall:", med(T[i, "device_wall"]), "| pinned in, device out:", med(T[i, "h2d_only"]), "| device in, pinned out:", med(T[i, "d2h_only"]), "| launches", exs[i].last_launch_count())
for ex in exs: ex.close()
if "--stages" in sys.argv:
    for env in ENVS: stages(env)
if "--stages" in sys.argv:
    for shape in [(480, 752, 1000), (376, 1241, 2000)]:
        stages({}, *shape)
