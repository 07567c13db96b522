#!/usr/bin/env python3
"""bench.py — ORB front-end throughput on B200 (metric of BASELINE.json).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

A "step" is one pass of the hot path (ORBextractor::operator()) over one batch of synthetic
752x480 frames per GPU (BASELINE.json configs[1]: EuRoC-shaped, 1000 keypoints, 8 levels, 1.2).
`value`   : frames/s, whole job, inputs already resident in HBM (device-pointer C-ABI call).
`e2e`     : frames/s through the host-buffer C-ABI call (orb_extract_batch): pinned host frames in,
            keypoints/descriptors/counts out, copies inside the timed region.
`roofline`: the dominant kernel (chosen from live per-stage CUDA-event timings) against the measured
            HBM copy bandwidth of MEASURED_PEAKS.json.
`matching`: Hamming kNN-2 (config 4: 2000x2000 per frame pair; config 5: 2000 queries against a
            10M-row DB sharded over the ranks, NCCL all-gather + exact merge) in descriptor pairs/s
            against the measured POPC-pipe peak.
Frames are sharded over ranks with no data-path collective (weak scaling: fixed frames per GPU).
The reference's CPU arm is oracle/_ref (its own src/ORBextractor.cc compiled against oracle/refshim/, built in the
build container and shipped as a built file), or the oracle port when that library is absent; timed on the host cores.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# rank 0 prints exactly one JSON line on stdout.  Native libraries write there too (NCCL's "NCCL version ..." banner appears at
# every debug level from VERSION up, WARN included), so file descriptor 1 is pointed at stderr for the whole run and the result
# line goes to the saved descriptor.
_RESULT_FD = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def capture_stdout():
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)

W, H, NFEAT, NLEVELS, SCALE, FAST_TH = 752, 480, 1000, 8, 1.2, 20
METRIC = "ORB frames/sec (752x480 EuRoC-shaped synthetic frames, 1000 kp, 8 levels, scale 1.2)"
WORKLOAD = "batched ORB extraction, 752x480 synthetic frames, 1000 kp, frame-sharded"


def level_pixels(w, h, nlevels=NLEVELS):
    inv = [np.float32(1)]
    s = np.float32(1.0 / float(np.float32(SCALE)))
    for _ in range(1, nlevels):
        inv.append(np.float32(inv[-1] * s))
    return [int(np.rint(np.float32(w) * v)) * int(np.rint(np.float32(h) * v)) for v in inv]


def algorithmic_bytes_per_frame(w, h, nkp):
    """every stage reads its input once and writes its output once."""
    P = level_pixels(w, h)
    return (P[0] + sum(P[:-1]) + sum(P[1:])) + sum(P) + 2 * sum(P) + nkp * 749 + nkp * 512 + nkp * 60


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """samples SM clocks / throttle reasons through NVML while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.02)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------- CPU arm
def cpu_kind():
    """"reference": oracle/_ref/libref_orbslam.so = the reference's own src/ORBextractor.cc compiled against oracle/refshim/ (its
    OpenCV primitives are the oracle's restatements); "port": the oracle restatement, when that library was not built."""
    from oracle import pyref
    return "reference" if pyref.available() else "port"


CPU_WHAT = {"reference": "the reference's own src/ORBextractor.cc compiled against oracle/refshim (OpenCV primitives = the oracle's restatements)",
            "port": "oracle port of src/ORBextractor.cc (oracle/_ref is not built on this box)"}


def cpu_extract_rate(frames, nthreads, seconds_hint=None):
    """frames/s of the reference extractor on the host with `nthreads` threads, one extractor instance per thread (the
    reference extractor is stateful, include/ORBextractor.h:74-75; ctypes releases the GIL during the call)."""
    from oracle import pyoracle as po
    from oracle import pyref
    po.lib()
    n = len(frames)
    if pyref.available():
        exs = [pyref.RefExtractor(NFEAT, SCALE, NLEVELS, 1, FAST_TH) for _ in range(nthreads)]
    else:
        exs = [po.OracleExtractor(NFEAT, SCALE, NLEVELS, 1, FAST_TH) for _ in range(nthreads)]
    exs[0](frames[0])                                     # warm
    nxt, lock, done = [0], threading.Lock(), [0]

    def work(ex):
        while True:
            with lock:
                i = nxt[0]
                nxt[0] += 1
            if i >= n:
                return
            ex(frames[i])
            with lock:
                done[0] += 1
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(e,)) for e in exs]
    [t.start() for t in th]
    [t.join() for t in th]
    dt = time.perf_counter() - t0
    return n / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from orbslam_jpminipc_b200.synth import synth_frames
    cores = os.cpu_count() or 1
    per_step = max(cores * 8, 32)
    frames = synth_frames(min(per_step, 32), H, W, 1000)
    frames = [frames[i % len(frames)] for i in range(per_step)]
    for _ in range(args.warmup):
        cpu_extract_rate(frames[:cores], cores)
    t = 0.0
    for _ in range(args.steps):
        _, dt = cpu_extract_rate(frames, cores)
        t += dt
    fps = per_step * args.steps / t
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": per_step, "width": W, "height": H, "nfeatures": NFEAT},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": cpu_kind(),
                             "sample": "%d frames per step x %d steps, frame-parallel over %d host threads; %s"
                                       % (per_step, args.steps, cores, CPU_WHAT[cpu_kind()])},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def bind_to_gpu_numa_node(index):
    """One process per GPU: run (and first-touch the pinned staging buffers) on the CPUs NVML reports as local to that GPU, so that
    eight ranks do not pull their frames through one socket's memory controllers.  Returns the CPU count bound to, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------- GPU arm
def run_matching(args, L, ex, dev, world, rank, stream, barrier, max_over_ranks, e0, e1):
    import torch
    import torch.distributed as dist
    from orbslam_jpminipc_b200._lib import check, ptr
    from orbslam_jpminipc_b200.synth import synth_descriptors
    popc = C.c_double(0)
    check(L.orb_measure_popc_peak(ex._h, C.byref(popc)), "orb_measure_popc_peak")
    NQ, ND, NPAIR = 2000, 2000, 64
    db, q = synth_descriptors(ND, NQ, seed_db=42 + rank, seed_q=43 + rank)
    d_q = torch.from_numpy(np.tile(q, (NPAIR, 1))).to(dev)
    d_db = torch.from_numpy(np.tile(db, (NPAIR, 1))).to(dev)
    o = [torch.zeros(NPAIR * NQ, dtype=torch.int32, device=dev) for _ in range(3)]

    def knn_pairs():
        check(L.orb_hamming_knn2_device(ex._h, ptr(d_q), NQ, ptr(d_db), ND, NPAIR, 0, ptr(o[0]), ptr(o[1]), ptr(o[2]),
                                        C.c_void_p(stream)), "orb_hamming_knn2_device")
    for _ in range(3):
        knn_pairs()
    barrier()
    reps = 20
    e0.record()
    for _ in range(reps):
        knn_pairs()
    e1.record()
    barrier()
    ms4 = max_over_ranks(e0.elapsed_time(e1))
    pairs4 = world * reps * NPAIR * NQ * ND / (ms4 * 1e-3)

    # config 5: 10M-row DB sharded over the ranks, queries replicated, NCCL all-gather + exact merge
    NDB = args.db_rows
    shard = NDB // world
    g = torch.Generator(device=dev)
    g.manual_seed(4242 + rank)
    d_shard = torch.randint(0, 256, (shard, 32), dtype=torch.uint8, device=dev, generator=g)
    _, q5 = synth_descriptors(0, NQ, seed_q=77)
    d_q5 = torch.from_numpy(q5).to(dev)
    part = torch.zeros(3 * NQ, dtype=torch.int32, device=dev)
    allp = torch.zeros(world * 3 * NQ, dtype=torch.int32, device=dev)
    fin = [torch.zeros(NQ, dtype=torch.int32, device=dev) for _ in range(3)]

    def knn_db():
        check(L.orb_hamming_knn2_device(ex._h, ptr(d_q5), NQ, ptr(d_shard), shard, 1, rank * shard,
                                        C.c_void_p(part.data_ptr()), C.c_void_p(part.data_ptr() + 4 * NQ),
                                        C.c_void_p(part.data_ptr() + 8 * NQ), C.c_void_p(stream)), "knn2 shard")
        if world > 1:
            dist.all_gather_into_tensor(allp, part)
            check(L.orb_knn2_merge_device(ex._h, ptr(allp), world, NQ, ptr(fin[0]), ptr(fin[1]), ptr(fin[2]),
                                          C.c_void_p(stream)), "merge")
    for _ in range(2):
        knn_db()
    barrier()
    reps5 = 5
    e0.record()
    for _ in range(reps5):
        knn_db()
    e1.record()
    barrier()
    ms5 = max_over_ranks(e0.elapsed_time(e1))
    pairs5 = reps5 * NQ * (shard * world) / (ms5 * 1e-3)
    matching = {"unit": "descriptor pairs/s", "popc_peak_gops": popc.value,
                "pair_blocks_2000x2000": {"pairs_per_s": pairs4, "queries_per_s": pairs4 / ND, "frame_pairs_per_s": pairs4 / (NQ * ND),
                                          "popc_frac": pairs4 / world * 8 / (popc.value * 1e9)},
                "db_sharded": {"db_rows": shard * world, "queries": NQ, "pairs_per_s": pairs5, "ms_per_query_batch": ms5 / reps5,
                               "popc_frac": pairs5 / world * 8 / (popc.value * 1e9), "merge": "nccl all_gather + k_knn2_merge" if world > 1 else "none (1 shard)"}}

    # config 3: KITTI-shaped 1241x376 pair, 2000 kp: extract both frames on the GPU, then frame-to-frame SearchByProjection
    if rank == 0:
        import orbslam_jpminipc_b200 as pkg
        from orbslam_jpminipc_b200.synth import synth_frame, shifted_frame
        h3, w3 = 376, 1241
        ex3 = pkg.ORBextractor(2000, SCALE, NLEVELS, 1, FAST_TH, device=torch.cuda.current_device(), max_width=w3, max_height=h3, max_batch=2)
        fa = synth_frame(h3, w3, 9000, quadrants=False)
        fb = shifted_frame(fa, 3, 2, 9001)
        (ka, da), (kb, db_) = ex3.extract_batch(np.stack([fa, fb]))
        m3 = pkg.ORBmatcher(0.9, True, extractor=ex3)
        fx = fy = 500.0
        rng = np.random.default_rng(9000)
        z = rng.uniform(2, 10, len(ka)).astype(np.float32)
        xyz = np.stack([(ka["x"] - w3 / 2) / fx * z, (ka["y"] - h3 / 2) / fy * z, z], 1).astype(np.float32)
        Tcw = np.eye(4, dtype=np.float32)
        Tcw[:3, 3] = [0.03, 0.02, 0.01]
        has, outl = np.ones(len(ka), np.uint8), np.zeros(len(ka), np.uint8)
        cur = pkg.Frame(m3, kb, db_, w3, h3, fx, fy, w3 / 2, h3 / 2)
        last = pkg.Frame(m3, ka, da, w3, h3, fx, fy, w3 / 2, h3 / 2)
        nm, _ = m3.SearchByProjection(cur, last, 15.0, has, outl, xyz, Tcw)
        t0 = time.perf_counter()
        for _ in range(50):
            m3.SearchByProjection(cur, last, 15.0, has, outl, xyz, Tcw)
        sbp_ms = (time.perf_counter() - t0) / 50 * 1e3
        t0 = time.perf_counter()
        for _ in range(20):
            ex3.extract_batch(np.stack([fa, fb]))
        ext_ms = (time.perf_counter() - t0) / 20 * 1e3
        matching["search_by_projection_1241x376"] = {"keypoints": [int(len(ka)), int(len(kb))], "matches": int(nm), "th": 15,
                                                     "ms_per_pair_host_api": sbp_ms, "extract_two_frames_host_api_ms": ext_ms,
                                                     "note": "single frame pair, latency through the host-buffer C ABI (grid build excluded)"}
        matching["_sbp_inputs"] = (cur, last, has, outl, xyz, Tcw)
        # vocabulary transform (Frame::ComputeBoW, src/Frame.cc:279-287) on the reference's tree shape: k=10, L=6, levelsup=4
        from orbslam_jpminipc_b200.synth import synth_vocabulary_fast
        parent, vdesc, vweight = synth_vocabulary_fast(10, 6, seed=7)
        voc = pkg.ORBVocabulary(ex).create(10, 6, parent, vdesc, vweight)
        VB, VN = 256, 1000
        rng = np.random.default_rng(5)
        leaves = rng.integers(111111, 1111111, VB * VN)
        feats = vdesc[leaves] ^ (rng.integers(0, 256, (VB * VN, 32), dtype=np.uint8) & rng.integers(0, 256, (VB * VN, 32), dtype=np.uint8)
                                 & rng.integers(0, 256, (VB * VN, 32), dtype=np.uint8) & rng.integers(0, 256, (VB * VN, 32), dtype=np.uint8))
        d_feats = torch.from_numpy(feats).to(dev)
        d_cnt = torch.full((VB,), VN, dtype=torch.int32, device=dev)
        vo = {k_: torch.zeros(VB * (VN + 1), dtype=torch.int32, device=dev) for k_ in ("bw", "fn", "fs", "fi", "nb", "nf")}
        d_bv = torch.zeros(VB * VN, dtype=torch.float64, device=dev)

        def vocab_step():
            check(L.orb_vocab_transform_batch(ex._h, voc._v, ptr(d_feats), VN, ptr(d_cnt), VB, 4, VN, ptr(vo["bw"]), ptr(d_bv), ptr(vo["nb"]),
                                              ptr(vo["fn"]), ptr(vo["fs"]), ptr(vo["fi"]), ptr(vo["nf"])), "orb_vocab_transform_batch")
        for _ in range(3):
            vocab_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            vocab_step()
        torch.cuda.synchronize()
        vms = (time.perf_counter() - t0) / 10 * 1e3
        matching["vocabulary_transform_k10_L6"] = {"frames": VB, "features_per_frame": VN, "levelsup": 4, "ms_per_batch": vms,
                                                   "features_per_s": VB * VN / (vms * 1e-3), "frames_per_s": VB / (vms * 1e-3),
                                                   "mean_words_per_frame": float(vo["nb"][:VB].float().mean().item()),
                                                   "note": "device-resident descriptors, synthetic tree (ORBvoc.txt is not in the reference repository)"}
        matching["_vocab_inputs"] = (parent, vdesc, vweight, feats[:VN].copy())
    return matching


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import orbslam_jpminipc_b200 as pkg
    from orbslam_jpminipc_b200._lib import check, lib, ptr
    from orbslam_jpminipc_b200.synth import synth_frames, synth_descriptors

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    L = lib()
    B = args.batch
    CH = args.chunk if args.chunk > 0 else B
    ex = pkg.ORBextractor(NFEAT, SCALE, NLEVELS, 1, FAST_TH, device=local, max_width=W, max_height=H, max_batch=CH)
    cap = ex.capacity

    # synthetic frames: 32 distinct frames per rank, tiled to the batch (seeds differ per rank)
    base = synth_frames(min(B, 32), H, W, 1000 + 100 * rank)
    host = np.concatenate([base] * ((B + len(base) - 1) // len(base)))[:B].copy()
    pin = torch.from_numpy(host).pin_memory()
    d_img = pin.to(dev, non_blocking=False)
    d_kps = torch.zeros((B, cap, 7), dtype=torch.int32, device=dev)
    d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev)
    d_cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step_device():
        n = 0
        for f0 in range(0, B, CH):
            nb = min(CH, B - f0)
            check(L.orb_extract_batch_device(ex._h, C.c_void_p(d_img.data_ptr() + f0 * W * H), nb, W, H, W, W * H,
                                             C.c_void_p(d_kps.data_ptr() + f0 * cap * 28), C.c_void_p(d_desc.data_ptr() + f0 * cap * 32),
                                             cap, C.c_void_p(d_cnt.data_ptr() + f0 * 4), C.c_void_p(stream)), "orb_extract_batch_device")
            n += ex.last_launch_count()
        return n

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput (value) ----
    for _ in range(max(args.warmup, 3)):
        launches_per_step = step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop()
    # second timed region with per-stage CUDA events on the launching stream (stages serialised: the
    # blur/selection overlap of the production path is switched off while profiling)
    L.orb_profile_enable(ex._h, 1)
    step_device()                                   # warm the profiling path (buffers, event pool)
    torch.cuda.synchronize()
    check(L.orb_profile_read(ex._h, (C.c_double * 7)(), C.byref(C.c_int(0))), "orb_profile_read")
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms_profiled = e0.elapsed_time(e1)
    stage_ms = (C.c_double * 7)()
    ncalls = C.c_int(0)
    check(L.orb_profile_read(ex._h, stage_ms, C.byref(ncalls)), "orb_profile_read")
    L.orb_profile_enable(ex._h, 0)
    # per-stage device ms per STEP (a step is B/CH launches of every stage)
    stage = {L.orb_profile_stage_name(i).decode(): stage_ms[i] / args.steps for i in range(7)}
    frames_total = B * args.steps * world
    value = frames_total / (ms_total * 1e-3)
    nkp = float(d_cnt.float().mean().item())

    # ---- end to end through the host-buffer C-ABI call ----
    out_k = np.zeros((B, cap), pkg.KP_DTYPE)
    out_d = np.zeros((B, cap, 32), np.uint8)
    out_c = np.zeros(B, np.int32)
    pk, pd, pc = (torch.from_numpy(a.view(np.uint8).reshape(-1)).pin_memory() for a in (out_k, out_d, out_c))

    # a context whose max_batch is a fraction of the call's batch makes orb_extract_batch pipeline
    # H2D(k+1) | kernels(k) | D2H(k-1) over its two internal streams
    ex_h = pkg.ORBextractor(NFEAT, SCALE, NLEVELS, 1, FAST_TH, device=local, max_width=W, max_height=H, max_batch=args.e2e_chunk)

    def step_host():
        check(L.orb_extract_batch(ex_h._h, ptr(pin), B, W, H, W, W * H, C.c_void_p(pk.data_ptr()), C.c_void_p(pd.data_ptr()),
                                  cap, C.c_void_p(pc.data_ptr())), "orb_extract_batch")
    for _ in range(2):
        step_host()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_host()
    e1.record()
    barrier()
    e2e_sync_ms = max_over_ranks(e0.elapsed_time(e1))     # the call is synchronous: events bracket H2D + kernels + D2H
    e2e_launches = ex_h.last_launch_count()

    # streaming form of the same call: step i is enqueued (orb_extract_batch_async) before step i-1 is waited for (orb_wait), with
    # two sets of pinned output buffers, so the H2D of a step overlaps the kernels of the previous one.  Every step still copies its
    # frames host->device and its keypoints / descriptors / counts device->host inside the timed region.  Uses the context whose
    # max_batch is the whole step (one chunk per call, consecutive calls alternate the two work sets).
    outs = [(pk, pd, pc), tuple(torch.empty_like(t_).pin_memory() for t_ in (pk, pd, pc))]

    def run_stream(steps):
        prev = None
        for i in range(steps):
            ok_, od_, oc_ = outs[i & 1]
            tk = C.c_longlong(-1)
            check(L.orb_extract_batch_async(ex._h, ptr(pin), B, W, H, W, W * H, C.c_void_p(ok_.data_ptr()), C.c_void_p(od_.data_ptr()),
                                            cap, C.c_void_p(oc_.data_ptr()), C.byref(tk)), "orb_extract_batch_async")
            if prev is not None:
                check(L.orb_wait(ex._h, prev), "orb_wait")
            prev = tk.value
        check(L.orb_wait(ex._h, prev), "orb_wait")
    run_stream(3)
    barrier()
    e0.record()
    run_stream(args.steps)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = frames_total / (e2e_ms * 1e-3)
    e2e_stream_launches = ex.last_launch_count()
    same = all(torch.equal(a_, b_) for a_, b_ in zip(outs[0][2:], outs[1][2:]))    # both buffer sets hold the same counts
    assert same, "streaming call: the two output buffer sets disagree"

    # ---- the metric's other named shape: 640x480 / 1000 kp (BASELINE.json configs[0], the reference's own CPU-runnable case) ----
    W0, H0 = 640, 480
    ex0 = pkg.ORBextractor(NFEAT, SCALE, NLEVELS, 1, FAST_TH, device=local, max_width=W0, max_height=H0, max_batch=B)
    base0 = synth_frames(min(B, 32), H0, W0, 1000 + 100 * rank)
    d_img0 = torch.from_numpy(np.concatenate([base0] * ((B + len(base0) - 1) // len(base0)))[:B].copy()).to(dev)

    def step0():
        check(L.orb_extract_batch_device(ex0._h, ptr(d_img0), B, W0, H0, W0, W0 * H0, ptr(d_kps), ptr(d_desc), cap, ptr(d_cnt),
                                         C.c_void_p(stream)), "orb_extract_batch_device 640x480")
    assert ex0.capacity <= cap
    for _ in range(3):
        step0()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step0()
    e1.record()
    barrier()
    ms0 = max_over_ranks(e0.elapsed_time(e1))
    nkp0 = float(d_cnt.float().mean().item())
    hbm0, _ = measured_peaks()
    config0 = {"workload": "batched ORB extraction, 640x480 synthetic frames, 1000 kp", "value": frames_total / (ms0 * 1e-3), "unit": "frames/s",
               "ms_per_step": ms0 / args.steps, "mean_keypoints": nkp0,
               "pipeline_frac": (frames_total / world / (ms0 * 1e-3)) * algorithmic_bytes_per_frame(W0, H0, nkp0) / (hbm0 * 1e9)}
    step_device()                                          # leave the 752x480 results in the output buffers
    torch.cuda.synchronize()

    # ---- single-frame latency through orb_extract (how the tracking thread calls the extractor: one frame, host buffers in and
    #      out, blocking); with and without the CUDA-graph replay of the pass ----
    def frame_latency(graph):
        old = os.environ.get("ORB_GRAPH")
        os.environ["ORB_GRAPH"] = "1" if graph else "0"
        ex1 = pkg.ORBextractor(NFEAT, SCALE, NLEVELS, 1, FAST_TH, device=local, max_width=W, max_height=H, max_batch=1)
        if old is None:
            del os.environ["ORB_GRAPH"]
        else:
            os.environ["ORB_GRAPH"] = old
        for i in range(10):
            ex1(base[i % len(base)])
        ts = []
        for i in range(200):
            t0 = time.perf_counter()
            ex1(base[i % len(base)])
            ts.append(time.perf_counter() - t0)
        return {"median_ms": float(np.median(ts) * 1e3), "p90_ms": float(np.percentile(ts, 90) * 1e3)}
    latency = {"api": "orb_extract (one 752x480 frame per blocking call, pageable host buffers, python ctypes caller)",
               "graph_replay": frame_latency(True), "plain_launches": frame_latency(False)} if rank == 0 else None

    # ---- roofline of the dominant kernel ----
    hbm, hbm_src = measured_peaks()
    dom = max(stage, key=stage.get)
    P = level_pixels(W, H)
    alg = {"k_level0": P[0] * 2, "k_resize(x7)": sum(P[:-1]) + sum(P[1:]), "k_fast_nms": sum(P), "k_cell_compact": sum(P),
           "k_select": nkp * 8, "k_blur": 2 * sum(P), "k_describe": nkp * (749 + 512 + 60)}
    nlaunch = (B + CH - 1) // CH                       # launches of each stage per step
    bytes_per_launch = alg[dom] * B / nlaunch
    achieved = bytes_per_launch / (stage[dom] / nlaunch * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum per frame from the committed `ncu --set full` capture
    # (profiles/r1l_ncu_full_summary.md, 64-frame launches), scaled to this run's launch size
    ncu_mb_per_frame = {"k_fast_nms": (81.4 + 33.0) / 64, "k_blur": (85.6 + 37.3) / 64, "k_describe": (117.3 + 5.2) / 64,
                        "k_cell_compact": (64.7 + 4.3) / 64, "k_select": 6.6 / 64, "k_level0": 23.2 / 64}
    traffic = ncu_mb_per_frame[dom] * 1e6 * B / nlaunch if dom in ncu_mb_per_frame and W == 752 else None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                "traffic": traffic, "alu_pipe_pct_ncu": 87.7 if dom == "k_fast_nms" else None,
                "note": "k_fast_nms is integer-ALU bound (ncu, profiles/r1l_ncu_full_summary.md: ALU pipe 87.7 % of peak, DRAM 5.3 %); the HBM fraction is the required yardstick, not its limiter", "peak_source": hbm_src, "algorithmic_bytes_per_launch": bytes_per_launch,
                "kernel_ms_per_launch": stage[dom] / nlaunch, "stage_ms_per_step": stage, "profiled_ms_per_step": ms_profiled / args.steps,
                "pipeline_bytes_per_frame": algorithmic_bytes_per_frame(W, H, nkp),
                "pipeline_frac": (value / world) * algorithmic_bytes_per_frame(W, H, nkp) / (hbm * 1e9)}

    # ---- matching (Hamming kNN-2) ----
    matching = None
    if not args.skip_matching:
        matching = run_matching(args, L, ex, dev, world, rank, stream, barrier, max_over_ranks, e0, e1)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": B, "frames_per_launch": CH, "width": W, "height": H, "nfeatures": NFEAT,
                       "nlevels": NLEVELS, "scale": SCALE, "fast_th": FAST_TH, "mean_keypoints": nkp,
                       "l2_policy": "no flush needed: per-step working set (frames + pyramids + score maps, %.0f MB) exceeds the 126 MB L2"
                                    % (B * (W * H + 2 * 1.45e6) / 1e6),
                       "parallelism": "frames sharded over %d GPU(s), no collective on the extraction path" % world,
                       "cpu_affinity": ("rank bound to the %d CPUs local to its GPU (NVML)" % numa) if numa else "unbound"},
            "clocks": clocks, "gpu_launches": launches_per_step * args.steps,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(B * W * H),
                    "d2h_bytes_per_step": int(B * cap * 60 + B * 4), "ms_per_step": e2e_ms / args.steps,
                    "gpu_launches_per_step": e2e_stream_launches, "chunk": min(B, CH),
                    "api": "orb_extract_batch_async + orb_wait, two steps in flight (pinned host buffers in and out, calls alternate two work sets)",
                    "synchronous_call": {"value": frames_total / (e2e_sync_ms * 1e-3), "ms_per_step": e2e_sync_ms / args.steps,
                                         "gpu_launches_per_step": e2e_launches, "chunk": args.e2e_chunk,
                                         "api": "orb_extract_batch (one blocking call per step, internally chunked + double-buffered)"}},
            "roofline": roofline, "config0_640x480": config0, "single_frame_latency": latency, "matching": matching}
    sbp_inputs = matching.pop("_sbp_inputs", None) if matching else None
    vocab_inputs = matching.pop("_vocab_inputs", None) if matching else None
    if args.cpu_baseline:
        cores = os.cpu_count() or 1
        nfr = 64 * cores                                   # ~12 core-seconds of CPU work at ~80 frames/s/core
        fr = [base[i % len(base)] for i in range(nfr)]
        fps1, dt1 = cpu_extract_rate(fr[:48], 1)
        fpsN, dtN = cpu_extract_rate(fr, cores)
        line["cpu_baseline"] = {"value": fpsN, "unit": "frames/s", "cores": cores, "kind": cpu_kind(), "single_thread_value": fps1,
                                "sample": "%d frames of the same workload, frame-parallel over %d host threads (%.1f s wall); "
                                          "single-thread figure on 48 frames (%.1f s); %s" % (nfr, cores, dtN, dt1, CPU_WHAT[cpu_kind()])}
        if sbp_inputs is not None:
            from oracle import pyoracle as po
            cur, last, has, outl, xyz, Tcw = sbp_inputs
            oc = po.OracleFrame(cur.kps, cur.desc, cur.width, cur.height, cur.fx, cur.fy, cur.cx, cur.cy)
            ol = po.OracleFrame(last.kps, last.desc, last.width, last.height, last.fx, last.fy, last.cx, last.cy)
            t0 = time.perf_counter()
            for _ in range(20):
                po.search_by_projection(oc, ol, has, outl, xyz, Tcw, 15.0, True)
            line["cpu_baseline"]["search_by_projection_1241x376_ms_per_pair"] = (time.perf_counter() - t0) / 20 * 1e3
            db4, q4 = synth_descriptors(2000, 2000)
            t0 = time.perf_counter()
            po.knn2(q4, db4)
            line["cpu_baseline"]["knn2_2000x2000_pairs_per_s_1thread"] = 4e6 / (time.perf_counter() - t0)
        if vocab_inputs is not None:
            from oracle import pyoracle as po
            parent, vdesc, vweight, f1 = vocab_inputs
            ov = po.OracleVocabulary(10, 6, parent, vdesc, vweight)
            ov.transform(f1, 4)
            t0 = time.perf_counter()
            for _ in range(10):
                ov.transform(f1, 4)
            line["cpu_baseline"]["vocabulary_transform_k10_L6_ms_per_frame_1thread"] = (time.perf_counter() - t0) / 10 * 1e3
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="frames per GPU per step")
    ap.add_argument("--db-rows", type=int, default=10_000_000)
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--skip-matching", action="store_true")
    ap.add_argument("--e2e-chunk", type=int, default=64)
    ap.add_argument("--chunk", type=int, default=0, help="frames per kernel launch (context max_batch); 0 = batch")
    args = ap.parse_args()
    capture_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
