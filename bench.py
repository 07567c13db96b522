#!/usr/bin/env python3
"""bench.py — the metric of BASELINE.json: ORB frames/s @640x480 / 1000 kp AND Hamming matches/s, 1/2/4/8 B200 vs host CPU.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU implementation on the host cores

A "step" is `--step-launches` (default 8) passes of the hot path (ORBextractor::operator()) over one batch of `--batch` (1024)
synthetic 640x480 frames per GPU = 8192 frames per GPU and step (>= 50 ms of device work, so that the timed region of K = 20 steps
is > 1 s and the clock sampler sees it).
`value`        : frames/s, whole job, inputs already resident in HBM (device-pointer C-ABI call orb_extract_batch_device).
`e2e`          : frames/s through the host-buffer C-ABI calls (orb_extract_batch_async + orb_wait): pinned host frames in,
                 keypoints / descriptors / counts out, every launch's H2D + D2H inside the timed region; next to it the N-rank raw
                 copy ceiling of exactly those bytes (no kernels) measured in the same run.
`roofline`     : the dominant kernel (from live per-stage CUDA-event timings on the launching stream) against the measured HBM copy
                 bandwidth of MEASURED_PEAKS.json; `traffic` / ALU-pipe figures come from the committed ncu capture named next to them.
`matches_per_s`: Hamming kNN-2 (config 4: 2000 x 2000 descriptor blocks per frame pair; config 5: 2000 queries against a 10 M-row
                 database sharded over the ranks, NCCL all-gather + exact merge inside the library) in descriptor pairs/s against the
                 POPC-pipe peak measured in the same run and the theoretical one.
`verified`     : what was compared bit for bit with the CPU oracle / a single-GPU scan AFTER the timed regions, on the buffers the
                 timed calls wrote.  Any mismatch exits non-zero without a result line.
`config1_752x480`: the same extraction measurements on BASELINE.json configs[1] (752x480 EuRoC-shaped frames).
Frames are sharded over ranks with no data-path collective (weak scaling: fixed frames per GPU).
The CPU arm is oracle/_ref (the reference's own src/ORBextractor.cc compiled against oracle/refshim/, built in the build container
and shipped as a built file), or the oracle port when that library is absent; protocol of BASELINE.md §2.
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


NFEAT, NLEVELS, SCALE, FAST_TH = 1000, 8, 1.2, 20
W0, H0 = 640, 480            # BASELINE.json metric / configs[0]: the headline
W1, H1 = 752, 480            # configs[1]
METRIC = "ORB frames/sec @640x480 1000kp and Hamming matches/sec"
WORKLOAD = "batched ORB extraction, 640x480 synthetic frames, 1000 kp, 8 levels, scale 1.2, frame-sharded (BASELINE.json metric shape, configs[0]); Hamming kNN-2 in matches_per_s"
POPC_THEORETICAL_GOPS = 16 * 148 * 1.965       # 16 POPC/clk/SM x 148 SMs x 1.965 GHz


def shared_config():
    """identical in both arms (the driver compares them)"""
    return {"workload": WORKLOAD, "width": W0, "height": H0, "nfeatures": NFEAT, "nlevels": NLEVELS, "scale": SCALE, "fast_th": FAST_TH,
            "l2_policy": "inputs larger than L2: every launch streams its whole batch of frames + their pyramids and score maps (~1 GB per 256 frames) through a 126 MB L2"}


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
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_record(w, h):
    """per-kernel dram bytes / pipe figures of the committed `ncu --set full` capture (written by tools/ncu_traffic.py), or None"""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    return d if (d.get("width"), d.get("height")) == (w, h) else None


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


def pcie_info(index):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        return {"gen": pynvml.nvmlDeviceGetCurrPcieLinkGeneration(h), "width": pynvml.nvmlDeviceGetCurrPcieLinkWidth(h),
                "max_gen": pynvml.nvmlDeviceGetMaxPcieLinkGeneration(h), "max_width": pynvml.nvmlDeviceGetMaxPcieLinkWidth(h)}
    except Exception:
        return None


# ------------------------------------------------------------------------------- CPU arm
def cpu_kind():
    """"reference": oracle/_ref/libref_orbslam.so = the reference's own src/ORBextractor.cc compiled against oracle/refshim/ (its
    OpenCV primitives are the oracle's restatements); "port": the oracle restatement, when that library was not built."""
    from oracle import pyref
    return "reference" if pyref.available() else "port"


CPU_WHAT = {"reference": "the reference's own src/ORBextractor.cc compiled against oracle/refshim (OpenCV primitives = the oracle's restatements)",
            "port": "oracle port of src/ORBextractor.cc (oracle/_ref is not built on this box)"}


def _pct(fps):
    a = np.asarray(fps, np.float64)
    return {"median": float(np.median(a)), "p10": float(np.percentile(a, 10)), "p90": float(np.percentile(a, 90))}


def cpu_extract_protocol(frames, nthreads, timed, warm=20):
    """`warm` untimed frames, then `timed` frames, frame-parallel over `nthreads` threads with one extractor instance per
    thread (the reference extractor is stateful, include/ORBextractor.h:74-75; ctypes releases the GIL during the call).
    -> (frames/s over the wall clock of the timed part, percentiles of nthreads / per-frame time, seconds)"""
    from oracle import pyoracle as po
    from oracle import pyref
    po.lib()
    mk = (lambda: pyref.RefExtractor(NFEAT, SCALE, NLEVELS, 1, FAST_TH)) if pyref.available() else (lambda: po.OracleExtractor(NFEAT, SCALE, NLEVELS, 1, FAST_TH))
    exs = [mk() for _ in range(nthreads)]
    n = len(frames)
    per_thread = [[] for _ in range(nthreads)]
    start = threading.Barrier(nthreads + 1)
    w_each, t_each = -(-warm // nthreads), -(-timed // nthreads)

    def work(k):
        ex = exs[k]
        for i in range(w_each):
            ex(frames[(k + i * nthreads) % n])
        start.wait()
        for i in range(t_each):
            t0 = time.perf_counter()
            ex(frames[(k + i * nthreads) % n])
            per_thread[k].append(time.perf_counter() - t0)
    th = [threading.Thread(target=work, args=(k,)) for k in range(nthreads)]
    [t.start() for t in th]
    start.wait()
    t0 = time.perf_counter()
    [t.join() for t in th]
    dt = time.perf_counter() - t0
    ts = np.concatenate([np.asarray(p) for p in per_thread])
    return t_each * nthreads / dt, _pct(nthreads / ts), dt, t_each * nthreads


def cpu_knn_protocol(q, db, nthreads, use_popcnt, reps=3):
    """descriptor pairs/s of the brute-force best / second-best scan on the host: queries split over `nthreads` threads"""
    from oracle import pyoracle as po
    parts = np.array_split(np.arange(len(q)), nthreads)
    rates = []
    for _ in range(reps):
        th = [threading.Thread(target=lambda p=p: po.knn2(q[p], db, use_popcnt)) for p in parts if len(p)]
        t0 = time.perf_counter()
        [t.start() for t in th]
        [t.join() for t in th]
        rates.append(len(q) * len(db) / (time.perf_counter() - t0))
    return _pct(rates)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from orbslam_jpminipc_b200.synth import synth_frames
    cores = os.cpu_count() or 1
    per_step = max(cores * 32, 256)          # ~0.4 s of all-core work per step: thread start-up no longer weighs on the rate
    frames = synth_frames(32, H0, W0, 1000)
    for _ in range(args.warmup):
        cpu_extract_protocol(frames, cores, cores, warm=0)
    t, fr, pcts = 0.0, 0, []
    for _ in range(args.steps):
        _, pc, dt, n = cpu_extract_protocol(frames, cores, per_step, warm=0)
        t += dt
        fr += n
        pcts.append(pc["median"])
    fps = fr / t
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": shared_config(),
            "step": {"frames": fr // args.steps, "what": "bounded sample of the workload: %d frames per step, frame-parallel over %d host threads" % (fr // args.steps, cores)},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": cpu_kind(),
                             "sample": "%d frames per step x %d steps, frame-parallel over %d host threads, one extractor per thread; %s"
                                       % (fr // args.steps, args.steps, cores, CPU_WHAT[cpu_kind()])},
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


class Fail(SystemExit):
    pass


def require(cond, what):
    if not cond:
        sys.stderr.write("bench.py: VERIFICATION FAILED: %s\n" % what)
        sys.stderr.flush()
        os._exit(3)


# ------------------------------------------------------------------------------- GPU arm
class Env:
    pass


def measure_extraction(E, w, h, steps, warmup, launches, full):
    """device-resident value, per-stage profile, e2e (streaming + blocking), raw copy ceiling and verification for one frame shape"""
    import torch
    import orbslam_jpminipc_b200 as pkg
    from orbslam_jpminipc_b200._lib import check, ptr
    from orbslam_jpminipc_b200.synth import synth_frames
    L, dev, B, world, rank = E.L, E.dev, E.B, E.world, E.rank
    ex = pkg.ORBextractor(NFEAT, SCALE, NLEVELS, 1, FAST_TH, device=E.local, max_width=w, max_height=h, max_batch=B)
    cap = ex.capacity
    base = synth_frames(min(B, 32), h, w, 1000 + 100 * rank)            # 32 distinct frames per rank, tiled to the batch
    host = np.concatenate([base] * ((B + len(base) - 1) // len(base)))[:B].copy()
    if E.args.wc_input:
        # write-combined pinned input (orb_host_alloc_input): the CPU only writes the frames, the copy engine reads them without cache snoops
        base_ptr = L.orb_host_alloc_input(host.nbytes)
        require(bool(base_ptr), "orb_host_alloc_input")
        wc = np.ctypeslib.as_array(C.cast(base_ptr, C.POINTER(C.c_uint8)), shape=(host.nbytes,)).reshape(host.shape)
        wc[...] = host
        pin = torch.from_numpy(wc)
    else:
        pin = torch.from_numpy(host).pin_memory()
    d_img = pin.to(dev, non_blocking=False)
    d_kps = torch.zeros((B, cap, 7), dtype=torch.int32, device=dev)
    d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev)
    d_cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    stream = E.stream
    e0, e1 = E.e0, E.e1
    nstage = 7

    def launch_device():
        check(L.orb_extract_batch_device(ex._h, ptr(d_img), B, w, h, w, w * h, ptr(d_kps), ptr(d_desc), cap, ptr(d_cnt), C.c_void_p(stream)),
              "orb_extract_batch_device")

    def step_device():
        for _ in range(launches):
            launch_device()
    for _ in range(warmup):
        step_device()
    launches_per_call = ex.last_launch_count()
    E.barrier()
    sampler = ClockSampler(E.local) if full else None
    if sampler:
        sampler.start()
    e0.record()
    for _ in range(steps):
        step_device()
    e1.record()
    E.barrier()
    ms_total = E.max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    frames_total = B * launches * steps * world
    value = frames_total / (ms_total * 1e-3)
    nkp = float(d_cnt.float().mean().item())
    res = {"value": value, "ms_per_step": ms_total / steps, "mean_keypoints": nkp, "clocks": clocks, "cap": cap,
           "gpu_launches": launches_per_call * launches * steps, "launches_per_call": launches_per_call}

    # second timed region with per-stage CUDA events on the launching stream (stages serialised: the blur / selection overlap of the
    # production path is switched off while profiling)
    L.orb_profile_enable(ex._h, 1)
    launch_device()                                  # warm the profiling path (buffers, event pool)
    torch.cuda.synchronize()
    check(L.orb_profile_read(ex._h, (C.c_double * nstage)(), C.byref(C.c_int(0))), "orb_profile_read")
    prof_launches = max(8, min(launches * 2, 64))
    e0.record()
    for _ in range(prof_launches):
        launch_device()
    e1.record()
    E.barrier()
    ms_profiled = e0.elapsed_time(e1)
    stage_ms = (C.c_double * nstage)()
    ncalls = C.c_int(0)
    check(L.orb_profile_read(ex._h, stage_ms, C.byref(ncalls)), "orb_profile_read")
    L.orb_profile_enable(ex._h, 0)
    stage = {L.orb_profile_stage_name(i).decode(): stage_ms[i] / prof_launches for i in range(nstage)}     # device ms per launch of B frames
    res["stage_ms_per_launch"] = stage
    res["profiled_ms_per_launch"] = ms_profiled / prof_launches

    # ---- end to end through the host-buffer C-ABI calls ----
    out_k = np.zeros((B, cap), pkg.KP_DTYPE)
    out_d = np.zeros((B, cap, 32), np.uint8)
    out_c = np.zeros(B, np.int32)
    pk, pd, pc = (torch.from_numpy(a.view(np.uint8).reshape(-1)).pin_memory() for a in (out_k, out_d, out_c))
    DEPTH = 3                      # calls in flight: launch i is enqueued before launch i-2 is waited for, each with its own pinned outputs
    outs = [(pk, pd, pc)] + [tuple(torch.empty_like(t_).pin_memory() for t_ in (pk, pd, pc)) for _ in range(DEPTH - 1)]
    h2d, d2h = int(B * w * h), int(B * cap * 60 + B * 4)

    # streaming form: launch i is enqueued (orb_extract_batch_async) before launch i-2 is waited for (orb_wait), with three sets of pinned
    # output buffers, so the H2D of a launch overlaps the kernels of the previous one.  Every launch still copies its frames
    # host->device and its keypoints / descriptors / counts device->host inside the timed region.
    def run_stream(n):
        pending = []
        for i in range(n):
            ok_, od_, oc_ = outs[i % DEPTH]
            tk = C.c_longlong(-1)
            check(L.orb_extract_batch_async(ex._h, ptr(pin), B, w, h, w, w * h, C.c_void_p(ok_.data_ptr()), C.c_void_p(od_.data_ptr()),
                                            cap, C.c_void_p(oc_.data_ptr()), C.byref(tk)), "orb_extract_batch_async")
            pending.append(tk.value)
            if len(pending) == DEPTH:                    # the call that owns the pinned outputs the next launch will write
                check(L.orb_wait(ex._h, pending.pop(0)), "orb_wait")
        for tk_ in pending:
            check(L.orb_wait(ex._h, tk_), "orb_wait")
    run_stream(2 * DEPTH)
    E.barrier()
    e0.record()
    run_stream(launches * steps)
    e1.record()
    E.barrier()
    e2e_ms = E.max_over_ranks(e0.elapsed_time(e1))
    e2e_value = frames_total / (e2e_ms * 1e-3)
    res["e2e"] = {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d * launches, "d2h_bytes_per_step": d2h * launches,
                  "ms_per_step": e2e_ms / steps, "gpu_launches_per_step": ex.last_launch_count() * launches,
                  "frames_per_call": B, "calls_per_step": launches,
                  "api": "orb_extract_batch_async + orb_wait, three calls in flight (pinned host buffers in and out, one output set per call in flight; calls alternate two device work sets, D2H on its own stream)"}

    if full:
        # one blocking call per launch, internally chunked (a context whose max_batch is a fraction of the call's batch makes
        # orb_extract_batch pipeline H2D(k+1) | kernels(k) | D2H(k-1) over its two internal streams)
        ex_h = pkg.ORBextractor(NFEAT, SCALE, NLEVELS, 1, FAST_TH, device=E.local, max_width=w, max_height=h, max_batch=E.args.e2e_chunk)

        def call_host():
            check(L.orb_extract_batch(ex_h._h, ptr(pin), B, w, h, w, w * h, C.c_void_p(pk.data_ptr()), C.c_void_p(pd.data_ptr()),
                                      cap, C.c_void_p(pc.data_ptr())), "orb_extract_batch")
        for _ in range(2):
            call_host()
        nsync = max(steps, 8)
        E.barrier()
        e0.record()
        for _ in range(nsync):
            call_host()
        e1.record()
        E.barrier()
        sync_ms = E.max_over_ranks(e0.elapsed_time(e1))
        res["e2e"]["synchronous_call"] = {"value": B * nsync * world / (sync_ms * 1e-3), "ms_per_call": sync_ms / nsync, "chunk": E.args.e2e_chunk,
                                          "api": "orb_extract_batch (one blocking call per %d frames, internally chunked + double-buffered)" % B}
        ex_h.close()

        # raw copy ceiling: the same H2D + D2H bytes per launch, no kernels, all ranks at once, H2D and D2H on two streams (PCIe is full
        # duplex) -> what the box's PCIe / host-memory fabric allows for this copy pattern at this rank count
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
        d_in2 = torch.empty_like(d_img)
        d_out2 = torch.zeros(d2h, dtype=torch.uint8, device=dev)
        h_out2 = torch.empty(d2h, dtype=torch.uint8).pin_memory()

        def copies(n):
            for _ in range(n):
                with torch.cuda.stream(s_up):
                    d_in2.copy_(pin, non_blocking=True)
                with torch.cuda.stream(s_dn):
                    h_out2.copy_(d_out2, non_blocking=True)
        copies(4)
        E.barrier()
        ncopy = min(launches * steps, 128)
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t2 = torch.cuda.Event(enable_timing=True)
        t0.record(s_up)
        s_dn.wait_event(t0)
        copies(ncopy)
        t1.record(s_up)
        t2.record(s_dn)
        E.barrier()
        cms = E.max_over_ranks(max(t0.elapsed_time(t1), t0.elapsed_time(t2)))
        ceil_fps = B * ncopy * world / (cms * 1e-3)
        res["e2e"]["copy_ceiling"] = {"frames_per_s": ceil_fps, "h2d_gbs_per_gpu": h2d * ncopy / (t0.elapsed_time(t1) * 1e-3) / 1e9,
                                      "d2h_gbs_per_gpu": d2h * ncopy / (t0.elapsed_time(t2) * 1e-3) / 1e9, "pcie_link": pcie_info(E.local),
                                      "what": "%d ranks concurrently copying the same bytes per launch (%.1f MB H2D + %.1f MB D2H), pinned memory, no kernels"
                                              % (world, h2d / 1e6, d2h / 1e6)}
        res["e2e"]["frac_of_copy_ceiling"] = e2e_value / ceil_fps
        res["e2e"]["frac_of_device_value"] = e2e_value / value
        # what bounds the end-to-end number at this rank count: the kernels (device value) or the box's copy fabric (ceiling)
        res["e2e"]["bound"] = "kernels (device-resident value)" if value < ceil_fps else "host<->device copies (copy_ceiling)"
        res["e2e"]["frac_of_bound"] = e2e_value / min(value, ceil_fps)

    # ---- verification (after the timed regions, on what the timed calls wrote) ----
    if E.args.verify:
        from oracle import pyoracle as po
        launch_device()
        torch.cuda.synchronize()
        cnt = d_cnt.cpu().numpy()
        kps_h = d_kps.cpu().numpy().view(np.uint8).reshape(B, cap, 28)
        desc_h = d_desc.cpu().numpy()
        orc = po.OracleExtractor(NFEAT, SCALE, NLEVELS, 1, FAST_TH)
        checked = sorted({0, min(B - 1, 31), B - 1})
        for i in checked:
            rk, rd = orc(host[i])
            n = int(cnt[i])
            require(n == len(rk), "%dx%d frame %d: %d keypoints, oracle %d" % (w, h, i, n, len(rk)))
            require(np.array_equal(kps_h[i, :n].reshape(-1), rk.view(np.uint8).reshape(-1)), "%dx%d frame %d: keypoints differ from the oracle" % (w, h, i))
            require(np.array_equal(desc_h[i, :n], rd), "%dx%d frame %d: descriptors differ from the oracle" % (w, h, i))
        # the host-buffer (e2e) path wrote the same frames: both pinned output sets must equal the device-resident result
        for ok_, od_, oc_ in outs:
            require(np.array_equal(oc_.numpy().view(np.int32), cnt), "e2e counts differ from the device-resident run")
            k2 = ok_.numpy().reshape(B, cap, 28)
            d2 = od_.numpy().reshape(B, cap, 32)
            for i in checked:
                n = int(cnt[i])
                require(np.array_equal(k2[i, :n], kps_h[i, :n]) and np.array_equal(d2[i, :n], desc_h[i, :n]), "e2e frame %d differs from the device-resident run" % i)
        res["verified"] = "frames %s of the timed batch: keypoints (all 7 fields, angle bit patterns) and descriptors bit-exact vs the CPU oracle; e2e output buffers equal the device-resident run" % checked
        res["_frame0"] = (int(cnt[0]), kps_h[0, :int(cnt[0])].copy(), desc_h[0, :int(cnt[0])].copy())
    res["_base"] = base
    res["_ex"] = ex
    return res


def roofline_of(res, w, h, B):
    hbm, hbm_src = measured_peaks()
    stage = res["stage_ms_per_launch"]
    nkp = res["mean_keypoints"]
    dom = max(stage, key=stage.get)
    P = level_pixels(w, h)
    alg = {"k_level0": P[0] * 2, "k_resize(x7)": sum(P[:-1]) + sum(P[1:]), "k_pyramid": P[0] + sum(P[:-1]) + sum(P[1:]), "k_fast_nms": sum(P), "k_cell_compact": sum(P) / 8,
           "k_select": nkp * 8, "k_blur": 2 * sum(P), "k_describe": nkp * (749 + 512 + 60)}
    bytes_per_launch = alg.get(dom, 0) * B
    achieved = bytes_per_launch / (stage[dom] * 1e-3) / 1e9
    rec = ncu_record(w, h)
    k = (rec or {}).get("kernels", {}).get(dom.split("(")[0])
    traffic = k["dram_bytes_per_frame"] * B if k else None
    return {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
            "traffic": traffic, "traffic_source": (rec or {}).get("source") if k else None,
            "alu_pipe_pct_ncu": k.get("alu_pipe_pct") if k else None, "warp_inst_per_pixel_ncu": k.get("warp_inst_per_pixel") if k else None,
            "note": "k_fast_nms is bound by instruction issue on the integer / half2 pipes (ncu figures in alu_pipe_pct_ncu / traffic_source: ALU pipe ~71 %, FMA pipe ~20 %, issue ~66 %, DRAM ~6 %); the HBM fraction is the required yardstick, not its limiter",
            "peak_source": hbm_src, "algorithmic_bytes_per_launch": bytes_per_launch, "frames_per_launch": B,
            "kernel_ms_per_launch": stage[dom], "stage_ms_per_launch": stage, "stage_ms_per_step": stage, "profiled_ms_per_launch": res["profiled_ms_per_launch"],
            "pipeline_bytes_per_frame": algorithmic_bytes_per_frame(w, h, nkp)}


def run_matching(E):
    import torch
    import torch.distributed as dist
    from orbslam_jpminipc_b200._lib import check, ptr
    from orbslam_jpminipc_b200.sharding import RankComm, shard_range
    from orbslam_jpminipc_b200.synth import db_queries, db_rows_torch, synth_descriptors
    from oracle import pyoracle as po
    L, dev, world, rank, stream, ex = E.L, E.dev, E.world, E.rank, E.stream, E.ex
    e0, e1 = E.e0, E.e1
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
    E.barrier()
    reps = E.args.match_reps
    e0.record()
    for _ in range(reps):
        knn_pairs()
    e1.record()
    E.barrier()
    ms4 = E.max_over_ranks(e0.elapsed_time(e1))
    pairs4 = world * reps * NPAIR * NQ * ND / (ms4 * 1e-3)
    verified = {}
    if E.args.verify:
        r = po.knn2(q, db)
        for blk in (0, NPAIR - 1):
            got = [t[blk * NQ:(blk + 1) * NQ].cpu().numpy() for t in o]
            require(all(np.array_equal(a, b) for a, b in zip(got, r)), "kNN-2 2000x2000 block %d differs from the oracle" % blk)
        verified["pair_blocks_2000x2000"] = "blocks 0 and %d of the timed launch: (idx1, d1, d2) bit-exact vs the CPU oracle" % (NPAIR - 1)

    # config 5: a 10 M-row database sharded over the ranks (counter-based generator, SURVEY.md §8d: planted neighbours whose exact
    # duplicates always lie in ANOTHER shard), queries replicated; kNN + NCCL all-gather + exact merge all inside the library
    # (orb_knn2_sharded_device) on one stream
    NDB, SEED5 = E.args.db_rows, 4242
    lo, hi = shard_range(NDB, rank, world)
    d_shard = db_rows_torch(lo, hi, NDB, SEED5, dev)
    q5, planted = db_queries(NQ, NDB, SEED5)
    d_q5 = torch.from_numpy(q5).to(dev)
    comm = RankComm(ex)
    fin = None
    for _ in range(2):
        fin = comm.knn2_sharded(d_q5, d_shard, lo, stream=stream)
    E.barrier()
    reps5 = max(5, min(E.args.steps, 10))
    e0.record()
    for _ in range(reps5):
        fin = comm.knn2_sharded(d_q5, d_shard, lo, stream=stream)
    e1.record()
    E.barrier()
    ms5 = E.max_over_ranks(e0.elapsed_time(e1))
    pairs5 = reps5 * NQ * NDB / (ms5 * 1e-3)
    fin_h = [t.cpu().numpy() for t in fin]
    # ---- A/B: the same two workloads on the tensor-core engine (csrc/orb_match_tc.cu: +-1 int8, tcgen05.mma kind::i8, accumulator in
    #      TMEM).  Exact integer arithmetic, so the results must be bit-identical to the POPC engine's; any difference fails the run.
    tensor = None
    if E.args.tensor_ab:
        check(L.orb_set_knn_engine(ex._h, 1), "orb_set_knn_engine")
        popc4 = [t.clone() for t in o]
        for _ in range(3):
            knn_pairs()
        E.barrier()
        e0.record()
        for _ in range(reps):
            knn_pairs()
        e1.record()
        E.barrier()
        tms4 = E.max_over_ranks(e0.elapsed_time(e1))
        require(all(bool(torch.equal(a, b)) for a, b in zip(popc4, o)), "tensor-core kNN differs from the POPC engine on the 2000x2000 blocks")
        for _ in range(2):
            tfin = comm.knn2_sharded(d_q5, d_shard, lo, stream=stream)
        E.barrier()
        e0.record()
        for _ in range(reps5):
            tfin = comm.knn2_sharded(d_q5, d_shard, lo, stream=stream)
        e1.record()
        E.barrier()
        tms5 = E.max_over_ranks(e0.elapsed_time(e1))
        require(all(np.array_equal(a, b.cpu().numpy()) for a, b in zip(fin_h, tfin)), "tensor-core kNN differs from the POPC engine on the sharded database")
        check(L.orb_set_knn_engine(ex._h, 0), "orb_set_knn_engine")
        tp4, tp5 = world * reps * NPAIR * NQ * ND / (tms4 * 1e-3), reps5 * NQ * NDB / (tms5 * 1e-3)
        # int8 tensor peak: nominal 4.5 POP/s dense per GPU (B200_PROFILING.md family figure: 2x the bf16 2.25 PFLOP/s); each pair is 256 MACs
        tensor = {"engine": "ORB_KNN_TENSOR: descriptor bits as +-1 int8, tcgen05.mma kind::i8 M128 N128 K32 x 8 for two query tiles per expanded database tile, accumulators in TMEM, best/second-best scan on tcgen05.ld with 16-bit packed keys; Hamming = (256 - dot) / 2",
                  "bit_exact_vs_popc_engine": True, "default": False,
                  "pair_blocks_2000x2000": {"pairs_per_s": tp4, "speedup_vs_popc": tp4 / pairs4, "timed_ms": tms4},
                  "db_sharded_10M": {"pairs_per_s": tp5, "ms_per_query_batch": tms5 / reps5, "speedup_vs_popc": tp5 / pairs5,
                                     "int8_tensor_frac_nominal": tp5 / world * 512 / 4.5e15},
                  "note": "experiment (VERDICT r1 item 9); the POPC engine stays the default because the path's contract (BASELINE.json north_star) names integer-pipe kernels"}
    merge_check = None
    if E.args.verify:
        # every rank holds the merged result: all ranks must agree, and rank 0 compares it with ONE scan over the whole database
        # regenerated on its own GPU
        ev = planted >= 0
        require(np.array_equal(fin_h[0][ev], planted[ev]) and np.array_equal(fin_h[1][ev], fin_h[2][ev]),
                "config 5: a planted query did not return the lower copy of its duplicated row with d2 == d1")
        if world > 1:
            mine = torch.stack([t.to(torch.int32) for t in fin]).contiguous()
            allr = torch.empty((world,) + tuple(mine.shape), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(allr, mine)
            require(bool((allr == allr[0:1]).all().item()), "config 5: ranks disagree on the merged result")
        if rank == 0:
            del d_shard
            whole = db_rows_torch(0, NDB, NDB, SEED5, dev)
            s = [torch.zeros(NQ, dtype=torch.int32, device=dev) for _ in range(3)]
            check(L.orb_hamming_knn2_device(ex._h, ptr(d_q5), NQ, ptr(whole), NDB, 1, 0, ptr(s[0]), ptr(s[1]), ptr(s[2]), C.c_void_p(stream)), "single scan")
            torch.cuda.synchronize()
            require(all(np.array_equal(a, b.cpu().numpy()) for a, b in zip(fin_h, s)), "config 5: merged sharded result differs from the single scan")
            del whole
        merge_check = ("bit-exact vs single scan (%d shards, NCCL all-gather + k_knn2_merge in orb_knn2_sharded_device; all ranks agree)" % world) if world > 1 \
            else "1 shard: bit-exact vs the planted neighbours (lower copy wins, d2 == d1)"
        verified["db_sharded"] = merge_check
    transport = comm.transport
    comm.close()

    pk = popc.value * 1e9
    matching = {"unit": "descriptor pairs/s", "popc_peak_gops": popc.value, "popc_peak_source": "measured in this run by the library's register-resident __popc micro-kernel (k_popc_bench); theoretical 16 POPC/clk/SM x 148 SM x 1.965 GHz = %.0f G/s" % POPC_THEORETICAL_GOPS,
                "pair_blocks_2000x2000": {"pairs_per_s": pairs4, "queries_per_s": pairs4 / ND, "frame_pairs_per_s": pairs4 / (NQ * ND),
                                          "popc_frac": pairs4 / world * 8 / pk, "popc_frac_theoretical": pairs4 / world * 8 / (POPC_THEORETICAL_GOPS * 1e9),
                                          "timed_ms": ms4, "launches": reps},
                "db_sharded": {"db_rows": NDB, "queries": NQ, "pairs_per_s": pairs5, "queries_per_s": pairs5 / NDB, "ms_per_query_batch": ms5 / reps5,
                               "popc_frac": pairs5 / world * 8 / pk, "popc_frac_theoretical": pairs5 / world * 8 / (POPC_THEORETICAL_GOPS * 1e9),
                               "merge": ("orb_knn2_sharded_device: k_knn2 + ncclAllGather (%s) + k_knn2_merge on one stream" % transport) if world > 1 else "none (1 shard)",
                               "merge_check": merge_check},
                "tensor_core_experiment": tensor,
                "verified": verified}
    return matching


def run_tracking_extras(E, matching):
    """config 3 latency (SearchByProjection on a KITTI-shaped pair) and the vocabulary transform; rank 0 only"""
    import torch
    import orbslam_jpminipc_b200 as pkg
    from orbslam_jpminipc_b200._lib import check, ptr
    from orbslam_jpminipc_b200.synth import shifted_frame, synth_frame, synth_vocabulary_fast
    L, dev, ex = E.L, E.dev, E.ex
    h3, w3 = 376, 1241
    ex3 = pkg.ORBextractor(2000, SCALE, NLEVELS, 1, FAST_TH, device=E.local, max_width=w3, max_height=h3, max_batch=2)
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
    nm, mt = m3.SearchByProjection(cur, last, 15.0, has, outl, xyz, Tcw)
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
    cpu = {}
    if E.args.verify or E.args.cpu_baseline:
        from oracle import pyoracle as po
        oc = po.OracleFrame(cur.kps, cur.desc, cur.width, cur.height, cur.fx, cur.fy, cur.cx, cur.cy)
        ol = po.OracleFrame(last.kps, last.desc, last.width, last.height, last.fx, last.fy, last.cx, last.cy)
        rn, rmt = po.search_by_projection(oc, ol, has, outl, xyz, Tcw, 15.0, True)
        if E.args.verify:
            require(int(rn) == int(nm) and np.array_equal(np.asarray(rmt), np.asarray(mt)), "SearchByProjection 1241x376 differs from the oracle")
            matching["verified"]["search_by_projection_1241x376"] = "match vector and count bit-exact vs the CPU oracle"
        if E.args.cpu_baseline:
            t0 = time.perf_counter()
            for _ in range(20):
                po.search_by_projection(oc, ol, has, outl, xyz, Tcw, 15.0, True)
            cpu["search_by_projection_1241x376_ms_per_pair"] = (time.perf_counter() - t0) / 20 * 1e3
    # vocabulary transform (Frame::ComputeBoW, src/Frame.cc:279-287) on the reference's tree shape: k=10, L=6, levelsup=4
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
    if E.args.cpu_baseline:
        from oracle import pyoracle as po
        ov = po.OracleVocabulary(10, 6, parent, vdesc, vweight)
        f1 = feats[:VN].copy()
        ov.transform(f1, 4)
        t0 = time.perf_counter()
        for _ in range(10):
            ov.transform(f1, 4)
        cpu["vocabulary_transform_k10_L6_ms_per_frame_1thread"] = (time.perf_counter() - t0) / 10 * 1e3
    return cpu


SMALL_CALL_OFF = {"ORB_SMALL_CALL": "0", "ORB_SELECT_WIDE": "0", "ORB_COMPACT_WIDE": "0", "ORB_FAST_WIDE": "0", "ORB_PDL": "0", "ORB_STAGE_SMALL": "0"}


def frame_latency(E, base, w, h):
    """Blocking single-frame calls (what the reference does once per camera frame, src/Frame.cc:60) through orb_extract: pageable and
    pinned host buffers, the library's default against the same call with every small-call form switched off (the batch kernels and
    launch scheme, i.e. where the round started); the variants are measured in interleaved rounds."""
    import torch
    import orbslam_jpminipc_b200 as pkg
    from orbslam_jpminipc_b200._lib import check, lib, ptr
    L = lib()

    def ctx(env):
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            return pkg.ORBextractor(NFEAT, SCALE, NLEVELS, 1, FAST_TH, device=E.local, max_width=w, max_height=h, max_batch=1)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v

    exs = {"default": ctx({}), "small_call_forms_off": ctx(SMALL_CALL_OFF)}
    cap = exs["default"].capacity
    n = C.c_int(0)
    bufs = {"pageable": (np.stack(base[:8]), np.zeros(cap, pkg.KP_DTYPE), np.zeros((cap, 32), np.uint8)),
            "pinned": (torch.from_numpy(np.stack(base[:8])).pin_memory(), torch.zeros((cap, 7), dtype=torch.int32).pin_memory(),
                       torch.zeros((cap, 32), dtype=torch.uint8).pin_memory())}
    T = {(a, b): [] for a in exs for b in bufs}
    for rnd in range(5):
        for a, ex in exs.items():
            for b, (fr, k, d) in bufs.items():
                for i in range(48):
                    t0 = time.perf_counter()
                    check(L.orb_extract(ex._h, ptr(fr[i % 8]), w, h, w, ptr(k), ptr(d), cap, C.byref(n)), "orb_extract")
                    if i >= 8:
                        T[a, b].append(time.perf_counter() - t0)
    out = {a: {b: {"median_ms": float(np.median(T[a, b]) * 1e3), "p90_ms": float(np.percentile(T[a, b], 90) * 1e3)} for b in bufs} for a in exs}
    for ex in exs.values():
        ex.close()
    return out


def check_frame_sharding(E, res):
    """rank r's result for its frame 0 equals what rank 0 gets when it extracts that same frame itself (frames are regenerated from the
    rank's seed); everything gathered with one all_gather of the (padded) frame-0 outputs"""
    import torch
    import torch.distributed as dist
    from orbslam_jpminipc_b200.synth import synth_frame
    world, rank, dev = E.world, E.rank, E.dev
    n0, k0, d0 = res["_frame0"]
    cap = res["cap"]
    buf = torch.zeros(4 + cap * 60, dtype=torch.uint8)
    buf[:4] = torch.from_numpy(np.array([n0], np.int32).view(np.uint8))
    buf[4:4 + n0 * 28] = torch.from_numpy(k0.reshape(-1))
    buf[4 + cap * 28:4 + cap * 28 + n0 * 32] = torch.from_numpy(d0.reshape(-1))
    mine = buf.to(dev)
    allb = torch.empty((world, buf.numel()), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allb, mine)
    if rank != 0:
        return None
    allb = allb.cpu().numpy()
    ex = res["_ex"]
    for r in range(1, world):
        k, d = ex(synth_frame(H0, W0, 1000 + 100 * r))
        n = int(allb[r, :4].view(np.int32)[0])
        require(n == len(k), "frame sharding: rank %d frame 0 has %d keypoints, rank 0 gets %d for the same frame" % (r, n, len(k)))
        require(np.array_equal(allb[r, 4:4 + n * 28], k.view(np.uint8).reshape(-1)) and
                np.array_equal(allb[r, 4 + cap * 28:4 + cap * 28 + n * 32], d.reshape(-1)), "frame sharding: rank %d frame 0 differs from rank 0's extraction of the same frame" % r)
    return "frame 0 of every rank's timed batch equals rank 0's own extraction of that frame (keypoints + descriptors, bit-exact)"


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import orbslam_jpminipc_b200 as pkg
    from orbslam_jpminipc_b200._lib import lib
    from orbslam_jpminipc_b200.synth import synth_descriptors

    E = Env()
    E.args = args
    E.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    E.rank = rank = int(os.environ.get("RANK", "0"))
    E.local = local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    E.dev = torch.device("cuda", local)
    E.L = lib()
    E.B = args.batch
    E.stream = torch.cuda.current_stream().cuda_stream
    E.e0, E.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=E.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    E.barrier, E.max_over_ranks = barrier, max_over_ranks
    warm = max(args.warmup, 3)
    SL = args.step_launches

    # ---- headline: 640x480 / 1000 kp ----
    r0 = measure_extraction(E, W0, H0, args.steps, warm, SL, full=True)
    E.ex = r0["_ex"]
    roofline = roofline_of(r0, W0, H0, E.B)
    roofline["pipeline_frac"] = (r0["value"] / world) * roofline["pipeline_bytes_per_frame"] / (roofline["peak"] * 1e9)
    verified = {"extraction_640x480": r0.get("verified")} if args.verify else {}
    if args.verify and world > 1:
        verified["frame_sharding"] = check_frame_sharding(E, r0)
    latency = None
    if rank == 0 and not args.quick:
        latency = {"api": "orb_extract (one 640x480 frame per blocking call, python ctypes caller, 200 timed calls per variant in interleaved rounds)",
                   "what": "default = the small-call forms of the pass (calls of <= 12 frames: short resize tiles, CTA-per-cell compaction, 32-warp selection, 512-thread FAST CTAs, programmatic dependent launch inside the replayed graph, one staged result block for pageable outputs); small_call_forms_off = the batch kernels and launch scheme for one frame",
                   **frame_latency(E, r0["_base"], W0, H0)}

    # ---- configs[1]: 752x480 ----
    config1 = None
    if not args.quick:
        r1 = measure_extraction(E, W1, H1, max(args.steps // 2, 2), 3, SL, full=False)
        rf1 = roofline_of(r1, W1, H1, E.B)
        config1 = {"workload": "batched ORB extraction, 752x480 EuRoC-shaped synthetic frames, 1000 kp (BASELINE.json configs[1])",
                   "value": r1["value"], "unit": "frames/s", "ms_per_step": r1["ms_per_step"], "mean_keypoints": r1["mean_keypoints"],
                   "e2e": r1["e2e"], "roofline": {k: rf1[k] for k in ("kernel", "achieved", "peak", "frac", "kernel_ms_per_launch", "stage_ms_per_launch")},
                   "pipeline_frac": (r1["value"] / world) * rf1["pipeline_bytes_per_frame"] / (rf1["peak"] * 1e9)}
        if args.verify:
            verified["extraction_752x480"] = r1.get("verified")
        r1["_ex"].close()

    # ---- matching (Hamming kNN-2) ----
    matching, cpu_extra = None, {}
    if not args.skip_matching:
        matching = run_matching(E)
        if rank == 0 and not args.quick:
            cpu_extra = run_tracking_extras(E, matching)
        verified.update({"matching_" + k: v for k, v in matching.pop("verified").items()})
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cfg = shared_config()
    line = {"metric": METRIC, "value": r0["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": r0["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": cfg,
            "step": {"frames_per_gpu": E.B * SL, "launches": SL, "frames_per_launch": E.B, "mean_keypoints": r0["mean_keypoints"],
                     "sharding": "frames sharded over %d GPU(s), no collective on the extraction path" % world,
                     "cpu_affinity": ("rank bound to the %d CPUs local to its GPU (NVML)" % numa) if numa else "unbound"},
            "clocks": r0["clocks"], "gpu_launches": r0["gpu_launches"], "e2e": r0["e2e"], "roofline": roofline,
            "matches_per_s": None if matching is None else {
                "unit": "descriptor pairs/s (a 'match' = one query's best / second-best over the rows: queries_per_s)",
                "pair_blocks_2000x2000": matching["pair_blocks_2000x2000"], "db_sharded_10M": matching["db_sharded"],
                "tensor_core_experiment": matching["tensor_core_experiment"],
                "popc_peak_gops_measured": matching["popc_peak_gops"], "popc_peak_gops_theoretical": POPC_THEORETICAL_GOPS,
                "popc_peak_source": matching["popc_peak_source"],
                "roofline": "POPC pipe: 8 POPC32 per descriptor pair (SURVEY.md §8d)"},
            "verified": (verified if args.verify else False), "merge_check": (matching or {}).get("db_sharded", {}).get("merge_check"),
            "config1_752x480": config1, "single_frame_latency": latency,
            "tracking_extras": None if matching is None else {k: matching[k] for k in ("search_by_projection_1241x376", "vocabulary_transform_k10_L6") if k in matching}}
    if args.cpu_baseline:
        cores = os.cpu_count() or 1
        base = r0["_base"]
        fps1, pc1, dt1, n1 = cpu_extract_protocol(base, 1, 200, warm=20)
        fpsN, pcN, dtN, nN = cpu_extract_protocol(base, cores, 200 * cores, warm=20 * cores)       # the protocol's 20 + 200 frames on EVERY thread
        db4, q4 = synth_descriptors(20000, 2000)
        cb = {"value": fpsN, "unit": "frames/s", "cores": cores, "kind": cpu_kind(),
              "sample": "BASELINE.md §2 protocol on the headline workload (640x480, 1000 kp): %d timed frames after 20 warm-up on 1 thread (%.1f s), %d timed frames "
                        "after %d warm-up frame-parallel over %d host threads, one extractor per thread (%.1f s); %s"
                        % (n1, dt1, nN, 20 * cores, cores, dtN, CPU_WHAT[cpu_kind()]),
              "single_thread": dict(pc1, value=fps1, frames=n1), "all_cores": dict(pcN, value=fpsN, frames=nN),
              "single_thread_value": fps1}
        if not args.skip_matching:
            cb["matcher_pairs_per_s"] = {
                "workload": "brute-force best / second-best, 2000 queries x 20000 rows (256-bit), 3 repetitions; kind: port (oracle restatement of src/ORBmatcher.cc:197-222 + :1794-1810; the reference has no stand-alone kNN entry point)",
                "bit_hack_1thread": cpu_knn_protocol(q4, db4, 1, False), "popcountll_1thread": cpu_knn_protocol(q4, db4, 1, True),
                "bit_hack_all_cores": cpu_knn_protocol(q4, db4, cores, False), "popcountll_all_cores": cpu_knn_protocol(q4, db4, cores, True)}
        cb.update(cpu_extra)
        line["cpu_baseline"] = cb
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="frames per GPU per launch (measured on B200, 640x480: 256 -> 148.3 K, 512 -> 153.7 K, 1024 -> 156.7 K frames/s: fewer stage boundaries per frame)")
    ap.add_argument("--step-launches", type=int, default=8, help="launches of --batch frames per step (8 x 1024 = 8192 frames per GPU and step)")
    ap.add_argument("--db-rows", type=int, default=10_000_000)
    ap.add_argument("--match-reps", type=int, default=400)
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-verify", dest="verify", action="store_false")
    ap.add_argument("--skip-matching", action="store_true")
    ap.add_argument("--no-tensor-ab", dest="tensor_ab", action="store_false", help="skip the tensor-core kNN A/B")
    ap.add_argument("--quick", action="store_true", help="headline measurement only (profiling runs): no 752x480 pass, no latency / tracking extras")
    ap.add_argument("--e2e-chunk", type=int, default=64)
    ap.add_argument("--wc-input", action="store_true", help="input frames in write-combined pinned memory (orb_host_alloc_input)")
    args = ap.parse_args()
    capture_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
