"""Fuzz sweep of the host-buffer pipeline (orb_extract_batch_async / orb_wait / orb_extract_batch / orb_extract_batch_device on one
context): random context batch size, call sizes above and below it (chunking over both work sets, the out streams), several calls in
flight, waits in random order, device-pointer calls in between; every frame must equal its blocking single-frame extraction:
   gpurun -- 'python tools/fuzz_async.py 0 60'"""
import sys, time, ctypes as C
import numpy as np
sys.path.insert(0, ".")
import torch
import orbslam_jpminipc_b200 as pkg
from orbslam_jpminipc_b200._lib import check, lib, ptr
from orbslam_jpminipc_b200.synth import synth_frames
L = lib()
bad = 0; n = 0; t0 = time.time()
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(20000 + seed)
    h, w = [(240, 320), (200, 300), (300, 400)][int(rng.integers(0, 3))]
    nf = int(rng.choice([200, 300, 400]))
    pool = np.stack(synth_frames(24, h, w, seed0=3000 + seed))
    one = pkg.ORBextractor(nf, 1.2, 8, 1, 20, max_width=w, max_height=h, max_batch=1)
    ref = [one(f) for f in pool]
    mb = int(rng.integers(1, 17))
    ex = pkg.ORBextractor(nf, 1.2, 8, 1, 20, max_width=w, max_height=h, max_batch=mb)
    cap = ex.capacity
    calls = []
    for c in range(int(rng.integers(2, 9))):
        nimg = int(rng.integers(1, 3 * mb + 2))
        sel = rng.integers(0, len(pool), nimg)
        fr = torch.from_numpy(np.ascontiguousarray(pool[sel])).pin_memory()
        k = np.zeros((nimg, cap), pkg.KP_DTYPE); d = np.zeros((nimg, cap, 32), np.uint8); cnt = np.zeros(nimg, np.int32)
        kind = rng.random()
        if kind < 0.7:
            t = C.c_longlong(-1)
            check(L.orb_extract_batch_async(ex._h, ptr(fr), nimg, w, h, w, w * h, ptr(k), ptr(d), cap, ptr(cnt), C.byref(t)), "async")
            calls.append((sel, k, d, cnt, t.value, fr))
        elif kind < 0.85:
            check(L.orb_extract_batch(ex._h, ptr(fr), nimg, w, h, w, w * h, ptr(k), ptr(d), cap, ptr(cnt)), "sync")
            calls.append((sel, k, d, cnt, None, fr))
        else:                                       # device pointers on torch's stream, at most max_batch frames
            m_ = min(nimg, mb); sel = sel[:m_]
            dfr = fr[:m_].cuda()
            dk = torch.zeros((m_, cap, 7), dtype=torch.int32, device="cuda"); dd = torch.zeros((m_, cap, 32), dtype=torch.uint8, device="cuda")
            dc = torch.zeros(m_, dtype=torch.int32, device="cuda")
            check(L.orb_extract_batch_device(ex._h, ptr(dfr), m_, w, h, w, w * h, ptr(dk), ptr(dd), cap, ptr(dc), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "device")
            torch.cuda.synchronize()
            calls.append((sel, dk.cpu().numpy().view(np.uint8).reshape(m_, cap, 28).copy().view(pkg.KP_DTYPE).reshape(m_, cap), dd.cpu().numpy(), dc.cpu().numpy(), None, None))
        if rng.random() < 0.3 and calls:            # wait for a random earlier ticket now
            j = int(rng.integers(0, len(calls)))
            if calls[j][4] is not None: check(L.orb_wait(ex._h, calls[j][4]), "wait")
    order = [c for c in calls if c[4] is not None]
    rng.shuffle(order)
    for c in order: check(L.orb_wait(ex._h, c[4]), "wait")
    for ci, (sel, k, d, cnt, tk, _) in enumerate(calls):
        for i, src in enumerate(sel):
            rk, rd = ref[src]
            n += 1
            ok = cnt[i] == len(rk) and np.array_equal(k[i, :cnt[i]].view(np.uint8), rk.view(np.uint8)) and np.array_equal(d[i, :cnt[i]], rd)
            if not ok:
                bad += 1; print("MISMATCH seed", seed, "max_batch", mb, "call", ci, "frame", i, "kind", "async" if tk is not None else "sync/device", int(cnt[i]), len(rk))
    ex.close(); one.close()
print("frames checked", n, "bad", bad, "%.1f s" % (time.time() - t0))
