"""Host models of the integer identities the extraction kernels lean on (orbslam_jpminipc_b200/csrc/orb_extract.cu), checked
exhaustively with numpy so that a change to one of the constants fails on the CPU before it reaches a GPU:

* k_fast_nms NMS decision: in a 16-bit lane, `s - min(s, m) + 0x7fff` has its sign bit set exactly when s > m (s, m in 0..255, no
  carry into the neighbouring lane); a sign-replicating PRMT (selector 0xfdb9) turns the four sign bits into byte masks; the four
  mask bytes are gathered into four bitmap bits by `(mask & 0x01010101) * 0x01020408 >> 24`; `c + (th-1) * 0x01010101` adds th-1 to
  every score byte without a carry while score + th - 1 <= 254.
* task numbering of partly filled tiles: floor(task / n) == task * ceil(2^20 / n) >> 20 (k_fast_nms) and
  task * ceil(2^16 / n) >> 16 (k_blur) over the ranges those kernels use, in 32-bit arithmetic.
* k_fast_nms epilogue: the packed excess max(Mn - v - th, v - th - Mx, 0) equals max(T - th, 0) with T the FAST strength."""
import numpy as np


def test_nms_lane_sign_trick_exhaustive():
    s, m = np.meshgrid(np.arange(256, dtype=np.uint32), np.arange(256, dtype=np.uint32), indexing="ij")
    y = s - np.minimum(s, m) + 0x7fff
    assert y.max() < 0x10000                                   # stays inside the 16-bit lane
    assert np.array_equal((y >> 15) & 1, (s > m).astype(np.uint32))
    # two lanes packed in one 32-bit register behave independently
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, (100000, 2), dtype=np.uint32)
    b = rng.integers(0, 256, (100000, 2), dtype=np.uint32)
    packed_s = a[:, 0] | (a[:, 1] << 16)
    packed_min = np.minimum(a[:, 0], b[:, 0]) | (np.minimum(a[:, 1], b[:, 1]) << 16)
    y2 = (packed_s - packed_min + np.uint32(0x7fff7fff)) & np.uint32(0xffffffff)
    assert np.array_equal((y2 >> 15) & 1, (a[:, 0] > b[:, 0]).astype(np.uint32))
    assert np.array_equal((y2 >> 31) & 1, (a[:, 1] > b[:, 1]).astype(np.uint32))


def _prmt(a, b, sel):
    """PTX prmt.b32 (generic mode) on python ints."""
    src = [(a >> (8 * i)) & 0xff for i in range(4)] + [(b >> (8 * i)) & 0xff for i in range(4)]
    out = 0
    for i in range(4):
        n = (sel >> (4 * i)) & 0xf
        byte = src[n & 7]
        if n & 8:
            byte = 0xff if byte & 0x80 else 0x00
        out |= byte << (8 * i)
    return out


def test_nms_mask_and_bitmap_bits():
    for surv in range(16):                                      # which of the four pixels survive
        ylo = (0x8001 if surv & 1 else 0x7fff) | ((0x80fe if surv & 2 else 0x7fff) << 16)
        yhi = (0xffff if surv & 4 else 0x0000) | ((0x8000 if surv & 8 else 0x7fff) << 16)
        m4 = _prmt(ylo, yhi, 0xfdb9)
        assert m4 == sum(0xff << (8 * i) for i in range(4) if surv >> i & 1)
        bits = (((m4 & 0x01010101) * 0x01020408) & 0xffffffff) >> 24
        assert bits == surv
    for th in range(1, 8):                                      # th = max(min(fastTh, 7), 1)
        c = np.arange(0, 256 - th, dtype=np.uint32)             # excess = T - th <= 255 - th
        word = c | (c << 8) | (c << 16) | (c << 24)
        out = (word + np.uint32((th - 1) * 0x01010101)) & np.uint32(0xffffffff)
        for i in range(4):
            assert np.array_equal((out >> (8 * i)) & 0xff, c + th - 1)


def test_multiply_shift_division_is_exact():
    for n in range(1, 67):
        inv = ((1 << 20) + n - 1) // n
        t = np.arange(0, 4000, dtype=np.uint64)
        assert (t * inv).max() < 2 ** 32
        assert np.array_equal((t * inv) >> 20, t // n), n
    for n in range(1, 17):
        inv = ((1 << 16) + n - 1) // n
        t = np.arange(0, 62 * 16 + 256, dtype=np.uint64)
        assert np.array_equal((t * inv) >> 16, t // n), n


def test_excess_epilogue_equals_strength_minus_threshold():
    rng = np.random.default_rng(2)
    v = rng.integers(0, 256, 200000).astype(np.int64)
    mn = rng.integers(0, 256, 200000).astype(np.int64)          # max over arcs of the arc minimum
    mx = rng.integers(0, 256, 200000).astype(np.int64)          # min over arcs of the arc maximum
    for th in (1, 7):
        strength = np.maximum(mn - v, v - mx)
        want = np.maximum(strength - th, 0)
        nvp, vm1 = -(v + th), v - th + 1                        # per pixel pair, once
        qb = np.maximum(mn + nvp, 0)                            # VIADDMNMX.RELU
        got = np.maximum((-mx - 1) + vm1, qb)                   # ~Mx = -Mx - 1 in a 16-bit lane
        assert np.array_equal(got, want)


# ---- high-byte-lane NMS (the default NMS lane layout of k_fast_nms; -DORB_NMS_LOBYTE selects the older PRMT-unpacking form) ---------------------------------
# numpy restatement of the exact lane operations of that variant, compared with the plain definition of the masked strict
# 8-neighbour test on tie-heavy random score words.
U=np.uint32
def lanes(x): return (x & U(0xffff)), (x >> U(16))
def pack(lo,hi): return (lo | (hi << U(16))).astype(U)
def vmax(a,b):
    al,ah=lanes(a); bl,bh=lanes(b); return pack(np.maximum(al,bl), np.maximum(ah,bh))
def vmin(a,b):
    al,ah=lanes(a); bl,bh=lanes(b); return pack(np.minimum(al,bl), np.minimum(ah,bh))
def vmax3(a,b,c): return vmax(vmax(a,b),c)
def funnel_r(lo,hi,sh):   # __funnelshift_r(lo, hi, sh): (hi:lo) >> sh, low 32 bits
    v=(hi.astype(np.uint64)<<np.uint64(32))|lo.astype(np.uint64)
    return ((v>>np.uint64(sh)) & np.uint64(0xffffffff)).astype(U)
def prmt_sign(a,b,sel):
    src=[(a>>U(8*i))&U(0xff) for i in range(4)]+[(b>>U(8*i))&U(0xff) for i in range(4)]
    out=np.zeros_like(a)
    for i in range(4):
        n=(sel>>(4*i))&0xf
        byte=src[n&7]
        if n&8: byte=np.where(byte&U(0x80),U(0xff),U(0)).astype(U)
        out|=byte<<U(8*i)
    return out
def hibyte_nms(pc,c,nc,ul,uc,ur,dl,dc,dr,ml,mr,up,dn):
    z=U(0)
    ul=np.where(up,ul,z); uc=np.where(up,uc,z); ur=np.where(up,ur,z)
    dl=np.where(dn,dl,z); dc=np.where(dn,dc,z); dr=np.where(dn,dr,z)
    c8=(c<<U(8)); u8=(uc<<U(8)); d8=(dc<<U(8)); ml8=ml<<U(8); mr8=mr<<U(8)
    Lo=vmax3(c8,u8,d8)&ml
    Ro=vmax3(funnel_r(c,nc,8),funnel_r(uc,ur,8),funnel_r(dc,dr,8))&mr
    Mo=vmax3(Lo,Ro,vmax(uc,dc))
    Le=vmax3(funnel_r(pc,c,16),funnel_r(ul,uc,16),funnel_r(dl,dc,16))&ml8
    Re=vmax3(c,uc,dc)&mr8
    Me=vmax3(Le,Re,vmax(u8,d8))
    ao=(c>>U(1))&U(0x7f807f80); bo=((Mo>>U(1))|U(0x007f007f))&U(0x7fff7fff)
    ae=(c8>>U(1))&U(0x7f807f80); be=((Me>>U(1))|U(0x007f007f))&U(0x7fff7fff)
    yo=(ao-vmin(ao,bo)+U(0x7fff7fff)).astype(U); ye=(ae-vmin(ae,be)+U(0x7fff7fff)).astype(U)
    return prmt_sign(ye,yo,0xfbd9)
def direct(pc,c,nc,ul,uc,ur,dl,dc,dr,ml,mr,up,dn):
    def row(p,x,n):   # 12 bytes: p(4) x(4) n(4)
        return np.stack([(w>>U(8*i))&U(0xff) for w in (p,x,n) for i in range(4)],-1).astype(np.int64)
    R=row(pc,c,nc); Uu=row(ul,uc,ur)*up[...,None]; D=row(dl,dc,dr)*dn[...,None]
    out=np.zeros_like(c)
    for i in range(4):
        k=4+i
        l=((ml>>U(8*i))&U(1)).astype(np.int64); r=((mr>>U(8*i))&U(1)).astype(np.int64)
        nb=np.stack([R[...,k-1]*l,R[...,k+1]*r,Uu[...,k],Uu[...,k-1]*l,Uu[...,k+1]*r,D[...,k],D[...,k-1]*l,D[...,k+1]*r],-1).max(-1)
        out|=np.where(R[...,k]>nb,U(0xff),U(0)).astype(U)<<U(8*i)
    return out


def test_nms_high_byte_lanes_model():
    rng = np.random.default_rng(0)
    n = 100000
    for trial in range(8):
        def word(vals):
            return (vals[:, 0] | (vals[:, 1] << U(8)) | (vals[:, 2] << U(16)) | (vals[:, 3] << U(24))).astype(U)
        def score_word():      # tie-heavy small scores or the full byte range, 40 % zeros
            b = rng.integers(0, 6 if trial % 2 else 256, (n, 4)).astype(U) * (rng.random((n, 4)) < 0.6)
            return word(b.astype(U))
        def mask_word():
            return word(np.where(rng.random((n, 4)) < 0.85, 0xff, 0).astype(U))
        W = [score_word() for _ in range(9)]
        ml, mr = mask_word(), mask_word()
        up, dn = rng.random(n) < 0.9, rng.random(n) < 0.9
        assert np.array_equal(hibyte_nms(*W, ml, mr, up, dn), direct(*W, ml, mr, up, dn)), trial


# ---- k_knn2: carry-save compression of the eight XOR words + packed (distance, row) keys (csrc/orb_match.cu) -------------------------
def test_knn2_carry_save_popcount_and_packed_key_model():
    """numpy model of the default k_knn2 inner loop: three carry-save adders (LOP3 0x96 / 0xE8) turn 8 POPC into 5, and the running
    best / second best are two packed keys with  m2 = min(m2, max(m1, key)); m1 = min(m1, key).  Checked against the plain
    definition (sum of 8 popcounts; strict '<' scan: first minimum wins, d2 = second order statistic of the distance multiset)."""
    rng = np.random.default_rng(12)
    popc = lambda a: np.unpackbits(np.ascontiguousarray(a).view(np.uint8).reshape(a.shape + (4,)), axis=-1).sum(-1).astype(np.int64)
    maj = lambda a, b, c: (a & b) | (a & c) | (b & c)
    for trial in range(20):
        nrows = int(rng.integers(1, 700))
        q = rng.integers(0, 2**32, 8, dtype=np.uint64).astype(np.uint32)
        db = rng.integers(0, 2**32, (nrows, 8), dtype=np.uint64).astype(np.uint32)
        if trial % 3 == 0:                                  # ties: copies of a few rows, and near copies of the query
            db[rng.integers(0, nrows, nrows // 2 + 1)] = db[0]
            db[rng.integers(0, nrows, 3)] = q ^ np.uint32(1 << int(rng.integers(0, 32)))
        x = db ^ q[None, :]
        s1, c1 = x[:, 0] ^ x[:, 1] ^ x[:, 2], maj(x[:, 0], x[:, 1], x[:, 2])
        s2, c2 = x[:, 3] ^ x[:, 4] ^ x[:, 5], maj(x[:, 3], x[:, 4], x[:, 5])
        s3, c3 = s1 ^ s2 ^ x[:, 6], maj(s1, s2, x[:, 6])
        d = popc(s3) + popc(x[:, 7]) + 2 * (popc(c1) + popc(c2) + popc(c3))
        assert np.array_equal(d, popc(x).sum(1))
        m1 = m2 = 2**31 - 1
        for r in range(nrows):
            key = int(d[r]) * (1 << 22) + r
            m2 = min(m2, max(m1, key))
            m1 = min(m1, key)
        e1, e2, ei = 2**31 - 1, 2**31 - 1, -1                # the reference scan, src/ORBmatcher.cc:197-222
        for r in range(nrows):
            if d[r] < e1:
                e2, e1, ei = e1, int(d[r]), r
            elif d[r] < e2:
                e2 = int(d[r])
        assert (m1 >> 22, m1 & (2**22 - 1)) == (e1, ei)
        assert (2**31 - 1 if m2 == 2**31 - 1 else m2 >> 22) == e2


def test_border_ring_covers_everything_the_path_reads():
    """csrc/orb_internal.h ORB_RING = 4: k_border writes only a 4 px reflect-101 ring instead of the reference's 16 px frame.  That is
    enough iff nothing reads further than 3 px past the ROI: keypoints lie >= 16 px inside (EDGE_THRESHOLD), the rotated rBRIEF pattern
    reaches at most ceil(max |p|) px from a keypoint, GaussianBlur 7x7 reaches 3 px from a ROI pixel."""
    import math
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "orbslam_jpminipc_b200", "csrc", "orb_internal.h")).read()
    ring = int(re.search(r"#define ORB_RING (\d+)", hdr).group(1))
    edge = int(re.search(r"#define ORB_EDGE (\d+)", hdr).group(1))
    lines = [l for l in open(os.path.join(root, "orbslam_jpminipc_b200", "csrc", "orb_pattern.inc")) if not l.strip().startswith("//")]
    nums = [int(x) for x in re.findall(r"-?\d+", " ".join(lines))]
    assert len(nums) == 1024
    reach = max(math.hypot(nums[i], nums[i + 1]) for i in range(0, 1024, 2))
    # cvRound(x*b + y*a) of a point at distance r from the keypoint is at most round(r) away on either axis
    assert round(reach) - edge <= ring - 1 and 3 <= ring - 1 + 0 and ring % 4 == 0


def test_quota_loop_bound():
    """orb_plan.cu sizes the per-level keypoint list as ncells * nfCell + ncells (+ margin): the quota redistribution of
    src/ORBextractor.cc:622-670 never leaves more, whatever the per-cell key counts (random and adversarial), and the bound is tight
    to within two records."""
    import math
    rng = np.random.default_rng(5)

    def quota(t, q):
        n = len(t)
        retain = np.zeros(n, np.int64); closed = np.zeros(n, bool)
        d = 0
        for c in range(n):
            if t[c] > q: retain[c] = q
            else: retain[c] = t[c]; d += q - t[c]; closed[c] = True
        while d > 0 and closed.sum() < n:
            new = q + int(math.ceil(np.float32(d) / np.float32(n - closed.sum())))
            d = 0
            for c in range(n):
                if closed[c]: continue
                if t[c] > new: retain[c] = new
                else: retain[c] = t[c]; d += new - t[c]; closed[c] = True
        return int(retain.sum())
    worst = -10 ** 9
    for trial in range(4000):
        n = int(rng.integers(1, 60)); nd = int(rng.integers(1, 600)); q = -(-nd // n)
        mode = trial % 4
        if mode == 0: t = rng.integers(0, 3 * q + 2, n)
        elif mode == 1: t = rng.integers(0, 50 * q + 2, n) * (rng.random(n) < 0.3)
        elif mode == 2: t = np.sort(rng.integers(0, q * n + 5, n))
        else: t = (q + np.arange(n) * int(rng.integers(0, 4)) + rng.integers(-2, 3, n)).clip(0)
        total = quota(t, q)
        assert total <= n * q + n, (n, q, list(t))
        worst = max(worst, total - n * q - n)
    assert worst >= -4                       # tight: some input comes within a few records of the bound


def test_half_lane_min_max_model():
    """k_fast_nms round 2: a lane holds 0x6400 | pixel = the fp16 number 1024 + pixel.  (1) positive halves order like their bit
    patterns, so the integer 16-bit min / max keep working; (2) d = relu(a - b), max = b + d, min = a - d are exact in fp16 for every
    pixel pair; (3) the epilogue's lane constants absorb the 0x6400 without leaving a signed 16-bit lane."""
    px = np.arange(256, dtype=np.uint16)
    bits = (0x6400 | px).astype(np.uint16)
    half = bits.view(np.float16)
    assert np.array_equal(half.astype(np.float32), 1024.0 + px)
    a, b = np.meshgrid(half, half, indexing="ij")
    d = np.maximum((a - b).astype(np.float16), np.float16(0))                # fma.rn.relu.f16x2(b, -1, a): one rounding, exact here
    assert np.array_equal((a.astype(np.float64) - b.astype(np.float64)).clip(0), d.astype(np.float64))
    mx, mn = (b + d).astype(np.float16), (a - d).astype(np.float16)
    assert np.array_equal(mx.view(np.uint16), np.maximum(a.view(np.uint16), b.view(np.uint16)))
    assert np.array_equal(mn.view(np.uint16), np.minimum(a.view(np.uint16), b.view(np.uint16)))
    # epilogue: q = max(Mn + nvp', ~Mx + vm1', 0) with nvp' = -(v + th) - 0x6400, vm1' = v - th + 1 + 0x6400 in int16 lanes
    for th in (1, 7, 20, 255):
        v, m = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
        Mn = (0x6400 + m).astype(np.int32); Mx = Mn
        nvp = (-(v + th) - 0x6400).astype(np.int32); vm1 = (v - th + 1 + 0x6400).astype(np.int32)
        for term in (Mn + nvp, (-Mx - 1) + vm1):
            assert term.min() >= -32768 and term.max() <= 32767 and nvp.min() >= -32768 and vm1.max() <= 32767
        assert np.array_equal(np.maximum(Mn + nvp, 0), np.maximum(m - v - th, 0))
        assert np.array_equal(np.maximum((-Mx - 1) + vm1, 0), np.maximum(v - m - th, 0))


# ---------------------------------------------------------------- round 2, last session: matcher / small-call restructurings
def test_grid_column_runs_equal_the_cell_by_cell_walk():
    """warp_window_walk (orb_match.cu): Frame::GetFeaturesInArea scans ix outer, iy inner, insertion order (src/Frame.cc:233-259).  The
    grid's CSR is indexed by cell id = ix * 48 + iy, so the cells iy = y0..y1 of one column are ONE contiguous run of cell_items and
    the concatenation of the runs of columns x0..x1 is that scan order.  Also models the run lookup of the kernel: item t of the
    concatenation lies in run #(number of inclusive prefix sums <= t) at start_run + t with start_run = b_run - exclusive prefix."""
    rng = np.random.default_rng(11)
    COLS, ROWS = 64, 48
    for trial in range(200):
        n = int(rng.integers(0, 3000))
        cell = np.sort(rng.integers(0, COLS * ROWS, n)) if trial % 3 else np.sort(rng.integers(0, 40, n))   # crowded corner too
        cell_start = np.searchsorted(cell, np.arange(COLS * ROWS + 1)).astype(np.int64)
        items = rng.permutation(n)                                   # cell_items: any payload, the ORDER inside the CSR is what matters
        x0, x1 = sorted(rng.integers(0, COLS, 2)); y0, y1 = sorted(rng.integers(0, ROWS, 2))
        want = [items[j] for ix in range(x0, x1 + 1) for iy in range(y0, y1 + 1)
                for j in range(cell_start[ix * ROWS + iy], cell_start[ix * ROWS + iy + 1])]
        got = []
        for c0 in range(x0, x1 + 1, 32):                              # one trip of the kernel: up to 32 columns, one per lane
            cols = list(range(c0, min(c0 + 32, x1 + 1)))
            b = np.array([cell_start[c * ROWS + y0] for c in cols] + [0] * (32 - len(cols)))
            e = np.array([cell_start[c * ROWS + y1 + 1] for c in cols] + [0] * (32 - len(cols)))
            ln = e - b
            incl = np.cumsum(ln)
            start = b - (incl - ln)
            for t in range(int(incl[-1])):
                col = 0
                for step in (16, 8, 4, 2, 1):                         # the kernel's shuffle binary search
                    if incl[col + step - 1] <= t:
                        col += step
                assert col == int(np.searchsorted(incl, t, side="right"))
                got.append(items[start[col] + t])
        assert got == want


def test_sorted_candidate_list_gives_the_scan_order_best_and_second():
    """warp_sort_candidates + the resolution passes (orb_match.cu): the reference scans a query's candidates in order and keeps
    best / second best with strict '<' (src/ORBmatcher.cc:1559-1574, :466-473, :85-109), skipping claimed keypoints.  With the list
    sorted by (distance, scan position) the FIRST unclaimed entry is that best and the second unclaimed one is the second best
    (value and identity), for any set of claimed entries."""
    rng = np.random.default_rng(12)
    for trial in range(3000):
        n = int(rng.integers(1, 40))
        dist = rng.integers(0, 12, n) if trial % 2 else rng.integers(0, 257, n)      # tie-heavy and plain
        ids = rng.permutation(1000)[:n]
        claimed = rng.random(n) < rng.choice([0.0, 0.3, 0.8])
        # reference loop
        bd, bd2, bi, bi2 = 1 << 30, 1 << 30, -1, -1
        for p in range(n):
            if claimed[p]:
                continue
            if dist[p] < bd:
                bd2, bi2 = bd, bi
                bd, bi = dist[p], ids[p]
            elif dist[p] < bd2:
                bd2, bi2 = dist[p], ids[p]
        # rank sort by key (dist << 10 | position), as the kernel does by counting smaller keys
        key = (dist.astype(np.int64) << 10) | np.arange(n)
        rank = np.array([(key < k).sum() for k in key])
        assert sorted(rank) == list(range(n))
        order = np.empty(n, np.int64); order[rank] = np.arange(n)
        free = [p for p in order if not claimed[p]]
        gd, gi = (dist[free[0]], ids[free[0]]) if free else (1 << 30, -1)
        gd2, gi2 = (dist[free[1]], ids[free[1]]) if len(free) > 1 else (1 << 30, -1)
        assert (gd, gi, gd2) == (bd, bi, bd2)
        if len(free) > 1:
            assert gi2 == bi2                                         # the level-ratio rule (:114-117) reads the second's octave


def test_quota_redistribution_by_lanes_equals_the_serial_loop():
    """k_select_fast: the reference's quota loop (src/ORBextractor.cc:622-670) walks the cells serially, but a pass treats every cell
    independently given nNew and carries only two sums into the next pass; the kernel gives the cells to the 32 lanes and reduces
    the sums.  Same retained counts and the same offsets for any cell counts, skipped cells included."""
    import math
    rng = np.random.default_rng(13)
    for trial in range(2000):
        ncells = int(rng.integers(1, 200)); nfc = int(rng.integers(1, 20))
        total = rng.integers(0, 60, ncells) * (rng.random(ncells) < 0.8)
        skipped = rng.random(ncells) < 0.1
        total = np.where(skipped, 0, total)

        def serial():
            retain = np.zeros(ncells, np.int64); nomore = np.zeros(ncells, bool); nno = 0; ntd = 0
            for c in range(ncells):
                if skipped[c]:
                    continue
                if total[c] > nfc:
                    retain[c] = nfc
                else:
                    retain[c] = total[c]; ntd += nfc - total[c]; nomore[c] = True; nno += 1
            while ntd > 0 and nno < ncells:
                nnew = nfc + int(math.ceil(np.float32(ntd) / np.float32(ncells - nno)))
                ntd = 0
                for c in range(ncells):
                    if nomore[c]:
                        continue
                    if total[c] > nnew:
                        retain[c] = nnew
                    else:
                        retain[c] = total[c]; ntd += nnew - total[c]; nomore[c] = True; nno += 1
            return retain

        def by_lanes():
            retain = np.zeros(ncells, np.int64); nomore = np.zeros(ncells, bool)
            part_no = np.zeros(32, np.int64); part_td = np.zeros(32, np.int64)
            for lane in range(32):
                for c in range(lane, ncells, 32):
                    if skipped[c]:
                        continue
                    if total[c] > nfc:
                        retain[c] = nfc
                    else:
                        retain[c] = total[c]; part_td[lane] += nfc - total[c]; nomore[c] = True; part_no[lane] += 1
            nno, ntd = int(part_no.sum()), int(part_td.sum())
            while ntd > 0 and nno < ncells:
                nnew = nfc + int(math.ceil(np.float32(ntd) / np.float32(ncells - nno)))
                d = np.zeros(32, np.int64); m = np.zeros(32, np.int64)
                for lane in range(32):
                    for c in range(lane, ncells, 32):
                        if nomore[c]:
                            continue
                        if total[c] > nnew:
                            retain[c] = nnew
                        else:
                            retain[c] = total[c]; d[lane] += nnew - total[c]; nomore[c] = True; m[lane] += 1
                ntd = int(d.sum()); nno += int(m.sum())
            return retain
        a, b = serial(), by_lanes()
        assert np.array_equal(a, b)
