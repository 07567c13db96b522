"""Deterministic synthetic inputs (pure numpy, no OpenCV) for tests and bench.py.

Frames follow SURVEY.md §8d: band-limited noise texture, a low-contrast top-right quadrant
(FAST@20 starves, FAST@7 fires -> the reference's per-cell threshold fallback,
src/ORBextractor.cc:609-614) and a flat bottom-right quadrant (empty cells -> quota
redistribution, :644-670).  Descriptor sets follow the config-4/5 recipe: uniform random
256-bit rows with planted near neighbours and exact duplicates (tie-breaking).
"""
import numpy as np


def _gauss_sep(a, sigma, radius):
    k = np.exp(-0.5 * (np.arange(-radius, radius + 1, dtype=np.float64) / sigma) ** 2)
    k = (k / k.sum()).astype(np.float32)
    p = np.pad(a, ((0, 0), (radius, radius)), mode="reflect")
    out = np.zeros_like(a)
    for i in range(2 * radius + 1):
        out += k[i] * p[:, i:i + a.shape[1]]
    p = np.pad(out, ((radius, radius), (0, 0)), mode="reflect")
    out2 = np.zeros_like(a)
    for i in range(2 * radius + 1):
        out2 += k[i] * p[i:i + a.shape[0], :]
    return out2


def _bilinear_up(lo, h, w):
    ys = (np.arange(h, dtype=np.float32) + 0.5) * (lo.shape[0] - 1) / h
    xs = (np.arange(w, dtype=np.float32) + 0.5) * (lo.shape[1] - 1) / w
    y0 = np.floor(ys).astype(np.int64)
    x0 = np.floor(xs).astype(np.int64)
    y0 = np.clip(y0, 0, lo.shape[0] - 2)
    x0 = np.clip(x0, 0, lo.shape[1] - 2)
    fy = (ys - y0).astype(np.float32)[:, None]
    fx = (xs - x0).astype(np.float32)[None, :]
    a = lo[y0][:, x0]
    b = lo[y0][:, x0 + 1]
    c = lo[y0 + 1][:, x0]
    d = lo[y0 + 1][:, x0 + 1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def synth_frame(h, w, seed, quadrants=True):
    """One grayscale uint8 frame (h, w)."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, (h, w)).astype(np.float32)
    lo = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2)).astype(np.float32)
    b = _gauss_sep(a, 1.5, 4) + np.float32(0.5) * _bilinear_up(lo, h, w)
    b = (b - b.min()) / (b.max() - b.min()) * np.float32(255)
    if quadrants:
        y0, x0 = int(.45 * h), int(.55 * w)
        b[:y0, x0:] = 128 + (b[:y0, x0:] - 128) * np.float32(0.3)
        y1 = int(.55 * h)
        b[y1:, x0:] = 128 + rng.integers(-2, 3, (h - y1, w - x0)).astype(np.float32)
    return np.clip(np.rint(b), 0, 255).astype(np.uint8)


def synth_frames(n, h, w, seed0=1000):
    return np.stack([synth_frame(h, w, seed0 + k) for k in range(n)])


def shifted_frame(img, dx, dy, seed):
    """Frame B of a frame-to-frame pair: img translated by (dx, dy) with fresh +-2 noise."""
    rng = np.random.default_rng(seed)
    h, w = img.shape
    p = np.pad(img, ((abs(dy), abs(dy)), (abs(dx), abs(dx))), mode="reflect")
    b = p[abs(dy) - dy:abs(dy) - dy + h, abs(dx) - dx:abs(dx) - dx + w].astype(np.int32)
    b = b + rng.integers(-2, 3, (h, w))
    return np.clip(b, 0, 255).astype(np.uint8)


def synth_descriptors(n_db, n_q, seed_db=42, seed_q=43, flip_p=0.08, dup_frac=0.01):
    """(db[n_db,32], q[n_q,32]) uint8.  Even queries are a DB row with ~8 % flipped bits
    (d1 ~ 20), odd queries are random (d1 ~ 85-95); dup_frac of the DB rows are exact copies
    of an earlier row so that 'lowest index wins, d2 == d1' is exercised."""
    rng = np.random.default_rng(seed_db)
    db = rng.integers(0, 256, (n_db, 32), dtype=np.uint8)
    ndup = int(n_db * dup_frac)
    if ndup and n_db > 1:
        dst = rng.choice(np.arange(1, n_db), size=min(ndup, n_db - 1), replace=False)
        src = (dst * rng.random(len(dst))).astype(np.int64)
        db[dst] = db[src]
    rq = np.random.default_rng(seed_q)
    q = rq.integers(0, 256, (n_q, 32), dtype=np.uint8)
    if n_db:
        perm = rq.integers(0, n_db, (n_q + 1) // 2)
        flips = np.packbits((rq.random(((n_q + 1) // 2, 256)) < flip_p).astype(np.uint8), axis=1)
        q[0::2] = db[perm] ^ flips
    return db, q


def synth_vocabulary(k=10, L=6, seed=5, flip_bits=24, stop_frac=0.02, prune_frac=0.0, order="bfs"):
    """Synthetic vocabulary tree in the node order of DBoW2's text format (the reference's Data/ORBvoc.txt is not in the
    repository).  Children are their parent's descriptor with `flip_bits` random bits flipped, so the Hamming descent is
    meaningful; weights are idf-like positive doubles with a fraction `stop_frac` of zero-weight (stopped) words;
    `prune_frac` removes random subtrees (ragged tree: fewer than k children, leaves above level L);
    order="dfs" numbers nodes depth first (children of a node not consecutive).
    Returns (parent int32[n], desc uint8[n,32], weight float64[n]) with node 0 the root."""
    rng = np.random.default_rng(seed)
    parent = [0]
    bits = [np.zeros(256, np.uint8)]
    level = [0]
    root_children = rng.integers(0, 2, (k, 256), dtype=np.uint8)
    if order == "bfs":
        frontier = [0]
        for lv in range(1, L + 1):
            nxt = []
            for p in frontier:
                for c in range(k):
                    if lv > 1 and prune_frac > 0 and rng.random() < prune_frac:
                        continue
                    b = root_children[c].copy() if p == 0 else bits[p].copy()
                    if p != 0:
                        b[rng.choice(256, flip_bits, replace=False)] ^= 1
                    parent.append(p); bits.append(b); level.append(lv)
                    nxt.append(len(parent) - 1)
            frontier = nxt
    else:
        def grow(p, lv):
            if lv > L:
                return
            for c in range(k):
                if lv > 1 and prune_frac > 0 and rng.random() < prune_frac:
                    continue
                b = root_children[c].copy() if p == 0 else bits[p].copy()
                if p != 0:
                    b[rng.choice(256, flip_bits, replace=False)] ^= 1
                parent.append(p); bits.append(b); level.append(lv)
                grow(len(parent) - 1, lv + 1)
        grow(0, 1)
    n = len(parent)
    desc = np.packbits(np.stack(bits), axis=1, bitorder="little")
    weight = rng.uniform(0.5, 12.0, n)
    weight[rng.random(n) < stop_frac] = 0.0
    weight[0] = 0.0
    return np.asarray(parent, np.int32), desc, weight


def synth_vocabulary_fast(k=10, L=6, seed=5, flip_bits=24, stop_frac=0.02):
    """Full, balanced tree generated level by level with numpy (for k=10, L=6: 1 111 111 nodes), breadth-first order."""
    rng = np.random.default_rng(seed)
    parents = [np.zeros(1, np.int32)]
    descs = [np.zeros((1, 32), np.uint8)]
    prev_ids = np.zeros(1, np.int64)
    prev_desc = None
    next_id = 1
    for lv in range(1, L + 1):
        m = len(prev_ids) * k
        par = np.repeat(prev_ids, k)
        if lv == 1:
            d = rng.integers(0, 256, (m, 32), dtype=np.uint8)
        else:
            d = np.repeat(prev_desc, k, axis=0)
            # flip bits: xor with a sparse random mask (AND of j random bytes sets a bit with probability 2^-j)
            ands = max(1, int(round(np.log2(256.0 / flip_bits))))
            mask = rng.integers(0, 256, (m, 32), dtype=np.uint8)
            for _ in range(ands - 1):
                mask &= rng.integers(0, 256, (m, 32), dtype=np.uint8)
            d = d ^ mask
        parents.append(par.astype(np.int32)); descs.append(d)
        prev_ids = np.arange(next_id, next_id + m, dtype=np.int64)
        prev_desc = d
        next_id += m
    parent = np.concatenate(parents); desc = np.concatenate(descs)
    weight = rng.uniform(0.5, 12.0, len(parent))
    weight[rng.random(len(parent)) < stop_frac] = 0.0
    weight[0] = 0.0
    return parent, desc, weight


def write_vocabulary_text(path, k, L, parent, desc, weight, scoring=0, weighting=0, trailing_newline=True):
    """DBoW2 text format (TemplatedVocabulary::saveToTextFile / loadFromTextFile, TemplatedVocabulary.h:1338-1460).
    saveToTextFile ends the file with a newline; the reference's loader (`while(!f.eof())`, :1377) then reads one more, empty,
    line and hangs a phantom child with an uninitialised descriptor under the root.  trailing_newline=False writes the file the
    reference loads without that phantom (used when the reference's own DBoW2 is the checker, tests/test_ref_build.py)."""
    n = len(parent)
    has_child = np.zeros(n, bool)
    has_child[parent[1:]] = True
    with open(path, "w") as f:
        f.write("%d %d %d %d\n" % (k, L, scoring, weighting))
        for i in range(1, n):
            f.write("%d %d %s %s%s" % (parent[i], 0 if has_child[i] else 1, " ".join(str(int(b)) for b in desc[i]), repr(float(weight[i])),
                                       "\n" if (trailing_newline or i + 1 < n) else ""))


# ---- config 5: a keyframe-descriptor database that is the same at any GPU count ----
# Counter-based: row r is a pure function of (seed, r), so a rank generates exactly its row range on its own device and rank 0
# can regenerate everything for the single-scan check.  About 0.8 % of the rows in the upper half of the database are exact copies
# of the row half a database below them: the two copies ALWAYS lie in different shards (contiguous ranges, 2+ ranks), so "lowest
# global index wins, d2 == d1" is decided by the cross-shard merge, not inside one shard.  Even queries are planted next to such
# duplicated rows (Binomial(256, 0.08) flipped bits), odd queries are random.
_SM_GAMMA, _SM_M1, _SM_M2 = 0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB
_M64 = (1 << 64) - 1


def _splitmix_np(x):
    with np.errstate(over="ignore"):
        z = (x.astype(np.uint64) + np.uint64(_SM_GAMMA))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_SM_M1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_SM_M2)
        return z ^ (z >> np.uint64(31))


def db_is_dup_np(rows, ndb, seed):
    rows = np.asarray(rows, np.uint64)
    return (rows >= np.uint64(ndb // 2)) & ((_splitmix_np(rows ^ np.uint64((seed * 0x51ED27 + 0xABCD) & _M64)) & np.uint64(127)) == 0)


def db_rows_np(rows, ndb, seed):
    """descriptor rows (len(rows), 32) uint8 of the counter-based database, on the host"""
    rows = np.asarray(rows, np.uint64)
    eff = np.where(db_is_dup_np(rows, ndb, seed), rows - np.uint64(ndb // 2), rows)
    with np.errstate(over="ignore"):
        ctr = (eff[:, None] * np.uint64(4) + np.arange(4, dtype=np.uint64)[None, :]) + np.uint64((seed * 0x2545F491) & _M64)
    return _splitmix_np(ctr).astype("<u8").view(np.uint8).reshape(len(rows), 32)


def db_rows_torch(lo, hi, ndb, seed, device, chunk=1 << 20):
    """the same rows [lo, hi) generated on `device` (torch int64 arithmetic wraps like uint64; shifts are made logical)"""
    import torch

    def s64(c):
        c &= _M64
        return c - (1 << 64) if c >= (1 << 63) else c

    def lsr(z, k):
        return (z >> k) & ((1 << (64 - k)) - 1)

    def splitmix(x):
        z = x + s64(_SM_GAMMA)
        z = (z ^ lsr(z, 30)) * s64(_SM_M1)
        z = (z ^ lsr(z, 27)) * s64(_SM_M2)
        return z ^ lsr(z, 31)
    out = torch.empty((hi - lo, 32), dtype=torch.uint8, device=device)
    half = ndb // 2
    for a in range(lo, hi, chunk):
        b = min(a + chunk, hi)
        r = torch.arange(a, b, dtype=torch.int64, device=device)
        dup = (r >= half) & ((splitmix(r ^ s64(seed * 0x51ED27 + 0xABCD)) & 127) == 0)
        eff = torch.where(dup, r - half, r)
        ctr = eff[:, None] * 4 + torch.arange(4, dtype=torch.int64, device=device)[None, :] + s64(seed * 0x2545F491)
        out[a - lo:b - lo] = splitmix(ctr).view(torch.uint8).reshape(b - a, 32)
    return out


def db_queries(nq, ndb, seed, seed_q=43, flip_p=0.08):
    """queries for the counter-based database: even ones planted next to a DUPLICATED row (its two copies are half a database
    apart), odd ones random.  Returns (q [nq,32] uint8, planted_row int64 [nq], -1 for random queries)."""
    rq = np.random.default_rng(seed_q)
    q = rq.integers(0, 256, (nq, 32), dtype=np.uint8)
    planted = np.full(nq, -1, np.int64)
    if ndb >= 4:
        half = ndb // 2
        cand = np.arange(half, ndb, dtype=np.uint64)
        if len(cand) > (1 << 22):
            cand = cand[rq.integers(0, len(cand), 1 << 22)]
        dups = cand[db_is_dup_np(cand, ndb, seed)]
        ne = (nq + 1) // 2
        if len(dups):
            src = dups[rq.integers(0, len(dups), ne)].astype(np.int64) - half          # the lower copy: the index that must win
        else:
            src = rq.integers(0, ndb, ne)
        flips = np.packbits((rq.random((ne, 256)) < flip_p).astype(np.uint8), axis=1)
        q[0::2] = db_rows_np(src, ndb, seed) ^ flips
        planted[0::2] = src
    return q, planted
