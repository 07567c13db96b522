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
