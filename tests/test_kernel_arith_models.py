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
