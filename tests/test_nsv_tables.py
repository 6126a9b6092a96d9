"""Host-side emulation of the minima tables and the `first body >= j0 whose shared-level byte is <= l`
query that cells_kernel uses for the end of a cell's run (gravity.cu: unit_kernel, nsv_level2_block,
first_le_16, nsv_in_block, nsv_descend, nsv_next_le), statement for statement, checked against brute
force for window / block / super-block edge sizes and for runs of every length class."""
import numpy as np

NONE = 255


def build(A):
    n = len(A)
    nb = (n + 255) // 256
    n_pad = nb * 256
    t1 = np.full(n_pad, NONE, np.uint8)
    t1[:n] = A
    t16 = t1.reshape(-1, 16).min(axis=1)                       # aligned windows of 16 bodies
    nblocks = nb
    b_pad = (nblocks + 255) // 256 * 256
    nsuper = b_pad // 256
    # level 0 of t2: block minima; entries past nblocks are NOT initialised on the device: poison them
    t2 = np.zeros((9, b_pad), np.uint8)
    t2[0, :nblocks] = t1.reshape(-1, 256).min(axis=1)
    t3 = np.full(nsuper, NONE, np.uint8)
    for sb in range(nsuper):                                   # nsv_level2_block
        tab = np.full(384, NONE, np.uint8)
        lo = sb * 256
        valid = max(0, min(256, nblocks - lo))
        tab[:valid] = t2[0, lo:lo + valid]
        for k in range(1, 9):
            new = np.full(384, NONE, np.uint8)
            new[:256] = np.minimum(tab[:256], tab[(1 << (k - 1)):(1 << (k - 1)) + 256])
            t2[k, lo:lo + 256] = new[:256]
            tab = new
        t3[sb] = tab[0]
    return dict(t1=t1, t16=t16, t2=t2, t3=t3, n=n, n_pad=n_pad, nblocks=nblocks, b_pad=b_pad, nsuper=nsuper)


def first_le_16(w, l, off):
    for i in range(off, 16):
        if w[i] <= l:
            return i
    return 16


def descend(tab, j, end, l):
    for k in range(8, -1, -1):
        step = 1 << k
        if j + step <= end and tab[k, j] > l:
            j += step
    return j


def in_block(tv, b, l):
    w = first_le_16(tv["t16"][b * 16:b * 16 + 16], l, 0)
    assert w < 16
    jw = b * 256 + w * 16
    return jw + first_le_16(tv["t1"][jw:jw + 16], l, 0)


def next_le(tv, j0, l, stats=None):
    n, n_pad = tv["n"], tv["n_pad"]
    if j0 >= n:
        return n
    base, blk = j0 & ~15, j0 >> 8
    two = base + 32 <= n_pad
    pos = first_le_16(tv["t1"][base:base + 16], l, j0 - base)
    if pos < 16:
        return base + pos
    if two:
        pos = first_le_16(tv["t1"][base + 16:base + 32], l, 0)
        if pos < 16:
            return base + 16 + pos
    seen = base + (32 if two else 16)
    if (seen >> 8) == blk:
        pos = first_le_16(tv["t16"][blk * 16:blk * 16 + 16], l, (seen >> 4) - blk * 16)
        if pos < 16:
            jw = blk * 256 + pos * 16
            return jw + first_le_16(tv["t1"][jw:jw + 16], l, 0)
    b0 = blk + 1
    if b0 >= tv["nblocks"]:
        return n
    bb = b0 & ~15
    two_b = bb + 32 <= tv["b_pad"]
    pos = first_le_16(tv["t2"][0, bb:bb + 16], l, b0 - bb)
    if pos == 16:
        pos = first_le_16(tv["t2"][0, bb + 16:bb + 32], l, 0) if two_b else 16
        pos = pos + 16 if pos < 16 else 32
    b = bb + pos
    if pos == 32:
        if stats is not None:
            stats["descents"] = stats.get("descents", 0) + 1
        b = bb + (32 if two_b else 16)
        if b >= tv["nblocks"]:
            return n
        sb_end = min(tv["nblocks"], (b | 255) + 1)
        b = descend(tv["t2"], b, sb_end, l)
        if b == sb_end:
            if sb_end == tv["nblocks"]:
                return n
            q = sb_end >> 8
            while q < tv["nsuper"] and tv["t3"][q] > l:
                q += 1
            if q >= tv["nsuper"]:
                return n
            b = q << 8
            sb_end = min(tv["nblocks"], b + 256)
            b = descend(tv["t2"], b, sb_end, l)
            if b == sb_end:
                return n
    if b >= tv["nblocks"]:
        return n
    return in_block(tv, b, l)


def test_nsv_query_matches_brute_force():
    rng = np.random.default_rng(0)
    stats = {}
    for n in [1, 2, 15, 16, 17, 31, 32, 33, 67, 255, 256, 257, 511, 512, 1000, 4095, 4096, 4097, 8191, 8192,
              8200, 65535, 65536, 65537, 70000, 131072 + 300, 200000]:
        # mostly large values with rare small ones so answers are far away
        A = rng.integers(5, 40, n).astype(np.uint8)
        A[rng.random(n) < 0.3] = NONE
        for pos in rng.integers(0, n, max(1, n // 2000)):
            A[pos] = rng.integers(0, 6)
        tv = build(A)
        for _ in range(400):
            j0 = int(rng.integers(0, n + 1))
            l = int(rng.integers(0, 42))
            idx = np.nonzero(A[j0:] <= l)[0]
            want = j0 + int(idx[0]) if len(idx) else n
            got = next_le(tv, j0, l, stats)
            assert got == want, (n, j0, l, got, want)
    assert stats["descents"] > 50          # the long-run path was exercised too


def test_nsv_query_every_run_length_and_alignment():
    """one small value at distance d from j0, for every d up to past two blocks and every alignment of j0 inside a
    window / block, then a sweep of long distances (block-minima stage, table descent, super-block scan)"""
    n = 300000
    for d in list(range(0, 600)) + [4000, 4095, 4096, 4200, 8191, 8192, 9000, 65535, 65536, 70000, 140000]:
        for j0 in (0, 1, 15, 16, 17, 239, 240, 241, 255, 256, 257, 4095, 65535 - 16, 65536):
            if j0 + d >= n:
                continue
            A = np.full(n, 30, np.uint8)
            A[j0 + d] = 3
            if j0 > 0:
                A[j0 - 1] = 0                  # a hit just before the start must not be seen
            tv = build(A)
            assert next_le(tv, j0, 5) == j0 + d, (j0, d)
            assert next_le(tv, j0, 2) == n, (j0, d)


def test_nsv_query_no_hit_and_tail():
    for n in (1, 16, 17, 256, 257, 4096, 65536, 65537):
        A = np.full(n, 40, np.uint8)
        tv = build(A)
        for j0 in {0, n // 2, n - 1, n}:
            assert next_le(tv, j0, 39) == n
            assert next_le(tv, j0, 40) == min(j0, n)
