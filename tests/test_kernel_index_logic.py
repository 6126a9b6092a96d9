"""Host-side restatements of index arithmetic inside the CUDA kernels, checked exhaustively where the GPU
tests can only sample it (gravity.cu; every function names the device code it follows):

  * cells_kernel, phase 2: dealing the internal cells of a warp's 32 chains out to the lanes
    (inclusive prefix sum by shuffles + the five-step owner search);
  * unit_kernel: the source / destination windows and byte counts of the two bulk (TMA) copies of a tile,
    which must be 16-byte aligned and a multiple of 16 bytes long, and must cover the tile plus its halo;
  * kids_kernel / climb_kernel: which thread takes which entry of the 16 lists of climb starts.
"""
import numpy as np
import pytest


# ---- cells_kernel phase 2 ---------------------------------------------------------------------------
def warp_inclusive_scan(mine):
    """for (d = 1; d < 32; d <<= 1) { v = shfl_up(incl, d); if (lane >= d) incl += v; }"""
    incl = list(mine)
    d = 1
    while d < 32:
        prev = list(incl)
        for lane in range(32):
            if lane >= d:
                incl[lane] = prev[lane] + prev[lane - d]
        d <<= 1
    return incl


def owner_of(incl, t):
    """o = 0; for (step = 16; step > 0; step >>= 1) if (incl[o + step - 1] <= t) o += step;"""
    o = 0
    step = 16
    while step > 0:
        if incl[o + step - 1] <= t:
            o += step
        step >>= 1
    return o


@pytest.mark.parametrize("seed", range(6))
def test_internal_cells_are_dealt_out_exactly_once(seed):
    rng = np.random.default_rng(seed)
    for trial in range(300):
        kind = trial % 4
        if kind == 0:
            mine = rng.integers(0, 4, 32)                      # the usual: 0..3 internal cells per chain
        elif kind == 1:
            mine = np.zeros(32, int)
            mine[rng.integers(0, 32)] = rng.integers(1, 70)    # one deep chain, nothing else
        elif kind == 2:
            mine = np.where(rng.random(32) < 0.1, rng.integers(1, 40, 32), 0)
        else:
            mine = np.zeros(32, int)                           # (lanes past n, merged bodies: no tasks at all)
            if trial % 8 == 3:
                mine[31] = 5
        incl = warp_inclusive_scan(mine)
        assert incl == list(np.cumsum(mine))
        n_tasks = incl[31]
        excl = [incl[i] - mine[i] for i in range(32)]
        seen = set()
        for t0 in range(0, n_tasks, 32):
            for lane in range(32):
                t = t0 + lane
                o = owner_of(incl, t)
                assert 0 <= o <= 31                            # (a shuffle source lane, also for idle lanes)
                if t >= n_tasks:
                    continue
                k = t - excl[o]
                assert 0 <= k < mine[o], (mine, t, o, k)
                assert (o, k) not in seen
                seen.add((o, k))
        assert len(seen) == n_tasks


# ---- unit_kernel bulk copies ------------------------------------------------------------------------
TILE = 256


def tile_copies(n, tile):
    """issue(tile, buf) of unit_kernel: returns (sp_src, sp_dst_slot, sp_count, key_src, key_dst_slot, key_count)"""
    s0 = tile * TILE
    p0 = s0 - 1 if s0 else 0
    p1 = min(n, s0 + TILE + 1)
    k0 = s0 - 2 if s0 else 0
    k1 = min((n + 1) & ~1, s0 + TILE + 2)
    return p0, p0 - (s0 - 1), p1 - p0, k0, k0 - (s0 - 2), k1 - k0


@pytest.mark.parametrize("n", [1, 2, 3, 255, 256, 257, 258, 511, 512, 513, 1000, 4097, 65535, 65536, 65537, 1000004])
def test_unit_tile_windows(n):
    n_tiles = (n + TILE - 1) // TILE
    key_alloc = n + n // 8 + 32            # DevBuf::ensure: bytes + bytes / 8 + 256 (in keys)
    for tile in range(n_tiles):
        s0 = tile * TILE
        p0, pslot, pc, k0, kslot, kc = tile_copies(n, tile)
        # bulk copies: 16-byte aligned source and destination, size a multiple of 16 bytes, never empty
        assert pc > 0 and kc > 0
        assert (p0 * 32) % 16 == 0 and (pslot * 32) % 16 == 0 and (pc * 32) % 16 == 0
        assert (k0 * 8) % 16 == 0 and (kslot * 8) % 16 == 0 and (kc * 8) % 16 == 0
        # inside the shared-memory arrays t_sp[TILE + 2], t_key[TILE + 4] and inside the allocations
        assert 0 <= pslot and pslot + pc <= TILE + 2
        assert 0 <= kslot and kslot + kc <= TILE + 4
        assert p0 + pc <= n and k0 + kc <= key_alloc
        # thread i reads t_sp[i], [i + 1], [i + 2] = bodies s - 1, s, s + 1 and t_key[i + 1 .. i + 3]
        for i in range(min(TILE, n - s0)):
            s = s0 + i
            for slot, body, needed in ((i, s - 1, s > 0), (i + 1, s, True), (i + 2, s + 1, s + 1 < n)):
                if needed:
                    assert pslot <= slot < pslot + pc and p0 + (slot - pslot) == body
            for slot, body, needed in ((i + 1, s - 1, s > 0), (i + 2, s, True), (i + 3, s + 1, s + 1 < n)):
                if needed:
                    assert kslot <= slot < kslot + kc and k0 + (slot - kslot) == body


# ---- lists of climb starts --------------------------------------------------------------------------
LISTS = 16


def test_every_climb_start_is_taken_once():
    rng = np.random.default_rng(3)
    threads = 148 * 4 * 128
    for counts in ([0] * 16, [1] * 16, list(rng.integers(0, 5000, 16)), [threads // 16 + 7] + [0] * 15,
                   list(rng.integers(0, 3 * threads // 16, 16))):
        taken = [np.zeros(c, int) for c in counts]
        for gt in range(threads):
            l = gt % LISTS
            it = gt // LISTS
            while it < counts[l]:
                taken[l][it] += 1
                it += threads // LISTS
        for l in range(LISTS):
            assert (taken[l] == 1).all()


# ---- keys without the compare-and-halve chain (gravity.cu key_of) -----------------------------------------
def _spread(q, dim):
    out = np.zeros(len(q), dtype=np.uint64)
    lm = 31 if dim == 2 else 21
    for b in range(lm):
        out |= ((q >> np.uint64(b)) & np.uint64(1)) << np.uint64(dim * b)
    return out


def _key_quantised(s, ext, dim):
    """Host restatement of key_of's fast path: (keys, which bodies may take it)."""
    lm = 31 if dim == 2 else 21
    scale = np.float64(2.0 ** (lm - 1)) / np.float64(ext)
    guard = 1e-13 * 2.0 ** (lm - 1)
    fast = np.ones(len(s), dtype=bool)
    key = np.zeros(len(s), dtype=np.uint64)
    for a, f in enumerate(("x", "y", "z")[:dim]):
        v = (s[f] + ext) * scale
        fl = np.floor(v)
        e = v - fl
        fast &= (v > 0.0) & (v < 2.0 ** lm) & (e >= guard) & (e <= 1.0 - guard)
        key |= _spread(np.where(fast, fl, 0.0).astype(np.int64).astype(np.uint64), dim) << np.uint64(a)
    return key, fast


@pytest.mark.parametrize("dim", [2, 3])
def test_quantised_keys_equal_the_chain_outside_the_guard_band(dim):
    """The oracle's keys come from the reference's chain (octree.rs:160-215 centre +- extent/2 per level); the
    quantised form must give the same 62 / 63 bits for every body it accepts, and must REFUSE the bodies that sit
    on or within a few ulps of a cell boundary."""
    from oracle import binding as ob
    from physim_b200 import generators as gen
    from tests.test_gpu_parity import boundary_lattice
    for s, must_refuse in ((gen.cube(200_000, seed=4), False), (gen.readme_pipeline(50_000, seed=2), False),
                           (boundary_lattice(1.0, 5), True), (boundary_lattice(0.7368421052631579, 6), True)):
        o = ob.CellTable(dim, s)
        want = np.zeros(len(s), dtype=np.uint64)
        want[o.perm] = o.key
        got, fast = _key_quantised(s, o.extent, dim)
        assert np.array_equal(got[fast], want[fast])
        if must_refuse:
            assert (~fast).sum() > 0.9 * len(s)      # x is always near a level-12 boundary
        else:
            assert fast.mean() > 0.99


# ---- sort_local_kernel: the padding-free bitonic network that sorts a crowded bin ---------------------------
def _bitonic_all_ascending(items):
    """Host restatement of the network in sort_local_kernel: every exchange puts the smaller element at the lower
    index; the first step of each merge pairs i with its mirror in the block; a partner index >= m (the virtual
    +infinity padding up to the next power of two) never moves anything."""
    a = list(items)
    m = len(a)
    P = 1
    while P < m:
        P <<= 1

    def exchange(i, j):
        if j >= m:
            return
        if a[i] > a[j]:
            a[i], a[j] = a[j], a[i]

    kk = 2
    while kk <= P:
        hk = kk >> 1
        for t in range(P >> 1):
            base, off = (t & ~(hk - 1)) << 1, t & (hk - 1)
            exchange(base + off, base + kk - 1 - off)
        j2 = kk >> 2
        while j2 > 0:
            for t in range(P >> 1):
                i = ((t & ~(j2 - 1)) << 1) | (t & (j2 - 1))
                exchange(i, i + j2)
            j2 >>= 1
        kk <<= 1
    return a


def test_padding_free_bitonic_network_sorts_every_length():
    rng = np.random.default_rng(11)
    for m in list(range(1, 70)) + [100, 127, 128, 129, 513, 1000, 1025]:
        keys = rng.integers(0, max(2, m // 3), m)          # many equal keys: ties go by the index
        items = [(int(keys[i]), int(i)) for i in rng.permutation(m)]
        assert _bitonic_all_ascending(items) == sorted(items), m


# ---- the walk's link word: skip | level << 32 in the fourth double of a centre record --------------------------
def test_link_word_and_half_width_from_level():
    """pack_link / link_skip / link_level, and half_width_at: extent / 2^level taken from the exponent field equals
    the value `half *= 0.5` reaches after `level` steps (what cells_kernel's replay holds), for every level a tree
    can have, as long as the result is a normal double."""
    rng = np.random.default_rng(3)
    for _ in range(2000):
        skip = int(rng.integers(0, 2**32))
        level = int(rng.integers(0, 256))
        word = np.array([(level << 32) | skip], dtype=np.uint64).view(np.float64)[0]   # pack_link
        bits = int(np.array([word]).view(np.uint64)[0])
        assert bits & 0xFFFFFFFF == skip and (bits >> 32) & 0xFF == level
    for ext in list(rng.uniform(1e-3, 1e3, 200)) + [1.0, 0.7368421052631579, 5e-300, 1e300]:
        ext = np.float64(ext)
        half = ext
        for level in range(0, 80):
            hi = int(np.array([ext]).view(np.uint64)[0] >> np.uint64(32))
            if ((hi >> 20) & 0x7FF) > level + 1:               # the exponent trick's own condition
                tricked = (np.array([ext]).view(np.uint64) - np.uint64(level << 52)).view(np.float64)[0]
                assert tricked == half, (ext, level)
            assert np.ldexp(ext, -level) == half or half < 2.3e-308   # the fall-back (exact while normal)
            half = half * np.float64(0.5)
