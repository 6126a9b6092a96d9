"""numpy restatement of the HOST-VISIBLE logic of the sharded Barnes-Hut step (csrc/gravity.cu, "Sharded
Barnes-Hut"): which records a rank publishes for its level-K prefixes, how every rank rebuilds the cells above
level K from all ranks' records, and how the next cuts follow from the level-K histogram.  Test infrastructure:
the CUDA kernels (top_export_kernel / top_build_kernel) are the product; this model lets the exchange and the
rebuild be checked against the oracle's cell table on CPU ranks (tests/test_shard_model_gloo.py)."""
import numpy as np

K = {2: 6, 3: 4}          # TopTree<DIM>::K
LM = {2: 31, 3: 21}       # key levels


def offset(dim, level):
    return ((1 << (dim * level)) - 1) // ((1 << dim) - 1)


def first_cuts(dim, sorted_keys, world):
    """shard_plan_kernel: cut r = key of sorted body r n / world, rounded down to a level-K prefix."""
    shift = np.uint64(dim * (LM[dim] - K[dim]))
    n = len(sorted_keys)
    cuts = np.zeros(world + 1, dtype=np.uint64)
    for r in range(1, world):
        cuts[r] = (sorted_keys[(n * r) // world] >> shift) << shift
    cuts[world] = np.uint64(0xFFFFFFFFFFFFFFFF)
    return cuts


def export_slots(dim, tab, lo, hi):
    """top_export_kernel for the rank whose key range is [lo, hi): one record per level-K prefix it owns,
    taken from the cell table `tab` (oracle 2): the level-K cell, or the single-unit leaf above level K.
    Record = (count, units, cell index, X, Y, Z, M); zeros where the rank has no body."""
    k, shift = K[dim], np.uint64(dim * (LM[dim] - K[dim]))
    slots = 1 << (dim * k)
    rec = np.zeros((slots, 7))
    prefix = (tab.key[tab.head] >> shift).astype(np.int64)
    mine = (tab.key[tab.head] >= lo) & (tab.key[tab.head] < hi)
    leaf = tab.skip == np.arange(tab.n_cells) + 1
    for c in np.nonzero(mine & ((tab.level == k) | (leaf & (tab.level < k))))[0]:
        q = prefix[c]
        # a leaf above level K may hold (merged) bodies of several prefixes: it is recorded under its head's
        rec[q] = (tab.count[c], 1 if leaf[c] else 2, c, *tab.com_mass[c])
    return rec


def rebuild_top(dim, records, extent):
    """top_build_kernel: records [world, slots, 7] -> dense arrays over levels 0..K: units (0 / 1 / 2), count,
    com [.., 4]; cells above level K from their children in ascending digit order."""
    k, r = K[dim], 1 << dim
    cells = offset(dim, k + 1)
    units = np.zeros(cells, dtype=np.int64)
    count = np.zeros(cells, dtype=np.int64)
    com = np.zeros((cells, 4))
    merged = np.zeros_like(records[0])
    for rec in records:                         # at most one rank holds a prefix
        take = (rec[:, 0] != 0) & (merged[:, 0] == 0)
        merged[take] = rec[take]
    o = offset(dim, k)
    count[o:] = merged[:, 0]
    units[o:] = merged[:, 1]
    com[o:] = merged[:, 3:7]
    for l in range(k - 1, -1, -1):
        for p in range(1 << (dim * l)):
            kids = offset(dim, l + 1) + p * r + np.arange(r)
            kids = kids[units[kids] != 0]
            t = offset(dim, l) + p
            u = int(units[kids].sum())
            count[t] = count[kids].sum()
            if u == 1:
                units[t], com[t] = 1, com[kids[0]]
            elif u > 1:
                m = sx = sy = sz = 0.0
                for c in kids:                  # ComSum (fma on the device; plain products here: ~1 ulp apart)
                    m += com[c, 3]
                    sx += com[c, 3] * com[c, 0]
                    sy += com[c, 3] * com[c, 1]
                    sz += com[c, 3] * com[c, 2]
                units[t] = 2
                com[t] = (sx / m, sy / m, sz / m, m) if m != 0.0 else (0.0, 0.0, 0.0, 0.0)
    return units, count, com


def next_cuts(dim, count, world, n_total):
    """The cuts for the next build: rank r starts at the first level-K prefix whose running count reaches r n / world."""
    k = K[dim]
    shift = np.uint64(dim * (LM[dim] - k))
    per = count[offset(dim, k):]
    cum = np.concatenate([[0], np.cumsum(per)])        # bodies in prefixes < q
    cuts = np.zeros(world + 1, dtype=np.uint64)
    for r in range(1, world):
        want = (n_total * r) // world
        q = int(np.searchsorted(cum, want, side="left"))
        cuts[r] = np.uint64(min(q, len(per))) << shift
    cuts[world] = np.uint64(0xFFFFFFFFFFFFFFFF)
    return cuts
