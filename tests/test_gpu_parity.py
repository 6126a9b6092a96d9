"""GPU parity: the CUDA path, called through the plugin C ABI (`*_get_api` vtables), against the
CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star):
  * keys, sort permutation, cell table (level, head, count, skip, parent, geometric centres and
    half-widths): bit-exact vs oracle 2, which tests/test_oracle_table.py ties to the reference's
    pointer tree;
  * per-target interaction counts: exact vs oracle 1 (the walk takes the reference's decisions);
  * accelerations (fp32 on device vs fp64 reference): relative error median <= 1e-5, p99 <= 1e-3.
"""
import numpy as np
import pytest

from oracle import binding as ob
from physim_b200 import api
from physim_b200 import generators as gen
from physim_b200.entity import accelerations, entities
from tests.util import assert_acc_parity, rel_err, vec

pytestmark = pytest.mark.gpu

DIM = {"astro": 2, "astro2": 3}


def with_merges(seed):
    rng = np.random.default_rng(seed)
    e = gen.cube(500, seed=seed)
    dup = e[:40].copy()
    near = e[40:80].copy()
    near["x"] += 3e-10
    far = e[80:120].copy()
    far["y"] += 2e-5
    out = np.concatenate([e, dup, near, far])
    rng.shuffle(out)
    return out


def bucket_twins(seed):
    s = gen.cube(400, seed=seed)
    twin = s[:20].copy()
    twin["x"] += 2e-8
    return np.concatenate([s, twin])


def boundary_lattice(ext, seed):
    """Bodies ON cell boundaries of every level and a few ulps either side of them (the keys' guard band: these
    take the reference's compare-and-halve chain on the device, everything else the quantised form), distinct in x
    by >= ext / 4096 so that nothing merges; one body pins the extent to `ext`."""
    rng = np.random.default_rng(seed)
    n = 1500

    def near_lattice(k, m):
        v = ext * (k / 2.0 ** m)
        d = rng.integers(-2, 3, len(v))
        for step in (1, 2):
            v = np.where(d >= step, np.nextafter(v, np.inf), v)
            v = np.where(d <= -step, np.nextafter(v, -np.inf), v)
        return v

    s = entities(n + 1)
    kx = rng.choice(np.arange(-4095, 4096), n, replace=False)
    s["x"][:n] = near_lattice(kx, np.full(n, 12))
    for f in ("y", "z"):
        m = rng.integers(1, 22, n)
        k = np.array([rng.integers(-(2 ** int(mm)) + 1, 2 ** int(mm)) for mm in m])
        s[f][:n] = near_lattice(k, m)
    s["mass"] = rng.uniform(0.5, 2.0, n + 1)
    s["x"][n], s["y"][n], s["z"][n] = -ext, ext, ext / 3
    return s


TREE_CASES = {
    "lattice_pow2": lambda: boundary_lattice(1.0, 5),
    "lattice_ugly": lambda: boundary_lattice(0.7368421052631579, 6),
    "cube20k": lambda: gen.readme_pipeline(20_000, seed=3),
    "merges": lambda: with_merges(7),
    "bucket": lambda: bucket_twins(9),
    "single": lambda: gen.star(x=0.3, y=-0.2, z=0.1, mass=2.0),
    "pair": lambda: np.concatenate([gen.star(x=1.0, mass=1.0), gen.star(x=-1.0, mass=3.0)]),
    "coincident10": lambda: np.concatenate([gen.star(mass=1.0)] * 10),
    "solar": lambda: gen.solar(),
    "ragged4097": lambda: gen.cube(4097, seed=21),
}


@pytest.mark.parametrize("name", ["astro2", "astro"])
@pytest.mark.parametrize("case", sorted(TREE_CASES))
def test_tree_is_bit_exact(name, case):
    s = TREE_CASES[case]()
    if name == "astro" and case == "solar":
        s = s.copy()
        s["x"] += 1e-3 * np.arange(len(s))  # moons share (x, y) structure; keep the quadtree finite
    el = api.TransformElement(name, theta=1.0, e=0.5)
    el.transform(s)
    t = el.debug_tree()
    o = ob.CellTable(DIM[name], s)
    assert t["extent"] == o.extent
    assert np.array_equal(t["key"], o.key)
    assert np.array_equal(t["perm"], o.perm)
    assert np.array_equal(t["cell_start"], o.cell_start)
    assert len(t["level"]) == o.n_cells
    for k in ("level", "head", "count", "skip", "parent"):
        assert np.array_equal(t[k], getattr(o, k)), k
    assert np.array_equal(t["centre_ext"], o.centre_ext)          # identical doubles
    leaf = o.skip == np.arange(o.n_cells) + 1
    assert np.array_equal(t["com_mass"][leaf], o.com_mass[leaf])  # leaves: the body itself
    np.testing.assert_allclose(t["com_mass"][:, 3], o.com_mass[:, 3], rtol=1e-13)
    np.testing.assert_allclose(t["com_mass"][:, :3], o.com_mass[:, :3], rtol=0, atol=1e-12 * max(o.extent, 1e-300))


@pytest.mark.parametrize("name", ["astro2", "astro"])
@pytest.mark.parametrize("theta", [0.5, 0.7, 1.0, 1.3, 1.5])
def test_barnes_hut_matches_reference(name, theta):
    s = gen.readme_pipeline(20_000, seed=5)
    s["fixed"][17] = True
    el = api.TransformElement(name, theta=theta, e=0.5)
    acc = el.transform(s)
    ref, cnt = ob.transform(name, s, theta, 0.5, counts=True)
    t = el.debug_tree()
    assert np.array_equal(t["counts"], cnt)               # the reference's interaction lists
    assert el.stats()["interactions"] == int(cnt.sum())
    assert acc["x"][17] == 0.0 and acc["y"][17] == 0.0    # fixed body untouched
    free = ~s["fixed"]
    assert_acc_parity(acc, ref, free)


def test_readme_config_c1():
    """BASELINE config 1: cube n=100000 seed=1 spin=1000 + 2 stars ! astro2 theta=1.5 e=0.5."""
    s = gen.readme_pipeline(100_000, seed=1, spin=1000.0)
    el = api.TransformElement("astro2", theta=1.5, e=0.5)
    acc = el.transform(s)
    ref, cnt = ob.transform("astro2", s, 1.5, 0.5, counts=True)
    assert np.array_equal(el.debug_tree()["counts"], cnt)
    r = assert_acc_parity(acc, ref)
    assert r.max() < 1e-2


@pytest.mark.parametrize("case", ["solar", "cube4096", "clumpy"])
def test_direct_sum_matches_reference(case):
    if case == "solar":
        s, e = gen.solar(), 0.1                     # BASELINE config 2 (example_pipelines/solar.toml)
    elif case == "cube4096":
        s, e = gen.readme_pipeline(4094, seed=2), 0.5
    else:
        s, e = gen.cube(3000, seed=8, size=50.0), 0.0
    el = api.TransformElement("simple_astro", e=e)
    acc = el.transform(s)
    ref = ob.transform("simple_astro", s, e=e)
    free = ~s["fixed"]
    assert_acc_parity(acc, ref, free)
    assert not vec(acc)[s["fixed"]].any()
    assert el.stats()["interactions"] == int(free.sum()) * len(s)


@pytest.mark.parametrize("packed", [False, True])
def test_direct_sum_far_from_the_origin(packed, monkeypatch):
    """A system offset from the origin (`cube centre=[1000,-2000,500]`): positions are differenced in fp64
    (small systems) or taken relative to the bounding-box centre in fp64 before the fp32 copy (packed kernel),
    so the offset costs no accuracy.  (Rounding the raw coordinates to fp32 first loses ~|x| 2^-24 = 1e-4 of
    every difference here: p99 error ~ 1e-2.)"""
    monkeypatch.setenv("PB200_DIRECT_SMALL", "0" if packed else "1")
    s = gen.cube(3000, seed=5, centre=(1000.0, -2000.0, 500.0))
    acc = api.TransformElement("simple_astro", e=0.5).transform(s)
    ref = ob.transform("simple_astro", s, e=0.5)
    r = assert_acc_parity(acc, ref)
    assert r.max() < 1e-4
    far = gen.cube(3000, seed=5)                      # the same system at the origin: same accuracy
    r0 = rel_err(api.TransformElement("simple_astro", e=0.5).transform(far), ob.transform("simple_astro", far, e=0.5))
    assert np.median(r) < 4 * np.median(r0) + 1e-7


def test_direct_sum_far_from_the_origin_packed_kernel_at_size():
    """The packed FP32 kernel at a size it is selected for by itself (n > 32768), offset system, sampled."""
    s = gen.cube(40_000, seed=6, centre=(-3000.0, 100.0, 7000.0))
    acc = api.TransformElement("astro2", theta=0.0, e=0.5).transform(s)
    pick = np.arange(0, len(s), 157)
    ref = accelerations(len(s))
    for i in pick:
        ob.direct_range(s, 0.5, int(i), int(i) + 1, acc=ref)
    r = assert_acc_parity(acc[pick], ref[pick])
    assert r.max() < 1e-4


def test_direct_sum_close_pairs_keep_their_direction():
    """Pairs separated by 1e-7 of the system size (below fp32 resolution of the coordinates): the small-system
    kernel differences in fp64, so each twin's acceleration - dominated by its partner - stays accurate."""
    s = gen.cube(2000, seed=9)
    rng = np.random.default_rng(1)
    d = rng.normal(size=(40, 3))
    d *= 1e-7 / np.linalg.norm(d, axis=1)[:, None]
    for k, axis in enumerate("xyz"):
        s[axis][40:80] = s[axis][:40] + d[:, k]
    s["mass"][:80] = 50.0                               # partner's pull ~ m / e dominates the rest (~ 1)
    acc = api.TransformElement("simple_astro", e=1e-3).transform(s)
    ref = ob.transform("simple_astro", s, e=1e-3)
    r = rel_err(acc, ref)
    assert r[:80].max() < 5e-6, r[:80].max()      # (a lost direction would be an error of order 1)
    assert_acc_parity(acc, ref)


@pytest.mark.parametrize("name", ["astro2", "astro"])
def test_nonpositive_theta_is_direct_sum(name):
    s = gen.readme_pipeline(3000, seed=12)
    ref = ob.transform(name, s, 0.0, 0.5)
    for theta in (0.0, -0.1):
        acc = api.TransformElement(name, theta=theta, e=0.5).transform(s)
        assert_acc_parity(acc, ref)


def test_direct_sum_sampled_at_c1_size():
    """N = 100 002 all-pairs on the GPU, checked on a 256-target sample of the O(N²) oracle."""
    s = gen.readme_pipeline(100_000, seed=1)
    acc = api.TransformElement("simple_astro", e=0.5).transform(s)
    pick = np.random.default_rng(0).choice(len(s), 256, replace=False)
    pick.sort()
    ref = accelerations(len(s))
    for i in pick:
        ob.direct_range(s, 0.5, int(i), int(i) + 1, acc=ref)
    assert_acc_parity(acc[pick], ref[pick])


def test_accumulates_and_respects_fixed_and_massless():
    s = gen.cube(2000, seed=4)
    s["fixed"][::7] = True
    s["mass"][5] = 0.0                 # massless free target: f/m = 0/0 = NaN in the reference
    base = accelerations(len(s))
    base["x"], base["y"], base["z"] = 1.0, -2.0, 3.0
    for name, kw in (("astro2", dict(theta=0.7, e=0.5)), ("simple_astro", dict(e=0.5)), ("astro", dict(theta=1.0))):
        acc = api.TransformElement(name, **kw).transform(s, base.copy())
        ref = ob.transform(name, s, kw.get("theta", 1.0), kw.get("e", 1.0), acc=base.copy())
        fx = s["fixed"]
        assert np.array_equal(vec(acc)[fx], vec(base)[fx])          # untouched
        assert np.isnan(acc["x"][5]) and np.isnan(ref["x"][5])
        ok = ~fx
        ok[5] = False
        shifted = accelerations(len(s))
        sref = accelerations(len(s))
        for k in "xyz":
            shifted[k] = acc[k] - base[k]
            sref[k] = ref[k] - base[k]
        assert_acc_parity(shifted, sref, ok, median=1e-4, p99=1e-2)  # looser: cancellation with base


def test_transform_twice_same_object_and_changing_n():
    el = api.TransformElement("astro2", theta=1.0, e=0.5)
    for n in (5000, 1200, 9000):
        s = gen.cube(n, seed=n)
        a1 = el.transform(s)
        a2 = el.transform(s)
        assert np.array_equal(vec(a1), vec(a2))                    # deterministic
        assert_acc_parity(a1, ob.transform("astro2", s, 1.0, 0.5))


@pytest.mark.parametrize("name", ["astro2", "astro"])
def test_validate_and_retry_paths(name):
    """The sort drops key bits below the tree depth seen last time and the cell table is sized from
    the last total; both guesses are validated after every build and the build re-runs when they
    were wrong.  Force both wrong guesses and require the bit-exact tree anyway."""
    s = bucket_twins(9) if name == "astro2" else with_merges(7)
    o = ob.CellTable(DIM[name], s)
    for sort_lo, cells in ((DIM[name] * (31 if name == "astro" else 21) - 2 * DIM[name], 0), (0, 1), (40, 1)):
        el = api.TransformElement(name, theta=1.0, e=0.5)
        el.debug_hint(sort_lo, cells)
        acc = el.transform(s)
        t = el.debug_tree()
        for k in ("key", "perm", "cell_start", "level", "head", "count", "skip", "parent"):
            assert np.array_equal(t[k], getattr(o, k)), (k, sort_lo, cells)
        assert_acc_parity(acc, ob.transform(name, s, 1.0, 0.5))
    # shallow data first, deep data second on the same object (the depth estimate is stale)
    el = api.TransformElement(name, theta=1.0, e=0.5)
    el.transform(gen.cube(3000, seed=1))
    el.transform(s)
    t = el.debug_tree()
    assert np.array_equal(t["perm"], o.perm) and np.array_equal(t["skip"], o.skip)


TREE_KEYS = ("key", "perm", "cell_start", "level", "head", "count", "skip", "parent")


@pytest.mark.parametrize("name", ["astro2", "astro"])
def test_bucket_local_sort_matches_global_sort(name):
    """From the second evaluation on, a body set whose top-8-bit key bins fit a shared-memory tile is
    sorted by the bucket sort (Pb200Stats.sort_mode 1..3: keys scattered to 256 buckets by the
    encoder, each bucket sorted in shared memory).  Same permutation, hence the same bit-exact tree,
    as the stable global LSD sort and the oracle."""
    s = np.concatenate([gen.cube(60_000, seed=3), with_merges(5)])
    o = ob.CellTable(DIM[name], s)
    ref = ob.transform(name, s, 1.0, 0.5)
    el = api.TransformElement(name, theta=1.0, e=0.5)
    el.transform(s)
    assert el.stats()["sort_mode"] == 0                      # nothing known about the buckets yet
    for forced in (None, 3, 2, 1):
        if forced is not None:
            el.debug_sort_mode(forced)
        acc = el.transform(s)
        st = el.stats()
        assert st["sort_mode"] == (forced or 1), st
        assert 0 < st["max_bucket"] <= 4608
        t = el.debug_tree()
        for k in TREE_KEYS:
            assert np.array_equal(t[k], getattr(o, k)), (k, forced)
        assert np.array_equal(t["centre_ext"], o.centre_ext)
        assert_acc_parity(acc, ref)


@pytest.mark.parametrize("name", ["astro2", "astro"])
def test_bucket_sort_overflow_and_skew_fall_back(name):
    """The buckets are cut at the previous evaluation's key quantiles.  A different body set on the
    same object makes them stale: one bucket overflows its tile, the kernel flags the build, the
    device skips it and the host re-runs with global passes.  Likewise when more bodies than a bin
    may hold agree on every sorted key bit."""
    el = api.TransformElement(name, theta=1.0, e=0.05)
    el.transform(gen.cube(30_000, seed=2))                   # splitters of a uniform cube
    s = plummer_like(30_000, 17)                             # concentrated far off-centre
    o = ob.CellTable(DIM[name], s)
    for forced in (1, 2):
        el.debug_sort_mode(forced)
        acc = el.transform(s)
        st = el.stats()
        t = el.debug_tree()
        for k in TREE_KEYS:
            assert np.array_equal(t[k], getattr(o, k)), (k, forced)
        assert_acc_parity(acc, ob.transform(name, s, 1.0, 0.05))
        if forced == 1:
            assert st["sort_mode"] == 0, st                                 # it did fall back
        el.transform(gen.cube(30_000, seed=2))
    # Crowds around one spot.  100 coincident bodies (one merged unit) and 1500 within 1e-5 are sorted by the
    # bucket sort: the bins follow the key range a bucket's bodies actually span, so a crowd spreads over them.
    # 3000 bodies within 1e-7 share one or two 63-bit octree keys (the key resolves extent * 2^-21 = 5e-7): one bin
    # of thousands, which the bucket's CTA sorts as a whole (bitonic network) instead of letting every member scan
    # it - the build stays on the bucket sort.  (The 62-bit quadtree key resolves 5e-10: no crowding.)
    for crowd, spread, want_mode in ((100, 0.0, 1), (1500, 1e-5, 1), (3000, 1e-7, 1), (6000, 1e-7, 1),
                                    # 6000 bodies on ONE octree key: a bucket of > 4608, the next capacity class
                                    (6000, 2e-8, 2 if name == "astro2" else 1)):
        twins = gen.cube(20_000, seed=8)
        jit = np.random.default_rng(crowd).random((crowd, 3)) * spread
        twins["x"][:crowd], twins["y"][:crowd], twins["z"][:crowd] = 0.3 + jit[:, 0], -0.2 + jit[:, 1], 0.6 + jit[:, 2]
        o = ob.CellTable(DIM[name], twins)
        el = api.TransformElement(name, theta=1.0, e=0.5)
        el.transform(twins)
        acc = el.transform(twins)                            # bucket sort attempted (abandoned for the 1500)
        t = el.debug_tree()
        for k in TREE_KEYS:
            assert np.array_equal(t[k], getattr(o, k)), (k, crowd, spread)
        assert el.stats()["sort_mode"] == want_mode, (crowd, spread)
        assert_acc_parity(acc, ob.transform(name, twins, 1.0, 0.5))


def plummer_like(n, seed):
    """Centrally concentrated 3-D cluster (deep, unbalanced tree), off-centre and large-scale."""
    rng = np.random.default_rng(seed)
    r = 1.0 / np.sqrt(rng.random(n) ** (-2.0 / 3.0) - 1.0 + 1e-12)
    r = np.minimum(r, 50.0)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    e = entities(n)
    p = d * r[:, None] * 20.0 + np.array([300.0, -150.0, 40.0])
    e["x"], e["y"], e["z"] = p.T
    e["mass"] = (rng.random(n) + 0.5) / n
    return e


@pytest.mark.parametrize("name", ["astro2", "astro"])
@pytest.mark.parametrize("theta", [0.5, 1.0])
def test_clustered_offcentre_distribution(name, theta):
    """Deep unbalanced tree, coordinates ~300 with structure down to ~1e-2: bit-exact tree, exact
    interaction lists, accelerations in tolerance."""
    s = plummer_like(30_000, 13)
    el = api.TransformElement(name, theta=theta, e=0.05)
    acc = el.transform(s)
    t = el.debug_tree()
    o = ob.CellTable(DIM[name], s)
    for k in ("key", "perm", "cell_start", "level", "head", "count", "skip", "parent"):
        assert np.array_equal(t[k], getattr(o, k)), k
    assert np.array_equal(t["centre_ext"], o.centre_ext)
    ref, cnt = ob.transform(name, s, theta, 0.05, counts=True)
    assert np.array_equal(t["counts"], cnt)
    assert_acc_parity(acc, ref)


def test_momentum_conservation_direct_sum():
    """Newton's third law as a size-independent property: sum_i m_i a_i ~ 0 for all-free bodies."""
    s = gen.cube(20_000, seed=6)
    acc = api.TransformElement("simple_astro", e=0.5).transform(s)
    p = (s["mass"][:, None] * vec(acc)).sum(axis=0)
    scale = (s["mass"][:, None] * np.abs(vec(acc))).sum()
    assert np.abs(p).max() / scale < 1e-5


def test_headline_size_properties():
    """BASELINE config 3 size (1 000 004 bodies, astro theta=1.3): properties that need no oracle
    run at this size, plus the oracle itself (a few seconds)."""
    s = gen.headline_pipeline(1_000_000, seed=1)
    el = api.TransformElement("astro", theta=1.3)
    acc = el.transform(s)
    t = el.debug_tree()
    n = len(s)
    assert (np.diff(t["key"].astype(np.uint64).view(np.int64)) >= 0).all()      # sorted
    assert np.array_equal(np.sort(t["perm"]), np.arange(n, dtype=np.uint32))    # a permutation
    leaf = t["skip"] == np.arange(len(t["skip"])) + 1
    assert t["count"][leaf].sum() == n and t["count"][0] == n
    np.testing.assert_allclose(t["com_mass"][0, 3], s["mass"].sum(), rtol=1e-12)
    com = (s["mass"][:, None] * np.stack([s["x"], s["y"], s["z"]], 1)).sum(0) / s["mass"].sum()
    np.testing.assert_allclose(t["com_mass"][0, :3], com, atol=1e-12)
    ref, cnt = ob.transform("astro", s, 1.3, 1.0, counts=True)
    assert np.array_equal(t["counts"], cnt)
    assert_acc_parity(acc, ref)
    # second evaluation: truncated key + bucket-local sort; same tree bit for bit
    acc2 = el.transform(s)
    assert el.stats()["sort_mode"] == 1, el.stats()
    t2 = el.debug_tree()
    for k in TREE_KEYS + ("counts",):
        assert np.array_equal(t[k], t2[k]), k
    assert np.array_equal(t["com_mass"], t2["com_mass"])
    assert np.array_equal(vec(acc), vec(acc2))


@pytest.mark.parametrize("name", ["astro2", "astro", "simple_astro"])
def test_tiny_and_degenerate_states(name):
    """n = 1, 2, 3; all bodies fixed; one massive + massless mix — the shapes discovery-time and toy
    pipelines produce (example_pipelines/shm.toml has a single star)."""
    kw = dict(theta=1.0, e=0.5) if name != "simple_astro" else dict(e=0.5)
    el = api.TransformElement(name, **kw)
    one = gen.star(x=0.3, y=0.1, z=0.2, mass=2.0)
    acc = el.transform(one)
    assert acc["x"][0] == 0.0 and acc["y"][0] == 0.0 and acc["z"][0] == 0.0   # nothing to attract it
    pair = np.concatenate([gen.star(x=1.0, y=0.5, mass=1.0), gen.star(x=-1.0, y=-0.25, z=0.1, mass=3.0)])
    ref = ob.transform(name, pair, kw.get("theta", 1.0), 0.5)
    assert_acc_parity(el.transform(pair), ref)
    three = np.concatenate([pair, gen.star(x=0.2, y=-0.9, z=0.5, mass=0.5, fixed=True)])
    ref = ob.transform(name, three, kw.get("theta", 1.0), 0.5)
    acc = el.transform(three)
    assert_acc_parity(acc[:2], ref[:2])
    assert acc["x"][2] == 0.0                                                  # fixed: untouched
    frozen = three.copy()
    frozen["fixed"][:] = True
    assert not vec(el.transform(frozen)).any()
    # the fused step and the resident loop on the same tiny states
    v = api.Verlet()
    out = v.integrate_fused(pair, el, 0.01)
    want = ob.Verlet().integrate(pair, lambda st, ac: ob.transform(name, st, kw.get("theta", 1.0), 0.5, acc=ac), 0.01)
    for k in ("x", "y", "z", "vx", "vy", "vz"):
        np.testing.assert_allclose(out[k], want[k], rtol=1e-6, atol=1e-12)
    sim = api.Sim(name, dt=0.01, **kw)
    sim.upload(one)
    sim.run(3)
    got = sim.download(one.copy())
    assert got["x"][0] == one["x"][0] and got["vx"][0] == 0.0
