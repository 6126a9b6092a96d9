"""The reference's own tree tests, restated against oracle 1 (pointer tree).

astro/src/octree.rs:218-519 (15 tests) and astro/src/quadtree.rs:197-492 (15 tests).  They assert
leaf counts and inequalities only, so they do not need the reference's ChaCha8 streams: where the
reference draws random bodies we draw from numpy with the same distribution.
"""
import numpy as np
import pytest

from oracle import binding as ob
from physim_b200.entity import entities

DIMS = [pytest.param(3, id="octree"), pytest.param(2, id="quadtree")]


def ent(x, y, z, mass=1.0):
    """Entity::new (physim-core/src/lib.rs:91-100)."""
    e = entities(1)
    e["x"], e["y"], e["z"], e["mass"] = x, y, z, mass
    e["radius"] = mass ** 0.33333
    return e


def random_entities(rng, n):
    """Entity::random (physim-core/src/lib.rs:115-128)."""
    e = entities(n)
    e["x"] = rng.uniform(-1.0, 1.0, n)
    e["y"] = rng.uniform(-1.0, 1.0, n)
    e["z"] = rng.uniform(0.0, 1.0, n)
    e["mass"], e["radius"] = 0.005, 0.02
    return e


@pytest.mark.parametrize("dim", DIMS)
def test_empty_tree(dim):
    t = ob.Tree(dim, extent=1.0)
    assert t.get_leaves_with_resolution([0, 0, 0], 0.5) == 0


@pytest.mark.parametrize("dim", DIMS)
def test_single_entity(dim):
    t = ob.Tree(dim, extent=1.0)
    t.push(ent(0, 0, 0))
    assert t.get_leaves_with_resolution([0, 0, 0], 0.5) == 1


@pytest.mark.parametrize("dim", DIMS)
def test_entities_at_origin(dim):
    t = ob.Tree(dim, extent=2.0)
    for _ in range(10):
        t.push(ent(0, 0, 0))
    out = t.get_leaves_with_resolution([0, 0, 0], 0.5, want=True)
    assert len(out) == 1
    assert out["mass"][0] == 10.0  # merged leaf sums the masses (octree.rs:69-80)


@pytest.mark.parametrize("dim", DIMS)
def test_entities_in_different_octants(dim):
    t = ob.Tree(dim, extent=2.0)
    if dim == 3:
        pos = [(sx * 0.5, sy * 0.5, sz * 0.5) for sz in (1, -1) for sy in (1, -1) for sx in (1, -1)]
    else:  # quadtree.rs: four quadrants, z = 0
        pos = [(0.5, 0.5, 0), (-0.5, 0.5, 0), (0.5, -0.5, 0), (-0.5, -0.5, 0)]
    for p in pos:
        t.push(ent(*p))
    assert t.get_leaves_with_resolution([0, 0, 0], 0.5) == len(pos)


@pytest.mark.parametrize("dim", DIMS)
def test_bh_factor_filtering(dim):
    t = ob.Tree(dim, extent=10.0)
    for x in (0.1, 5.0, 8.0):
        t.push(ent(x, 0, 0))
    assert t.get_leaves_with_resolution([0, 0, 0], -0.1) == 3
    assert t.get_leaves_with_resolution([0, 0, 0], 1.0) > 0


@pytest.mark.parametrize("dim", DIMS)
def test_boundary_conditions(dim):
    ext = 1.0
    t = ob.Tree(dim, extent=ext)
    pts = [(ext, 0, 0), (-ext, 0, 0), (0, ext, 0), (0, -ext, 0)]
    if dim == 3:
        pts += [(0, 0, ext), (0, 0, -ext)]
    for p in pts:
        t.push(ent(*p))
    assert t.get_leaves_with_resolution([0, 0, 0], -0.1) == len(pts)


@pytest.mark.parametrize("dim", DIMS)
def test_query_from_different_locations(dim):
    t = ob.Tree(dim, extent=10.0)
    for i in range(20):
        t.push(ent(i * 0.1, 0, 0))
    assert t.get_leaves_with_resolution([0, 0, 0], 0.5) > 0
    assert t.get_leaves_with_resolution([100.0, 0, 0], 0.5) > 0


@pytest.mark.parametrize("dim", DIMS)
def test_dense_cluster(dim):
    rng = np.random.default_rng(42)
    t = ob.Tree(dim, extent=1.0)
    e = entities(1000)
    e["x"], e["y"], e["z"] = ((rng.random((3, 1000)) - 0.5) * 0.2)
    e["mass"] = 1.0
    t.push(e)
    assert t.get_leaves_with_resolution([0, 0, 0], -0.1) == 1000


@pytest.mark.parametrize("dim", DIMS)
def test_sparse_distribution(dim):
    rng = np.random.default_rng(123)
    t = ob.Tree(dim, extent=100.0)
    e = entities(100)
    e["x"], e["y"], e["z"] = ((rng.random((3, 100)) - 0.5) * 200.0)
    e["mass"] = 1.0
    t.push(e)
    assert t.get_leaves_with_resolution([0, 0, 0], -0.1) == 100


@pytest.mark.parametrize("dim", DIMS)
def test_resolution_threshold(dim):
    t = ob.Tree(dim, extent=10.0)
    for x in (1.0, 5.0, 9.0):
        t.push(ent(x, 0, 0))
    loose = t.get_leaves_with_resolution([0, 0, 0], 0.1)
    medium = t.get_leaves_with_resolution([0, 0, 0], 0.5)
    strict = t.get_leaves_with_resolution([0, 0, 0], 2.0)
    assert loose >= medium >= strict


@pytest.mark.parametrize("dim", DIMS)
def test_lots(dim):
    n = 100_000
    t = ob.Tree(dim, extent=1.0)
    t.push(random_entities(np.random.default_rng(0), n))
    assert t.get_leaves_with_resolution([0, 0, 0], -0.1) == n


@pytest.mark.parametrize("dim", DIMS)
def test_extreme_coordinates(dim):
    t = ob.Tree(dim, extent=1000.0)
    t.push(ent(999.0, 999.0, 999.0))
    t.push(ent(-999.0, -999.0, -999.0))
    assert t.get_leaves_with_resolution([0, 0, 0], -0.1) == 2


@pytest.mark.parametrize("dim", DIMS)
def test_zero_extent(dim):
    t = ob.Tree(dim, extent=0.001)
    t.push(ent(0, 0, 0))
    assert t.get_leaves_with_resolution([0, 0, 0], -0.1) == 1


@pytest.mark.parametrize("dim", DIMS)
def test_query_outside_bounds(dim):
    t = ob.Tree(dim, extent=1.0)
    t.push(ent(0, 0, 0))
    assert t.get_leaves_with_resolution([1000.0, 1000.0, 1000.0], 0.5) > 0


@pytest.mark.parametrize("dim", DIMS)
def test_reproducibility(dim):
    t1, t2 = ob.Tree(dim, extent=1.0), ob.Tree(dim, extent=1.0)
    t1.push(random_entities(np.random.default_rng(999), 1000))
    t2.push(random_entities(np.random.default_rng(999), 1000))
    assert t1.get_leaves_with_resolution([0, 0, 0], -0.1) == t2.get_leaves_with_resolution([0, 0, 0], -0.1)


# ---- beyond the reference's tests: behaviours the restatement must keep ----------------------

def test_depth_guard_panics():
    """Distinct bodies closer than extent·2^-64 but further than 1e-9 cannot exist in fp64 at
    extent 1; at a huge extent they can: recursion deeper than 64 panics (octree.rs:60-62)."""
    t = ob.Tree(3, extent=1e30)
    t.push(ent(1.0, 1.0, 1.0))
    with pytest.raises(ob.OraclePanic):
        t.push(ent(1.0 + 1e-6, 1.0, 1.0))


def test_zero_mass_pair_panics():
    """centre_of_mass of two massless bodies is NaN and Star::fake panics (lib.rs:65-67)."""
    t = ob.Tree(3, extent=1.0)
    t.push(ent(0.5, 0.5, 0.5, mass=0.0))
    with pytest.raises(ob.OraclePanic):
        t.push(ent(-0.5, 0.5, 0.5, mass=0.0))


def test_acceptance_uses_half_width_and_geometric_centre():
    """octree.rs:137-144: accept iff extent / |p - centre| < theta, centre = cell centre."""
    t = ob.Tree(3, extent=1.0)
    t.push(ent(0.9, 0.9, 0.9))
    t.push(ent(-0.8, -0.7, -0.6))
    # root (extent 1, centre 0) seen from distance 4 along x: 1/4 < 0.3 -> one aggregated entity
    out = t.get_leaves_with_resolution([4.0, 0, 0], 0.3, want=True)
    assert len(out) == 1 and out["mass"][0] == 2.0
    # 1/4 is not < 0.25 -> opened
    assert t.get_leaves_with_resolution([4.0, 0, 0], 0.25) == 2
