"""Committed golden fixtures (tests/golden/*.npz, frozen oracle outputs on small seeded inputs).

CPU: the oracle must still reproduce them exactly.  GPU: the CUDA path must match them — tree
arrays bit for bit, interaction counts exactly, accelerations within the stated tolerance."""
import os

import numpy as np
import pytest

from oracle import binding as ob
from physim_b200.entity import ENTITY
from tests.util import MEDIAN_TOL, P99_TOL

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DIM = {"astro": 2, "astro2": 3}


def load(name):
    z = np.load(os.path.join(HERE, name))
    return z, z["state"].view(ENTITY)


def vec(a):
    return np.stack([a["x"], a["y"], a["z"]], 1)


FIELDS = ("x", "y", "z", "vx", "vy", "vz", "radius", "mass", "id", "fixed")


def same_entities(a, b):
    """Field-wise equality (the 7 padding bytes after `fixed` are unspecified)."""
    return all(np.array_equal(a[f], b[f]) for f in FIELDS)


def rel(a, ref):
    return np.linalg.norm(a - ref, axis=1) / np.linalg.norm(ref, axis=1)


def test_oracle_reproduces_solar():
    z, s = load("solar.npz")
    assert np.array_equal(vec(ob.transform("simple_astro", s, e=0.1)), z["acc"])
    final, _ = ob.run_pipeline("simple_astro", s, 1.0, 0.1, 0.01, 10)
    assert same_entities(final, z["final"].view(ENTITY))


@pytest.mark.parametrize("name", ["astro2", "astro"])
def test_oracle_reproduces_readme_small(name):
    z, s = load("readme_small.npz")
    tab = ob.CellTable(DIM[name], s)
    for k in ("key", "perm", "cell_start", "level", "head", "count", "skip", "parent", "centre_ext", "com_mass"):
        assert np.array_equal(getattr(tab, k), z[f"{name}_{k}"]), k
    for theta in (0.5, 1.0, 1.5):
        a, c = ob.transform(name, s, theta, 0.5, counts=True)
        assert np.array_equal(c, z[f"{name}_cnt_{theta}"])
        assert np.array_equal(vec(a), z[f"{name}_acc_{theta}"])
    fin, _ = ob.run_pipeline(name, s, 1.5, 0.5, 1e-5, 5)
    assert same_entities(fin, z[f"{name}_final"].view(ENTITY))


@pytest.mark.gpu
def test_gpu_matches_golden_solar():
    from physim_b200 import api
    z, s = load("solar.npz")
    acc = api.TransformElement("simple_astro", e=0.1).transform(s)
    free = ~s["fixed"]
    r = rel(vec(acc)[free], z["acc"][free])
    assert np.median(r) <= MEDIAN_TOL and np.percentile(r, 99) <= P99_TOL
    sim = api.Sim("simple_astro", e=0.1, dt=0.01)
    sim.upload(s)
    sim.run(10)
    out = sim.download(s.copy())
    ref = z["final"].view(ENTITY)
    for k in ("x", "y", "z", "vx", "vy", "vz"):
        np.testing.assert_allclose(out[k], ref[k], rtol=1e-5, atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["astro2", "astro"])
def test_gpu_matches_golden_readme_small(name):
    from physim_b200 import api
    z, s = load("readme_small.npz")
    for theta in (0.5, 1.0, 1.5):
        el = api.TransformElement(name, theta=theta, e=0.5)
        acc = el.transform(s)
        t = el.debug_tree()
        for k in ("key", "perm", "cell_start", "level", "head", "count", "skip", "parent", "centre_ext"):
            assert np.array_equal(t[k], z[f"{name}_{k}"]), k
        assert np.array_equal(t["counts"], z[f"{name}_cnt_{theta}"])
        r = rel(vec(acc), z[f"{name}_acc_{theta}"])
        assert np.median(r) <= MEDIAN_TOL and np.percentile(r, 99) <= P99_TOL
