"""Verlet restatement (integrators/src/verlet.rs:23-107) pinned against the analytic solution of
example_pipelines/shm.toml: one body, mass 1, x0 = 1, `shm mode=centre k=5`, dt = 0.01, 3140 steps
=> x(t) = cos(sqrt(k) t)."""
import numpy as np

from oracle import binding as ob
from physim_b200 import generators as gen


def shm_acc(k):
    def fn(state, acc):  # mechanics/src/shm.rs:87-95 (global centre, c = 0)
        acc["x"] += -k * state["x"] / state["mass"]
        acc["y"] += -k * state["y"] / state["mass"]
        acc["z"] += -k * state["z"] / state["mass"]
    return fn


def test_shm_analytic():
    k, dt, iters = 5.0, 0.01, 3140
    state = gen.star(x=1.0, mass=1.0, radius=0.2)
    v = ob.Verlet()
    xs = []
    for _ in range(iters):
        state = v.integrate(state, shm_acc(k), dt)
        xs.append(state["x"][0])
    t = dt * np.arange(1, iters + 1)
    err = np.abs(np.array(xs) - np.cos(np.sqrt(k) * t))
    # Störmer-Verlet phase error ~ (w dt)^2/24 * w t ≈ 1.5e-3 at t = 31.4
    assert err.max() < 2e-3
    assert err[:100].max() < 5e-5


def test_first_step_and_regular_step_formulas():
    s = gen.cube(16, seed=4, spin=3.0)
    a = np.random.default_rng(1).normal(size=(16, 3))

    def fn(state, acc):
        acc["x"] += a[:, 0]; acc["y"] += a[:, 1]; acc["z"] += a[:, 2]

    dt = 0.125
    v = ob.Verlet()
    s1 = v.integrate(s, fn, dt)
    np.testing.assert_array_equal(s1["x"], s["x"] + s["vx"] * dt + 0.5 * a[:, 0] * (dt * dt))
    np.testing.assert_array_equal(s1["vx"], s["vx"] + a[:, 0] * dt)
    s2 = v.integrate(s1, fn, dt)
    x2 = 2.0 * s1["x"] - s["x"] + a[:, 0] * (dt * dt)
    np.testing.assert_array_equal(s2["x"], x2)
    np.testing.assert_array_equal(s2["vx"], (x2 - s1["x"]) / dt)
    # radius / mass / id / fixed pass through untouched; `fixed` is NOT honoured by verlet
    for f in ("radius", "mass", "id", "fixed"):
        np.testing.assert_array_equal(s2[f], s[f])
    # a change of N re-runs the first-step formula (verlet.rs:102-106)
    s3 = v.integrate(s2[:8], lambda st, ac: None, dt)
    np.testing.assert_array_equal(s3["x"], s2["x"][:8] + s2["vx"][:8] * dt)


def test_pipeline_loop_matches_stepwise():
    s = gen.solar()
    final, secs = ob.run_pipeline("simple_astro", s, 1.0, 0.1, 0.01, 50)
    v = ob.Verlet()
    cur = s
    for _ in range(50):
        cur = v.integrate(cur, lambda st, ac: ob.transform("simple_astro", st, e=0.1, acc=ac), 0.01)
    for f in ("x", "y", "z", "vx", "vy", "vz"):
        np.testing.assert_array_equal(final[f], cur[f])
    assert secs.sum() > 0
