"""euler / rk4 restatements (integrators/src/euler.rs, rk4.rs) pinned against closed forms."""
import numpy as np

from oracle import binding as ob
from physim_b200 import generators as gen


def shm(k):
    def fn(state, acc):  # mechanics/src/shm.rs:87-95
        for c in "xyz":
            acc[c] += -k * state[c] / state["mass"]
    return fn


def test_rk4_is_fourth_order_on_shm():
    k = 5.0
    errs = []
    for dt in (0.02, 0.01):
        st = gen.star(x=1.0, mass=1.0)
        g = ob.Integrator("rk4")
        steps = int(round(2.0 / dt))
        for _ in range(steps):
            st = g.integrate(st, shm(k), dt)
        errs.append(abs(st["x"][0] - np.cos(np.sqrt(k) * 2.0)))
    assert errs[1] < 1e-8 and 12.0 < errs[0] / errs[1] < 20.0     # error ~ dt^4


def test_euler_is_verlets_first_step_every_step():
    s = gen.cube(32, seed=3, spin=2.0)
    a = np.random.default_rng(0).normal(size=(32, 3))

    def fn(state, acc):
        acc["x"] += a[:, 0]; acc["y"] += a[:, 1]; acc["z"] += a[:, 2]

    dt = 0.25
    e, v = ob.Integrator("euler"), ob.Integrator("verlet")
    s1 = e.integrate(s, fn, dt)
    assert np.array_equal(s1["x"], v.integrate(s, fn, dt)["x"])            # first verlet step
    s2 = e.integrate(s1, fn, dt)
    np.testing.assert_array_equal(s2["x"], s1["x"] + s1["vx"] * dt + 0.5 * a[:, 0] * (dt * dt))
    np.testing.assert_array_equal(s2["vx"], s1["vx"] + a[:, 0] * dt)


def test_rk4_formula_and_fixed_bodies():
    s = gen.cube(8, seed=1, spin=1.5)
    s["fixed"][2] = True
    g = 0.7

    def fn(state, acc):  # a = -g x  (so the four stages differ)
        acc["x"] += -g * state["x"]; acc["y"] += -g * state["y"]; acc["z"] += -g * state["z"]

    dt = 0.1
    out = ob.Integrator("rk4").integrate(s, fn, dt)
    x, v = s["x"], s["vx"]
    k1x, k1v = dt * v, dt * (-g * x)
    x2, v2 = x + 0.5 * k1x, v + 0.5 * k1v
    k2x, k2v = dt * v2, dt * (-g * x2)
    x3, v3 = x + 0.5 * k2x, v + 0.5 * k2v
    k3x, k3v = dt * v3, dt * (-g * x3)
    x4, v4 = x + k3x, v + k3v
    k4x, k4v = dt * v4, dt * (-g * x4)
    want_x = x + (k1x + 2.0 * k2x + 2.0 * k3x + k4x) / 6.0
    want_v = v + (k1v + 2.0 * k2v + 2.0 * k3v + k4v) / 6.0
    free = ~s["fixed"]
    np.testing.assert_array_equal(out["x"][free], want_x[free])
    np.testing.assert_array_equal(out["vx"][free], want_v[free])
    assert out["x"][2] == s["x"][2] and out["vx"][2] == 0.0 and out["vy"][2] == 0.0   # rk4.rs:163-170
    for f in ("mass", "radius", "id", "fixed"):
        assert np.array_equal(out[f], s[f])
