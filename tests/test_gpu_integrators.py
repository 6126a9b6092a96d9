"""GPU parity for euler and rk4: bit-exact given the same accelerations, and
within tolerance when composed with the GPU gravity transforms."""
import numpy as np
import pytest

from oracle import binding as ob
from physim_b200 import api
from physim_b200 import generators as gen

pytestmark = pytest.mark.gpu
FIELDS = ("x", "y", "z", "vx", "vy", "vz", "radius", "mass", "id", "fixed")


def same(a, b):
    return all(np.array_equal(a[f], b[f]) for f in FIELDS)


@pytest.mark.parametrize("name", ["euler", "rk4"])
def test_generic_step_is_bit_exact(name):
    s = gen.cube(3000, seed=4, spin=3.0)
    s["fixed"][7] = True
    g = 0.7

    def fn(state, acc):  # position-dependent, so rk4's four evaluation points matter
        acc["x"] += -g * state["x"] + 0.1 * state["vy"]
        acc["y"] += -g * state["y"]
        acc["z"] += -g * state["z"] + 0.3

    dev, ref = api.Verlet(name), ob.Integrator(name)
    sg = so = s
    for step in range(3):
        sg, so = dev.integrate(sg, fn, 0.05), ref.integrate(so, fn, 0.05)
        assert same(sg, so), f"{name} step {step}"
    sg, so = dev.integrate(sg[:501], fn, 0.05), ref.integrate(so[:501], fn, 0.05)
    assert same(sg, so)


@pytest.mark.parametrize("name", ["euler", "rk4"])
@pytest.mark.parametrize("element,theta,e", [("astro2", 1.5, 0.5), ("simple_astro", 1.0, 0.5)])
def test_fused_and_resident_match_reference_pipeline(name, element, theta, e):
    s = gen.readme_pipeline(8000, seed=2, spin=1000.0)
    s["fixed"][11] = True
    dt, steps = 1e-5, 4
    ref = ob.run_pipeline_with(name, element, s, theta, e, dt, steps)
    disp = np.abs(np.stack([ref[k] - s[k] for k in "xyz"], 1)).max()
    vref = np.stack([ref[k] for k in ("vx", "vy", "vz")], 1)

    def check(out):
        assert np.abs(np.stack([out[k] - ref[k] for k in "xyz"], 1)).max() <= 1e-6 * disp
        assert np.abs(np.stack([out[k] for k in ("vx", "vy", "vz")], 1) - vref).max() <= 1e-6 * np.abs(vref).max()
        for f in ("radius", "mass", "id", "fixed"):
            assert np.array_equal(out[f], s[f])

    el, g = api.TransformElement(element, theta=theta, e=e), api.Verlet(name)
    cur = s
    for _ in range(steps):
        cur = g.integrate_fused(cur, el, dt)
    check(cur)
    sim = api.Sim(element, theta=theta, e=e, dt=dt)
    sim.set_integrator(name)
    sim.upload(s)
    sim.run(steps)
    check(sim.download(s.copy()))
    if name == "rk4":   # rk4.rs:163-170: a fixed body keeps its place and loses its velocity
        assert cur["x"][11] == s["x"][11] and cur["vx"][11] == 0.0


def test_rk4_with_plugin_transform_as_callback():
    """The composition physim runs for `astro2 ! rk4` (example_pipelines/energysink.toml): the
    integrator calls the transform four times per step through the C ABI with host temporaries."""
    s = gen.readme_pipeline(3000, seed=6, spin=1000.0)
    el = api.TransformElement("astro2", theta=1.0, e=0.5)
    dev, ref = api.Verlet("rk4"), ob.Integrator("rk4")
    sg = so = s
    for _ in range(3):
        sg = dev.integrate(sg, lambda st, ac: el.transform(st, ac), 1e-5)
        so = ref.integrate(so, lambda st, ac: ob.transform("astro2", st, 1.0, 0.5, acc=ac), 1e-5)
    disp = np.abs(so["x"] - s["x"]).max()
    assert np.abs(sg["x"] - so["x"]).max() <= 1e-6 * disp


def test_device_cube_generator_equals_host_chacha8_stream():
    """`cube` made on the device (csrc/generate.cu) against the host restatement of the reference's
    ChaCha8 stream (physim_b200.generators.cube_chacha8), bit for bit, ragged n included."""
    from physim_b200 import generators as gen
    for n, seed, spin, size, centre in ((8, 1, 0.0, 1.0, (0, 0, 0)), (100_003, 1, 1000.0, 1.0, (0, 0, 0)),
                                        (4099, 77, 3.5, 2.5, (10.0, -20.0, 0.125))):
        want = gen.cube_chacha8(n, seed=seed, spin=spin, mass=2.0, size=size, centre=centre)
        sim = api.Sim("astro2", theta=1.0, e=0.5, dt=1e-5)
        sim.generate_cube(n, seed=seed, spin=spin, mass=2.0, size=size, centre=centre)
        got = sim.download(np.zeros(n, dtype=want.dtype))
        for k in ("x", "y", "z", "vx", "vy", "vz"):
            assert np.array_equal(got[k], want[k]), (k, n)
        sim.run(2)                                   # and the state is usable: two steps from it
        ref = api.Sim("astro2", theta=1.0, e=0.5, dt=1e-5)
        ref.upload(want)
        ref.run(2)
        a, b = sim.download(want.copy()), ref.download(want.copy())
        for k in ("x", "y", "z"):
            assert np.array_equal(a[k], b[k]), k
