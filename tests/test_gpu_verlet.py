"""GPU parity for the verlet step and the composed `transform ! verlet` loop."""
import numpy as np
import pytest

from oracle import binding as ob
from physim_b200 import api
from physim_b200 import generators as gen
from tests.util import assert_acc_parity, vec

pytestmark = pytest.mark.gpu

POS = ("x", "y", "z", "vx", "vy", "vz")


def test_generic_step_is_bit_exact():
    """Same accelerations in -> identical doubles out (first step, regular steps, N change)."""
    s = gen.cube(5000, seed=4, spin=3.0)
    s["fixed"][3] = True
    a = np.random.default_rng(1).normal(size=(5000, 3)) * 10.0

    def fn(state, acc):
        acc["x"] += a[: len(state), 0]; acc["y"] += a[: len(state), 1]; acc["z"] += a[: len(state), 2]

    g, o = api.Verlet(), ob.Verlet()
    sg, so = s, s
    for step in range(4):
        sg, so = g.integrate(sg, fn, 0.125), o.integrate(so, fn, 0.125)
        assert sg.tobytes() == so.tobytes(), f"step {step}"
    sg, so = g.integrate(sg[:777], fn, 0.125), o.integrate(so[:777], fn, 0.125)  # N changed: first-step rule
    assert sg.tobytes() == so.tobytes()


def test_shm_analytic_kat():
    """example_pipelines/shm.toml: x(t) = cos(sqrt(5) t), dt = 0.01, 3140 steps."""
    k, dt, iters = 5.0, 0.01, 3140

    def fn(state, acc):
        for c in "xyz":
            acc[c] += -k * state[c] / state["mass"]

    st = gen.star(x=1.0, mass=1.0, radius=0.2)
    v = api.Verlet()
    xs = []
    for _ in range(iters):
        st = v.integrate(st, fn, dt)
        xs.append(st["x"][0])
    t = dt * np.arange(1, iters + 1)
    assert np.abs(np.array(xs) - np.cos(np.sqrt(k) * t)).max() < 2e-3


@pytest.mark.parametrize("name,theta,e", [("astro2", 1.5, 0.5), ("astro", 1.3, 1.0), ("simple_astro", 1.0, 0.5)])
def test_fused_step_matches_reference_pipeline(name, theta, e):
    s = gen.readme_pipeline(20_000, seed=2, spin=1000.0)
    dt, steps = 1e-5, 5
    ref, _ = ob.run_pipeline(name, s, theta, e, dt, steps)
    el, v = api.TransformElement(name, theta=theta, e=e), api.Verlet()
    cur = s
    for _ in range(steps):
        cur = v.integrate_fused(cur, el, dt)
    for f in ("radius", "mass", "id", "fixed"):
        assert np.array_equal(cur[f], s[f])
    dx = np.abs(np.stack([cur[k] - ref[k] for k in "xyz"], 1)).max()
    disp = np.abs(np.stack([ref[k] - s[k] for k in "xyz"], 1)).max()
    assert dx <= 1e-6 * disp
    vref = np.stack([ref[k] for k in ("vx", "vy", "vz")], 1)
    dv = np.abs(np.stack([cur[k] for k in ("vx", "vy", "vz")], 1) - vref).max()
    assert dv <= 1e-6 * np.abs(vref).max()


def test_generic_step_with_plugin_transform_as_callback():
    """The drop-in composition physim runs: verlet.integrate(acc_fn = our transform via the C ABI)."""
    s = gen.solar()
    el = api.TransformElement("simple_astro", e=0.1)
    g, o = api.Verlet(), ob.Verlet()
    sg = so = s
    for _ in range(20):
        sg = g.integrate(sg, lambda st, ac: el.transform(st, ac), 0.01)
        so = o.integrate(so, lambda st, ac: ob.transform("simple_astro", st, e=0.1, acc=ac), 0.01)
    for k in POS:
        np.testing.assert_allclose(sg[k], so[k], rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("name,theta,e", [("astro2", 1.5, 0.5), ("simple_astro", 1.0, 0.5)])
def test_device_resident_sim_matches_reference_pipeline(name, theta, e):
    s = gen.readme_pipeline(10_000, seed=9, spin=1000.0)
    dt, steps = 1e-5, 10
    ref, _ = ob.run_pipeline(name, s, theta, e, dt, steps)
    sim = api.Sim(name, theta=theta, e=e, dt=dt)
    sim.upload(s)
    sim.run(steps)
    out = sim.download(s.copy())
    disp = np.abs(np.stack([ref[k] - s[k] for k in "xyz"], 1)).max()
    assert np.abs(np.stack([out[k] - ref[k] for k in "xyz"], 1)).max() <= 1e-6 * disp
    # sharded evaluation (world = 2) over one device: each rank's slice equals the full run's slice
    outs = []
    for rank in range(2):
        sh = api.Sim(name, theta=theta, e=e, dt=dt, rank=rank, world=2)
        sh.upload(s)
        sh.step_local()
        outs.append(sh.download(s.copy()))
    one = api.Sim(name, theta=theta, e=e, dt=dt)
    one.upload(s)
    one.run(1)
    full = one.download(s.copy())
    h = len(s) // 2
    for k in POS:
        assert np.array_equal(outs[0][k][:h], full[k][:h])
        assert np.array_equal(outs[1][k][h:], full[k][h:])


@pytest.mark.parametrize("name,theta", [("astro", 1.3), ("astro2", 0.7)])
def test_unverified_chunks_equal_checked_steps(name, theta):
    """pb200_sim_run enqueues 32-step chunks whose tree builds are verified only at the chunk's end
    (cell-table size, truncated-sort depth) and replays a chunk that fails.  Whatever happens, the
    result must be bit-identical to taking the same steps with a host check after every build."""
    s = gen.readme_pipeline(30_000, seed=8, spin=1000.0)
    steps, dt = 100, 2e-5          # long enough for the closest pairs (tree depth) to change
    a = api.Sim(name, theta=theta, e=0.5, dt=dt)
    a.upload(s)
    a.run(steps)
    out_a = a.download(s.copy())
    b = api.Sim(name, theta=theta, e=0.5, dt=dt)
    b.upload(s)
    for _ in range(steps):
        b.step_local()             # host-checked every step
    out_b = b.download(s.copy())
    for k in POS:
        assert np.array_equal(out_a[k], out_b[k]), k
    st = a.stats()
    assert st["sort_bits"] <= 63 and st["replays"] <= 2


@pytest.mark.parametrize("name,theta", [("astro", 1.3), ("astro2", 0.7), ("simple_astro", 1.0)])
def test_lean_resident_steps_equal_general_kernel(name, theta, monkeypatch):
    """The device-resident loop's lean verlet step (x_{n+1} written over x_{n-1}, buffers swapped, velocity
    derived on demand, extent of the new positions handed to the next tree build) must leave the same bits
    as the general kernel, also when the state is read out mid-run and when the integrator changes."""
    s = gen.readme_pipeline(20_000, seed=12, spin=1000.0)
    dt = 2e-5

    def run(lean):
        monkeypatch.setenv("PB200_VERLET_LEAN", "1" if lean else "0")
        sim = api.Sim(name, theta=theta, e=0.5, dt=dt)
        sim.upload(s)
        outs = []
        sim.run(1)
        outs.append(sim.download(s.copy()))      # after the first-step formula
        sim.run(4)
        outs.append(sim.download(s.copy()))      # mid-run read-out (velocities materialised)
        sim.run(35)                              # crosses a 32-step checkpoint
        outs.append(sim.download(s.copy()))
        sim.set_integrator("euler")              # needs the velocities of the current state
        sim.run(2)
        sim.set_integrator("verlet")             # first-step formula again, then lean steps
        sim.run(3)
        outs.append(sim.download(s.copy()))
        return outs

    a, b = run(True), run(False)
    for i, (x, y) in enumerate(zip(a, b)):
        for k in POS:
            assert np.array_equal(x[k], y[k]), (i, k)


def test_stats_extent_after_lean_steps():
    """Pb200Stats.extent must be the root half-width the LAST build used, also in the lean resident loop
    where the build takes its extent from the slot the previous verlet step reduced into."""
    s = gen.readme_pipeline(20_000, seed=3, spin=1000.0)
    dt = 1e-5
    sim = api.Sim("astro2", theta=1.5, e=0.5, dt=dt)
    sim.upload(s)
    sim.run(5)
    before = sim.download(s.copy())        # state the 6th build will see
    sim.run(1)
    want = max(np.abs(before[k]).max() for k in ("x", "y", "z"))
    assert sim.stats()["extent"] == want


def potential(s, e):
    """Conserved potential of the reference's force law: U = -mi mj (pi/2 - atan(r/sqrt(e)))/sqrt(e)."""
    p = np.stack([s["x"], s["y"], s["z"]], 1)
    m = s["mass"]
    r = np.linalg.norm(p[:, None, :] - p[None, :, :], axis=2)
    iu = np.triu_indices(len(s), 1)
    se = np.sqrt(e)
    return -(m[iu[0]] * m[iu[1]] * (np.pi / 2 - np.arctan(r[iu] / se)) / se).sum()


def energy(s, e):
    k = 0.5 * (s["mass"] * (s["vx"] ** 2 + s["vy"] ** 2 + s["vz"] ** 2)).sum()
    return k + potential(s, e)


def test_energy_and_momentum_drift_solar():
    """BASELINE config 2 (solar.toml: simple_astro e=0.1, verlet, dt=0.01), 1000 steps, sun free so
    that momentum is conserved: GPU drift within 2x of the oracle's and small in absolute terms."""
    s = gen.solar()
    s["fixed"][:] = False
    e, dt, steps = 0.1, 0.01, 1000
    e0 = energy(s, e)
    ref, _ = ob.run_pipeline("simple_astro", s, 1.0, e, dt, steps)
    sim = api.Sim("simple_astro", e=e, dt=dt)
    sim.upload(s)
    sim.run(steps)
    out = sim.download(s.copy())
    d_ref = abs(energy(ref, e) - e0) / abs(e0)
    d_gpu = abs(energy(out, e) - e0) / abs(e0)
    assert d_gpu <= max(2.0 * d_ref, 1e-6) and d_gpu < 1e-3

    def mom(x):
        return (x["mass"][:, None] * np.stack([x["vx"], x["vy"], x["vz"]], 1)).sum(0)

    scale = (s["mass"] * np.sqrt(s["vx"] ** 2 + s["vy"] ** 2 + s["vz"] ** 2)).sum()
    assert np.abs(mom(out) - mom(s)).max() / scale < 1e-6 + 10 * np.abs(mom(ref) - mom(s)).max() / scale


def test_device_resident_run_feeds_csvsink(tmp_path):
    """`cube ! astro2 ! verlet ! csvsink print_n=2`: the device-resident loop as the producer of the
    renderer's states (pipeline.rs:129-131,179); rows against the oracle's pipeline."""
    s = gen.readme_pipeline(3000, seed=4, spin=1000.0)
    dt, steps, print_n = 1e-5, 6, 2
    sim = api.Sim("astro2", theta=1.5, e=0.5, dt=dt)
    sim.upload(s)
    path = tmp_path / "run.csv"
    sink = api.CsvSink(str(path), print_n)
    sim.run_csvsink(steps, sink)
    assert sink.count() == steps + 1
    sink.close()
    lines = path.read_text().split("\n")
    assert lines[-1] == "" and len(lines) - 1 == 1 + steps // print_n
    printed = [0] + [k for k in range(1, steps + 1) if k % print_n == 0]
    for line, k in zip(lines, printed):
        want = s if k == 0 else ob.run_pipeline("astro2", s, 1.5, 0.5, dt, k)[0]
        row = np.array(line.split(",")[:-1], dtype=np.float64).reshape(-1, 3)
        ref = np.stack([want["x"], want["y"], want["z"]], 1)
        if k == 0:
            assert np.array_equal(row, ref)                      # shortest digits round-trip exactly
        else:
            disp = np.abs(ref - np.stack([s["x"], s["y"], s["z"]], 1)).max()
            assert np.abs(row - ref).max() <= 1e-5 * disp


def test_resident_fused_steps_equal_copy_every_step():
    """pb200_verlet_set_resident: same bits as the fused step that copies the state in and out every call; the
    upload is skipped once the input is the previous output, and a replaced state is noticed."""
    s = gen.readme_pipeline(30_000, seed=6, spin=1000.0)
    s["radius"] = np.linspace(0.01, 0.02, len(s))          # pass-through fields must come back untouched
    s["id"] = np.arange(len(s)) % 7
    s["fixed"][11] = True
    dt = 1e-5
    el_a, el_b = api.TransformElement("astro2", theta=1.0, e=0.5), api.TransformElement("astro2", theta=1.0, e=0.5)
    va, vb = api.Verlet(), api.Verlet()
    vb.set_resident(True)
    bufs_a, bufs_b = [s.copy(), s.copy()], [s.copy(), s.copy()]
    k = 0
    for step in range(6):
        va.integrate_fused(bufs_a[k], el_a, dt, out=bufs_a[k ^ 1])
        vb.integrate_fused(bufs_b[k], el_b, dt, out=bufs_b[k ^ 1])
        k ^= 1
        for f in ("x", "y", "z", "vx", "vy", "vz", "radius", "mass", "id", "fixed"):
            assert np.array_equal(bufs_a[k][f], bufs_b[k][f]), (step, f)
    hits, misses = vb.resident_counts()
    # call 1 uploads; call 2 writes into the other buffer (newly page-locked: uploads); from then on the buffers
    # alternate and the input of a call is NOT the buffer the previous call wrote... unless the caller clones, as
    # physim does: emulate that from here on
    state, new_state = bufs_b[k].copy(), bufs_b[k].copy()
    ref_state = bufs_a[k].copy()
    vb.integrate_fused(state, el_b, dt, out=new_state)     # new output buffer: full upload
    h0, m0 = vb.resident_counts()
    for step in range(5):
        state = new_state.copy()                           # pipeline.rs:173
        vb.integrate_fused(state, el_b, dt, out=new_state)
    h1, m1 = vb.resident_counts()
    assert (h1 - h0, m1 - m0) == (5, 0)
    ref_new = ref_state.copy()
    for step in range(6):
        va.integrate_fused(ref_state, el_a, dt, out=ref_new)
        ref_state = ref_new.copy()
    for f in ("x", "y", "z", "vx", "vy", "vz", "radius", "id", "fixed"):
        assert np.array_equal(new_state[f], ref_new[f]), f
    # a different state in the same buffers: noticed by the sample, uploaded in full
    other = gen.readme_pipeline(30_000, seed=99, spin=10.0)
    vb.integrate_fused(other, el_b, dt, out=new_state)
    h2, m2 = vb.resident_counts()
    assert m2 == m1 + 1
    # same n: verlet.rs:52-82 still takes the PREVIOUS call's input as x_{n-1} (only a change of n restarts the
    # scheme), so the reference result is the copy-every-step handle with the same history given the same input
    want = va.integrate_fused(other, el_a, dt)
    for f in ("x", "y", "z", "vx", "vy", "vz"):
        assert np.array_equal(new_state[f], want[f]), f


def test_dropin_composition_equals_fused_within_acc_rounding():
    """IntegratorElement::integrate with an acc_fn that calls the plugin transform through its vtable (what stock
    physim runs): same accelerations (fp32 -> fp64 on the host), same verlet arithmetic as the fused step."""
    s = gen.readme_pipeline(20_000, seed=2, spin=1000.0)
    el = api.TransformElement("astro2", theta=1.0, e=0.5)
    a = api.Verlet().integrate_dropin(s, el, 1e-5)
    b = api.Verlet().integrate_fused(s, el, 1e-5)
    for f in ("x", "y", "z", "vx", "vy", "vz"):
        assert np.array_equal(a[f], b[f]), f
