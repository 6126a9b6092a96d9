"""BASELINE configs[3] / configs[4] at (or as near as a test allows to) full size, checked on target samples.

C4  cube n = 2^24, astro2 theta=0 (all pairs) e=0.5: a 4096-target slice of the 2^24-source evaluation on the GPU,
    256 of those targets against the oracle's fp64 all-pairs sum over all 2^24 sources (~35 s of CPU).
C5  cube n = 2^26, astro2 theta=0.7 e=0.5: the reference's pointer tree for 2^26 bodies needs ~18 GB and minutes
    of CPU, so the test runs the same configuration at 2^22 bodies (stated; the bench's bh_large extra runs 2^26):
    4096 sampled targets, interaction counts exactly and accelerations within tolerance against oracle 1
    (astro/src/octree.rs:130-158 walk on the oracle's pointer tree, force law astro/src/lib.rs:84-113)."""
import numpy as np
import pytest

from oracle import binding as ob
from physim_b200 import api
from physim_b200 import generators as gen
from physim_b200.entity import accelerations
from tests.util import assert_acc_parity

pytestmark = pytest.mark.gpu


def test_c4_direct_sum_full_source_set_sampled_targets():
    n = 1 << 24
    sim = api.Sim("astro2", theta=0.0, e=0.5, dt=1e-6)
    sim.generate_cube(n, seed=1)                         # the reference's own `cube seed=1` stream, made on the device
    t0, n_t = 5_000_000, 4096
    sim.set_targets(t0, t0 + n_t)
    state = sim.download(np.zeros(n, dtype=gen.entities(1).dtype))
    state["mass"] = 1.0 / n
    sim.run(1)
    acc = sim.last_accelerations()
    pick = t0 + np.arange(0, n_t, 16)                    # 256 targets
    ref = accelerations(n)
    for i in pick:
        ob.direct_range(state, 0.5, int(i), int(i) + 1, acc=ref)
    r = assert_acc_parity(acc[pick], ref[pick])
    assert r.max() < 1e-4
    outside = np.r_[acc["x"][:t0], acc["x"][t0 + n_t:]]
    assert not outside.any()                             # targets outside the slice untouched


def test_c5_barnes_hut_theta07_sampled_targets_at_4m():
    n = 1 << 22
    s = gen.cube_chacha8(n, seed=1)
    theta, e = 0.7, 0.5
    el = api.TransformElement("astro2", theta=theta, e=e)
    acc = el.transform(s)
    counts = el.debug_tree()["counts"]
    tree = ob.Tree(3, extent=ob.state_extent(s))
    tree.push(s)
    pick = np.random.default_rng(0).choice(n, 4096, replace=False)
    ref = accelerations(n)
    want_counts = np.zeros(len(pick), dtype=np.int64)
    px, py, pz = s["x"], s["y"], s["z"]
    for k, i in enumerate(pick):
        leaves = tree.get_leaves_with_resolution((px[i], py[i], pz[i]), theta, want=True)
        want_counts[k] = len(leaves)
        dx, dy, dz = leaves["x"] - px[i], leaves["y"] - py[i], leaves["z"] - pz[i]
        r2 = dx * dx + dy * dy + dz * dz
        keep = r2 > 0                                    # transformers.rs:145-147 (same position: skipped)
        w = leaves["mass"][keep] / (np.sqrt(r2[keep]) * (r2[keep] + e))
        ref["x"][i], ref["y"][i], ref["z"][i] = (w * dx[keep]).sum(), (w * dy[keep]).sum(), (w * dz[keep]).sum()
    assert np.array_equal(counts[pick], want_counts)
    assert 30 < want_counts.mean() < 120                 # the theta = 0.7 regime (SURVEY: ~44-53 per target at 1 M)
    assert_acc_parity(acc[pick], ref[pick])
