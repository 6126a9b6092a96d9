"""Several GPUs of one box (csrc/multi.cu): the sharded run must equal the single-GPU run BIT FOR BIT.

Skipped with fewer than two devices (run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).
Single process, one rank per device - the shape the physim plugin has; tools/multi_rank_check.py runs the same
comparison with one process per rank under torchrun (NCCL across processes, CUDA IPC peer mappings)."""
import numpy as np
import pytest

from physim_b200 import api
from physim_b200 import generators as gen

pytestmark = pytest.mark.gpu
POS = ("x", "y", "z", "vx", "vy", "vz")


def n_devices():
    try:
        return api.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(n_devices() < 2, reason="needs two CUDA devices")


def single(name, theta, e, dt, s, steps):
    sim = api.Sim(name, theta=theta, e=e, dt=dt)
    sim.upload(s)
    sim.run(steps)
    return sim.download(s.copy()), sim.last_accelerations(), sim.stats()


@needs2
@pytest.mark.parametrize("name,theta", [("astro", 1.3), ("astro2", 0.7), ("astro2", 1.5)])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_equals_single_gpu_bitwise(name, theta, world):
    if n_devices() < world:
        pytest.skip(f"needs {world} devices")
    s = gen.readme_pipeline(60_000, seed=5, spin=1000.0)
    dt, steps = 1e-5, 10
    want, want_acc, st1 = single(name, theta, 0.5, dt, s, steps)
    ms = api.MultiSim(name, theta=theta, e=0.5, dt=dt, devices=list(range(world)))
    ms.upload(s)
    ms.run(steps)
    got = ms.download(s.copy())
    st = ms.stats()
    assert st["sharded_steps"] == steps - 1 and st["replicated_steps"] == 1 and st["replays"] == 0, st
    for k in POS:
        assert np.array_equal(got[k], want[k]), k
    acc = ms.last_accelerations()
    for k in "xyz":
        assert np.array_equal(acc[k], want_acc[k]), k
    assert st["interactions"] == st1["interactions"]
    assert ms.replicas_identical()
    bodies, cells = ms.rank_counts()
    assert bodies.sum() == len(s)
    assert bodies.max() <= 1.1 * len(s) / world + 2048      # cuts follow the level-K histogram
    ms.close()


@needs2
def test_sharded_run_crosses_checkpoints_and_merged_units():
    """70 steps (two 32-step checkpoints), coincident bodies (merged units) and a fixed body."""
    s = gen.readme_pipeline(30_000, seed=8, spin=1000.0)
    s[100:140] = s[:40]                       # coincident pairs: merged leaves
    s["fixed"][7] = True
    dt, steps = 1e-5, 70
    want, _, _ = single("astro2", 1.0, 0.5, dt, s, steps)
    ms = api.MultiSim("astro2", theta=1.0, e=0.5, dt=dt, devices=[0, 1])
    ms.upload(s)
    ms.run(30)
    ms.download(s.copy())                     # mid-run read-out (velocities materialised)
    ms.run(40)
    got = ms.download(s.copy())
    for k in POS:
        assert np.array_equal(got[k], want[k]), k
    ms.close()


@needs2
def test_direct_sum_sharded_by_target_equals_single_gpu():
    s = gen.cube(40_000, seed=2)
    want, want_acc, _ = single("astro2", 0.0, 0.5, 1e-5, s, 3)
    ms = api.MultiSim("astro2", theta=0.0, e=0.5, dt=1e-5, devices=[0, 1])
    ms.upload(s)
    ms.run(3)
    got = ms.download(s.copy())
    for k in POS:
        assert np.array_equal(got[k], want[k]), k
    ms.close()


def test_one_rank_msim_equals_sim():
    """world == 1 through the multi-rank handle: no NCCL, same bits as pb200_sim_*."""
    s = gen.readme_pipeline(20_000, seed=3, spin=1000.0)
    want, _, _ = single("astro", 1.3, 1.0, 1e-5, s, 6)
    ms = api.MultiSim("astro", theta=1.3, e=1.0, dt=1e-5, devices=[0])
    ms.upload(s)
    ms.run(6)
    got = ms.download(s.copy())
    for k in POS:
        assert np.array_equal(got[k], want[k]), k
    ms.close()
