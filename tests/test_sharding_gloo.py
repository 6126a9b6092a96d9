"""N > 1 host logic on CPU: two gloo ranks own body slices, compute their accelerations (CPU oracle
standing in for the kernels), update their slice, exchange positions in place — and must reproduce
the single-rank reference loop exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import binding as ob
from physim_b200 import generators as gen
from physim_b200.sharding import exchange, gather_elems, owned_range, slice_elems, slice_size


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_owned_ranges_tile_the_bodies():
    for n in (0, 1, 7, 67, 100_002, 1 << 24):
        for world in (1, 2, 3, 4, 8):
            edges = [owned_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            s = slice_size(n, world)
            assert all(sz == s for sz in sizes[:-1] if sz) and max(sizes, default=0) <= s
            # the gather buffer has world equal slots; only the tail past n is padding
            assert gather_elems(n, world) == world * s * 4 >= n * 4
            assert slice_elems(n, 1 % world, world) == (s * (1 % world) * 4, s * 4)


def _worker(rank, world, port, n_bodies, steps, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    state = gen.solar() if n_bodies == 67 else gen.readme_pipeline(n_bodies - 2, seed=3)
    n = len(state)
    t0, t1 = owned_range(n, rank, world)
    dt, e = 0.01, 0.1
    gathered = torch.zeros(gather_elems(n, world), dtype=torch.float64)   # {x, y, z, m} + padding
    g = gathered.view(-1, 4).numpy()[:n]
    g[:, 0], g[:, 1], g[:, 2], g[:, 3] = state["x"], state["y"], state["z"], state["mass"]
    prev = np.zeros((n, 3))
    vel = np.stack([state["vx"], state["vy"], state["vz"]], 1)
    cur = state.copy()
    for step in range(steps):
        cur["x"], cur["y"], cur["z"] = g[:, 0], g[:, 1], g[:, 2]
        acc = ob.direct_range(cur, e, t0, t1)                      # forces for the owned targets
        a = np.stack([acc["x"], acc["y"], acc["z"]], 1)[t0:t1]
        x = g[t0:t1, :3].copy()
        if step == 0:
            new = x + vel[t0:t1] * dt + 0.5 * a * (dt * dt)
            vel[t0:t1] = vel[t0:t1] + a * dt
        else:
            new = 2.0 * x - prev[t0:t1] + a * (dt * dt)
            vel[t0:t1] = (new - x) / dt
        prev[t0:t1] = x
        g[t0:t1, :3] = new
        exchange(gathered, n, rank, world)
    if rank == 0:
        np.save(out, g.copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_bodies", [67, 302])
def test_two_ranks_reproduce_the_single_rank_loop(tmp_path, n_bodies):
    steps = 5
    out = str(tmp_path / "g.npy")
    mp.spawn(_worker, args=(2, free_port(), n_bodies, steps, out), nprocs=2, join=True)
    got = np.load(out)
    state = gen.solar() if n_bodies == 67 else gen.readme_pipeline(n_bodies - 2, seed=3)
    ref, _ = ob.run_pipeline("simple_astro", state, 1.0, 0.1, 0.01, steps)
    for k, name in enumerate("xyz"):
        np.testing.assert_allclose(got[:, k], ref[name], rtol=1e-13, atol=1e-15)
