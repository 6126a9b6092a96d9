"""N > 1 host logic on CPU (gloo, world_size 2 and 3): every rank publishes the records of the level-K key
prefixes in its range, the records are all-gathered, and every rank rebuilds the cells above level K and the next
cuts from them - the exchange of csrc/multi.cu's sharded step, with tests/shard_model.py standing in for
top_export_kernel / top_build_kernel and the oracle's cell table for the rank's tree.  Checked: all ranks end
with identical top trees and cuts; the rebuilt cells equal the oracle's own cells above level K (existence,
leaf-ness, body counts, centres of mass); the cuts balance the bodies."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import binding as ob
from physim_b200 import generators as gen
from tests import shard_model as sm


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def make_state(case):
    if case == "cube":
        return gen.readme_pipeline(6000, seed=3, spin=1000.0)
    if case == "sparse":                      # leaves above level K, empty prefixes, a merged pair
        s = gen.cube(40, seed=5)
        s[1] = s[0]
        return s
    s = gen.cube(3000, seed=9)                # clustered: most bodies in a few prefixes
    s["x"][:2500] *= 0.01
    s["y"][:2500] *= 0.01
    s["z"][:2500] *= 0.01
    return s


def _worker(rank, world, port, dim, case, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = make_state(case)
    tab = ob.CellTable(dim, s)                # every rank sees every body (replicated state)
    cuts = sm.first_cuts(dim, tab.key, world)
    mine = torch.from_numpy(sm.export_slots(dim, tab, cuts[rank], cuts[rank + 1]))
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)           # collective 1 of the sharded step
    units, count, com = sm.rebuild_top(dim, [g.numpy() for g in gathered], tab.extent)
    nxt = sm.next_cuts(dim, count, world, len(s))
    # every rank must hold the same top tree and the same cuts: compare with rank 0's
    blob = torch.from_numpy(np.concatenate([units.astype(np.float64), count.astype(np.float64), com.ravel(),
                                            nxt.astype(np.float64)]))
    ref = blob.clone()
    dist.broadcast(ref, 0)
    same = bool(torch.equal(ref, blob))
    if rank == 0:
        np.savez(out, units=units, count=count, com=com, cuts=cuts, nxt=nxt, same=same)
    flags = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(flags, torch.tensor([1.0 if same else 0.0]))
    assert all(f.item() == 1.0 for f in flags)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,dim,case", [(2, 2, "cube"), (2, 3, "cube"), (2, 3, "sparse"), (2, 2, "sparse"),
                                            (2, 3, "clustered"), (3, 2, "clustered"), (3, 3, "cube")])
def test_ranks_rebuild_the_oracles_top_tree(world, dim, case, tmp_path):
    out = str(tmp_path / "top.npz")
    mp.spawn(_worker, args=(world, free_port(), dim, case, out), nprocs=world, join=True)
    z = np.load(out)
    assert z["same"]
    s = make_state(case)
    tab = ob.CellTable(dim, s)
    k, shift = sm.K[dim], np.uint64(dim * (sm.LM[dim] - sm.K[dim]))
    units, count, com = z["units"], z["count"], z["com"]
    # every oracle cell at a level <= K appears with the same body count, leaf-ness and centre of mass
    seen = np.zeros(len(units), dtype=bool)
    for c in np.nonzero(tab.level <= k)[0]:
        l = int(tab.level[c])
        p = int(tab.key[tab.head[c]] >> np.uint64(dim * (sm.LM[dim] - l))) if l else 0
        t = sm.offset(dim, l) + p
        seen[t] = True
        leaf = tab.skip[c] == c + 1
        assert count[t] == tab.count[c], (c, l)
        assert (units[t] == 1) == leaf, (c, l)
        assert np.allclose(com[t], tab.com_mass[c], rtol=1e-12, atol=1e-12 * tab.extent), (c, l)
    # ... and nothing else does, except below a leaf (a unit's record is carried down its prefix path)
    for t in np.nonzero((units != 0) & ~seen)[0]:
        assert units[t] == 1
    # the cuts partition the bodies evenly at level-K granularity (clustered sets: as evenly as a prefix allows)
    nxt = z["nxt"]
    sizes = [int(((tab.key >= nxt[r]) & (tab.key < nxt[r + 1])).sum()) for r in range(world)]
    assert sum(sizes) == len(s)
    biggest_prefix = int(count[sm.offset(dim, k):].max())
    assert max(sizes) <= len(s) / world + biggest_prefix
