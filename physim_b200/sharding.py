"""Target sharding across ranks (one process per GPU).

The path shards by target body (astro/src/transformers.rs:138-159 is an independent loop over i):
rank r owns bodies [r*n/world, (r+1)*n/world) — forces and verlet state — and needs every body's
{x, y, z, m} each step, exchanged with one all-gather of the owned slices, in place on the buffer
the kernels read.  These helpers mirror `sim_upload` in csrc/engine.cu and wrap the collective.
"""


def owned_range(n, rank, world):
    """[t0, t1) of the bodies rank `rank` owns; identical to engine.cu (n*rank/world)."""
    return n * rank // world, n * (rank + 1) // world


def slice_elems(n, rank, world, width=4):
    """(offset, count) in elements of the rank's slice of an n x width buffer."""
    t0, t1 = owned_range(n, rank, world)
    return t0 * width, (t1 - t0) * width


def exchange(gathered, n, rank, world, width=4):
    """All-gather of the owned slices, in place on `gathered` (a flat torch tensor of n*width
    elements whose owned slice holds this rank's new positions).  Uneven slices (n % world != 0)
    use all_gather with per-rank views."""
    import torch.distributed as dist

    if world == 1:
        return
    if n % world == 0:
        off, cnt = slice_elems(n, rank, world, width)
        dist.all_gather_into_tensor(gathered, gathered[off:off + cnt].clone() if gathered.device.type == "cpu"
                                    else gathered[off:off + cnt])
    else:
        views = []
        for r in range(world):
            o, c = slice_elems(n, r, world, width)
            views.append(gathered[o:o + c])
        off, cnt = slice_elems(n, rank, world, width)
        # ranks own different counts: broadcast each slice from its owner
        for r in range(world):
            dist.broadcast(views[r], src=r)
