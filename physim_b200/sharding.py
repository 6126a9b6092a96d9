"""Target sharding across ranks (one process per GPU).

The path shards by target body (astro/src/transformers.rs:138-159 is an independent loop over i):
rank r owns bodies [r*n/world, (r+1)*n/world) — forces and verlet state — and needs every body's
{x, y, z, m} each step, exchanged with one all-gather of the owned slices, in place on the buffer
the kernels read.  These helpers mirror `sim_upload` in csrc/engine.cu and wrap the collective.
"""


def slice_size(n, world):
    """S = ceil(n / world): every rank's slice of the gather buffer holds S bodies."""
    return (n + world - 1) // world


def owned_range(n, rank, world):
    """[t0, t1) of the bodies rank `rank` owns; identical to sim_upload in csrc/engine.cu:
    equal slices of S bodies, the last one shorter (possibly empty)."""
    s = slice_size(n, world)
    return min(n, s * rank), min(n, s * (rank + 1))


def gather_elems(n, world, width=4):
    """Elements of the gather buffer: world * S records (>= n: the tail is padding)."""
    return slice_size(n, world) * world * width


def slice_elems(n, rank, world, width=4):
    """(offset, count) in elements of the rank's slice of the gather buffer."""
    s = slice_size(n, world)
    return s * rank * width, s * width


def exchange(gathered, n, rank, world, width=4):
    """One all-gather of the equal-sized slices, in place on `gathered` (a flat torch tensor of
    gather_elems(n, world) elements whose owned slice holds this rank's new positions)."""
    import torch.distributed as dist

    if world == 1:
        return
    off, cnt = slice_elems(n, rank, world, width)
    mine = gathered[off:off + cnt]
    dist.all_gather_into_tensor(gathered, mine.clone() if gathered.device.type == "cpu" else mine)
