"""Bit-identity check between two builds of the library (run under gpurun):

    python tools/ab_bits.py > a.txt;  PB200_LIB_PATH=tools/scratch/libbase.so python tools/ab_bits.py > b.txt;  diff a.txt b.txt

Prints one sha256 per (case, array): the inspected cell table of a single evaluation (keys, permutation,
cell_start, level, head, count, skip, parent, centres, centres of mass, per-target interaction counts), its
accelerations, and the state after a few resident verlet steps.  A kernel change that claims "same operations
in the same order" must leave every line unchanged."""
import hashlib
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from physim_b200 import api, generators as gen  # noqa: E402


def h(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def cases():
    yield "quad_200k", "astro", dict(theta=1.3, e=1.0), gen.headline_pipeline(200_000, seed=3)
    yield "oct_150k", "astro2", dict(theta=0.7, e=0.5), gen.readme_pipeline(150_000, seed=5, spin=1000.0)
    yield "quad_1M", "astro", dict(theta=1.3, e=1.0), gen.headline_pipeline(1_000_000, seed=1)
    # merged units and a crowd (pseudo levels, > 2^DIM children, slow summation paths)
    s = gen.cube(30_000, seed=2)
    jit = np.random.default_rng(7).random((1500, 3)) * 1e-5
    s["x"][:1500], s["y"][:1500], s["z"][:1500] = 0.3 + jit[:, 0], -0.2 + jit[:, 1], 0.6 + jit[:, 2]
    s["x"][2000:2100], s["y"][2000:2100], s["z"][2000:2100] = 0.1, 0.1, 0.1
    yield "crowd_30k", "astro2", dict(theta=1.0, e=0.5), s
    yield "tiny_7", "astro2", dict(theta=1.5, e=0.5), gen.cube(7, seed=9)


def main():
    for name, kind, prm, s in cases():
        el = api.TransformElement(kind, **prm)
        acc = el.transform(s)
        t = el.debug_tree()
        for k in ("key", "perm", "cell_start", "level", "head", "count", "skip", "parent", "centre_ext",
                  "com_mass", "counts"):
            print(name, k, h(t[k]))
        print(name, "acc", h(np.stack([acc[k] for k in ("x", "y", "z")], 1)))
        st = el.stats()
        print(name, "n_cells", st["n_cells"], "sort_mode", st["sort_mode"])
        sim = api.Sim(kind, dt=1e-5, **prm)
        sim.upload(s)
        sim.run(12)
        out = sim.download(s.copy())
        print(name, "state12", h(np.stack([out[k] for k in ("x", "y", "z", "vx", "vy", "vz")], 1)))
        a12 = sim.last_accelerations()
        print(name, "acc12", h(np.stack([a12[k] for k in ("x", "y", "z")], 1)))


if __name__ == "__main__":
    main()
