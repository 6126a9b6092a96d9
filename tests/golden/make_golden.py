"""Generates tests/golden/*.npz from the CPU oracle (oracle/physim_oracle.cpp).

The reference ships no golden vectors for this path and cannot be built offline, so
these fixtures freeze the ORACLE's outputs on small seeded inputs: they guard the oracle against
drift between rounds and give the GPU tests fixed files to compare with.  Regenerate only when the
oracle is deliberately changed:   python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402
from physim_b200 import generators as gen  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def vec(a):
    return np.stack([a["x"], a["y"], a["z"]], 1)


def main():
    # 1. example_pipelines/solar.toml (BASELINE config 2): simple_astro e=0.1, verlet dt=0.01
    s = gen.solar(planets=8, asteroids=50, seed=2)
    acc = ob.transform("simple_astro", s, e=0.1)
    final, _ = ob.run_pipeline("simple_astro", s, 1.0, 0.1, 0.01, 10)
    np.savez_compressed(os.path.join(HERE, "solar.npz"), state=s.view(np.uint8), acc=vec(acc),
                        final=final.view(np.uint8))
    # 2. README pipeline in miniature: cube 510 + 2 stars, both trees, several thetas
    s = gen.readme_pipeline(510, seed=1, spin=1000.0)
    out = {"state": s.view(np.uint8)}
    for name, dim in (("astro2", 3), ("astro", 2)):
        tab = ob.CellTable(dim, s)
        for k in ("key", "perm", "cell_start", "level", "head", "count", "skip", "parent", "centre_ext", "com_mass"):
            out[f"{name}_{k}"] = getattr(tab, k)
        for theta in (0.5, 1.0, 1.5):
            a, c = ob.transform(name, s, theta, 0.5, counts=True)
            out[f"{name}_acc_{theta}"] = vec(a)
            out[f"{name}_cnt_{theta}"] = c
        fin, _ = ob.run_pipeline(name, s, 1.5, 0.5, 1e-5, 5)
        out[f"{name}_final"] = fin.view(np.uint8)
    np.savez_compressed(os.path.join(HERE, "readme_small.npz"), **out)
    print("written", [f for f in os.listdir(HERE) if f.endswith(".npz")])


if __name__ == "__main__":
    main()
