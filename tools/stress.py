"""Large-N robustness / memory-sizing run (BASELINE configs[4] scaled to one GPU): N bodies,
astro2 theta=0.7 e=0.5, a few device-resident steps; prints the tree size, interactions and timing
and checks size-independent properties of the tree."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from physim_b200 import api, generators as gen

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
t0 = time.time()
state = gen.cube(n, seed=1)
print("generated", n, "bodies in %.1f s" % (time.time() - t0), flush=True)
sim = api.Sim("astro2", theta=0.7, e=0.5, dt=1e-6)
sim.upload(state)
ms = sim.run_timed(1)
ms = sim.run_timed(steps)
st = sim.stats()
print("ms/step %.3f  particle-steps/s %.4g  cells %d (%.3f/body)  interactions/target %.1f  launches %d"
      % (ms / steps, n * steps / (ms * 1e-3), st["n_cells"], st["n_cells"] / n, st["interactions"] / n,
         st["kernel_launches"]), flush=True)
assert 1.2 * n < st["n_cells"] < 2.0 * n
out = sim.download(state.copy())
d = np.abs(out["x"] - state["x"]).max()
assert np.isfinite(out["x"]).all() and np.isfinite(out["vx"]).all() and 0 < d < 1e-3, d
# momentum: sum m a ~ 0 (monopole forces are not pairwise antisymmetric, so only loosely)
acc = sim.last_accelerations()
p = (state["mass"] * acc["x"]).sum()
scale = (state["mass"] * np.abs(acc["x"])).sum()
print("sum m a_x / sum m |a_x| = %.2e" % (p / scale))
assert abs(p / scale) < 5e-2
print("stress ok")
