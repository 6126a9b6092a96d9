"""Device-resident c3 steps in consecutive chunks: ms/step and the loop's replay / sort statistics per chunk."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from physim_b200 import api

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
sim = api.Sim(w["element"], theta=w["theta"], e=w["e"], dt=w["dt"])
sim.upload(bench.make_state(w))
sim.run_timed(3)
for chunk in (32, 32, 32, 32, 64, 64, 64, 64):
    ms = sim.run_timed(chunk)
    st = sim.stats()
    print(chunk, "steps: %.4f ms/step" % (ms / chunk), {k: st[k] for k in ("replays", "sort_bits", "sort_mode", "max_bucket", "n_cells", "extent") if k in st}, flush=True)
