import sys, os
sys.path.insert(0, "/root/repo")
import bench
from physim_b200 import api
w = bench.WORKLOADS[sys.argv[1]]
state = bench.make_state(w)
sim = api.Sim(w["element"], theta=w["theta"], e=w["e"], dt=w["dt"])
sim.upload(state)
for i in range(12):
    sim.run(int(sys.argv[2]) if len(sys.argv) > 2 else 2)
    st = sim.stats()
    print(i, "mode", st["sort_mode"], "max_bucket", st["max_bucket"], "extent %.4f" % st["extent"], "cells", st["n_cells"], "replays", st["replays"], "bits", st["sort_bits"])
