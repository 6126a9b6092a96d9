"""Per-kernel device times of a workload after it has evolved: python tools/late_profile.py c3o 300"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from physim_b200 import api

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3o"]
pre = int(sys.argv[2]) if len(sys.argv) > 2 else 300
sim = api.Sim(w["element"], theta=w["theta"], e=w["e"], dt=w["dt"])
sim.upload(bench.make_state(w))
for stage, k in (("start", 8), ("late", pre)):
    sim.run(k)
    ms = sim.run_timed(32) / 32
    sim.profile(True)
    sim.run(4)
    rep = sim.profile_report()
    sim.profile(False)
    st = sim.stats()
    print(stage, "%.4f ms/step" % ms, "interactions/target %.1f" % (st["interactions"] / st["n_bodies"]), "cells", st["n_cells"],
          [(r["kernel"], round(r["ms"] / r["launches"] * 1e3, 1)) for r in sorted(rep, key=lambda r: -r["ms"])], flush=True)
