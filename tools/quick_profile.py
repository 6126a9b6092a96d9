"""Per-kernel device times of a few device-resident steps (debug aid): workload, steps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from physim_b200 import api

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
state = bench.make_state(w)
sim = api.Sim(w["element"], theta=w["theta"], e=w["e"], dt=w["dt"])
sim.upload(state)
sim.run(2)
sim.profile(True)
ms = sim.run_timed(steps)
rep = sim.profile_report()
print("ms/step %.4f" % (ms / steps))
for k in sorted(rep, key=lambda k: -k["ms"]):
    print("  %-22s %6.1f launches/step  %9.4f ms/step" % (k["kernel"], k["launches"] / steps, k["ms"] / steps))
