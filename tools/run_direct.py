"""Runs the tiled direct-sum kernel alone (for ncu / variant sweeps): N bodies, all targets."""
import sys, os, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from physim_b200 import api, generators as gen

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
clocks = []
stop = False
def sample():
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        while not stop:
            clocks.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
            time.sleep(0.01)
    except Exception as e:
        clocks.append(("err", str(e)))
sim = api.Sim("simple_astro", e=0.5, dt=1e-6)
sim.upload(gen.cube(n, seed=1))
ms = sim.run_timed(1)
th = threading.Thread(target=sample); th.start()
ms = sim.run_timed(reps)
stop = True; th.join()
mhz = sorted(c[0] for c in clocks if isinstance(c[0], int))
pw = [c[1] for c in clocks if isinstance(c[0], int)]
print("n", n, "ms/eval %.2f" % (ms / reps),
      "interactions/s %.4g" % (n * n * reps / (ms * 1e-3)), "sm_mhz median", mhz[len(mhz) // 2] if mhz else None,
      "power max", max(pw) if pw else None, "samples", len(mhz))
