"""Which kernels of the step gain from programmatic dependent launch (run under gpurun).

Every kernel of the chain is launched with the programmatic-serialization attribute; PB200_PDL_OFF=name[,name]
launches the named ones plainly.  Measures c3 with each kernel switched off in turn (and PB200_CELLS_BLOCKS=5),
combines whatever beat the all-on step by more than the run-to-run noise, and runs that combination on the other
workloads next to the default."""
import json
import os
import subprocess
import sys

KERNELS = ["encode_bucket_kernel", "sort_local_kernel", "unit_kernel", "scan_lookback_kernel", "cells_kernel",
           "kids_kernel", "climb_kernel", "walk_kernel", "verlet_lean_kernel"]


def run(workload, env):
    e = dict(os.environ)
    e.update(env)
    out = subprocess.run([sys.executable, "bench.py", "--workload", workload, "--skip-extras", "--steps", "96"],
                         env=e, capture_output=True, text=True, timeout=200).stdout
    d = json.loads(out.strip().splitlines()[-1])
    ks = {k["kernel"]: round(k["ms_per_step"] * 1e3, 1) for k in d["roofline"]["kernels"]}
    print("==", workload, env, round(d["ms_per_step"] * 1e3, 1), ks, flush=True)
    return d["ms_per_step"] * 1e3


def main():
    base = run("c3", {})
    base = min(base, run("c3", {}))
    gains = {}
    for k in KERNELS:
        gains[k] = base - run("c3", {"PB200_PDL_OFF": k})
    b5 = base - run("c3", {"PB200_CELLS_BLOCKS": "5"})
    off = [k for k in KERNELS if gains[k] > 1.5]
    env = {}
    if off:
        env["PB200_PDL_OFF"] = ",".join(off)
    if b5 > 1.0:
        env["PB200_CELLS_BLOCKS"] = "5"
    print("gains (us):", {k: round(v, 1) for k, v in gains.items()}, "cells_blocks=5:", round(b5, 1), "->", env, flush=True)
    if env:
        run("c3", env)
        for w in ("c1", "c3o", "c5s"):
            run(w, {})
            run(w, env)


if __name__ == "__main__":
    main()
