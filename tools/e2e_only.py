"""The e2e leg of bench.py alone (host Entity[n] -> host Entity[n] through the C ABI), for host-pipeline A/B runs.

  PB200_CHUNKS=16 PB200_HOST_PIPE=perchunk PB200_HOST_THREADS=16 python tools/e2e_only.py [workload] [steps]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from physim_b200 import api  # noqa: E402

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
args = argparse.Namespace(steps=int(sys.argv[2]) if len(sys.argv) > 2 else 50, warmup=3)
r = bench.measure_e2e(api, bench.make_state(w), w, args)
print(json.dumps({k: r[k] for k in ("ms_per_step", "device_ms", "host_ms")}))
print("resident", r["resident"]["ms_per_step"], "dropin", r["dropin"]["ms_per_step"], "transform_only", r["transform_only"]["ms_per_call"])
