set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02h.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/pytest_gpu_r02h.log
timeout 900 python bench.py --steps 100 > gpurun_out/bench_r02h_c3.json 2> gpurun_out/bench_r02h_c3.err; echo bench rc=$?; tail -5 gpurun_out/bench_r02h_c3.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r02h_c3.json"))
print(round(d["ms_per_step"],4), [(k["kernel"],round(k["ms_per_step"]*1e3,1)) for k in d["roofline"]["kernels"]])
print("e2e", d["e2e"]["ms_per_step"], "resident", d["e2e"]["resident"]["ms_per_step"], "dropin", d["e2e"]["dropin"]["ms_per_step"])
print("direct", d["direct_sum"]["interactions_per_s"], d["direct_sum"]["roofline"]["frac"])
print("cpu", d["cpu_baseline"]["value"])
print("other", json.dumps(d.get("other_workloads"))[:1500])
print("bh_large", d.get("bh_large"))
PY
