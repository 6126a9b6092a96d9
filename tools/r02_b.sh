set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02b.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu_r02b.log
PB200_DEBUG_CHECK=1 timeout 200 python bench.py --workload c3o --steps 50 --skip-extras > gpurun_out/bench_r02b_c3o.json 2> gpurun_out/bench_r02b_c3o.err; tail -12 gpurun_out/bench_r02b_c3o.err
timeout 200 python bench.py --steps 100 --skip-extras > gpurun_out/bench_r02b_c3.json 2> gpurun_out/bench_r02b_c3.err
python - <<'PY'
import json
for w in ("c3o","c3"):
    d=json.load(open(f"gpurun_out/bench_r02b_{w}.json"))
    print(w, d["ms_per_step"], [(k["kernel"],round(k["ms_per_step"]*1e3,1)) for k in d["roofline"]["kernels"]])
PY
