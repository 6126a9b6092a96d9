set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_verlet.py tests/test_gpu_integrators.py tests/test_gpu_host_harness.py -m gpu -q -x > gpurun_out/pytest_gpu_r02p.log 2>&1; echo pytest rc=$?; tail -6 gpurun_out/pytest_gpu_r02p.log
for w in c3 c5s c1; do
timeout 300 python bench.py --steps 100 --skip-extras --workload $w > gpurun_out/bench_r02p_${w}.json 2> gpurun_out/bench_r02p_${w}.err; echo bench $w rc=$?
done
timeout 200 python tools/e2e_only.py > gpurun_out/e2e_r02p.log 2>&1; tail -5 gpurun_out/e2e_r02p.log | cut -c1-1500
python - <<'PY'
import json
for f in ("c3","c5s","c1"):
    try:
        d=json.load(open(f"gpurun_out/bench_r02p_{f}.json"))
        print(f, round(d["ms_per_step"],4), [(k["kernel"],round(k["ms_per_step"]*1e3,1)) for k in d["roofline"]["kernels"]])
    except Exception as e: print(f, "ERR", e)
PY
