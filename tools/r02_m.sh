set -x
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r02m.log 2>&1; echo pytest rc=$?; tail -6 gpurun_out/pytest_gpu_r02m.log
for pf in 1 0; do
for w in c3 c5s; do
PB200_WALK_PF=$pf timeout 300 python bench.py --steps 100 --skip-extras --workload $w > gpurun_out/bench_r02m_${w}_pf$pf.json 2> gpurun_out/bench_r02m_${w}_pf$pf.err; echo bench $w pf=$pf rc=$?
done
done
timeout 200 python tools/steps_diag.py c3o > gpurun_out/steps_diag_c3o_r02m.log 2>&1; cat gpurun_out/steps_diag_c3o_r02m.log | cut -c1-300
python - <<'PY'
import json
for f in ("c3_pf1","c3_pf0","c5s_pf1","c5s_pf0"):
    try:
        d=json.load(open(f"gpurun_out/bench_r02m_{f}.json"))
        print(f, round(d["ms_per_step"],4), [(k["kernel"],round(k["ms_per_step"]*1e3,1)) for k in d["roofline"]["kernels"]])
    except Exception as e: print(f, "ERR", e)
PY
