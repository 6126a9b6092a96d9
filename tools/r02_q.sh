set -x
for v in 0 1; do
for w in c3 c5s c3o; do
PB200_WALK_VARIANT=$v timeout 300 python bench.py --steps 60 --skip-extras --workload $w > gpurun_out/bench_r02q_${w}_v$v.json 2> gpurun_out/bench_r02q_${w}_v$v.err; echo bench $w v=$v rc=$?
done
done
python - <<'PY'
import json
for v in (0,1):
  for f in ("c3","c5s","c3o"):
    try:
        d=json.load(open(f"gpurun_out/bench_r02q_{f}_v{v}.json"))
        print(f, v, round(d["ms_per_step"],4), [(k["kernel"],round(k["ms_per_step"]*1e3,1)) for k in d["roofline"]["kernels"] if k["kernel"] in ("walk_kernel","cells_kernel")])
    except Exception as e: print(f, "ERR", e)
PY
