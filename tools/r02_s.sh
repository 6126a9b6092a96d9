set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
for lib in main w6; do
for w in c3 c5s; do
L=physim_b200/libphysim_b200.so; [ $lib = w6 ] && L=physim_b200/libphysim_b200_w6.so
PB200_LIB_PATH=$PWD/$L timeout 300 python bench.py --steps 60 --skip-extras --workload $w > gpurun_out/bench_r02s_${w}_$lib.json 2> gpurun_out/bench_r02s_${w}_$lib.err; echo bench $w $lib rc=$?
done
done
python - <<'PY'
import json
for v in ("main","w6"):
  for f in ("c3","c5s"):
    try:
        d=json.load(open(f"gpurun_out/bench_r02s_{f}_{v}.json"))
        print(f, v, round(d["ms_per_step"],4), [(k["kernel"],round(k["ms_per_step"]*1e3,1)) for k in d["roofline"]["kernels"] if k["kernel"] in ("walk_kernel","cells_kernel")])
    except Exception as e: print(f, "ERR", e)
PY
