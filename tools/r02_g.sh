set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02g.log 2>&1; echo pytest rc=$?; tail -12 gpurun_out/pytest_gpu_r02g.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_rank_check.py 200000 40 astro2 0.7 > gpurun_out/mrc_r02g.log 2>&1; echo mrc rc=$?; tail -2 gpurun_out/mrc_r02g.log | cut -c1-600
for w in c3 c3o; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 5 --skip-extras --workload $w > gpurun_out/bench_r02g_${w}_g2.json 2> gpurun_out/bench_r02g_${w}_g2.err; echo bench2 $w rc=$?
done
timeout 200 python bench.py --steps 100 --skip-extras > gpurun_out/bench_r02g_c3.json 2> gpurun_out/bench_r02g_c3.err
python - <<'PY'
import json
for f in ("c3_g2","c3o_g2","c3"):
    try:
        d=json.load(open(f"gpurun_out/bench_r02g_{f}.json"))
        print(f, round(d["ms_per_step"],4), d.get("sharding"), [(k["kernel"],k["launches_per_step"],round(k["ms_per_step"]*1e3,1)) for k in d["roofline"]["kernels"]])
    except Exception as e: print(f, "ERR", e)
PY
