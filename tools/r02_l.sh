set -x
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/shard_debug.py 16777216 3 > gpurun_out/sd_r02l_2.log 2>&1; echo sd2 rc=$?; grep -E "chunk" gpurun_out/sd_r02l_2.log | cut -c1-700 | tail -12
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -5
for w in c3 c3o; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 5 --skip-extras --workload $w > gpurun_out/bench_r02l_${w}_g2.json 2> gpurun_out/bench_r02l_${w}_g2.err; echo bench2 $w rc=$?
done
python - <<'PY'
import json
for f in ("c3_g2","c3o_g2"):
    try:
        d=json.load(open(f"gpurun_out/bench_r02l_{f}.json"))
        print(f, d["config"]["n_bodies"], round(d["ms_per_step"],4), d["value"], d.get("sharding"), [(k["kernel"],round(k["ms_per_step"]*1e3,1)) for k in d["roofline"]["kernels"]])
    except Exception as e: print(f, "ERR", e)
PY
