set -x
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -3
for k in local exchange; do
PB200_SHARD_KEYS=$k timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 5 --skip-extras --workload c3 > gpurun_out/bench_r02r_c3_g2_$k.json 2> gpurun_out/bench_r02r_c3_g2_$k.err; echo bench2 $k rc=$?
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_rank_check.py 8000000 6 astro2 0.7 2>&1 | tail -1 | cut -c1-300
python - <<'PY'
import json
for f in ("local","exchange"):
    try:
        d=json.load(open(f"gpurun_out/bench_r02r_c3_g2_{f}.json"))
        print(f, d["config"]["n_bodies"], round(d["ms_per_step"],4), d["value"], d.get("sharding")["replays"], [(k["kernel"],round(k["ms_per_step"]*1e3,1)) for k in d["roofline"]["kernels"]])
    except Exception as e: print(f, "ERR", e)
PY
