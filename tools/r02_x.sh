timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 5 --skip-extras > gpurun_out/bench_r02x_c3_g2.json 2> gpurun_out/bench_r02x_c3_g2.err; echo bench2 rc=$?
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_rank_check.py 16000000 6 astro2 0.7 2>&1 | tail -1 | cut -c1-250
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r02x_c3_g2.json"))
print(d["config"]["n_bodies"], round(d["ms_per_step"],4), d["value"], d["sharding"]["replays"], [(k["kernel"],round(k["ms_per_step"]*1e3,1)) for k in d["roofline"]["kernels"]])
PY
