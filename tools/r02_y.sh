timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/shard_debug.py 16777216 2 2>&1 | grep -E "chunk 1|kernel *\[" | cut -c1-200 | tail -14
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 5 --skip-extras > gpurun_out/bench_r02y_c3_g2.json 2> gpurun_out/bench_r02y_c3_g2.err; echo bench2 rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r02y_c3_g2.json"))
print(d["config"]["n_bodies"], round(d["ms_per_step"],4), d["value"], [(k["kernel"],round(k["ms_per_step"]*1e3,1)) for k in d["roofline"]["kernels"][:6]])
PY
