set -x
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_gpu_multi_r02_8gpu.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/pytest_gpu_multi_r02_8gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/bench_r02_8gpu_c3_g8.json 2> gpurun_out/bench_r02_8gpu_c3_g8.err; echo bench8 rc=$?; tail -3 gpurun_out/bench_r02_8gpu_c3_g8.err | cut -c1-400
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r02_8gpu_c3_g8.json"))
print(d["n_gpus"], d["config"]["n_bodies"], round(d["ms_per_step"],4), d["value"], d.get("sharding"))
print([(k["kernel"],k["launches_per_step"],round(k["ms_per_step"]*1e3,1)) for k in d["roofline"]["kernels"]])
print("direct", d.get("direct_sum"))
print("bh_large", d.get("bh_large"))
PY
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r02_8gpu_c3_g8.json"))
print("bh kernels", d.get("bh_large",{}).get("kernels_ms_per_step"))
PY
