set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02d.log 2>&1; echo pytest rc=$?; tail -12 gpurun_out/pytest_gpu_r02d.log
for w in c3 c3o; do
PB200_DEBUG_CHECK=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 5 --skip-extras --workload $w > gpurun_out/bench_r02d_${w}_g2.json 2> gpurun_out/bench_r02d_${w}_g2.err; echo bench2 $w rc=$?; grep -c "check:" gpurun_out/bench_r02d_${w}_g2.err; grep "check:" gpurun_out/bench_r02d_${w}_g2.err | grep -v "overflow 0 short 0 sort_error 0 bucket_overflow 0" | head -5
done
for w in c3o c5s c1; do timeout 200 python bench.py --workload $w --steps 50 --skip-extras > gpurun_out/bench_r02d_$w.json 2> gpurun_out/bench_r02d_$w.err; done
python - <<'PY'
import json
for f in ("c3_g2","c3o_g2","c3o","c5s","c1"):
    try:
        d=json.load(open(f"gpurun_out/bench_r02d_{f}.json"))
        print(f, round(d["ms_per_step"],4), d.get("sharding"), [(k["kernel"],k["launches_per_step"],round(k["ms_per_step"]*1e3,1)) for k in d["roofline"]["kernels"]])
    except Exception as e: print(f, "ERR", e)
PY
