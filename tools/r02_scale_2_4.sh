set -x
for g in 2 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2952$g bench.py --gpus $g --steps 100 --warmup 5 > gpurun_out/bench_r02v_c3_g$g.json 2> gpurun_out/bench_r02v_c3_g$g.err; echo bench$g rc=$?; tail -2 gpurun_out/bench_r02v_c3_g$g.err | cut -c1-300
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29529 bench.py --impl reference --gpus 4 --steps 3 --warmup 1 > gpurun_out/bench_r02v_ref_g4.json 2> gpurun_out/bench_r02v_ref_g4.err; echo ref4 rc=$?; cut -c1-600 gpurun_out/bench_r02v_ref_g4.json
python - <<'PY'
import json
for g in (2,4):
    try:
        d=json.load(open(f"gpurun_out/bench_r02v_c3_g{g}.json"))
        print(g, d["config"]["n_bodies"], d["scaling"], round(d["ms_per_step"],4), d["value"], d.get("sharding"))
        print("   direct", d.get("direct_sum",{}).get("interactions_per_s"), d.get("direct_sum",{}).get("sample"), "bh", d.get("bh_large"), "e2e", d.get("e2e"))
    except Exception as e: print(g, "ERR", e)
PY
