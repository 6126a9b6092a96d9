set -x
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu_final.log
timeout 400 python bench.py > gpurun_out/bench_r01b_c3.json 2> gpurun_out/bench_r01b_c3.err; echo bench rc=$?
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r01b_ref.json 2>/dev/null; echo ref rc=$?
for w in c1 c3o c5s; do timeout 200 python bench.py --workload $w --skip-extras --steps 50 > gpurun_out/bench_r01b_$w.json 2>/dev/null; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 51 -c 111 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 5 --warmup 3 --skip-extras > gpurun_out/ncu_r01b_list.log 2>&1; echo ncu1 rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -s 51 -c 11 -o gpurun_out/prof_r01b -f python bench.py --steps 5 --warmup 3 --skip-extras > gpurun_out/ncu_r01b_full.log 2>&1; echo ncu2 rc=$?
timeout 200 ncu --set full --clock-control none --import-source on -k regex:direct_kernel_x2 -c 1 -o gpurun_out/prof_r01b_direct -f python tools/run_direct.py 1048576 1 > gpurun_out/ncu_r01b_direct.log 2>&1; echo ncu3 rc=$?
python -c "import json;d=json.load(open('gpurun_out/bench_r01b_c3.json'));print(d['ms_per_step'],d['value'],d['e2e']['value'],d['direct_sum']['interactions_per_s'],d['roofline']['kernel'],d['roofline']['frac'],d['cpu_baseline']['value'])"
