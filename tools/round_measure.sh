# usage: bash tools/round_measure.sh TAG [full]   (run under gpurun; writes gpurun_out/*_TAG.*)
T=${1:-r01c}
set -x
date +%s > gpurun_out/t0_$T
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$T.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu_$T.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; echo smoke rc=$?
timeout 400 python bench.py > gpurun_out/bench_${T}_c3.json 2> gpurun_out/bench_${T}_c3.err; echo bench rc=$?
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${T}_ref.json 2>/dev/null; echo ref rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 108 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 5 --warmup 3 --skip-extras > gpurun_out/ncu_${T}_list.log 2>&1; echo ncu1 rc=$?
if [ "$2" = full ]; then
for w in c1 c3o c5s; do timeout 200 python bench.py --workload $w --skip-extras --steps 50 > gpurun_out/bench_${T}_$w.json 2>/dev/null; done
timeout 150 python tools/stress.py 16777216 3 > gpurun_out/stress_${T}.log 2>&1; echo stress rc=$?; tail -3 gpurun_out/stress_${T}.log
timeout 400 ncu --set full --clock-control none --import-source on -s 60 -c 9 -o gpurun_out/prof_$T -f python bench.py --steps 5 --warmup 3 --skip-extras > gpurun_out/ncu_${T}_full.log 2>&1; echo ncu2 rc=$?
fi
python -c "import json;d=json.load(open('gpurun_out/bench_${T}_c3.json'));print(d['ms_per_step'],d['value'],d['e2e']['value'],d['direct_sum']['interactions_per_s'],d['roofline']['kernel'],d['roofline']['frac'],d['cpu_baseline']['value'])"
echo elapsed $(( $(date +%s) - $(cat gpurun_out/t0_$T) ))
