# usage (under gpurun): bash tools/r02_o.sh TAG   - the round's record: smoke, bench, reference arm, ncu launch list, ncu full set
T=${1:-r02}
set -x
date +%s > gpurun_out/t0_$T
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; echo smoke rc=$?; tail -1 gpurun_out/smoke_$T.log
timeout 900 python bench.py > gpurun_out/bench_${T}_c3.json 2> gpurun_out/bench_${T}_c3.err; echo bench rc=$?
timeout 300 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_${T}_ref.json 2>/dev/null; echo ref rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 108 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 5 --warmup 3 --skip-extras > gpurun_out/ncu_${T}_list.log 2>&1; echo ncu1 rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -s 60 -c 9 -o gpurun_out/prof_$T -f python bench.py --steps 5 --warmup 3 --skip-extras > gpurun_out/ncu_${T}_full.log 2>&1; echo ncu2 rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 4 -c 1 -o gpurun_out/prof_${T}_walk_c5s -f python bench.py --workload c5s --steps 3 --warmup 3 --skip-extras > gpurun_out/ncu_${T}_walk.log 2>&1; echo ncu3 rc=$?
python -c "import json;d=json.load(open('gpurun_out/bench_${T}_c3.json'));print(d['ms_per_step'],d['value'],d['e2e']['value'],d['direct_sum']['interactions_per_s'],d['roofline']['kernel'],d['roofline']['frac'],d['cpu_baseline']['value'], d['bh_large'])"
echo elapsed $(( $(date +%s) - $(cat gpurun_out/t0_$T) ))
