# usage (under gpurun): bash tools/quick_ab.sh TAG "ENV1=..;ENV2=.." [workloads]  — parity tests on the default build, then short benches per env setting
T=${1:-ab}
W=${3:-"c3 c1"}
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_$T.log
IFS=';' read -ra SETS <<< "$2"
for e in "default" "${SETS[@]}"; do
  for w in $W; do
    if [ "$e" = default ]; then envs=""; else envs="$e"; fi
    echo "== $e $w"
    env $envs timeout 200 python bench.py --workload $w --skip-extras --steps 100 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['ms_per_step'], [(k['kernel'],round(k['ms_per_step']*1e3,1)) for k in d['roofline']['kernels']])"
  done
done
if [ -n "$NCU_K" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$NCU_K" -s ${NCU_S:-12} -c ${NCU_C:-4} -o gpurun_out/prof_$T -f python bench.py --steps 5 --warmup 3 --skip-extras > gpurun_out/ncu_$T.log 2>&1; echo ncu rc=$?
fi
