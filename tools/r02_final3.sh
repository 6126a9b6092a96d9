# usage (under gpurun): bash tools/r02_final3.sh — cells_kernel at 5 CTAs/SM: parity subset, bit identity with the previous build, smoke, bench line
BASE=$PWD/tools/scratch/libbase.so
timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_verlet.py -m gpu -x -q > gpurun_out/pytest_final3.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/pytest_final3.log
timeout 60 python tools/ab_bits.py > gpurun_out/bits3_new.txt 2> gpurun_out/bits3_new.err; echo bits_new rc=$?
PB200_LIB_PATH=$BASE timeout 60 python tools/ab_bits.py > gpurun_out/bits3_base.txt 2> gpurun_out/bits3_base.err; echo bits_base rc=$?
diff gpurun_out/bits3_new.txt gpurun_out/bits3_base.txt > gpurun_out/bits3_diff.txt; echo "bits diff rc=$? lines=$(wc -l < gpurun_out/bits3_diff.txt) of $(wc -l < gpurun_out/bits3_new.txt)"
timeout 60 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
timeout 100 python bench.py --skip-extras > gpurun_out/bench_r02_final3_c3.json 2> gpurun_out/bench_r02_final3_c3.err; echo bench rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_final3_c3.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline'].get('avg_launch_ms'), d.get('roofline_step'))"
