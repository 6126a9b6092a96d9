# round 2, first GPU call: baseline tests, c3o sort diagnosis, ncu of the direct kernel
set -x
date +%s > gpurun_out/t0_r02a
nvidia-smi -L
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02a.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu_r02a.log
PB200_DEBUG_CHECK=1 timeout 200 python tools/sort_diag.py c3o 2 > gpurun_out/sortdiag_c3o_r02a.log 2>&1; tail -30 gpurun_out/sortdiag_c3o_r02a.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:direct_kernel -s 1 -c 1 -o gpurun_out/prof_direct_r02a -f python tools/run_direct.py 262144 2 > gpurun_out/ncu_direct_r02a.log 2>&1; echo ncu rc=$?; tail -3 gpurun_out/ncu_direct_r02a.log
timeout 100 python tools/run_direct.py 1048576 3 > gpurun_out/direct_1M_r02a.log 2>&1; cat gpurun_out/direct_1M_r02a.log
echo elapsed $(( $(date +%s) - $(cat gpurun_out/t0_r02a) ))
