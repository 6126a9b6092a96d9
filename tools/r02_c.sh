set -x
nvidia-smi -L; nvidia-smi topo -m 2>/dev/null | head -12
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi_r02c.log 2>&1; echo pytest rc=$?; tail -25 gpurun_out/pytest_multi_r02c.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_rank_check.py 200000 40 astro 1.3 > gpurun_out/mrc_r02c.log 2>&1; echo mrc rc=$?; tail -8 gpurun_out/mrc_r02c.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/multi_rank_check.py 1000000 64 astro 1.3 > gpurun_out/mrc2_r02c.log 2>&1; echo mrc2 rc=$?; tail -4 gpurun_out/mrc2_r02c.log
PB200_LIB_PATH=$PWD/tools/scratch/libdebug_skew.so timeout 200 python tools/sort_diag.py c3o 32 > gpurun_out/skew_r02c.log 2>&1; grep -m 20 skew gpurun_out/skew_r02c.log; tail -3 gpurun_out/skew_r02c.log
