set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02final.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/pytest_gpu_r02final.log
bash tools/r02_o.sh r02f
