set -x
export PB200_DEBUG_CHECK=1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/shard_debug.py 16777216 3 > gpurun_out/sd_r02k_2.log 2>&1; echo sd2 rc=$?; grep -E "chunk|check:" gpurun_out/sd_r02k_2.log | cut -c1-900 | tail -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 tools/shard_debug.py 33554432 3 > gpurun_out/sd_r02k_4.log 2>&1; echo sd4 rc=$?; grep -E "chunk|check:" gpurun_out/sd_r02k_4.log | cut -c1-900 | tail -24
