set -x
for n in 4000000 8000000 16000000; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_rank_check.py $n 6 astro2 0.7 > gpurun_out/mrc_r02j_$n.log 2>&1; echo mrc $n rc=$?; tail -2 gpurun_out/mrc_r02j_$n.log | cut -c1-700
done
