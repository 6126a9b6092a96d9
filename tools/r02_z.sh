timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 5 --skip-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['alg_bytes_per_launch'], d['roofline_step'], d['e2e']['ms_per_step'], d['e2e']['d2h_bytes_per_step'])
print([(k['kernel'], round(k['ms_per_step']*1e3,1), None if k['gbs'] is None else round(k['gbs'])) for k in d['roofline']['kernels']])"
