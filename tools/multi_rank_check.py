"""torchrun --nproc-per-node G tools/multi_rank_check.py [n] [steps] [element] [theta]
One process per rank (NCCL across processes, CUDA IPC peer mappings): every rank runs the sharded simulation;
rank 0 also runs the single-GPU simulation and requires bit-equal positions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from physim_b200 import api, generators as gen

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
name = sys.argv[3] if len(sys.argv) > 3 else "astro"
theta = float(sys.argv[4]) if len(sys.argv) > 4 else 1.3
torch.cuda.set_device(local)
dist.init_process_group("gloo")
idt = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    idt = torch.from_numpy(api.comm_unique_id().copy())
dist.broadcast(idt, 0)
s = gen.readme_pipeline(n, seed=1, spin=1000.0)
ms = api.MultiSim(name, theta=theta, e=0.5, dt=1e-5, world=world, rank=rank, device=local, comm_id=idt.numpy())
ms.upload(s)
ms.run(3)
t = ms.run_timed(steps)
got = ms.download(s.copy())
st = ms.stats()
bodies, cells = ms.rank_counts()
ok = True
if rank == 0:
    sim = api.Sim(name, theta=theta, e=0.5, dt=1e-5, device=local)
    sim.upload(s)
    sim.run(3)
    t1 = sim.run_timed(steps)
    want = sim.download(s.copy())
    ok = all(np.array_equal(got[k], want[k]) for k in ("x", "y", "z", "vx", "vy", "vz"))
    print(f"world {world} n {len(s)} {name} theta {theta}: sharded {t / steps:.4f} ms/step, single {t1 / steps:.4f} ms/step, "
          f"bit-identical {ok}, bodies/rank {bodies}, stats {st}", flush=True)
flag = torch.tensor([1 if ok else 0])
dist.broadcast(flag, 0)
ms.close()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
