"""torchrun --nproc-per-node G tools/shard_debug.py n [chunks] [element] [theta]: device-made cube, sharded run in
chunks of 3 steps, per chunk: ms/step, stats, bodies per rank, the cuts (rank 0 prints)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from physim_b200 import api

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
n = int(sys.argv[1])
chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 3
name = sys.argv[3] if len(sys.argv) > 3 else "astro2"
theta = float(sys.argv[4]) if len(sys.argv) > 4 else 0.7
torch.cuda.set_device(local)
dist.init_process_group("gloo")
idt = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    idt = torch.from_numpy(api.comm_unique_id().copy())
dist.broadcast(idt, 0)
ms = api.MultiSim(name, theta=theta, e=0.5, dt=1e-6, world=world, rank=rank, device=local, comm_id=idt.numpy())
ms.generate_cube(n, seed=1)
for c in range(chunks):
    t = ms.run_timed(3)
    st = ms.stats()
    bodies, cells = ms.rank_counts()
    dbg = ms.debug_shard()
    if rank == 0:
        print(f"chunk {c}: {t / 3:.3f} ms/step sharded {st['sharded_steps']} replicated {st['replicated_steps']} replays {st['replays']} "
              f"cells {st['n_cells']} inter/target {st['interactions'] / n:.1f} bodies/rank {bodies} dbg {dbg}", flush=True)
# per-kernel device times of EVERY rank (event pairs around each launch), 3 profiled steps
ms.profile(True)
ms.run(3)
rep = {k["kernel"]: round(k["ms"] / 3, 3) for k in ms.profile_report()}
ms.profile(False)
allrep = [None] * world
dist.all_gather_object(allrep, rep)
if rank == 0:
    for name in sorted(allrep[0], key=lambda k: -allrep[0][k]):
        print("%-24s" % name, [r.get(name) for r in allrep], flush=True)
ms.close()
dist.destroy_process_group()
