"""BASELINE config 4 across GPUs: trajectories sharded over the ranks (no data-path collective), every rank solves its shard as
ONE block-diagonal system (vus_set_components), one all-gather of the per-trajectory summary at the end.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/c4_sharded.py [n_traj] [n_poses]"""
import os
import sys
import time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from visual_underwater_slam_b200 import synthetic, parallel

n_traj = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n_poses = int(sys.argv[2]) if len(sys.argv) > 2 else 500
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))


def make(t):
    d = synthetic.make_trajectory_graph(n_poses, seed=4 + t, n_loops=5, loop_min_gap=100 if n_poses >= 300 else n_poses // 3)
    return d["graph"].to_problem(d["initial"])


first, last = parallel.shard_range(n_traj, rank, world)
cache = {t: make(t) for t in range(first, last)}            # graph construction is outside the timed region
for rep in range(2):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    table, loc, _ = parallel.solve_sharded(lambda t: cache[t], n_traj, device=local, batched=True)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        it = table[:, parallel.SUMMARY_FIELDS.index("iterations")]
        print("rep %d: %d trajectories x %d poses on %d GPU(s): %.3f s (max over ranks) = %.1f trajectories/s; iterations %d..%d; all errors decreased: %s" % (
            rep, n_traj, n_poses, world, dt.item(), n_traj / dt.item(), it.min(), it.max(), bool(np.all(table[:, 0] < table[:, 4]))), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
