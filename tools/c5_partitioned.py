"""BASELINE config 5 in the small: ONE pose graph split by contiguous pose range over the ranks (parallel.partition_pose_graph,
PartitionedSolver: halo exchange + all-reduced PCG / LM scalars through torch.distributed / NCCL).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/c5_partitioned.py <poses> <loops>"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, '.')
import torch, torch.distributed as dist
from visual_underwater_slam_b200 import parallel, synthetic
from visual_underwater_slam_b200.optimizer import Session, LevenbergMarquardtParams
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = int(sys.argv[1]); loops = int(sys.argv[2])
t = time.time()
d = synthetic.make_pose_graph(n, seed=5, n_loops=loops, noise_scale=float(os.environ.get("NOISE", "0.05")))
prob = d["graph"].to_problem(d["initial"])
if rank == 0: print("gen %.1f s, factors %d" % (time.time() - t, prob["n_factors"]), flush=True)
t = time.time()
part = parallel.partition_pose_graph(prob, world)[rank]
if rank == 0: print("partition %.1f s" % (time.time() - t), "halo", len(part["halo_global"]), flush=True)
p = LevenbergMarquardtParams(); p.pcgMaxIterations = 2000
if len(sys.argv) > 3: p.pcgRelTol = float(sys.argv[3])
ps = parallel.PartitionedSolver(part, p, device=lr)
dist.barrier(); torch.cuda.synchronize(); t = time.time()
res = ps.optimize()
torch.cuda.synchronize(); dist.barrier(); dt = time.time() - t
poses = ps.gather_poses()
if rank == 0:
    print(json.dumps(dict(world=world, n=n, seconds=dt, iterations=res["iterations"], tries=res["inner_iterations"], pcg=res["pcg_iterations"],
                          initial_error=res["initial_error"], final_error=res["final_error"], comm=ps.comm_calls,
                          factors_per_s=prob["n_factors"] * res["linearizations"] / dt)), flush=True)
    if world == 1 and n <= 500000: np.save("gpurun_out/part_ref_%d.npy" % n, poses)      # reference for the next (multi-rank) run
    elif os.path.exists("gpurun_out/part_ref_%d.npy" % n):
        ref = np.load("gpurun_out/part_ref_%d.npy" % n)
        print("max pose diff vs 1 rank: t %.3e R %.3e" % (np.abs(ref[:, 9:] - poses[:, 9:]).max(), np.abs(ref[:, :9] - poses[:, :9]).max()))
ps.session.close()
dist.destroy_process_group()
