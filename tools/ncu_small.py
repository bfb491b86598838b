"""One batched C4 solve, nothing else -- the target of the ncu captures of the small-block kernels:
    ncu --set full --clock-control none --import-source on --kernel-name-base demangled --kernel-name regex:SmallBwdBody \
        --launch-skip 400 --launch-count 12 -o gpurun_out/r1_small_bwd python tools/ncu_small.py 256 500"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from visual_underwater_slam_b200 import synthetic, parallel

T = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 500
probs = []
for t in range(T):
    d = synthetic.make_trajectory_graph(n, seed=4 + t, n_loops=5, loop_min_gap=100 if n >= 300 else n // 3)
    probs.append(d["graph"].to_problem(d["initial"]))
stats = {}
res = parallel.solve_batched(probs, keep_values=False, stats=stats)
print("solved", len(res), "trajectories; rounds", stats["inner_iterations"])
