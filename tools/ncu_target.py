"""A short single-solve workload for ncu: python tools/ncu_target.py POSES [band_chunks]  (C3 generator settings scaled to POSES)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visual_underwater_slam_b200 import synthetic
from visual_underwater_slam_b200.optimizer import Session, LevenbergMarquardtParams

n = int(sys.argv[1])
kw = dict(synthetic.CONFIGS["C3"])
kw.update(n_poses=n, n_landmarks=2 * n)
d = synthetic.make_trajectory_graph(**kw)
p = LevenbergMarquardtParams()
if len(sys.argv) > 2:
    p.bandChunks = int(sys.argv[2])
s = Session(d["graph"].to_problem(d["initial"]), p)
res = s.optimize()
print(res["iterations"], res["inner_iterations"], res["final_error"], res["ms_total"])
