"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
Covers: stereo + IMU/DVL + loop-closure LM (supernode kernels, Schur, border), a batched solve (small-block kernels, per-component
reductions), marginals, IMU preintegration and stereo back-projection."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from visual_underwater_slam_b200 import synthetic, parallel
from visual_underwater_slam_b200.optimizer import Session, LevenbergMarquardtParams, backproject_stereo

p = LevenbergMarquardtParams()
p.maxIterations = 3
d = synthetic.make_trajectory_graph(60, seed=1, n_landmarks=100, n_loops=2, loop_min_gap=20, pixel_noise=1.0)
prob = d["graph"].to_problem(d["initial"])
s = Session(prob, p)
r = s.optimize()
print("stereo graph:", r["iterations"], r["final_error"])
cov = s.marginal_covariance([("pose", 10), ("lm", 3), ("bias", 0)])
print("marginals:", cov.shape, float(np.trace(cov)))
pts = backproject_stereo(s, np.arange(5, dtype=np.int32), prob["stereo"]["meas"][:5])
print("backproject:", pts.shape)
s.close()
probs = []
for t in range(3):
    dd = synthetic.make_trajectory_graph(50 + 5 * t, seed=10 + t, n_loops=2, loop_min_gap=15)
    probs.append(dd["graph"].to_problem(dd["initial"]))
res = parallel.solve_batched(probs, params=p)
print("batched:", [(x["iterations"], round(x["final_error"], 3)) for x in res])
dp = synthetic.make_pose_graph(80, seed=5, n_loops=6)
sp = Session(dp["graph"].to_problem(dp["initial"]), p)
print("pose graph:", sp.optimize()["final_error"])
sp.close()
print("sanitize smoke done")
