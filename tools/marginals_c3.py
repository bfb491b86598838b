"""Cost of gtsam.Marginals queries at BASELINE config 3 size (100 000 poses, 2 M stereo factors) after optimize()."""
import os
import sys
import time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from visual_underwater_slam_b200 import synthetic
from visual_underwater_slam_b200.optimizer import Session, LevenbergMarquardtParams

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
d = synthetic.make_trajectory_graph(n, seed=3, n_landmarks=2 * n, pixel_noise=1.0, drift_scale=0.1)
s = Session(d["graph"].to_problem(d["initial"]), LevenbergMarquardtParams())
r = s.optimize()
print("optimize: %d iterations, %.1f ms" % (r["iterations"], r["ms_total"]))
for q in ([("pose", n // 2)], [("pose", n - 1), ("vel", n - 1), ("bias", 0)], [("lm", 12345 % (2 * n))],
          [("pose", n // 3), ("lm", 777 % (2 * n)), ("pose", 2 * n // 3)]):
    s.marginal_covariance(q)                                   # warm (graph capture of the 1-vector solve)
    t0 = time.perf_counter()
    c = s.marginal_covariance(q)
    dt = time.perf_counter() - t0
    sd = np.sqrt(np.diag(c))
    print("%-60s %3d columns  %.1f ms   sigma range %.2e .. %.2e" % (q, c.shape[0], 1e3 * dt, sd.min(), sd.max()))
s.close()
