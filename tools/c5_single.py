"""C5-style pose graph (prior + odometry + skip + loop closures) on ONE GPU: time, LM path, PCG iterations.
usage: python tools/c5_single.py [n_poses] [n_loops]"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from visual_underwater_slam_b200 import synthetic
from visual_underwater_slam_b200.optimizer import Session, LevenbergMarquardtParams

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
loops = int(sys.argv[2]) if len(sys.argv) > 2 else n // 2
t0 = time.time()
d = synthetic.make_pose_graph(n, seed=5, n_loops=loops)
prob = d["graph"].to_problem(d["initial"])
print("generated %d poses / %d factors in %.1f s" % (n, prob["n_factors"], time.time() - t0), flush=True)
p = LevenbergMarquardtParams()
p.verbosityLM = "SUMMARY"
t0 = time.perf_counter()
s = Session(prob, p)
t1 = time.perf_counter()
print("session (upload + analyze) %.2f s, layout %s" % (t1 - t0, s.layout()), flush=True)
r = s.optimize()
t2 = time.perf_counter()
print("optimize %.2f s: iterations %d, tries %d, pcg iterations %d, error %.6e -> %.6e, launches %d" % (
    t2 - t1, r["iterations"], r["inner_iterations"], r["pcg_iterations"], r["initial_error"], r["final_error"], r["kernel_launches"]))
print("factors linearized / s: %.3e" % (r["factors_linearized"] / (t2 - t1)))
s.close()
