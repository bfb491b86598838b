"""LM behaviour of one generator setting at a given size on the GPU: tries, accepted steps, chi2 per factor.
usage: python tools/noise_sweep.py POSES 'key=value,...' ['key=value,...' ...]"""
import ast
import sys
import os
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from visual_underwater_slam_b200 import synthetic
from visual_underwater_slam_b200.optimizer import Session, LevenbergMarquardtParams

n = int(sys.argv[1])
for spec in sys.argv[2:]:
    kw = {}
    for item in spec.split(","):
        if item:
            k, v = item.split("=")
            try:
                kw[k] = ast.literal_eval(v)
            except Exception:
                kw[k] = v
    lm_ratio = kw.pop("lm_ratio", 2)
    t0 = time.time()
    d = synthetic.make_trajectory_graph(n, seed=kw.pop("seed", 3), n_landmarks=lm_ratio * n, **kw)
    prob = d["graph"].to_problem(d["initial"])
    tg = time.time() - t0
    s = Session(prob, LevenbergMarquardtParams())
    t0 = time.time()
    res = s.optimize()
    ts = time.time() - t0
    tr = s.trace()
    v = s.values()
    T = d["truth"]["poses"]
    rmse = np.sqrt(((v["poses"][:, 9:] - T[:, 9:]) ** 2).sum(1).mean())
    nf = d["meta"]["n_factors"]
    print(f"n={n} {spec}: iters {res['iterations']} tries {res['inner_iterations']} pcg {res['pcg_iterations']} err0 {res['initial_error']:.3e} "
          f"final {res['final_error']:.4e} chi2/factor {2 * res['final_error'] / nf:.3f} pos-rmse {rmse:.3f} m  solve {ts:.2f} s (gen {tg:.0f} s)", flush=True)
    print("   ", "".join("%d%s " % (round(np.log10(t["lam"])), "+" if t["success"] else "-") for t in tr)[:600], flush=True)
    s.close()
