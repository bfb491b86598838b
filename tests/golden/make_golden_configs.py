"""Generates tests/golden/lm_c1.npz and lm_c2s.npz from the CPU oracle: BASELINE.json config 1 at full size
(2 000 poses, IMU + DVL chain, 50 loop closures) and a reduced config 2 (1 000 poses, 4 000 landmarks x 10
observations).  PARITY UNPINNED by the reference: the fixtures freeze the oracle's LM path (error after
every accepted step, lambda tries, final poses) so that neither oracle nor kernels can drift silently."""
import json
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from visual_underwater_slam_b200 import synthetic  # noqa: E402
from oracle import lm  # noqa: E402

CASES = {
    "lm_c1": dict(n_poses=2000, seed=1, n_loops=50, drift_scale=0.1),
    "lm_c2s": dict(n_poses=1000, seed=2, n_landmarks=4000, pixel_noise=1.0, drift_scale=0.1),
}
for name, kw in CASES.items():
    d = synthetic.make_trajectory_graph(**kw)
    prob = d["graph"].to_problem(d["initial"])
    vals, info = lm.lm_optimize(prob)
    tries = [(t["lam"], bool(t["success"])) for t in info["trace"]["tries"]]
    meta = dict(make=kw, iterations=info["iterations"], final_error=info["error"], final_lambda=info["lam"],
                errors=info["trace"]["errors"], tries=tries, n_factors=d["meta"]["n_factors"], preintegration="manifold")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), poses=vals["poses"], vels=vals["vels"],
                        biases=vals["biases"], lms=vals["lms"][::max(1, len(vals["lms"]) // 500)], meta=json.dumps(meta))
    print(name, {k: meta[k] for k in ("iterations", "final_error", "final_lambda")}, len(tries), "tries")
