"""Generates the BASELINE.json config fixtures from the CPU oracle (run here, committed with its output):

  lm_c1.npz   config 1 at full size (2 000 poses, IMU + DVL chain, 50 loop closures)
  lm_c2s.npz  a reduced config 2 (1 000 poses, 4 000 landmarks x 10 observations, soft noise: the round-1 fixture)
  lm_c2.npz   config 2 at full size (5 000 poses, 20 000 landmarks x 10 observations)
  lm_c3.npz   config 3 at full size (100 000 poses, 200 000 landmarks x 10 observations; ~15 min of oracle time, once)

PARITY UNPINNED by the reference: the fixtures freeze the oracle's LM path (error after every accepted step,
every lambda try with its trial error, final error / lambda, a strided subsample of the final values) so that neither oracle
nor kernels can drift silently.  usage: python tests/golden/make_golden_configs.py [names...]"""
import json
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from visual_underwater_slam_b200 import synthetic  # noqa: E402
from oracle import lm  # noqa: E402

CASES = {
    "lm_c1": dict(synthetic.CONFIGS["C1"]),
    "lm_c2s": dict(n_poses=1000, seed=2, n_landmarks=4000, pixel_noise=1.0, drift_scale=0.1),
    "lm_c2": dict(synthetic.CONFIGS["C2"]),
    "lm_c3": dict(synthetic.CONFIGS["C3"]),
}
for name in (sys.argv[1:] or ["lm_c1", "lm_c2s", "lm_c2"]):
    kw = CASES[name]
    t0 = time.time()
    d = synthetic.make_trajectory_graph(**kw)
    prob = d["graph"].to_problem(d["initial"])
    vals, info = lm.lm_optimize(prob, verbose=True)
    tries = [(t["lam"], bool(t["success"]), bool(t["solved"]), float(t["new_err"])) for t in info["trace"]["tries"]]
    stride = max(1, kw["n_poses"] // 2000)
    meta = dict(make=kw, iterations=info["iterations"], final_error=info["error"], final_lambda=info["lam"],
                errors=info["trace"]["errors"], tries=tries, n_factors=d["meta"]["n_factors"], pose_stride=stride,
                preintegration=d["meta"]["preintegration"], options=prob["options"], oracle_seconds=time.time() - t0)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), poses=vals["poses"][::stride], vels=vals["vels"][::stride],
                        biases=vals["biases"], lms=vals["lms"][::max(1, len(vals["lms"]) // 500)], meta=json.dumps(meta))
    print(name, {k: meta[k] for k in ("iterations", "final_error", "final_lambda")}, len(tries), "tries", "%.0f s" % (time.time() - t0), flush=True)
