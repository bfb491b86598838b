"""Generates tests/golden/rows_marginals.npz and rows_batched.npz from the CPU oracle (run here, committed with its output):
frozen answers for the rows built around the LM path -- marginal covariances (gtsam.Marginals) and a batch of
independent trajectories (BASELINE config 4 in the small).  PARITY UNPINNED by the reference (no gtsam installable, no
reference tests): the fixtures freeze the oracle's own answers so that later changes cannot drift silently."""
import json
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_common as pc  # noqa: E402
from oracle import lm  # noqa: E402

# ---- marginals at the INITIAL values of a stereo + loop-closure graph
make = dict(n_poses=80, n_lm=120, n_loops=2, seed=11, loop_min_gap=30)
queries = [("pose", 0), ("pose", 79), ("vel", 40), ("bias", 0), ("lm", 17)]
_, prob = pc.make(**make)
short = {"pose": "x", "vel": "v", "bias": "b", "lm": "l"}
cov = lm.marginal_covariance(prob, lm.values_of(prob), [(short[k], i) for k, i in queries])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "rows_marginals.npz"), cov=cov,
                    meta=json.dumps(dict(make=make, queries=queries)))
print("marginals", cov.shape, float(np.trace(cov)))

# ---- a batch of independent trajectories, each solved by its own oracle LM
batch = [dict(n_poses=70 + 15 * t, n_loops=3, seed=90 + t, loop_min_gap=20) for t in range(4)]
rows = []
poses = []
for mk in batch:
    _, p = pc.make(**mk)
    vals, info = lm.lm_optimize(p)
    rows.append([info["iterations"], len(info["trace"]["tries"]), info["error"], info["lam"]])
    poses.append(vals["poses"])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "rows_batched.npz"), summary=np.array(rows),
                    poses=np.concatenate(poses, 0), meta=json.dumps(dict(batch=batch)))
print("batched", rows)
