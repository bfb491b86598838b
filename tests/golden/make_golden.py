"""Generates tests/golden/lm_small.npz from the CPU oracle (run here, committed with its output).
PARITY UNPINNED by the reference (no gtsam installable, no reference tests): the fixture freezes the
oracle's own answer so that later changes to oracle or kernels cannot drift silently."""
import json
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_common as pc  # noqa: E402
from oracle import lm  # noqa: E402

make = dict(n_poses=150, n_lm=300, n_loops=2, seed=7, loop_min_gap=50)
_, prob = pc.make(**make)
vals0 = lm.values_of(prob)
fe0 = lm.factor_errors(prob, vals0)
vals, info = lm.lm_optimize(prob)
meta = dict(make=make, iterations=info["iterations"], final_error=info["error"], final_lambda=info["lam"],
            errors=info["trace"]["errors"], options=prob["options"])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "lm_small.npz"), factor_errors_initial=fe0,
                    poses=vals["poses"], vels=vals["vels"], biases=vals["biases"], lms=vals["lms"], meta=json.dumps(meta))
print(meta)
