"""How far the CUDA path ends from a frozen oracle run (tests/golden/lm_*.npz): python tools/golden_diff.py lm_c3"""
import json
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from visual_underwater_slam_b200 import synthetic
from visual_underwater_slam_b200.optimizer import Session, LevenbergMarquardtParams

g = np.load(os.path.join(ROOT, "tests", "golden", sys.argv[1] + ".npz"))
meta = json.loads(str(g["meta"]))
d = synthetic.make_trajectory_graph(**meta["make"])
prob = d["graph"].to_problem(d["initial"])
s = Session(prob, LevenbergMarquardtParams())
res = s.optimize()
v = s.values()
st = meta.get("pose_stride", 1)
dp = v["poses"][::st, 9:] - g["poses"][:, 9:]
err = np.sqrt((dp ** 2).sum(1))
print("final error", res["final_error"], "golden", meta["final_error"], "rel", abs(res["final_error"] - meta["final_error"]) / meta["final_error"])
print("pose rmse", np.sqrt((err ** 2).mean()), "max", err.max(), "at strided index", int(err.argmax()), "of", len(err))
print("first quarter rmse", np.sqrt((err[:len(err) // 4] ** 2).mean()), "last quarter", np.sqrt((err[-len(err) // 4:] ** 2).mean()))
print("rot max", np.abs(v["poses"][::st, :9] - g["poses"][:, :9]).max(), "vel max", np.abs(v["vels"][::st] - g["vels"]).max(), "bias", np.abs(v["biases"] - g["biases"]).max())
T = d["truth"]["poses"][::st, 9:]
print("distance to truth: golden rmse", np.sqrt(((g["poses"][:, 9:] - T) ** 2).sum(1).mean()), "mine", np.sqrt(((v["poses"][::st, 9:] - T) ** 2).sum(1).mean()))
for t, (lam, suc, sol, ne) in zip(s.trace(), meta["tries"]):
    print("  lam %.0e  mine %.10e  golden %.10e  rel %.2e" % (lam, t["new_err"], ne, abs(t["new_err"] - ne) / ne))
