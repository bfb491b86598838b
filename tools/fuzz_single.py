"""Randomised check of the single-graph path against the CPU oracle through the host emulation (no GPU): random stereo / chain /
loop-closure graphs with ragged tracks (observations removed), tracks longer than the band and landmarks seen twice from one
pose; LM trajectory try by try (parity_common.check_lm_parity) and marginal covariances of random variables.
    python tools/fuzz_single.py [seconds]"""
import os
import sys
import time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import parity_common as pc
from visual_underwater_slam_b200 import _native, synthetic

emu = _native.bind(os.path.join(ROOT, 'tests', 'emu', 'libvus_emu.so'))
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 420.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)


def mutate_stereo(prob, drop, ndup):
    st = {k: np.asarray(v) for k, v in prob["stereo"].items()}
    n0 = len(st["orig"])
    if n0 == 0:
        return prob
    keep = rng.random(n0) >= drop
    for l in range(len(prob["lm_keys"])):
        rows = np.nonzero(st["l"] == l)[0]
        if keep[rows].sum() < 2:
            keep[rows[:2]] = True
    st = {k: v[keep].copy() for k, v in st.items()}
    if ndup:
        dup = rng.integers(0, len(st["orig"]), ndup)
        for k in st:
            st[k] = np.concatenate([st[k], st[k][dup]], 0)
        st["meas"][-ndup:] += rng.normal(0, 0.3, (ndup, 3))
        st["orig"][-ndup:] = 10 ** 9 + np.arange(ndup)
    prob = dict(prob)
    others = [np.asarray(prob[t]["orig"]) for t in ("prior_pose", "prior_vel", "between", "dvl", "imu")]
    order = np.sort(np.concatenate(others + [st["orig"]]))
    remap = {int(o): i for i, o in enumerate(order)}
    for t in ("prior_pose", "prior_vel", "between", "dvl", "imu"):
        f = dict(prob[t]); f["orig"] = np.array([remap[int(o)] for o in f["orig"]], dtype=np.int64); prob[t] = f
    st["orig"] = np.array([remap[int(o)] for o in st["orig"]], dtype=np.int64)
    prob["stereo"] = st
    prob["n_factors"] = len(order)
    return prob


t0 = time.time(); cases = ok = refused = 0
while time.time() - t0 < budget:
    cases += 1
    n = int(rng.integers(20, 80)); nlm = int(rng.integers(0, 90)); opl = int(rng.integers(2, 15)); loops = int(rng.integers(0, 4))
    seed = int(rng.integers(0, 10000)); drop = float(rng.choice([0.0, 0.0, 0.2, 0.4])); ndup = int(rng.choice([0, 0, 0, 2, 5]))
    desc = dict(n=n, nlm=nlm, opl=min(opl, n - 1), loops=loops, seed=seed, drop=drop, ndup=ndup)
    try:
        d = synthetic.make_trajectory_graph(n, seed=seed, n_landmarks=nlm, obs_per_landmark=min(opl, n - 1), n_loops=loops,
                                            loop_min_gap=max(5, n // 4), pixel_noise=1.0)
        prob = mutate_stereo(d["graph"].to_problem(d["initial"]), drop, ndup if nlm else 0)
        pc.check_factor_parity(emu, prob)
        pc.check_lm_parity(emu, prob)
        q = [("pose", int(rng.integers(0, n))), ("vel", int(rng.integers(0, n))), ("bias", 0)]
        if nlm:
            q.append(("lm", int(rng.integers(0, nlm))))
        try:
            pc.check_marginals(emu, prob, q, rtol=1e-5)
        except RuntimeError as e:                     # the library refuses a column whose undamped solve did not converge
            if "did not converge" not in str(e):
                raise
            refused += 1
        ok += 1
    except Exception as e:
        print("FAIL", desc, repr(e)[:300], flush=True)
print("cases", cases, "ok", ok, "marginal queries refused (undamped solve not converged)", refused)
