"""Randomised check of batched mode against the CPU oracle through the host emulation (no GPU): batches of 1-5 random
trajectories (chain + bias or pose graph, 25-110 poses, 0-9 loop closures) must follow the oracle's LM path per trajectory.
    python tools/fuzz_batched.py        # ~8 minutes, 343 cases / 0 failures at the end of round 1"""
import sys, time, numpy as np, traceback
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import parity_common as pc
from visual_underwater_slam_b200 import _native, synthetic
emu = _native.bind(os.path.join(ROOT, 'tests', 'emu', 'libvus_emu.so'))
rng = np.random.default_rng(123)
t0=time.time(); n_ok=0; fails=[]
case=0
while time.time()-t0 < 500:
    case+=1
    T = int(rng.integers(1,6)); kind = rng.random() < 0.7
    probs=[]; desc=[]
    for t in range(T):
        n = int(rng.integers(25,110)); loops = int(rng.integers(0,10)); seed=int(rng.integers(0,10000))
        gap = max(5, n//4)
        if kind:
            d = synthetic.make_trajectory_graph(n, seed=seed, n_loops=loops, loop_min_gap=gap)
        else:
            d = synthetic.make_pose_graph(max(n,110) if loops else n, seed=seed, n_loops=loops)
        probs.append(d['graph'].to_problem(d['initial'])); desc.append((n,loops,seed))
    try:
        pc.check_batched_parity(emu, probs); n_ok+=1
    except Exception as e:
        fails.append((kind, desc, repr(e)[:200])); print('FAIL', kind, desc, repr(e)[:300], flush=True)
print('cases', case, 'ok', n_ok, 'fails', len(fails))
