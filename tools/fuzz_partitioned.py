"""Randomised check of the pose-range partition on CPU (host emulation + gloo): random chain lengths, loop-closure counts, rank
counts and with / without skip factors (k = 2 / k = 1); the N-rank solve must take the 1-rank LM path and end on the 1-rank
poses whenever the 1-rank solve itself converged every damped system (closure-dominated graphs whose PCG stalls are skipped:
their steps are inexact on any rank count).
    python tools/fuzz_partitioned.py [seconds] [seed]"""
import json
import os
import sys
import tempfile
import time
from pathlib import Path
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import test_parallel as tp

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 300.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 3)
tmp = Path(tempfile.mkdtemp())
t0 = time.time(); cases = ok = skipped = 0
while time.time() - t0 < budget:
    n = int(rng.integers(8, 60)) * 16 + int(rng.integers(0, 16))
    loops = int(rng.integers(0, 6))
    world = int(rng.integers(2, 5))
    extra = ("noskip",) if rng.random() < 0.4 else ()
    desc = dict(n=n, loops=loops, world=world, noskip=bool(extra))
    cases += 1
    try:
        one, many = tmp / "a.npy", tmp / "b.npy"
        tp._run_part(1, n, loops, one, tmp, extra=extra)
        m1 = json.load(open(str(one) + ".json"))
        if m1["pcg"] > 150 * m1["tries"]:                   # the 1-rank PCG itself stalls: no exact steps to compare
            skipped += 1
            continue
        tp._run_part(world, n, loops, many, tmp, extra=extra)
        mw = json.load(open(str(many) + ".json"))
        p1, pw = np.load(one), np.load(many)
        assert m1["iterations"] == mw["iterations"] and m1["tries"] == mw["tries"], (m1, mw)
        assert np.abs(p1 - pw).max() < 1e-6, np.abs(p1 - pw).max()
        assert mw["pcg"] <= 1.3 * m1["pcg"] + 10, (m1["pcg"], mw["pcg"])
        ok += 1
        print("ok", desc, "pcg", m1["pcg"], mw["pcg"], "diff %.1e" % np.abs(p1 - pw).max(), flush=True)
    except Exception as e:
        print("FAIL", desc, repr(e)[:300], flush=True)
print("cases", cases, "ok", ok, "skipped (1-rank PCG stalls)", skipped)
