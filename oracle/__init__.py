"""TEST INFRASTRUCTURE ONLY: a CPU restatement of the gtsam arithmetic behind /root/reference/batch.py:337 (lie.py, factors.py,
preint.py, lm.py).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it; the product never does.
Parity is UNPINNED against gtsam itself: real_gtsam() is the run-time probe for the day one is installed."""


def real_gtsam():
    """The real gtsam Python module if one is importable (and is not this repo's gtsam-named package), else None."""
    try:
        import importlib
        mod = importlib.import_module("gtsam")
    except Exception:
        return None
    if not hasattr(mod, "Pose3") or "visual_underwater_slam_b200" in (getattr(mod, "__file__", "") or ""):
        return None
    return mod
