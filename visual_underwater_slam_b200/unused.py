"""Names /root/reference/batch.py imports from gtsam (batch.py:20-27) but never uses on the batch LM path.  They exist so
that `from visual_underwater_slam_b200 import (...)` with batch.py's own import list succeeds; constructing one raises,
because nothing behind it is part of the path this package replaces."""


def _placeholder(name, where):
    class _Unused:
        __doc__ = f"gtsam.{name}: imported by {where} but not used on the batch LM path; not implemented here."

        def __init__(self, *args, **kwargs):
            raise NotImplementedError(f"gtsam.{name} is imported by {where} but is not on the batch Levenberg-Marquardt path "
                                      "this package replaces (DESIGN.md 6)")
    _Unused.__name__ = _Unused.__qualname__ = name
    return _Unused


BetweenFactorConstantBias = _placeholder("BetweenFactorConstantBias", "batch.py:20")
Cal3_S2 = _placeholder("Cal3_S2", "batch.py:20")
ConstantTwistScenario = _placeholder("ConstantTwistScenario", "batch.py:21")
PinholeCameraCal3_S2 = _placeholder("PinholeCameraCal3_S2", "batch.py:22")
PriorFactorConstantBias = _placeholder("PriorFactorConstantBias", "batch.py:23")
PriorFactorPoint3 = _placeholder("PriorFactorPoint3", "batch.py:24")
NavState = _placeholder("NavState", "batch.py:24")
