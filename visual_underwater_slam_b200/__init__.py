"""B200-native drop-in for the batch Levenberg-Marquardt path of hvak/visual-underwater-slam.

`import visual_underwater_slam_b200 as gtsam` exposes the gtsam names /root/reference/batch.py
uses (batch.py:19-27): build a NonlinearFactorGraph + Values, then
LevenbergMarquardtOptimizer(graph, initial, LevenbergMarquardtParams()).optimize()  (batch.py:337).
All arithmetic runs in hand-written sm_100a CUDA kernels behind the C-ABI in include/vus.h;
there is no CPU fallback.
"""
from .config import gtsam_build, set_gtsam_build
from .symbol import Symbol, symbol, symbolChr, symbolIndex
from . import symbol_shorthand
from .geometry import Point3, Rot3, Pose3, StereoPoint2, Cal3_S2Stereo
from .noise import noiseModel
from .navigation import (PreintegrationParams, PreintegratedImuMeasurements, imuBias, ConstantBias,
                         preintegrate_batch)
from .factors import (PriorFactorPose3, PriorFactorVector, BetweenFactorPose3, DvlVelocityFactor,
                      GenericStereoFactor3D, ImuFactor, CustomFactor)
from .values import Values
from .graph import NonlinearFactorGraph
from .optimizer import LevenbergMarquardtParams, LevenbergMarquardtOptimizer, Marginals, JointMarginal, optimize_many
from .incremental import ISAM2, ISAM2Result
from .unused import (BetweenFactorConstantBias, Cal3_S2, ConstantTwistScenario, PinholeCameraCal3_S2,
                     PriorFactorConstantBias, PriorFactorPoint3, NavState)
from . import utils

__all__ = [n for n in dir() if not n.startswith("_")]
