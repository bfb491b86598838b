"""Host-side factor carriers with gtsam's constructor signatures.

Each class only records keys, measurement and noise; NonlinearFactorGraph.to_problem() packs them
into per-type structure-of-arrays tables (with the original insertion index) for the C-ABI.  The
arithmetic (residual, Jacobians, whitening) is in csrc/linearize.cuh.

Reference call sites: PriorFactorPose3 / PriorFactorVector batch.py:281-282; ImuFactor :238;
CustomFactor (DVL) :245-249; GenericStereoFactor3D :300-304; BetweenFactorPose3 is named by
BASELINE.json north_star (loop closures, configs 1 and 5).
"""
import functools
import numpy as np
from .geometry import Pose3, StereoPoint2, Cal3_S2Stereo
from .symbol import symbolChr
from .navigation import PreintegratedImuMeasurements
from .noise import upper_sqrt_information


def _diag_sqrt_info(noise, dim, what):
    if noise.dim() != dim:
        raise ValueError(f"{what}: noise model dimension {noise.dim()} != {dim}")
    d = noise.sqrt_info_diag()
    if d is None:
        raise NotImplementedError(f"{what}: only Diagonal/Isotropic noise is supported on the B200 path")
    return d


class _Factor:
    ftype = None

    def keys(self):
        return list(self._keys)

    def size(self):
        return len(self._keys)


class PriorFactorPose3(_Factor):
    ftype = "prior_pose"

    def __init__(self, key, prior, noise):
        self._keys = [int(key)]
        self.meas = Pose3(prior).as_row()
        self.sqrt_info = _diag_sqrt_info(noise, 6, "PriorFactorPose3")


class PriorFactorVector(_Factor):
    ftype = "prior_vel"

    def __init__(self, key, prior, noise):
        self._keys = [int(key)]
        self.meas = np.asarray(prior, dtype=np.float64).reshape(-1)
        if self.meas.size != 3:
            raise NotImplementedError("PriorFactorVector: only 3-vectors (velocity) are on the reference path")
        self.sqrt_info = _diag_sqrt_info(noise, 3, "PriorFactorVector")


class BetweenFactorPose3(_Factor):
    ftype = "between"

    def __init__(self, key1, key2, measured, noise):
        self._keys = [int(key1), int(key2)]
        self.meas = Pose3(measured).as_row()
        self.sqrt_info = _diag_sqrt_info(noise, 6, "BetweenFactorPose3")


class DvlVelocityFactor(_Factor):
    """Native stand-in for the reference's gtsam.CustomFactor DVL factor (batch.py:241-250).

    keys [V(i), X(i)]; e = R_i m - v_i (batch.py:213-229).  Jacobians are the analytically correct
    ones (de/dv = -I, de/dxi = [-R[m]x, 0]); the reference's own are defective.
    """
    ftype = "dvl"

    def __init__(self, noise, vKey, xKey, measurement):
        self._keys = [int(vKey), int(xKey)]
        self.meas = np.asarray(measurement, dtype=np.float64).reshape(-1)
        if self.meas.size != 3:
            raise ValueError("DvlVelocityFactor: measurement must have 3 entries")
        self.sqrt_info = _diag_sqrt_info(noise, 3, "DvlVelocityFactor")


class GenericStereoFactor3D(_Factor):
    ftype = "stereo"

    def __init__(self, measured, noise, poseKey, landmarkKey, K, body_P_sensor=None):
        if body_P_sensor is not None:
            raise NotImplementedError("body_P_sensor is not on the reference path (batch.py:300-304)")
        self._keys = [int(poseKey), int(landmarkKey)]
        self.meas = measured.vector() if isinstance(measured, StereoPoint2) else np.asarray(measured, float).reshape(3)
        self.sqrt_info = _diag_sqrt_info(noise, 3, "GenericStereoFactor3D")
        self.K = K.vector() if isinstance(K, Cal3_S2Stereo) else np.asarray(K, dtype=np.float64).reshape(6)


class ImuFactor(_Factor):
    """ImuFactor(pose_i, vel_i, pose_j, vel_j, bias, pim) -- copies the PIM (batch.py:238, :291)."""
    ftype = "imu"

    def __init__(self, pose_i, vel_i, pose_j, vel_j, bias, pim):
        self._keys = [int(pose_i), int(vel_i), int(pose_j), int(vel_j), int(bias)]
        if not isinstance(pim, PreintegratedImuMeasurements):
            raise TypeError("ImuFactor expects a PreintegratedImuMeasurements")
        row, cov = pim.snapshot()
        self.pim = row
        Rm = upper_sqrt_information(cov[None])[0]
        iu = np.triu_indices(9)
        self.sqrt_info = Rm[iu[0], iu[1]]
        self.gravity = pim.params().n_gravity.copy()
        self.tangent = pim.tangent()


class CustomFactor(_Factor):
    """gtsam.CustomFactor(noise, keys, error_function) -- constructor-compatible only.

    A Python callback cannot run inside a CUDA kernel and this path has no CPU fallback, so an
    arbitrary CustomFactor raises at optimize()/error().  The one pattern the reference uses
    (batch.py:245-249: keys [V(i), X(i)], functools.partial(velocity_error, measurement)) is
    recognised and lowered to the native DvlVelocityFactor.
    """
    ftype = "custom"

    def __init__(self, noise, keys, error_function):
        self._keys = [int(k) for k in keys]
        self.noise = noise
        self.error_function = error_function

    def as_dvl(self):
        fn = self.error_function
        if (isinstance(fn, functools.partial) and len(fn.args) >= 1 and len(self._keys) == 2
                and symbolChr(self._keys[0]) == 'v' and symbolChr(self._keys[1]) == 'x'):
            m = np.asarray(fn.args[0], dtype=np.float64).reshape(-1)
            if m.size == 3 and self.noise.dim() == 3:
                return DvlVelocityFactor(self.noise, self._keys[0], self._keys[1], m)
        return None
