"""Host-side IMU preintegration and the gtsam navigation names the reference uses.

Mirrors (same names / argument meaning) gtsam.PreintegrationParams (batch.py:181-187),
gtsam.PreintegratedImuMeasurements (batch.py:91, :290, :293), gtsam.imuBias.ConstantBias
(batch.py:92) and gtsam.ImuFactor (batch.py:238).  Preintegration is sequential per factor and
O(#samples); it runs on the host, vectorised ACROSS factors (`preintegrate_batch`) so that
100k factors x 40 samples take ~1 s of numpy.  The factor residual/Jacobians run on the GPU.

Both build variants of gtsam are implemented (config.py): TANGENT preintegration
(gtsam TangentPreintegration.cpp, the default of the 4.0-4.2 wheels and the default here) and MANIFOLD
preintegration (Forster et al.; ManifoldPreintegration.cpp).  The variant is stated in every result file.

Packed PIM row (67 doubles, the layout include/vus.h documents for VUS_FACTOR_IMU):
  manifold: dR 9         | dP 3 | dV 3 | dt 1 | bias_hat(acc,gyro) 6 | dR/dbg 9     | dP/dba 9 | dP/dbg 9 | dV/dba 9 | dV/dbg 9
  tangent : theta 3, 0*6 | p 3  | v 3  | dt 1 | bias_hat(acc,gyro) 6 | dtheta/dbg 9 | dp/dba 9 | dp/dbg 9 | dv/dba 9 | dv/dbg 9
"""
import numpy as np
from . import config
from .noise import upper_sqrt_information

PIM_COLS = 67
_EPS = np.finfo(np.float64).eps


def _hat_b(w):
    n = w.shape[0]
    M = np.zeros((n, 3, 3))
    M[:, 0, 1], M[:, 0, 2] = -w[:, 2], w[:, 1]
    M[:, 1, 0], M[:, 1, 2] = w[:, 2], -w[:, 0]
    M[:, 2, 0], M[:, 2, 1] = -w[:, 1], w[:, 0]
    return M


def _exp_and_jr(phi):
    """Batched SO(3) exponential and right Jacobian of phi [n,3]."""
    n = phi.shape[0]
    th2 = np.sum(phi * phi, axis=1)
    small = th2 <= _EPS
    th = np.sqrt(np.where(small, 1.0, th2))
    K = _hat_b(phi / th[:, None])
    KK = K @ K
    I = np.broadcast_to(np.eye(3), (n, 3, 3))
    s = np.sin(th)
    omc = 2.0 * np.sin(0.5 * th) ** 2
    E = I + s[:, None, None] * K + omc[:, None, None] * KK
    Jr = I - (omc / th)[:, None, None] * K + (1.0 - s / th)[:, None, None] * KK
    if np.any(small):
        W = _hat_b(phi)
        E[small] = (I + W)[small]
        Jr[small] = (I - 0.5 * W)[small]
    return E, Jr


def _inv_dexp_and_derivative(theta, w):
    """c = Jr(theta)^-1 w with d c / d theta and Jr^-1 (gtsam so3::DexpFunctor::applyInvDexp), batched."""
    n = theta.shape[0]
    _, Jr = _exp_and_jr(theta)
    inv = np.linalg.inv(Jr)
    c = np.einsum('nij,nj->ni', inv, w)
    th2 = np.sum(theta * theta, axis=1)
    small = th2 <= _EPS
    t2 = np.where(small, 1.0, th2)
    th = np.sqrt(t2)
    sn = np.sin(th)
    omc = 2.0 * np.sin(0.5 * th) ** 2
    A = omc / t2
    B = (th - sn) / (t2 * th)
    dA = (th * sn - 2.0 * omc) / (t2 * th) / th            # (dA/dth) / th
    dB = (omc * th - 3.0 * (th - sn)) / (t2 * t2) / th
    txc = np.cross(theta, c)
    ttc = np.cross(theta, txc)
    tc = np.sum(theta * c, axis=1)
    I = np.broadcast_to(np.eye(3), (n, 3, 3))
    D = (-dA[:, None, None] * txc[:, :, None] * theta[:, None, :] + A[:, None, None] * _hat_b(c)
         + dB[:, None, None] * ttc[:, :, None] * theta[:, None, :]
         + B[:, None, None] * (tc[:, None, None] * I + theta[:, :, None] * c[:, None, :] - 2.0 * c[:, :, None] * theta[:, None, :]))
    if np.any(small):
        D[small] = (0.5 * _hat_b(c))[small]
    return c, -inv @ D, inv


class _PimState:
    """Batched running preintegration state (n factors advance in lock step)."""

    def __init__(self, n, bhat, tangent=None):
        self.n = n
        self.tangent = config.gtsam_build()["tangent_preintegration"] if tangent is None else bool(tangent)
        self.bhat = np.broadcast_to(np.asarray(bhat, dtype=np.float64), (n, 6)).copy()
        self.reset()

    def reset(self):
        n = self.n
        self.dR = np.broadcast_to(np.eye(3), (n, 3, 3)).copy()
        self.dP = np.zeros((n, 3))
        self.dV = np.zeros((n, 3))
        self.T = np.zeros(n)
        self.JRg = np.zeros((n, 3, 3))
        self.JPa = np.zeros((n, 3, 3))
        self.JPg = np.zeros((n, 3, 3))
        self.JVa = np.zeros((n, 3, 3))
        self.JVg = np.zeros((n, 3, 3))
        self.cov = np.zeros((n, 9, 9))
        self.theta = np.zeros((n, 3))            # tangent variant: dR = Exp(theta) is kept in step with it

    def _step_tangent(self, a, w, h, aC, wC, iC):
        """TangentPreintegration::UpdatePreintegrated + ::update + the covariance line of integrateMeasurement."""
        n = self.n
        hh = h[:, None, None]
        q = (0.5 * h * h)[:, None, None]
        wt, wt_H_theta, invH = _inv_dexp_and_derivative(self.theta, w)
        R, Jr = _exp_and_jr(self.theta)
        a_nav = np.einsum('nij,nj->ni', R, a)
        aH = -(R @ _hat_b(a)) @ Jr                        # d (R a) / d theta
        A = np.broadcast_to(np.eye(9), (n, 9, 9)).copy()
        A[:, 0:3, 0:3] += wt_H_theta * hh
        A[:, 3:6, 0:3] = aH * q
        A[:, 3:6, 6:9] = np.eye(3)[None] * hh
        A[:, 6:9, 0:3] = aH * hh
        Bm = np.concatenate([np.zeros((n, 3, 3)), R * q, R * hh], axis=1)
        Cm = np.concatenate([invH * hh, np.zeros((n, 6, 3))], axis=1)
        Ha = np.concatenate([np.zeros((n, 3, 3)), self.JPa, self.JVa], axis=1)
        Hg = np.concatenate([self.JRg, self.JPg, self.JVg], axis=1)
        Ha = A @ Ha - Bm
        Hg = A @ Hg - Cm
        self.JPa, self.JVa = Ha[:, 3:6], Ha[:, 6:9]
        self.JRg, self.JPg, self.JVg = Hg[:, 0:3], Hg[:, 3:6], Hg[:, 6:9]
        cov = A @ self.cov @ np.swapaxes(A, 1, 2)
        cov += Bm @ (aC[None] / hh) @ np.swapaxes(Bm, 1, 2)
        cov += Cm @ (wC[None] / hh) @ np.swapaxes(Cm, 1, 2)
        cov[:, 3:6, 3:6] += iC[None] * hh
        self.cov = cov
        self.dP = self.dP + self.dV * h[:, None] + a_nav * (0.5 * h * h)[:, None]
        self.dV = self.dV + a_nav * h[:, None]
        self.theta = self.theta + wt * h[:, None]
        self.dR, _ = _exp_and_jr(self.theta)
        self.T = self.T + h

    def step(self, acc, gyro, dt, aC, wC, iC):
        n = self.n
        h = np.broadcast_to(np.asarray(dt, dtype=np.float64), (n,))
        a = acc - self.bhat[:, :3]
        w = gyro - self.bhat[:, 3:]
        if self.tangent:
            self._step_tangent(a, w, h, aC, wC, iC)
            return
        inc, Jr = _exp_and_jr(w * h[:, None])
        incT = np.swapaxes(inc, 1, 2)
        ax = _hat_b(a)
        hh = h[:, None, None]
        q = (0.5 * h * h)[:, None, None]
        # first-order covariance propagation in the NavState chart [dR, dP_body, dV_body]
        iTa = incT @ ax
        A = np.zeros((n, 9, 9))
        A[:, 0:3, 0:3] = incT
        A[:, 3:6, 0:3] = -iTa * q
        A[:, 3:6, 3:6] = incT
        A[:, 3:6, 6:9] = incT * hh
        A[:, 6:9, 0:3] = -iTa * hh
        A[:, 6:9, 6:9] = incT
        Bm = np.concatenate([np.zeros((n, 3, 3)), incT * q, incT * hh], axis=1)
        Cm = Jr * hh
        cov = A @ self.cov @ np.swapaxes(A, 1, 2)
        cov += Bm @ (aC[None] / hh) @ np.swapaxes(Bm, 1, 2)
        cov[:, 0:3, 0:3] += Cm @ (wC[None] / hh) @ np.swapaxes(Cm, 1, 2)
        cov[:, 3:6, 3:6] += iC[None] * hh
        self.cov = cov
        # bias Jacobians (use the pre-update dR, dR/dbg)
        dacc_dbg = -(self.dR @ ax) @ self.JRg
        self.JPa += self.JVa * hh - q * self.dR
        self.JPg += self.JVg * hh + q * dacc_dbg
        self.JVa -= self.dR * hh
        self.JVg += dacc_dbg * hh
        self.JRg = incT @ self.JRg - Jr * hh
        # deltas
        Ra = np.einsum('nij,nj->ni', self.dR, a)
        self.dP += self.dV * h[:, None] + Ra * (0.5 * h * h)[:, None]
        self.dV += Ra * h[:, None]
        self.dR = self.dR @ inc
        self.T += h

    def pack(self):
        n = self.n
        head = np.concatenate([self.theta, np.zeros((n, 6))], axis=1) if self.tangent else self.dR.reshape(n, 9)
        return np.concatenate([head, self.dP, self.dV, self.T[:, None], self.bhat,
                               self.JRg.reshape(n, 9), self.JPa.reshape(n, 9), self.JPg.reshape(n, 9),
                               self.JVa.reshape(n, 9), self.JVg.reshape(n, 9)], axis=1)


class PreintegrationParams:
    """gtsam.PreintegrationParams; MakeSharedU(g) -> n_gravity = (0,0,-g) (batch.py:181)."""

    def __init__(self, n_gravity):
        self.n_gravity = np.asarray(n_gravity, dtype=np.float64).reshape(3).copy()
        self.accelerometerCovariance = np.eye(3)
        self.gyroscopeCovariance = np.eye(3)
        self.integrationCovariance = np.eye(3)
        self.use2ndOrderCoriolis = False
        self.omegaCoriolis = None

    @staticmethod
    def MakeSharedU(g=9.81):
        return PreintegrationParams([0.0, 0.0, -float(g)])

    @staticmethod
    def MakeSharedD(g=9.81):
        return PreintegrationParams([0.0, 0.0, float(g)])

    def setAccelerometerCovariance(self, c):
        self.accelerometerCovariance = np.asarray(c, dtype=np.float64).reshape(3, 3).copy()

    def setGyroscopeCovariance(self, c):
        self.gyroscopeCovariance = np.asarray(c, dtype=np.float64).reshape(3, 3).copy()

    def setIntegrationCovariance(self, c):
        self.integrationCovariance = np.asarray(c, dtype=np.float64).reshape(3, 3).copy()

    def setUse2ndOrderCoriolis(self, flag):
        if flag:
            raise NotImplementedError("2nd-order Coriolis is off on the reference path (batch.py:186)")
        self.use2ndOrderCoriolis = False

    def setOmegaCoriolis(self, w):
        w = np.asarray(w, dtype=np.float64).reshape(3)
        if np.any(w != 0.0):
            raise NotImplementedError("non-zero omegaCoriolis is not on the reference path (batch.py:187)")
        self.omegaCoriolis = w.copy()


class ConstantBias:
    """gtsam.imuBias.ConstantBias(acc, gyro); tangent order [acc; gyro]."""
    __slots__ = ("_b",)

    def __init__(self, biasAcc=None, biasGyro=None):
        a = np.zeros(3) if biasAcc is None else np.asarray(biasAcc, dtype=np.float64).reshape(3)
        g = np.zeros(3) if biasGyro is None else np.asarray(biasGyro, dtype=np.float64).reshape(3)
        self._b = np.concatenate([a, g])

    def accelerometer(self):
        return self._b[:3].copy()

    def gyroscope(self):
        return self._b[3:].copy()

    def vector(self):
        return self._b.copy()

    def __repr__(self):
        return f"ConstantBias(acc={self._b[:3]}, gyro={self._b[3:]})"


class imuBias:  # namespace: gtsam.imuBias.ConstantBias()
    ConstantBias = ConstantBias


class PreintegratedImuMeasurements:
    """gtsam.PreintegratedImuMeasurements(params, bias=ConstantBias()) (batch.py:91)."""

    def __init__(self, params, bias=None):
        self._p = params
        bhat = (bias or ConstantBias()).vector()
        self._s = _PimState(1, bhat[None])        # variant: config.gtsam_build() at construction, like the gtsam build

    def tangent(self):
        return self._s.tangent

    def integrateMeasurement(self, measuredAcc, measuredOmega, dt):
        """batch.py:290."""
        if dt <= 0:
            raise ValueError("dt <= 0 in PreintegratedImuMeasurements.integrateMeasurement")
        p = self._p
        self._s.step(np.asarray(measuredAcc, dtype=np.float64).reshape(1, 3),
                     np.asarray(measuredOmega, dtype=np.float64).reshape(1, 3), float(dt),
                     p.accelerometerCovariance, p.gyroscopeCovariance, p.integrationCovariance)

    def resetIntegration(self):
        """batch.py:293."""
        self._s.reset()

    def deltaTij(self):
        return float(self._s.T[0])

    def deltaRij(self):
        from .geometry import Rot3
        return Rot3(self._s.dR[0])

    def deltaPij(self):
        return self._s.dP[0].copy()

    def deltaVij(self):
        return self._s.dV[0].copy()

    def preintMeasCov(self):
        return self._s.cov[0].copy()

    def params(self):
        return self._p

    def snapshot(self):
        """(pim row [67], cov [9,9]) -- what ImuFactor copies at construction (the factor owns a copy)."""
        return self._s.pack()[0], self._s.cov[0].copy()


def preintegrate_batch(acc, gyro, dt, params, bias_hat=None, tangent=None):
    """Bulk API: acc, gyro [n,k,3]; dt scalar or [n,k]. -> (pim [n,67], sqrt_info_triu [n,45], cov [n,9,9]).
    tangent: None = config.gtsam_build()["tangent_preintegration"]."""
    acc = np.asarray(acc, dtype=np.float64)
    gyro = np.asarray(gyro, dtype=np.float64)
    n, k, _ = acc.shape
    dts = np.broadcast_to(np.asarray(dt, dtype=np.float64), (n, k))
    st = _PimState(n, np.zeros(6) if bias_hat is None else bias_hat, tangent)
    for s in range(k):
        st.step(acc[:, s], gyro[:, s], dts[:, s], params.accelerometerCovariance,
                params.gyroscopeCovariance, params.integrationCovariance)
    Rm = upper_sqrt_information(st.cov)
    iu = np.triu_indices(9)
    return st.pack(), Rm[:, iu[0], iu[1]], st.cov
