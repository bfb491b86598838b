"""TEST INFRASTRUCTURE ONLY -- CPU oracle, IMU preintegration.

Restates gtsam/navigation/ManifoldPreintegration.cpp::update, NavState.cpp::update and
ImuFactor.cpp::PreintegratedImuMeasurements::integrateMeasurement (upstream gtsam 4.x,
un-vendored) for the call sites /root/reference/batch.py:91 (construction), :290
(integrateMeasurement(acc, gyro, 0.005)) and :293 (resetIntegration), with the
parameters of batch.py:181-187 (MakeSharedU(9.81), no Coriolis, no body_P_sensor).

PARITY UNPINNED (no GTSAM here).  Both build variants of gtsam are restated:
`preintegrate` = ManifoldPreintegration (GTSAM_TANGENT_PREINTEGRATION=OFF), `preintegrate_tangent` =
TangentPreintegration.cpp::UpdatePreintegrated / ::update (the default of the gtsam 4.0-4.2 builds).
Pinned in tests by (i) exactness on constant-rate motion, (ii) finite differences of the
bias Jacobians, (iii) Monte-Carlo covariance, (iv) an independent autodiff derivation
(tests/independent_ref.py).

Batched over n factors, sequential over k samples (as the reference loop at batch.py:289).
"""
import numpy as np
from . import lie


def preintegrate(acc, gyro, dt, bhat, acc_cov, gyro_cov, int_cov):
    """acc, gyro: [n,k,3]; dt: float or [n,k]; bhat: [n,6] (acc, gyro).

    Returns (pim [n,67] in the layout of factors.PIM_COLS, cov [n,9,9]).
    """
    acc = np.asarray(acc, dtype=np.float64)
    gyro = np.asarray(gyro, dtype=np.float64)
    n, k, _ = acc.shape
    dts = np.broadcast_to(np.asarray(dt, dtype=np.float64), (n, k))
    bhat = np.broadcast_to(np.asarray(bhat, dtype=np.float64), (n, 6))
    I3 = np.broadcast_to(np.eye(3), (n, 3, 3))
    dR = I3.copy()
    dP = np.zeros((n, 3))
    dV = np.zeros((n, 3))
    T = np.zeros(n)
    JRg = np.zeros((n, 3, 3))
    JPa = np.zeros((n, 3, 3))
    JPg = np.zeros((n, 3, 3))
    JVa = np.zeros((n, 3, 3))
    JVg = np.zeros((n, 3, 3))
    cov = np.zeros((n, 9, 9))
    aC = np.asarray(acc_cov, dtype=np.float64)
    wC = np.asarray(gyro_cov, dtype=np.float64)
    iC = np.asarray(int_cov, dtype=np.float64)
    for s in range(k):
        h = dts[:, s]
        h1 = h[:, None]
        h2 = h[:, None, None]
        a = acc[:, s] - bhat[:, 0:3]
        w = gyro[:, s] - bhat[:, 3:6]
        inc = lie.so3_exp(w * h1)
        incT = np.swapaxes(inc, 1, 2)
        Jr = lie.so3_dexp(w * h1)
        A_sk = lie.skew(a)
        dt22 = 0.5 * h * h
        # covariance (A, B, C of NavState::update in the NavState chart)
        A = np.zeros((n, 9, 9))
        A[:, 0:3, 0:3] = incT
        A[:, 3:6, 0:3] = -(incT @ A_sk) * dt22[:, None, None]
        A[:, 3:6, 3:6] = incT
        A[:, 3:6, 6:9] = incT * h2
        A[:, 6:9, 0:3] = -(incT @ A_sk) * h2
        A[:, 6:9, 6:9] = incT
        B = np.zeros((n, 9, 3))
        B[:, 3:6] = incT * dt22[:, None, None]
        B[:, 6:9] = incT * h2
        C = np.zeros((n, 9, 3))
        C[:, 0:3] = Jr * h2
        cov = A @ cov @ np.swapaxes(A, 1, 2)
        cov = cov + B @ (aC[None] / h2) @ np.swapaxes(B, 1, 2)
        cov = cov + C @ (wC[None] / h2) @ np.swapaxes(C, 1, 2)
        cov[:, 3:6, 3:6] += iC[None] * h2
        # bias Jacobians use the pre-update dR and JRg
        D_acc_bg = -(dR @ A_sk) @ JRg
        JPa = JPa + JVa * h2 - dt22[:, None, None] * dR
        JPg = JPg + JVg * h2 + dt22[:, None, None] * D_acc_bg
        JVa = JVa - dR * h2
        JVg = JVg + D_acc_bg * h2
        JRg = incT @ JRg - Jr * h2
        # state
        Ra = np.einsum('nij,nj->ni', dR, a)
        dP = dP + dV * h1 + Ra * dt22[:, None]
        dV = dV + Ra * h1
        dR = dR @ inc
        T = T + h
    pim = np.concatenate([dR.reshape(n, 9), dP, dV, T[:, None], bhat, JRg.reshape(n, 9), JPa.reshape(n, 9),
                          JPg.reshape(n, 9), JVa.reshape(n, 9), JVg.reshape(n, 9)], axis=1)
    return pim, cov


def preintegrate_tangent(acc, gyro, dt, bhat, acc_cov, gyro_cov, int_cov):
    """TangentPreintegration: the preintegrated state is the 9-vector [theta, p, v] in the tangent space at the
    start of the interval.  Per sample (a = acc - bhat_a, w = gyro - bhat_g, R = Exp(theta)):
        theta += Jr(theta)^-1 w dt ;  p += v dt + R a dt^2/2 ;  v += R a dt
    with A = d new / d old (9x9), B = d new / d a (9x3), C = d new / d w (9x3):
        H_biasAcc <- A H_biasAcc - B ;  H_biasOmega <- A H_biasOmega - C
        cov <- A cov A^T + B (aCov/dt) B^T + C (wCov/dt) C^T ;  cov[3:6,3:6] += iCov dt
    Returns (pim [n,67], cov [n,9,9]); pim columns 0:3 hold theta (3:9 are zero), 22:31 hold d theta / d b_gyro,
    the p / v tables are as in the manifold layout (factors.PIM_COLS).
    """
    acc = np.asarray(acc, dtype=np.float64)
    gyro = np.asarray(gyro, dtype=np.float64)
    n, k, _ = acc.shape
    dts = np.broadcast_to(np.asarray(dt, dtype=np.float64), (n, k))
    bhat = np.broadcast_to(np.asarray(bhat, dtype=np.float64), (n, 6))
    x = np.zeros((n, 9))
    T = np.zeros(n)
    Ha = np.zeros((n, 9, 3))
    Hg = np.zeros((n, 9, 3))
    cov = np.zeros((n, 9, 9))
    aC = np.asarray(acc_cov, dtype=np.float64)
    wC = np.asarray(gyro_cov, dtype=np.float64)
    iC = np.asarray(int_cov, dtype=np.float64)
    I9 = np.broadcast_to(np.eye(9), (n, 9, 9))
    for s in range(k):
        h = dts[:, s]
        h1 = h[:, None]
        h2 = h[:, None, None]
        dt22 = (0.5 * h * h)
        a = acc[:, s] - bhat[:, 0:3]
        w = gyro[:, s] - bhat[:, 3:6]
        theta = x[:, 0:3]
        w_tan, w_tan_H_theta, invH = lie.so3_apply_inv_dexp(theta, w)
        R = lie.so3_exp(theta)
        a_nav = np.einsum('nij,nj->ni', R, a)
        a_nav_H_theta = R @ lie.skew(-a) @ lie.so3_dexp(theta)
        A = I9.copy()
        A[:, 0:3, 0:3] += w_tan_H_theta * h2
        A[:, 3:6, 0:3] = a_nav_H_theta * dt22[:, None, None]
        A[:, 3:6, 6:9] = np.eye(3)[None] * h2
        A[:, 6:9, 0:3] = a_nav_H_theta * h2
        B = np.zeros((n, 9, 3))
        B[:, 3:6] = R * dt22[:, None, None]
        B[:, 6:9] = R * h2
        C = np.zeros((n, 9, 3))
        C[:, 0:3] = invH * h2
        xn = np.concatenate([theta + w_tan * h1, x[:, 3:6] + x[:, 6:9] * h1 + a_nav * dt22[:, None],
                             x[:, 6:9] + a_nav * h1], axis=1)
        x = xn
        T = T + h
        Ha = A @ Ha - B
        Hg = A @ Hg - C
        cov = A @ cov @ np.swapaxes(A, 1, 2)
        cov = cov + B @ (aC[None] / h2) @ np.swapaxes(B, 1, 2)
        cov = cov + C @ (wC[None] / h2) @ np.swapaxes(C, 1, 2)
        cov[:, 3:6, 3:6] += iC[None] * h2
    pim = np.concatenate([x[:, 0:3], np.zeros((n, 6)), x[:, 3:6], x[:, 6:9], T[:, None], bhat,
                          Hg[:, 0:3].reshape(n, 9), Ha[:, 3:6].reshape(n, 9), Hg[:, 3:6].reshape(n, 9),
                          Ha[:, 6:9].reshape(n, 9), Hg[:, 6:9].reshape(n, 9)], axis=1)
    return pim, cov


def sqrt_info_upper(cov):
    """Gaussian::Covariance -> R upper with R^T R = cov^-1 (NoiseModel.cpp), packed row-major triu."""
    n, d, _ = cov.shape
    info = np.linalg.inv(cov)
    info = 0.5 * (info + np.swapaxes(info, 1, 2))
    L = np.linalg.cholesky(info)          # info = L L^T  ->  R = L^T
    Rm = np.swapaxes(L, 1, 2)
    iu = np.triu_indices(d)
    return Rm[:, iu[0], iu[1]]
