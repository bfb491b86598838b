"""TEST INFRASTRUCTURE ONLY -- CPU oracle, per-factor residuals and Jacobians.

NumPy FP64 restatement of the factor arithmetic GTSAM runs for the graph that
/root/reference/batch.py:270-305 builds and batch.py:337 optimises.  GTSAM is an
un-vendored, un-pinned dependency (README.md:18); upstream files restated:

  gtsam/nonlinear/PriorFactor.h       e = -Local(x, prior), H = I  (batch.py:281-282)
  gtsam/slam/BetweenFactor.h          e = Log(m^-1 T1^-1 T2)       (north_star)
  gtsam/slam/StereoFactor.h + gtsam/geometry/StereoCamera.cpp      (batch.py:300-305)
  gtsam/navigation/ImuFactor.cpp + ManifoldPreintegration.cpp + NavState.cpp (batch.py:238)
  gtsam/linear/NoiseModel.cpp         whitening                     (batch.py:95-98,:118)
  DVL velocity residual: defined by the reference itself, batch.py:196-233 (residual only;
  its Jacobians are defective, SURVEY.md Appendix B -- the analytic ones are used).

PARITY UNPINNED (no golden vectors in the reference).  Every Jacobian here is
pinned by central finite differences under GTSAM's retraction in tests/test_oracle_factors.py.

Conventions: every function is batched over n factors and returns
(r_whitened [n,m], [J_whitened_k [n,m,d_k] ...]) in the factor's key order.
Pose = (R [n,3,3], t [n,3]); pose tangent = [omega; v] with retract T*Exp(xi).
"""
import numpy as np
from . import lie


def _mv(A, x):
    return np.einsum('nij,nj->ni', A, x)


def _T(A):
    return np.swapaxes(A, -1, -2)


# ---------------------------------------------------------------- priors (PriorFactor.h)
def prior_pose(R, t, Rm, tm, sqrt_info):
    """PriorFactor<Pose3>: e = -Log(x^-1 prior); H = I6 (GTSAM leaves the exact dLog out)."""
    n = R.shape[0]
    e = -lie.pose_local(R, t, Rm, tm)
    H = np.broadcast_to(np.eye(6), (n, 6, 6))
    return sqrt_info * e, [sqrt_info[:, :, None] * H]


def prior_vec(v, vm, sqrt_info):
    """PriorFactor<Vector>: e = -(prior - x) = x - prior; H = I."""
    n, d = v.shape
    e = v - vm
    H = np.broadcast_to(np.eye(d), (n, d, d))
    return sqrt_info * e, [sqrt_info[:, :, None] * H]


# ---------------------------------------------------------------- BetweenFactor<Pose3>
def between(R1, t1, R2, t2, Rm, tm, sqrt_info, exact_jacobian=False):
    """hx = T1^-1 T2 ; e = Local(measured, hx) = Log(m^-1 hx).

    BetweenFactor.h::evaluateError has two builds:
      default (gtsam 4.1 / 4.2 wheels)            H1 = -Ad(hx^-1), H2 = I   -- the Jacobians of Between() only; the
                                                  derivative of Local() is left out (exact only at e = 0);
      GTSAM_SLOW_BUT_CORRECT_BETWEENFACTOR=ON     H1 = -dLog(e) Ad(hx^-1), H2 = dLog(e)   (exact_jacobian=True).
    """
    Rh, th = lie.pose_between(R1, t1, R2, t2)
    Re, te = lie.pose_between(Rm, tm, Rh, th)
    e = lie.pose_log(Re, te)
    n = R1.shape[0]
    D = lie.pose_dlog_xi(e) if exact_jacobian else np.broadcast_to(np.eye(6), (n, 6, 6))
    Rhi, thi = lie.pose_inverse(Rh, th)
    H1 = -D @ lie.pose_adjoint(Rhi, thi)
    H2 = D
    s = sqrt_info[:, :, None]
    return sqrt_info * e, [s * H1, s * H2]


# ---------------------------------------------------------------- DVL (batch.py:196-233)
def dvl(v, R, m, sqrt_info):
    """keys [V(i), X(i)] (batch.py:247): e = R m - v; de/dv = -I; de/dxi = [-R [m]x, 0]."""
    n = v.shape[0]
    e = _mv(R, m) - v
    Hv = np.broadcast_to(-np.eye(3), (n, 3, 3))
    Hx = np.zeros((n, 3, 6))
    Hx[:, :, :3] = -R @ lie.skew(m)
    s = sqrt_info[:, :, None]
    return sqrt_info * e, [s * Hv, s * Hx]


# ---------------------------------------------------------------- GenericStereoFactor
def stereo(R, t, l, z, K, sqrt_info):
    """StereoCamera::project2 with Cal3_S2Stereo K = (fx, fy, s, u0, v0, b); camera pose = X(i).

    Cheirality (q_z <= 0): e = 2 fx [1,1,1], zero Jacobians (StereoFactor.h catch branch).
    """
    fx, fy, _s, u0, v0, b = K
    n = R.shape[0]
    q = _mv(_T(R), l - t)
    bad = q[:, 2] <= 0.0
    qz = np.where(bad, 1.0, q[:, 2])
    d = 1.0 / qz
    x, y = q[:, 0], q[:, 1]
    uL = fx * x * d
    uR = fx * (x - b) * d
    vv = fy * y * d
    e = np.stack([u0 + uL - z[:, 0], u0 + uR - z[:, 1], v0 + vv - z[:, 2]], -1)
    v1 = vv / fy
    v2 = fx * v1
    dx = d * x
    zero = np.zeros(n)
    Hp = np.stack([
        np.stack([uL * v1, -fx - dx * uL, v2, -d * fx, zero, d * uL], -1),
        np.stack([uR * v1, -fx - dx * uR, v2, -d * fx, zero, d * uR], -1),
        np.stack([fy + vv * v1, -dx * vv, -x * d * fy, zero, -d * fy, d * vv], -1)], -2)
    # H_lm[:,k] = d * [fx R(k,0) - R(k,2) uL ; fx R(k,0) - R(k,2) uR ; fy R(k,1) - R(k,2) v]
    Hl = np.stack([
        d[:, None] * (fx * R[:, :, 0] - R[:, :, 2] * uL[:, None]),
        d[:, None] * (fx * R[:, :, 0] - R[:, :, 2] * uR[:, None]),
        d[:, None] * (fy * R[:, :, 1] - R[:, :, 2] * vv[:, None])], -2)
    e[bad] = 2.0 * fx
    Hp[bad] = 0.0
    Hl[bad] = 0.0
    s = sqrt_info[:, :, None]
    return sqrt_info * e, [s * Hp, s * Hl]


# ---------------------------------------------------------------- ImuFactor (Manifold)
PIM_COLS = 67  # dR 9 | dP 3 | dV 3 | dt 1 | bhat(acc,gyro) 6 | JRg 9 | JPa 9 | JPg 9 | JVa 9 | JVg 9


def unpack_pim(pim):
    n = pim.shape[0]
    return dict(dR=pim[:, 0:9].reshape(n, 3, 3), dP=pim[:, 9:12], dV=pim[:, 12:15], dt=pim[:, 15],
                bhat=pim[:, 16:22], JRg=pim[:, 22:31].reshape(n, 3, 3), JPa=pim[:, 31:40].reshape(n, 3, 3),
                JPg=pim[:, 40:49].reshape(n, 3, 3), JVa=pim[:, 49:58].reshape(n, 3, 3),
                JVg=pim[:, 58:67].reshape(n, 3, 3))


def unpack_triu(tri, d=9):
    """[n, d(d+1)/2] row-major upper triangle -> [n,d,d] upper-triangular matrix."""
    n = tri.shape[0]
    M = np.zeros((n, d, d))
    iu = np.triu_indices(d)
    M[:, iu[0], iu[1]] = tri
    return M


def imu(Ri, ti, vi, Rj, tj, vj, bias, pim, sqrt_info_triu, gravity, tangent=False):
    """ImuFactor, keys (X_i, V_i, X_j, V_j, B) (batch.py:238).

    tangent=False: ManifoldPreintegration::biasCorrectedDelta -- dR_c = dR Exp(JRg dbg);
    tangent=True : TangentPreintegration::biasCorrectedDelta  -- the preintegrated 9-vector is corrected LINEARLY,
                   theta_c = theta + (d theta / d bg) dbg, and PreintegrationBase::predict retracts it: dR_c = Exp(theta_c)
                   (pim columns 0:3 = theta, 22:31 = d theta / d bg; preint.preintegrate_tangent).
    Everything after the bias-corrected deltas is shared (PreintegrationBase::computeError / NavState).

    error = NavState_j.localCoordinates(predict(state_i, bias)) =
      [Log(Rj^T Ri Exp(th_c)); Rj^T(p_i + v_i dt + g dt^2/2 + Ri p_c - p_j); Rj^T(v_i + g dt + Ri v_c - v_j)]
    residual order [rot, pos, vel]; bias tangent [acc, gyro].
    Noise: Gaussian::Covariance(preintMeasCov): whiten with upper R, R^T R = Sigma^-1.
    """
    P = unpack_pim(pim)
    n = Ri.shape[0]
    g = np.asarray(gravity, dtype=np.float64)
    dt = P['dt'][:, None]
    dba = bias[:, 0:3] - P['bhat'][:, 0:3]
    dbg = bias[:, 3:6] - P['bhat'][:, 3:6]
    # bias-corrected deltas (ManifoldPreintegration::biasCorrectedDelta)
    if tangent:
        corr = pim[:, 0:3] + _mv(P['JRg'], dbg)        # theta_c
        dRc = lie.so3_exp(corr)
    else:
        corr = _mv(P['JRg'], dbg)
        dRc = P['dR'] @ lie.so3_exp(corr)
    pc = P['dP'] + _mv(P['JPa'], dba) + _mv(P['JPg'], dbg)
    vc = P['dV'] + _mv(P['JVa'], dba) + _mv(P['JVg'], dbg)
    RjT = _T(Rj)
    RjTRi = RjT @ Ri
    E = RjTRi @ dRc
    rR = lie.so3_log(E)
    rp = _mv(RjT, ti + vi * dt + 0.5 * g * dt * dt + _mv(Ri, pc) - tj)
    rv = _mv(RjT, vi + g * dt + _mv(Ri, vc) - vj)
    e = np.concatenate([rR, rp, rv], axis=1)

    dlog = lie.so3_dlog(rR)
    Z = np.zeros((n, 3, 3))
    I = np.broadcast_to(np.eye(3), (n, 3, 3))
    # d/dX_i  (9x6): columns [omega_i, rho_i]
    Hxi = np.concatenate([
        np.concatenate([dlog @ _T(dRc), Z], 2),
        np.concatenate([-RjTRi @ lie.skew(pc), RjTRi], 2),
        np.concatenate([-RjTRi @ lie.skew(vc), Z], 2)], 1)
    # d/dV_i (9x3)
    Hvi = np.concatenate([Z, RjT * dt[:, :, None], RjT], 1)
    # d/dX_j (9x6)
    Hxj = np.concatenate([
        np.concatenate([-dlog @ _T(E), Z], 2),
        np.concatenate([lie.skew(rp), -I], 2),
        np.concatenate([lie.skew(rv), Z], 2)], 1)
    # d/dV_j (9x3)
    Hvj = np.concatenate([Z, Z, -RjT], 1)
    # d/dB (9x6): [acc, gyro]
    Jr_corr = lie.so3_dexp(corr)
    Hb = np.concatenate([
        np.concatenate([Z, dlog @ Jr_corr @ P['JRg']], 2),
        np.concatenate([RjTRi @ P['JPa'], RjTRi @ P['JPg']], 2),
        np.concatenate([RjTRi @ P['JVa'], RjTRi @ P['JVg']], 2)], 1)
    W = unpack_triu(sqrt_info_triu, 9)
    return _mv(W, e), [W @ Hxi, W @ Hvi, W @ Hxj, W @ Hvj, W @ Hb]
