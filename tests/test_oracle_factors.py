"""Pins every analytic Jacobian of oracle/factors.py by central finite differences under GTSAM's
retraction T*Exp(xi), and the IMU preintegration restatement."""
import numpy as np
import pytest
from oracle import lie, factors as F, preint

rng = np.random.default_rng(7)


def rand_pose(n, rot_scale=1.0, t_scale=3.0):
    return lie.pose_exp(np.concatenate([rng.standard_normal((n, 3)) * rot_scale, rng.standard_normal((n, 3)) * t_scale], 1))


def fd_pose(fun, R, t, h=1e-6):
    cols = []
    for k in range(6):
        d = np.zeros((R.shape[0], 6))
        d[:, k] = h
        Rp, tp = lie.pose_retract(R, t, d)
        Rm, tm = lie.pose_retract(R, t, -d)
        cols.append((fun(Rp, tp) - fun(Rm, tm)) / (2 * h))
    return np.stack(cols, -1)


def fd_vec(fun, x, h=1e-6):
    cols = []
    for k in range(x.shape[1]):
        d = np.zeros_like(x)
        d[:, k] = h
        cols.append((fun(x + d) - fun(x - d)) / (2 * h))
    return np.stack(cols, -1)


def test_between_jacobians():
    n = 8
    R1, t1 = rand_pose(n)
    R2, t2 = rand_pose(n)
    Rm, tm = lie.pose_between(R1, t1, R2, t2)
    dR, dt = rand_pose(n, 0.2, 0.3)
    Rm, tm = lie.pose_compose(Rm, tm, dR, dt)
    s = np.abs(rng.standard_normal((n, 6))) + 0.5
    # GTSAM_SLOW_BUT_CORRECT_BETWEENFACTOR build: the exact Jacobians
    r, (H1, H2) = F.between(R1, t1, R2, t2, Rm, tm, s, exact_jacobian=True)
    n1 = fd_pose(lambda R, t: F.between(R, t, R2, t2, Rm, tm, s)[0], R1, t1)
    n2 = fd_pose(lambda R, t: F.between(R1, t1, R, t, Rm, tm, s)[0], R2, t2)
    assert np.allclose(H1, n1, atol=5e-8, rtol=1e-7)
    assert np.allclose(H2, n2, atol=5e-8, rtol=1e-7)
    # default build (gtsam 4.1 / 4.2 wheels): the Jacobians of Between() only -- exact where the residual is zero
    r0, (G1, G2) = F.between(R1, t1, R2, t2, Rm, tm, s)
    assert np.array_equal(r0, r) and not np.allclose(G2, n2, atol=1e-3)
    assert np.allclose(G2, s[:, :, None] * np.eye(6)[None])
    Rz, tz = lie.pose_between(R1, t1, R2, t2)
    _, (Z1, Z2) = F.between(R1, t1, R2, t2, Rz, tz, s)
    z1 = fd_pose(lambda R, t: F.between(R, t, R2, t2, Rz, tz, s)[0], R1, t1)
    assert np.allclose(Z1, z1, atol=5e-8, rtol=1e-7)


def test_prior_pose_matches_gtsam_identity_jacobian():
    n = 4
    R, t = rand_pose(n)
    dR, dt = rand_pose(n, 1e-4, 1e-4)
    Rm, tm = lie.pose_compose(R, t, dR, dt)
    s = np.ones((n, 6))
    r, (H,) = F.prior_pose(R, t, Rm, tm, s)
    assert np.allclose(H, np.eye(6)[None])                         # GTSAM: H = I exactly
    num = fd_pose(lambda R_, t_: F.prior_pose(R_, t_, Rm, tm, s)[0], R, t)
    assert np.allclose(num, np.eye(6)[None], atol=5e-4)            # ...which is exact only near the prior


def test_dvl_jacobians():
    n = 6
    R, t = rand_pose(n)
    v = rng.standard_normal((n, 3))
    m = rng.standard_normal((n, 3))
    s = np.full((n, 3), 10.0)
    r, (Hv, Hx) = F.dvl(v, R, m, s)
    assert np.allclose(r, s * (np.einsum('nij,nj->ni', R, m) - v))     # batch.py:213-229
    assert np.allclose(Hv, fd_vec(lambda x: F.dvl(x, R, m, s)[0], v), atol=1e-7)
    assert np.allclose(Hx, fd_pose(lambda R_, t_: F.dvl(v, R_, m, s)[0], R, t), atol=1e-6)


def test_stereo_jacobians_and_cheirality():
    n = 10
    K = (1827.0, 1827.5999755859375, 0.0, 968.9000244140625, 561.4000244140625, 0.063)
    R, t = rand_pose(n)
    q = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(1, 5, n)], -1)
    l = np.einsum('nij,nj->ni', R, q) + t
    z = rng.standard_normal((n, 3)) * 5 + 900
    s = np.full((n, 3), 0.1)
    r, (Hp, Hl) = F.stereo(R, t, l, z, K, s)
    # prediction equals the pinhole model with uR = fx (x-b)/z
    pred = np.stack([K[3] + K[0] * q[:, 0] / q[:, 2], K[3] + K[0] * (q[:, 0] - K[5]) / q[:, 2], K[4] + K[1] * q[:, 1] / q[:, 2]], -1)
    assert np.allclose(r, s * (pred - z))
    assert np.allclose(Hp, fd_pose(lambda R_, t_: F.stereo(R_, t_, l, z, K, s)[0], R, t), atol=1e-4, rtol=1e-6)
    assert np.allclose(Hl, fd_vec(lambda x: F.stereo(R, t, x, z, K, s)[0], l), atol=1e-4, rtol=1e-6)
    # cheirality: point behind the camera
    lb = np.einsum('nij,nj->ni', R, q * np.array([1, 1, -1.0])) + t
    rb, (Hpb, Hlb) = F.stereo(R, t, lb, z, K, s)
    assert np.allclose(rb, s * 2.0 * K[0])
    assert not Hpb.any() and not Hlb.any()


def _imu_case(n):
    k = 40
    acc = rng.standard_normal((n, k, 3)) * 0.5 + np.array([0, 0, 9.81])
    gyr = rng.standard_normal((n, k, 3)) * 0.2
    bhat = rng.standard_normal((n, 6)) * 0.01
    pim, cov = preint.preintegrate(acc, gyr, 0.005, bhat, np.eye(3) * 9e-8, np.eye(3) * 1.2e-7, np.eye(3) * 1e-7)
    return acc, gyr, bhat, pim, cov


def test_imu_jacobians():
    n = 5
    acc, gyr, bhat, pim, cov = _imu_case(n)
    W = preint.sqrt_info_upper(cov)
    g = np.array([0, 0, -9.81])
    Ri, ti = rand_pose(n)
    vi = rng.standard_normal((n, 3))
    P = F.unpack_pim(pim)
    # a nearby-consistent state j, perturbed
    Rj = Ri @ P['dR'] @ lie.so3_exp(rng.standard_normal((n, 3)) * 0.01)
    tj = ti + vi * 0.2 + 0.5 * g * 0.04 + np.einsum('nij,nj->ni', Ri, P['dP']) + rng.standard_normal((n, 3)) * 0.01
    vj = vi + g * 0.2 + np.einsum('nij,nj->ni', Ri, P['dV']) + rng.standard_normal((n, 3)) * 0.01
    b = bhat + rng.standard_normal((n, 6)) * 0.01
    # use unit whitening for a well-scaled comparison, then check whitening separately
    I45 = np.tile(np.eye(9)[np.triu_indices(9)], (n, 1))
    r, Hs = F.imu(Ri, ti, vi, Rj, tj, vj, b, pim, I45, g)
    f = lambda **kw: F.imu(kw.get('Ri', Ri), kw.get('ti', ti), kw.get('vi', vi), kw.get('Rj', Rj), kw.get('tj', tj),
                           kw.get('vj', vj), kw.get('b', b), pim, I45, g)[0]
    num = [fd_pose(lambda R_, t_: f(Ri=R_, ti=t_), Ri, ti), fd_vec(lambda x: f(vi=x), vi),
           fd_pose(lambda R_, t_: f(Rj=R_, tj=t_), Rj, tj), fd_vec(lambda x: f(vj=x), vj), fd_vec(lambda x: f(b=x), b)]
    for H, N in zip(Hs, num):
        assert np.allclose(H, N, atol=2e-7, rtol=1e-6)
    # whitening: R^T R = cov^-1 and residual is R e
    Wm = F.unpack_triu(W)
    assert np.allclose(np.swapaxes(Wm, 1, 2) @ Wm @ cov, np.eye(9)[None], atol=1e-6)
    rw, Hw = F.imu(Ri, ti, vi, Rj, tj, vj, b, pim, W, g)
    assert np.allclose(rw, np.einsum('nij,nj->ni', Wm, r))
    assert np.allclose(Hw[0], Wm @ Hs[0])


def test_preintegration_constant_motion_exact():
    """Constant body rate + constant specific force: closed-form deltas."""
    n, k, dt = 1, 40, 0.005
    w = np.array([0.1, -0.2, 0.3])
    a = np.array([0.3, -0.1, 9.7])
    pim, cov = preint.preintegrate(np.tile(a, (n, k, 1)), np.tile(w, (n, k, 1)), dt, np.zeros(6), np.eye(3) * 1e-8, np.eye(3) * 1e-8, np.eye(3) * 1e-8)
    P = F.unpack_pim(pim)
    assert np.allclose(P['dR'][0], lie.so3_exp((w * k * dt)[None])[0], atol=1e-12)
    assert np.isclose(P['dt'][0], k * dt)
    # replay the discrete recursion directly
    R = np.eye(3); p = np.zeros(3); v = np.zeros(3)
    for _ in range(k):
        p = p + v * dt + 0.5 * (R @ a) * dt * dt
        v = v + (R @ a) * dt
        R = R @ lie.so3_exp((w * dt)[None])[0]
    assert np.allclose(P['dP'][0], p, atol=1e-13) and np.allclose(P['dV'][0], v, atol=1e-13)


def test_preintegration_bias_jacobians_fd():
    n = 3
    acc, gyr, bhat, pim, cov = _imu_case(n)
    P = F.unpack_pim(pim)
    h = 1e-6
    for c in range(6):
        d = np.zeros(6); d[c] = h
        Pp = F.unpack_pim(preint.preintegrate(acc, gyr, 0.005, bhat + d, np.eye(3), np.eye(3), np.eye(3))[0])
        Pm = F.unpack_pim(preint.preintegrate(acc, gyr, 0.005, bhat - d, np.eye(3), np.eye(3), np.eye(3))[0])
        dP = (Pp['dP'] - Pm['dP']) / (2 * h)
        dV = (Pp['dV'] - Pm['dV']) / (2 * h)
        dth = lie.so3_log(np.swapaxes(Pm['dR'], 1, 2) @ Pp['dR']) / (2 * h)
        if c < 3:
            assert np.allclose(dP, P['JPa'][:, :, c], atol=1e-7) and np.allclose(dV, P['JVa'][:, :, c], atol=1e-7)
            assert np.allclose(dth, 0, atol=1e-9)
        else:
            assert np.allclose(dP, P['JPg'][:, :, c - 3], atol=1e-7) and np.allclose(dV, P['JVg'][:, :, c - 3], atol=1e-7)
            assert np.allclose(dth, P['JRg'][:, :, c - 3], atol=1e-7)


def test_preintegration_covariance_monte_carlo():
    """First-order covariance vs sampled noise (NavState chart at the noise-free delta)."""
    k, dt = 40, 0.005
    w0 = np.array([0.05, -0.1, 0.2]); a0 = np.array([0.2, 0.1, 9.8])
    aC, wC, iC = np.eye(3) * 1e-4, np.eye(3) * 1e-4, np.eye(3) * 0.0
    pim0, cov = preint.preintegrate(np.tile(a0, (1, k, 1)), np.tile(w0, (1, k, 1)), dt, np.zeros(6), aC, wC, iC)
    P0 = F.unpack_pim(pim0)
    m = 4000
    acc = a0 + rng.standard_normal((m, k, 3)) * np.sqrt(1e-4 / dt)
    gyr = w0 + rng.standard_normal((m, k, 3)) * np.sqrt(1e-4 / dt)
    P = F.unpack_pim(preint.preintegrate(acc, gyr, dt, np.zeros(6), aC, wC, iC)[0])
    R0T = P0['dR'][0].T
    e = np.concatenate([lie.so3_log(R0T[None] @ P['dR']), (P['dP'] - P0['dP']) @ R0T.T, (P['dV'] - P0['dV']) @ R0T.T], 1)
    emp = e.T @ e / m
    assert np.allclose(emp, cov[0], atol=0.15 * np.abs(cov[0]).max())
    assert np.allclose(np.diag(emp), np.diag(cov[0]), rtol=0.15)


def test_product_preintegration_matches_oracle():
    """The host-side product preintegrator (navigation.py) against the oracle restatement."""
    from visual_underwater_slam_b200.navigation import preintegrate_batch, PreintegrationParams
    n = 4
    acc, gyr, bhat, pim, cov = _imu_case(n)
    p = PreintegrationParams.MakeSharedU(9.81)
    p.setAccelerometerCovariance(np.eye(3) * 9e-8); p.setGyroscopeCovariance(np.eye(3) * 1.2e-7); p.setIntegrationCovariance(np.eye(3) * 1e-7)
    pim2, info2, cov2 = preintegrate_batch(acc, gyr, 0.005, p, bhat, tangent=False)
    assert np.allclose(pim2, pim, atol=1e-13)
    assert np.allclose(cov2, cov, rtol=1e-10, atol=1e-20)
    assert np.allclose(info2, preint.sqrt_info_upper(cov), rtol=1e-8)
    # the tangent build (the default, config.py)
    pim_t, cov_t = preint.preintegrate_tangent(acc, gyr, 0.005, bhat, np.eye(3) * 9e-8, np.eye(3) * 1.2e-7, np.eye(3) * 1e-7)
    pim3, info3, cov3 = preintegrate_batch(acc, gyr, 0.005, p, bhat)
    assert np.allclose(pim3, pim_t, atol=1e-13)
    assert np.allclose(cov3, cov_t, rtol=1e-10, atol=1e-20)
    assert np.allclose(info3, preint.sqrt_info_upper(cov_t), rtol=1e-8)


def test_oracle_reproduces_golden_small():
    """The committed golden fixture is the oracle's own frozen answer (tests/golden/make_golden.py): re-running the
    oracle must reproduce it, so oracle edits cannot drift silently."""
    import json
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    import parity_common as pc
    from oracle import lm
    g = np.load(os.path.join(here, "golden", "lm_small.npz"))
    meta = json.loads(str(g["meta"]))
    _, prob = pc.make(**meta["make"])
    vals, info = lm.lm_optimize(prob)
    assert info["iterations"] == meta["iterations"]
    assert abs(info["error"] - meta["final_error"]) <= 1e-9 * meta["final_error"]
    assert np.abs(vals["poses"] - g["poses"]).max() < 1e-8
