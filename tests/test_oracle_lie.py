"""Pins oracle/lie.py (the GTSAM SO3/Pose3 restatement) against scipy and finite differences.
SURVEY.md section 4 items 1-2; includes theta -> 0 and theta -> pi edge cases."""
import numpy as np
import pytest
from scipy.linalg import expm, logm
from scipy.spatial.transform import Rotation
from oracle import lie

rng = np.random.default_rng(0)


def hat6(xi):
    M = np.zeros((4, 4))
    M[:3, :3] = lie.skew(xi[:3])
    M[:3, 3] = xi[3:]
    return M


@pytest.mark.parametrize("scale", [1e-9, 1e-6, 1e-3, 0.3, 2.0, 3.1])
def test_so3_exp_log_vs_scipy(scale):
    w = rng.standard_normal((20, 3))
    w = w / np.linalg.norm(w, axis=1, keepdims=True) * scale
    R = lie.so3_exp(w)
    Rs = Rotation.from_rotvec(w).as_matrix()
    assert np.allclose(R, Rs, atol=1e-13)
    back = lie.so3_log(R)
    assert np.allclose(back, w, atol=1e-9 * max(1.0, scale), rtol=1e-7)


def test_so3_log_near_pi():
    for axis in (np.array([0, 0, 1.0]), np.array([0, 1.0, 0]), np.array([1.0, 0, 0]), np.array([1, 2, 3.0]) / np.sqrt(14)):
        for th in (np.pi, np.pi - 1e-5, np.pi - 1e-3):
            R = Rotation.from_rotvec(axis * th).as_matrix()
            w = lie.so3_log(R[None])[0]
            R2 = lie.so3_exp(w[None])[0]
            assert np.allclose(R, R2, atol=2e-6)


@pytest.mark.parametrize("scale", [1e-9, 1e-4, 0.5, 2.5])
def test_pose_exp_log_vs_expm(scale):
    xi = rng.standard_normal((10, 6))
    xi[:, :3] *= scale / np.linalg.norm(xi[:, :3], axis=1, keepdims=True)
    R, t = lie.pose_exp(xi)
    for k in range(10):
        M = expm(hat6(xi[k]))
        assert np.allclose(R[k], M[:3, :3], atol=1e-12)
        assert np.allclose(t[k], M[:3, 3], atol=1e-12)
    back = lie.pose_log(R, t)
    assert np.allclose(back, xi, atol=1e-8)


def _fd(f, x, h=1e-6):
    f0 = f(x)
    J = np.zeros((f0.size, x.size))
    for k in range(x.size):
        d = np.zeros_like(x)
        d[k] = h
        J[:, k] = (f(x + d) - f(x - d)) / (2 * h)
    return J


@pytest.mark.parametrize("scale", [1e-7, 1e-3, 0.7, 2.0])
def test_so3_dexp_dlog(scale):
    w = rng.standard_normal(3)
    w *= scale / np.linalg.norm(w)
    Jr = lie.so3_dexp(w[None])[0]
    R0 = lie.so3_exp(w[None])[0]
    num = _fd(lambda d: lie.so3_log((R0.T @ lie.so3_exp((w + d)[None])[0])[None])[0], np.zeros(3))
    assert np.allclose(Jr, num, atol=1e-8)
    assert np.allclose(lie.so3_dlog(w[None])[0] @ Jr, np.eye(3), atol=1e-9)


@pytest.mark.parametrize("scale", [1e-7, 1e-4, 0.3, 1.5])
def test_pose_dexp_dlog(scale):
    xi = rng.standard_normal(6)
    xi[:3] *= scale / np.linalg.norm(xi[:3])
    R0, t0 = lie.pose_exp(xi[None])
    J = lie.pose_dexp(xi[None])[0]

    def f(d):
        R1, t1 = lie.pose_exp((xi + d)[None])
        return lie.pose_local(R0, t0, R1, t1)[0]
    if scale > 1e-6:   # below that GTSAM's own (w x v - R w x v)/theta^2 cancellation swamps a finite difference
        num = _fd(f, np.zeros(6), h=1e-4 if scale < 1e-2 else 1e-6)
        assert np.allclose(J, num, atol=2e-8)
    assert np.allclose(lie.pose_dlog_xi(xi[None])[0] @ J, np.eye(6), atol=1e-8)


def test_adjoint_identity():
    xi = rng.standard_normal((5, 6))
    R, t = lie.pose_exp(xi)
    d = rng.standard_normal((5, 6)) * 1e-6
    # T Exp(d) T^-1 = Exp(Ad_T d)
    Ad = lie.pose_adjoint(R, t)
    Rd, td = lie.pose_exp(d)
    Ri, ti = lie.pose_inverse(R, t)
    Rl, tl = lie.pose_compose(*lie.pose_compose(R, t, Rd, td), Ri, ti)
    lhs = lie.pose_log(Rl, tl)
    assert np.allclose(lhs, np.einsum('nij,nj->ni', Ad, d), atol=1e-11)


def test_quaternion_w_first():
    q = Rotation.from_rotvec([0.3, -0.2, 0.5]).as_quat()  # x y z w
    assert np.allclose(lie.quat_to_rot(q[3], q[0], q[1], q[2]), Rotation.from_quat(q).as_matrix(), atol=1e-14)
