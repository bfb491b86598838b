"""The oracle (and, further down, the library behind the C-ABI) against tests/independent_ref.py: an autodiff derivation of
every factor from its definition that shares no code with oracle/ or with the kernels ("next round" item 1).

Tolerances: the two derivations are both FP64 and differ by rounding in a handful of 3x3 products -> 1e-9 relative on
Jacobians whose entries reach 1e3 (stereo) and 1e-10 elsewhere.  Preintegrated covariances go through 40 congruence
updates on each side -> 1e-9 relative.
"""
import numpy as np
import pytest
import scipy.linalg
import torch

import independent_ref as ind
from oracle import lie, factors as F, preint

RNG = np.random.default_rng(20261018)
T = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)
N = lambda t: t.detach().numpy()

ACC_COV = 8.999999999999999e-08 * np.eye(3)         # batch.py:183-185
GYRO_COV = 1.2184696791468346e-07 * np.eye(3)
INT_COV = 1e-07 * np.eye(3)
K_CAL = (1827.0, 1827.5999755859375, 0.0, 968.9000244140625, 561.4000244140625, 0.063)   # batch.py:110-115


def rand_rot(scale):
    """rotation vector with |w| < 2.5 rad (inside the principal branch of the logarithm)."""
    w = RNG.standard_normal(3) * scale
    return w * min(1.0, 2.5 / np.linalg.norm(w))


def rand_pose(scale_rot=1.0, scale_t=3.0):
    xi = np.concatenate([rand_rot(scale_rot), RNG.standard_normal(3) * scale_t])
    R, t = lie.pose_exp(xi[None])
    return R[0], t[0]


def P4(R, t):
    return ind.pose(T(R), T(t))


def close(a, b, rtol):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.abs(a - b).max() <= rtol * max(1.0, np.abs(b).max()), (np.abs(a - b).max(), np.abs(b).max())


# ------------------------------------------------------------------ the independent maps against the true definitions
def test_independent_maps_against_scipy():
    for _ in range(5):
        w = rand_rot(0.9)
        R = N(ind.Exp3(T(w)))
        close(R, scipy.linalg.expm(N(ind.hat3(T(w)))), 1e-13)
        close(N(ind.Log3(T(R))), w, 1e-12)
        xi = np.concatenate([rand_rot(0.8), RNG.standard_normal(3) * 2])
        M = scipy.linalg.expm(N(ind.hat6(T(xi))))
        close(N(ind.Log6(T(M))), xi, 1e-11)
        L = scipy.linalg.logm(M).real
        close(N(ind.Log6(T(M))), np.array([L[2, 1], L[0, 2], L[1, 0], L[0, 3], L[1, 3], L[2, 3]]), 1e-10)
        # right Jacobian from the augmented exponential vs its definition d Log(Exp(w)^-1 Exp(w + d)) / dd
        Jr = ind.jac(lambda d: ind.Log3(ind.Exp3(T(w)).T @ ind.Exp3(T(w) + d)), 3)
        close(N(ind.right_jacobian(T(w))), N(Jr), 1e-10)


def test_oracle_lie_maps_match_the_independent_ones():
    for _ in range(5):
        xi = np.concatenate([rand_rot(1.2), RNG.standard_normal(3) * 3])
        R, t = lie.pose_exp(xi[None])
        close(np.block([[R[0], t[0][:, None]], [np.zeros((1, 3)), 1]]), N(ind.Exp6(T(xi))), 1e-13)
        close(lie.pose_log(R, t)[0], N(ind.Log6(ind.Exp6(T(xi)))), 1e-11)
        close(lie.so3_dexp(xi[None, :3])[0], N(ind.right_jacobian(T(xi[:3]))), 1e-12)
        close(lie.pose_adjoint(R, t)[0], N(ind.adjoint(P4(R[0], t[0]))), 1e-12)


# ------------------------------------------------------------------ factors
def test_prior_pose_residual_and_the_identity_jacobian_convention():
    R, t = rand_pose()
    # near the prior and far from it
    for scale in (1e-3, 0.5):
        d = RNG.standard_normal(6) * scale
        Rp, tp = lie.pose_retract(R[None], t[None], d[None])
        e, H_gtsam, H_true = ind.prior_pose(P4(R, t), P4(Rp[0], tp[0]))
        r, J = F.prior_pose(R[None], t[None], Rp, tp, np.ones((1, 6)))
        close(r[0], N(e), 1e-11)
        close(J[0][0], N(H_gtsam), 0)                   # gtsam: H = I, NOT the derivative
        dev = np.abs(N(H_true) - np.eye(6)).max()
        assert dev < 2 * scale and (scale > 0.1) == (dev > 0.05)     # exact only at the prior


@pytest.mark.parametrize("slow", [False, True])
def test_between(slow):
    for _ in range(4):
        R1, t1 = rand_pose()
        R2, t2 = rand_pose()
        d = RNG.standard_normal(6) * 0.3
        Rh, th = lie.pose_between(R1[None], t1[None], R2[None], t2[None])
        Rm, tm = lie.pose_retract(Rh, th, d[None])
        e, H1, H2 = ind.between(P4(R1, t1), P4(R2, t2), P4(Rm[0], tm[0]), slow)
        s = RNG.uniform(0.5, 20, (1, 6))
        r, J = F.between(R1[None], t1[None], R2[None], t2[None], Rm, tm, s, exact_jacobian=slow)
        close(r[0], s[0] * N(e), 1e-10)
        close(J[0][0], s[0][:, None] * N(H1), 1e-10)
        close(J[1][0], s[0][:, None] * N(H2), 1e-10)


def test_dvl():
    for _ in range(4):
        R, t = rand_pose()
        v, m = RNG.standard_normal(3), RNG.standard_normal(3)
        e, Hv, Hx = ind.dvl(T(v), P4(R, t), T(m))
        r, J = F.dvl(v[None], R[None], m[None], np.full((1, 3), 10.0))
        close(r[0], 10 * N(e), 1e-12)
        close(J[0][0], 10 * N(Hv), 1e-12)
        close(J[1][0], 10 * N(Hx), 1e-12)


def test_stereo_and_cheirality():
    for trial in range(6):
        R, t = rand_pose()
        q = np.array([RNG.uniform(-1, 1), RNG.uniform(-1, 1), RNG.uniform(1, 5)])
        if trial == 5:
            q[2] = -1.0                                  # behind the camera
        l = R @ q + t
        z = np.array([900.0, 880.0, 500.0]) + RNG.standard_normal(3) * 10
        e, Hx, Hl = ind.stereo(P4(R, t), T(l), T(z), K_CAL)
        r, J = F.stereo(R[None], t[None], l[None], z[None], K_CAL, np.full((1, 3), 0.1))
        close(r[0], 0.1 * N(e), 1e-10)
        close(J[0][0], 0.1 * N(Hx), 1e-10)
        close(J[1][0], 0.1 * N(Hl), 1e-10)
    assert np.all(N(e) == 2 * K_CAL[0])


# ------------------------------------------------------------------ IMU: preintegration and factor, both gtsam builds
def imu_samples(k=40):
    acc = RNG.standard_normal((k, 3)) * 0.4 + np.array([0.1, -0.2, 9.81])
    gyr = RNG.standard_normal((k, 3)) * 0.05 + np.array([0.05, -0.1, 0.2])
    return acc, gyr


BHAT = np.array([0.02, -0.01, 0.015, 0.002, -0.001, 0.0015])


def test_preintegration_manifold():
    acc, gyr = imu_samples()
    ref = ind.preintegrate_manifold(T(acc), T(gyr), 0.005, T(BHAT), T(ACC_COV), T(GYRO_COV), T(INT_COV))
    pim, cov = preint.preintegrate(acc[None], gyr[None], 0.005, BHAT, ACC_COV, GYRO_COV, INT_COV)
    P = F.unpack_pim(pim)
    close(P["dR"][0], N(ref["dR"]), 1e-12)
    close(P["dP"][0], N(ref["dP"]), 1e-12)
    close(P["dV"][0], N(ref["dV"]), 1e-12)
    for k in ("JRg", "JPa", "JPg", "JVa", "JVg"):
        close(P[k][0], N(ref[k]), 1e-10)
    assert np.abs(N(ref["JRa"])).max() < 1e-14           # rotation does not depend on the accelerometer bias
    assert np.abs(cov[0] - N(ref["cov"])).max() <= 1e-9 * np.abs(cov[0]).max()


def test_preintegration_tangent():
    acc, gyr = imu_samples()
    ref = ind.preintegrate_tangent(T(acc), T(gyr), 0.005, T(BHAT), T(ACC_COV), T(GYRO_COV), T(INT_COV))
    pim, cov = preint.preintegrate_tangent(acc[None], gyr[None], 0.005, BHAT, ACC_COV, GYRO_COV, INT_COV)
    x, Hb = N(ref["x"]), N(ref["Hb"])
    close(pim[0, 0:3], x[0:3], 1e-12)
    assert np.all(pim[0, 3:9] == 0.0)
    close(pim[0, 9:12], x[3:6], 1e-12)
    close(pim[0, 12:15], x[6:9], 1e-12)
    close(pim[0, 22:31].reshape(3, 3), Hb[0:3, 3:6], 1e-10)
    close(pim[0, 31:40].reshape(3, 3), Hb[3:6, 0:3], 1e-10)
    close(pim[0, 40:49].reshape(3, 3), Hb[3:6, 3:6], 1e-10)
    close(pim[0, 49:58].reshape(3, 3), Hb[6:9, 0:3], 1e-10)
    close(pim[0, 58:67].reshape(3, 3), Hb[6:9, 3:6], 1e-10)
    assert np.abs(Hb[0:3, 0:3]).max() < 1e-14
    assert np.abs(cov[0] - N(ref["cov"])).max() <= 1e-9 * np.abs(cov[0]).max()


def test_the_two_preintegration_builds_differ_only_at_second_order():
    acc, gyr = imu_samples()
    pm, cm = preint.preintegrate(acc[None], gyr[None], 0.005, BHAT, ACC_COV, GYRO_COV, INT_COV)
    pt, ct = preint.preintegrate_tangent(acc[None], gyr[None], 0.005, BHAT, ACC_COV, GYRO_COV, INT_COV)
    Rm = pm[0, 0:9].reshape(3, 3)
    Rt = lie.so3_exp(pt[:, 0:3])[0]
    d = np.abs(lie.so3_log((Rm.T @ Rt)[None])).max()
    assert 1e-12 < d < 1e-4                              # not the same numbers: the variant matters for parity with a given gtsam


@pytest.mark.parametrize("tangent", [False, True])
def test_imu_factor(tangent):
    g = np.array([0.0, 0.0, -9.81])
    for _ in range(3):
        acc, gyr = imu_samples()
        if tangent:
            pim, cov = preint.preintegrate_tangent(acc[None], gyr[None], 0.005, BHAT, ACC_COV, GYRO_COV, INT_COV)
            ref = dict(x=T(np.concatenate([pim[0, 0:3], pim[0, 9:15]])), dt=float(pim[0, 15]), bhat=T(BHAT))
            Hb = np.zeros((9, 6))
            Hb[0:3, 3:6] = pim[0, 22:31].reshape(3, 3)
            Hb[3:6, 0:3], Hb[3:6, 3:6] = pim[0, 31:40].reshape(3, 3), pim[0, 40:49].reshape(3, 3)
            Hb[6:9, 0:3], Hb[6:9, 3:6] = pim[0, 49:58].reshape(3, 3), pim[0, 58:67].reshape(3, 3)
            ref["Hb"] = T(Hb)
        else:
            pim, cov = preint.preintegrate(acc[None], gyr[None], 0.005, BHAT, ACC_COV, GYRO_COV, INT_COV)
            P = F.unpack_pim(pim)
            ref = {k: T(P[k][0]) for k in ("dR", "dP", "dV", "JRg", "JPa", "JPg", "JVa", "JVg")}
            ref.update(dt=float(P["dt"][0]), bhat=T(BHAT))
        Ri, ti = rand_pose()
        vi = RNG.standard_normal(3)
        # state j = prediction + a perturbation, bias away from bhat
        d = RNG.standard_normal(6) * np.array([0.05] * 3 + [0.2] * 3)
        Rj, tj = lie.pose_retract(Ri[None], ti[None] + vi * 0.2, d[None])
        vj = vi + RNG.standard_normal(3) * 0.1
        bias = BHAT + RNG.standard_normal(6) * np.array([0.01] * 3 + [0.002] * 3)
        e, Hs = ind.imu_factor(P4(Ri, ti), T(vi), P4(Rj[0], tj[0]), T(vj), T(bias), ref, T(g), tangent)
        W = np.eye(9)[np.triu_indices(9)][None]         # identity sqrt information, packed upper triangle
        r, J = F.imu(Ri[None], ti[None], vi[None], Rj, tj, vj[None], bias[None], pim, W, g, tangent=tangent)
        close(r[0], N(e), 1e-10)
        for Jo, Hi in zip(J, Hs):
            close(Jo[0], N(Hi), 1e-9)


# ------------------------------------------------------------------ the library behind the C-ABI against the autodiff derivation
def _pim_ref(row, tangent):
    """packed PIM row (include/vus.h) -> the dict independent_ref.imu_factor takes."""
    bhat = T(row[16:22])
    if tangent:
        Hb = np.zeros((9, 6))
        Hb[0:3, 3:6] = row[22:31].reshape(3, 3)
        Hb[3:6, 0:3], Hb[3:6, 3:6] = row[31:40].reshape(3, 3), row[40:49].reshape(3, 3)
        Hb[6:9, 0:3], Hb[6:9, 3:6] = row[49:58].reshape(3, 3), row[58:67].reshape(3, 3)
        return dict(x=T(np.concatenate([row[0:3], row[9:15]])), dt=float(row[15]), bhat=bhat, Hb=T(Hb))
    return dict(dR=T(row[0:9].reshape(3, 3)), dP=T(row[9:12]), dV=T(row[12:15]), dt=float(row[15]), bhat=bhat,
                JRg=T(row[22:31].reshape(3, 3)), JPa=T(row[31:40].reshape(3, 3)), JPg=T(row[40:49].reshape(3, 3)),
                JVa=T(row[49:58].reshape(3, 3)), JVg=T(row[58:67].reshape(3, 3)))


def check_library_against_autodiff(lib, tangent, slow, n_each=3):
    """Every factor type of a small DVL / IMU / stereo / loop-closure graph, linearized by the library at a perturbed estimate,
    against independent_ref: whitened residuals and Jacobians in the node-ordered layout of include/vus.h."""
    from visual_underwater_slam_b200 import synthetic
    from visual_underwater_slam_b200.optimizer import Session
    import visual_underwater_slam_b200 as gtsam
    prev = gtsam.gtsam_build()
    gtsam.set_gtsam_build(tangent_preintegration=tangent, slow_but_correct_betweenfactor=slow)
    try:
        d = synthetic.make_trajectory_graph(30, seed=5, n_landmarks=40, n_loops=3, loop_min_gap=8, pixel_noise=1.0)
        prob = d["graph"].to_problem(d["initial"])
    finally:
        gtsam.set_gtsam_build(**prev)
    assert prob["options"] == dict(tangent_preintegration=tangent, slow_but_correct_betweenfactor=slow)
    prob = dict(prob)
    prob["biases"] = prob["biases"] + np.array([[0.01, -0.02, 0.005, 0.001, -0.002, 0.0015]])      # away from bias_hat
    prob["vels"] = prob["vels"] + 0.3
    s = Session(prob, lib=lib)
    try:
        out = {name: s.linearize(name) for name in ("prior_pose", "prior_vel", "between", "dvl", "stereo", "imu")}
    finally:
        s.close()
    P = lambda i: P4(prob["poses"][i, :9].reshape(3, 3), prob["poses"][i, 9:])
    g = T(prob["gravity"])
    for name, (r, J) in out.items():
        f = prob[name]
        for q in range(min(n_each, len(f["orig"]))):
            si = f["sqrt_info"][q]
            if name == "prior_pose":
                e, H, _ = ind.prior_pose(P(f["x"][q]), P4(f["meas"][q, :9].reshape(3, 3), f["meas"][q, 9:]))
                Hs = [H]
            elif name == "prior_vel":
                e, Hs = T(prob["vels"][f["v"][q]] - f["meas"][q]), [torch.eye(3)]
            elif name == "between":
                e, H1, H2 = ind.between(P(f["x1"][q]), P(f["x2"][q]), P4(f["meas"][q, :9].reshape(3, 3), f["meas"][q, 9:]), slow)
                Hs = [H1, H2]
            elif name == "dvl":
                e, Hv, Hx = ind.dvl(T(prob["vels"][f["v"][q]]), P(f["x"][q]), T(f["meas"][q]))
                Hs = [Hx, Hv]                                # node order: pose columns before velocity columns
            elif name == "stereo":
                e, Hx, Hl = ind.stereo(P(f["x"][q]), T(prob["lms"][f["l"][q]]), T(f["meas"][q]), prob["calib"])
                Hs = [Hx, Hl]
            else:
                e, Hs = ind.imu_factor(P(f["xi"][q]), T(prob["vels"][f["vi"][q]]), P(f["xj"][q]), T(prob["vels"][f["vj"][q]]),
                                       T(prob["biases"][f["b"][q]]), _pim_ref(f["meas"][q], tangent), g, tangent)
            if name == "imu":
                W = np.zeros((9, 9))
                W[np.triu_indices(9)] = si
            else:
                W = np.diag(si)
            close(r[q], W @ N(e), 1e-9)
            close(J[q], W @ np.concatenate([N(h) for h in Hs], 1), 1e-9)


@pytest.fixture(scope="module")
def emu():
    import os
    import subprocess
    from visual_underwater_slam_b200 import _native
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-s", "-C", os.path.join(root, "visual_underwater_slam_b200", "csrc"), "emu"], check=True)
    return _native.bind(os.path.join(root, "tests", "emu", "libvus_emu.so"))


@pytest.mark.parametrize("tangent,slow", [(True, False), (False, True)])
def test_kernel_bodies_against_autodiff_on_the_emulation(emu, tangent, slow):
    check_library_against_autodiff(emu, tangent, slow)


@pytest.mark.gpu
@pytest.mark.parametrize("tangent,slow", [(True, False), (False, False), (True, True), (False, True)])
def test_cuda_factors_against_autodiff(tangent, slow):
    from visual_underwater_slam_b200 import _native
    check_library_against_autodiff(_native.load(), tangent, slow)
