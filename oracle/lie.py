"""TEST INFRASTRUCTURE ONLY -- CPU oracle, SO(3)/SE(3) arithmetic.

NumPy FP64 restatement of the Lie-group functions GTSAM evaluates underneath
`gtsam.LevenbergMarquardtOptimizer(...).optimize()` (reference call site
/root/reference/batch.py:337).  GTSAM itself (borglab/gtsam, un-pinned, README.md:18)
is NOT vendored in /root/reference and is not installable here, so this file restates
the published algorithm of upstream gtsam 4.1/4.2:

  gtsam/geometry/SO3.cpp   ExpmapFunctor / DexpFunctor / Logmap / LogmapDerivative
  gtsam/geometry/Pose3.cpp Expmap / Logmap / AdjointMap / ExpmapDerivative /
                           LogmapDerivative / computeQforExpmapDerivative

PARITY UNPINNED: the reference holds no tests or golden vectors for this path
(SURVEY.md 8c).  The formulas are pinned instead against scipy.linalg.expm/logm and
central finite differences in tests/test_oracle_lie.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  Every function is batched: leading axis N.
Tangent order is GTSAM's: xi = [omega(3); v(3)]  (rotation first, batch.py:95).
"""
import numpy as np

EPS = np.finfo(np.float64).eps


def skew(w):
    w = np.asarray(w, dtype=np.float64)
    z = np.zeros(w.shape[:-1])
    return np.stack([
        np.stack([z, -w[..., 2], w[..., 1]], -1),
        np.stack([w[..., 2], z, -w[..., 0]], -1),
        np.stack([-w[..., 1], w[..., 0], z], -1)], -2)


def _eye(n):
    return np.broadcast_to(np.eye(3), (n, 3, 3)).copy()


def so3_exp(w):
    """SO3::Expmap (SO3.cpp ExpmapFunctor): Rodrigues; theta^2 <= eps -> I + W."""
    w = np.atleast_2d(np.asarray(w, dtype=np.float64))
    n = w.shape[0]
    th2 = np.einsum('ni,ni->n', w, w)
    near = th2 <= EPS
    th = np.sqrt(np.where(near, 1.0, th2))
    W = skew(w)
    K = W / th[:, None, None]
    KK = K @ K
    s = np.sin(th)
    s2 = np.sin(0.5 * th)
    omc = 2.0 * s2 * s2
    R = _eye(n) + s[:, None, None] * K + omc[:, None, None] * KK
    R[near] = (_eye(n) + W)[near]
    return R


def so3_log(R):
    """SO3::Logmap (SO3.cpp, gtsam 4.1 thresholds)."""
    R = np.asarray(R, dtype=np.float64).reshape(-1, 3, 3)
    n = R.shape[0]
    tr = R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]
    out = np.empty((n, 3))
    vee = np.stack([R[:, 2, 1] - R[:, 1, 2], R[:, 0, 2] - R[:, 2, 0], R[:, 1, 0] - R[:, 0, 1]], -1)
    tr3 = tr - 3.0
    far = tr3 < -1e-6
    c = np.clip((tr - 1.0) * 0.5, -1.0, 1.0)
    th = np.arccos(c)
    with np.errstate(divide='ignore', invalid='ignore'):
        mag_far = th / (2.0 * np.sin(th))
    mag_near = 0.5 - tr3 / 12.0 + tr3 * tr3 / 60.0
    mag = np.where(far, mag_far, mag_near)
    out[:] = mag[:, None] * vee
    pi_case = np.abs(tr + 1.0) < 1e-10      # gtsam 4.1 threshold (4.2 widens it to 1e-3 with a refined formula)
    for k in np.nonzero(pi_case)[0]:
        r = R[k]
        if abs(r[2, 2] + 1.0) > 1e-5:
            out[k] = (np.pi / np.sqrt(2.0 + 2.0 * r[2, 2])) * np.array([r[0, 2], r[1, 2], 1.0 + r[2, 2]])
        elif abs(r[1, 1] + 1.0) > 1e-5:
            out[k] = (np.pi / np.sqrt(2.0 + 2.0 * r[1, 1])) * np.array([r[0, 1], 1.0 + r[1, 1], r[2, 1]])
        else:
            out[k] = (np.pi / np.sqrt(2.0 + 2.0 * r[0, 0])) * np.array([1.0 + r[0, 0], r[1, 0], r[2, 0]])
    return out


def so3_dexp(w):
    """SO3::ExpmapDerivative = right Jacobian Jr (DexpFunctor::dexp)."""
    w = np.atleast_2d(np.asarray(w, dtype=np.float64))
    n = w.shape[0]
    th2 = np.einsum('ni,ni->n', w, w)
    near = th2 <= EPS
    th = np.sqrt(np.where(near, 1.0, th2))
    W = skew(w)
    K = W / th[:, None, None]
    KK = K @ K
    s2 = np.sin(0.5 * th)
    omc = 2.0 * s2 * s2
    a = omc / th
    b = 1.0 - np.sin(th) / th
    J = _eye(n) - a[:, None, None] * K + b[:, None, None] * KK
    J[near] = (_eye(n) - 0.5 * W)[near]
    return J


def so3_dlog(w):
    """SO3::LogmapDerivative = Jr^{-1}; theta^2 <= eps -> I."""
    w = np.atleast_2d(np.asarray(w, dtype=np.float64))
    n = w.shape[0]
    th2 = np.einsum('ni,ni->n', w, w)
    near = th2 <= EPS
    th2s = np.where(near, 1.0, th2)
    th = np.sqrt(th2s)
    W = skew(w)
    WW = W @ W
    coef = 1.0 / th2s - (1.0 + np.cos(th)) / (2.0 * th * np.sin(th))
    J = _eye(n) + 0.5 * W + coef[:, None, None] * WW
    J[near] = _eye(n)[near]
    return J


def so3_apply_dexp_deriv(w, c):
    """d/dw [ Jr(w) c ] at fixed c  (SO3.cpp DexpFunctor::applyDexp, H1), batched [n,3,3].

    Jr(w) c = c - A (w x c) + B (w x (w x c)),  A = (1 - cos th)/th^2,  B = (th - sin th)/th^3;  th^2 <= eps -> [c]x / 2.
    """
    w = np.atleast_2d(np.asarray(w, dtype=np.float64))
    c = np.atleast_2d(np.asarray(c, dtype=np.float64))
    n = w.shape[0]
    th2 = np.einsum('ni,ni->n', w, w)
    near = th2 <= EPS
    th2s = np.where(near, 1.0, th2)
    th = np.sqrt(th2s)
    s, co = np.sin(th), np.cos(th)
    omc = 2.0 * np.sin(0.5 * th) ** 2
    A = omc / th2s
    B = (th - s) / (th2s * th)
    dA = (th * s - 2.0 * omc) / (th2s * th)              # dA/dth
    dB = (omc * th - 3.0 * (th - s)) / (th2s * th2s)     # dB/dth
    wxc = np.cross(w, c)
    wwc = np.cross(w, wxc)
    wc = np.einsum('ni,ni->n', w, c)
    outer = lambda a, b: a[:, :, None] * b[:, None, :]
    D = (-(dA / th)[:, None, None] * outer(wxc, w) + A[:, None, None] * skew(c)
         + (dB / th)[:, None, None] * outer(wwc, w)
         + B[:, None, None] * (wc[:, None, None] * _eye(n) + outer(w, c) - 2.0 * outer(c, w)))
    D[near] = (0.5 * skew(c))[near]
    return D


def so3_apply_inv_dexp(w, v):
    """DexpFunctor::applyInvDexp: c = Jr(w)^-1 v, H_w = -Jr^-1 d/dw[Jr(w) c], H_v = Jr^-1.  -> (c, H_w, H_v)"""
    Jr = so3_dexp(w)
    inv = np.linalg.inv(Jr)
    c = np.einsum('nij,nj->ni', inv, np.atleast_2d(v))
    return c, -inv @ so3_apply_dexp_deriv(w, c), inv


def pose_compose(Ra, ta, Rb, tb):
    return Ra @ Rb, ta + np.einsum('nij,nj->ni', Ra, tb)


def pose_inverse(R, t):
    Rt = np.swapaxes(R, -1, -2)
    return Rt, -np.einsum('nij,nj->ni', Rt, t)


def pose_between(Ra, ta, Rb, tb):
    """a^{-1} b"""
    Rat = np.swapaxes(Ra, -1, -2)
    return Rat @ Rb, np.einsum('nij,nj->ni', Rat, tb - ta)


def pose_exp(xi):
    """Pose3::Expmap."""
    xi = np.atleast_2d(np.asarray(xi, dtype=np.float64))
    w, v = xi[:, :3], xi[:, 3:]
    R = so3_exp(w)
    th2 = np.einsum('ni,ni->n', w, w)
    big = th2 > EPS
    tpar = w * np.einsum('ni,ni->n', w, v)[:, None]
    wxv = np.cross(w, v)
    t = (wxv - np.einsum('nij,nj->ni', R, wxv) + tpar) / np.where(big, th2, 1.0)[:, None]
    t = np.where(big[:, None], t, v)
    return R, t


def pose_log(R, t):
    """Pose3::Logmap."""
    R = np.asarray(R, dtype=np.float64).reshape(-1, 3, 3)
    t = np.asarray(t, dtype=np.float64).reshape(-1, 3)
    w = so3_log(R)
    th = np.linalg.norm(w, axis=1)
    small = th < 1e-10
    ths = np.where(small, 1.0, th)
    W = skew(w / ths[:, None])
    WT = np.einsum('nij,nj->ni', W, t)
    tan = np.tan(0.5 * ths)
    u = t - (0.5 * ths)[:, None] * WT + (1.0 - ths / (2.0 * tan))[:, None] * np.einsum('nij,nj->ni', W, WT)
    u = np.where(small[:, None], t, u)
    return np.concatenate([w, u], axis=1)


def pose_adjoint(R, t):
    """Pose3::AdjointMap = [[R,0],[ [t]x R, R ]]."""
    n = R.shape[0]
    A = np.zeros((n, 6, 6))
    A[:, :3, :3] = R
    A[:, 3:, 3:] = R
    A[:, 3:, :3] = skew(t) @ R
    return A


def pose_Q(xi):
    """Pose3::computeQforExpmapDerivative (Barfoot eq. 102, sign-adjusted as in GTSAM)."""
    xi = np.atleast_2d(np.asarray(xi, dtype=np.float64))
    w, v = xi[:, :3], xi[:, 3:]
    V = skew(v)
    W = skew(w)
    phi = np.linalg.norm(w, axis=1)
    WVW = W @ V @ W
    big = np.abs(phi) > 1e-5
    p = np.where(big, phi, 1.0)
    s, c = np.sin(p), np.cos(p)
    p2 = p * p
    p3 = p2 * p
    p4 = p3 * p
    p5 = p4 * p
    c1 = np.where(big, (p - s) / p3, 1.0 / 6.0)
    c2 = np.where(big, (1.0 - p2 / 2.0 - c) / p4, -1.0 / 24.0)
    c3 = np.where(big, -0.5 * ((1.0 - p2 / 2.0 - c) / p4 - 3.0 * (p - s - p3 / 6.0) / p5),
                  0.5 * (1.0 / 24.0 + 3.0 / 120.0))
    T1 = W @ V + V @ W - WVW
    T2 = W @ W @ V + V @ W @ W - 3.0 * WVW
    T3 = WVW @ W + W @ WVW
    return -0.5 * V + c1[:, None, None] * T1 + c2[:, None, None] * T2 + c3[:, None, None] * T3


def pose_dexp(xi):
    """Pose3::ExpmapDerivative."""
    xi = np.atleast_2d(np.asarray(xi, dtype=np.float64))
    n = xi.shape[0]
    Jr = so3_dexp(xi[:, :3])
    J = np.zeros((n, 6, 6))
    J[:, :3, :3] = Jr
    J[:, 3:, 3:] = Jr
    J[:, 3:, :3] = pose_Q(xi)
    return J


def pose_dlog_xi(xi):
    """Pose3::LogmapDerivative evaluated at xi = Logmap(pose)."""
    xi = np.atleast_2d(np.asarray(xi, dtype=np.float64))
    n = xi.shape[0]
    Jw = so3_dlog(xi[:, :3])
    Q = pose_Q(xi)
    Q2 = -Jw @ Q @ Jw
    J = np.zeros((n, 6, 6))
    J[:, :3, :3] = Jw
    J[:, 3:, 3:] = Jw
    J[:, 3:, :3] = Q2
    return J


def pose_retract(R, t, xi):
    """Pose3::retract with GTSAM_POSE3_EXPMAP: T * Expmap(xi)."""
    dR, dt = pose_exp(xi)
    return pose_compose(R, t, dR, dt)


def pose_local(Ra, ta, Rb, tb):
    """Pose3::localCoordinates: Logmap(a^{-1} b)."""
    R, t = pose_between(Ra, ta, Rb, tb)
    return pose_log(R, t)


def quat_to_rot(w, x, y, z):
    """Rot3::Quaternion(w,x,y,z) (w first, batch.py:47,:131)."""
    q = np.array([w, x, y, z], dtype=np.float64)
    q = q / np.linalg.norm(q)
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
