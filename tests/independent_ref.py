"""TEST INFRASTRUCTURE ONLY -- a second, code-independent derivation of every factor on the batch.py:337 path.

Nothing here is shared with oracle/ (or with the product): every residual is written from its DEFINITION with the
matrix exponential (the power series itself, `expm` below, float64 -- `torch.linalg.matrix_exp` is only good to ~1e-11 for
generators of norm 0.01-0.04, measured here, which is not enough to referee 1e-10 comparisons) and every Jacobian is taken by
automatic differentiation
(`torch.func.jacrev`) under gtsam's retraction  Pose3: T * Exp(xi), xi = [omega; v];  vectors: additive.
Closed forms that the oracle and the kernels use (Rodrigues, the right Jacobian Jr, Pose3's Q matrix, the IMU bias
Jacobian recursions, the A / B / C covariance propagation matrices) do not appear: Jr and the SE(3) "V" matrix come out of
the exponential of an augmented generator, the preintegration Jacobians out of differentiating the integration loop itself.

What this pins (tests/test_independent_pin.py): the oracle and the CUDA path against a derivation that can only share a
misreading of WHAT gtsam computes (conventions: residual frame / sign, tangent order, which Jacobians gtsam approximates),
not a slip in HOW it is computed.  Those conventions are listed explicitly at each function, with the gtsam file they
restate.  gtsam itself is still absent: parity stays "unpinned" until a real gtsam is run against these.
"""
import torch

torch.set_default_dtype(torch.float64)
DT = torch.float64


# ------------------------------------------------------------------ Lie groups from the matrix exponential
def expm(M):
    """exp(M) = sum M^k / k!, by scaling (norm <= 1/4), 20 series terms (remainder < 1e-30) and repeated squaring."""
    nrm = float(torch.linalg.matrix_norm(M.detach(), ord=1))
    sq = 0
    while nrm > 0.25:
        nrm *= 0.5
        sq += 1
    A = M / (2.0 ** sq)
    E = torch.eye(M.shape[0], dtype=DT)
    term = torch.eye(M.shape[0], dtype=DT)
    for k in range(1, 21):
        term = term @ A / k
        E = E + term
    for _ in range(sq):
        E = E @ E
    return E


def hat3(w):
    z = torch.zeros((), dtype=DT)
    return torch.stack([torch.stack([z, -w[2], w[1]]), torch.stack([w[2], z, -w[0]]), torch.stack([-w[1], w[0], z])])


def hat6(xi):
    """se(3) generator of xi = [omega; v] (gtsam Pose3 tangent order, rotation first)."""
    top = torch.cat([hat3(xi[:3]), xi[3:].reshape(3, 1)], 1)
    return torch.cat([top, torch.zeros(1, 4, dtype=DT)], 0)


def Exp3(w):
    return expm(hat3(w))


def Exp6(xi):
    return expm(hat6(xi))


def left_jacobian(w):
    """V(w) = int_0^1 Exp(s w) ds, read off exp([[W, I], [0, 0]]) -- no closed form."""
    G = torch.zeros(6, 6, dtype=DT)
    G = G + torch.cat([torch.cat([hat3(w), torch.eye(3, dtype=DT)], 1), torch.zeros(3, 6, dtype=DT)], 0)
    return expm(G)[:3, 3:]


def right_jacobian(w):
    return left_jacobian(-w)


def Log3(R):
    """SO(3) logarithm (principal branch, |angle| < pi); checked against scipy.linalg.logm in the tests."""
    v = 0.5 * torch.stack([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = torch.sqrt(torch.sum(v * v) + 1e-300)
    c = 0.5 * (torch.trace(R) - 1.0)
    th = torch.atan2(s, c)
    # th / s -> 1 + s^2/6 near the identity (keeps the expression differentiable there)
    k = torch.where(s > 1e-8, th / s, 1.0 + s * s / 6.0)
    return v * k


def Log6(T):
    w = Log3(T[:3, :3])
    v = torch.linalg.solve(left_jacobian(w), T[:3, 3])
    return torch.cat([w, v])


def pose(R, t):
    top = torch.cat([R, t.reshape(3, 1)], 1)
    return torch.cat([top, torch.tensor([[0.0, 0.0, 0.0, 1.0]], dtype=DT)], 0)


def inv(T):
    R, t = T[:3, :3], T[:3, 3]
    return pose(R.T, -R.T @ t)


def retract(T, xi):
    return T @ Exp6(xi)


def jac(f, n):
    """Jacobian of f(delta) at delta = 0 in R^n."""
    return torch.func.jacrev(f)(torch.zeros(n, dtype=DT))


def adjoint(T):
    """Ad_T from its definition  T hat(xi) T^-1 = hat(Ad_T xi)."""
    def f(xi):
        M = T @ hat6(xi) @ inv(T)
        return torch.stack([M[2, 1], M[0, 2], M[1, 0], M[0, 3], M[1, 3], M[2, 3]])
    return jac(f, 6)


# ------------------------------------------------------------------ factors (unwhitened residual + Jacobians)
def prior_pose(T, Tp):
    """gtsam/nonlinear/PriorFactor.h: e = -Local(x, prior) = -Log(x^-1 prior).  gtsam returns H = I (it omits the derivative
    of Local); the TRUE derivative is returned second so the tests can show the approximation for what it is."""
    e = -Log6(inv(T) @ Tp)
    H_true = jac(lambda d: -Log6(inv(retract(T, d)) @ Tp), 6)
    return e, torch.eye(6, dtype=DT), H_true


def between(T1, T2, Tm, slow_but_correct):
    """gtsam/slam/BetweenFactor.h: hx = T1^-1 T2, e = Local(measured, hx) = Log(measured^-1 hx).
    GTSAM_SLOW_BUT_CORRECT_BETWEENFACTOR: exact Jacobians; default build: those of Between() alone, H1 = -Ad(hx^-1), H2 = I."""
    e = Log6(inv(Tm) @ inv(T1) @ T2)
    if slow_but_correct:
        H1 = jac(lambda d: Log6(inv(Tm) @ inv(retract(T1, d)) @ T2), 6)
        H2 = jac(lambda d: Log6(inv(Tm) @ inv(T1) @ retract(T2, d)), 6)
    else:
        hx = inv(T1) @ T2
        # Jacobians of hx = between(T1, T2) in hx's own chart: d Local(hx, between(T1 (+) d, T2)) / d d
        H1 = jac(lambda d: Log6(inv(hx) @ inv(retract(T1, d)) @ T2), 6)
        H2 = jac(lambda d: Log6(inv(hx) @ inv(T1) @ retract(T2, d)), 6)
    return e, H1, H2


def dvl(v, T, m):
    """batch.py:208-229: keys [V(i), X(i)], e = R_i m - v_i (the residual only; the reference's Jacobians are defective)."""
    e = T[:3, :3] @ m - v
    Hv = jac(lambda d: T[:3, :3] @ m - (v + d), 3)
    Hx = jac(lambda d: retract(T, d)[:3, :3] @ m - v, 6)
    return e, Hv, Hx


def stereo(T, l, z, K):
    """gtsam/geometry/StereoCamera.cpp::project2 + StereoFactor.h: q = R^T (l - t); (uL, uR, v) = (cx + fx x/z, cx + fx (x - b)/z,
    cy + fy y/z); e = projection - measured; cheirality (z <= 0): e = 2 fx (1,1,1), zero Jacobians."""
    fx, fy, _s, cx, cy, b = [float(k) for k in K]

    def proj(Tc, lm):
        q = Tc[:3, :3].T @ (lm - Tc[:3, 3])
        return torch.stack([cx + fx * q[0] / q[2], cx + fx * (q[0] - b) / q[2], cy + fy * q[1] / q[2]])
    q = T[:3, :3].T @ (l - T[:3, 3])
    if q[2] <= 0:
        return torch.full((3,), 2.0 * fx, dtype=DT), torch.zeros(3, 6, dtype=DT), torch.zeros(3, 3, dtype=DT)
    e = proj(T, l) - z
    Hx = jac(lambda d: proj(retract(T, d), l) - z, 6)
    Hl = jac(lambda d: proj(T, l + d) - z, 3)
    return e, Hx, Hl


# ------------------------------------------------------------------ IMU preintegration, both gtsam builds
def _nav_local(R0, p0, v0, R1, p1, v1):
    """NavState::localCoordinates (NavState.cpp): [Log(R0^T R1); R0^T (p1 - p0); R0^T (v1 - v0)]."""
    return torch.cat([Log3(R0.T @ R1), R0.T @ (p1 - p0), R0.T @ (v1 - v0)])


def _manifold_step(R, p, v, a, w, h):
    """ManifoldPreintegration::update -> NavState::update (body-frame increments)."""
    return R @ Exp3(w * h), p + v * h + R @ a * (0.5 * h * h), v + R @ a * h


def preintegrate_manifold(acc, gyro, h, bhat, aC, wC, iC):
    """-> dict(dR, dP, dV, dt, JRg, JPa, JPg, JVa, JVg, cov).  Bias Jacobians: derivative of the whole integration loop
    with respect to the bias it subtracts (rotation in the chart of the result: d Log(dR^T dR(b)) / db).  Covariance:
    cov <- A cov A^T + B (aC/h) B^T + C (wC/h) C^T, cov[3:6,3:6] += iC h  (ImuFactor.cpp integrateMeasurement), with A, B, C the
    derivatives of ONE step in the NavState chart, by autodiff."""
    k = acc.shape[0]

    def run(b):
        R, p, v = torch.eye(3, dtype=DT), torch.zeros(3, dtype=DT), torch.zeros(3, dtype=DT)
        for s in range(k):
            R, p, v = _manifold_step(R, p, v, acc[s] - b[:3], gyro[s] - b[3:], h)
        return R, p, v
    R, p, v = run(bhat)
    Jb = jac(lambda d: _pvec(run(bhat + d), R), 6)
    cov = torch.zeros(9, 9, dtype=DT)
    Rs, ps, vs = torch.eye(3, dtype=DT), torch.zeros(3, dtype=DT), torch.zeros(3, dtype=DT)
    for s in range(k):
        a, w = acc[s] - bhat[:3], gyro[s] - bhat[3:]
        Rn, pn, vn = _manifold_step(Rs, ps, vs, a, w, h)

        def step_local(dx, da, dw):
            R0, p0, v0 = Rs @ Exp3(dx[:3]), ps + Rs @ dx[3:6], vs + Rs @ dx[6:9]          # NavState::retract
            return _nav_local(Rn, pn, vn, *_manifold_step(R0, p0, v0, a + da, w + dw, h))
        z9, z3 = torch.zeros(9, dtype=DT), torch.zeros(3, dtype=DT)
        A = torch.func.jacrev(step_local, 0)(z9, z3, z3)
        B = torch.func.jacrev(step_local, 1)(z9, z3, z3)
        C = torch.func.jacrev(step_local, 2)(z9, z3, z3)
        cov = A @ cov @ A.T + B @ (aC / h) @ B.T + C @ (wC / h) @ C.T
        cov[3:6, 3:6] = cov[3:6, 3:6] + iC * h
        Rs, ps, vs = Rn, pn, vn
    return dict(dR=R, dP=p, dV=v, dt=k * h, JRg=Jb[0:3, 3:6], JPa=Jb[3:6, 0:3], JPg=Jb[3:6, 3:6], JVa=Jb[6:9, 0:3],
                JVg=Jb[6:9, 3:6], JRa=Jb[0:3, 0:3], cov=cov)


def _pvec(state, Rref):
    R, p, v = state
    return torch.cat([Log3(Rref.T @ R), p, v])


def _tangent_step(x, a, w, h):
    """TangentPreintegration::UpdatePreintegrated: theta += Jr(theta)^-1 w h; p += v h + Exp(theta) a h^2/2; v += Exp(theta) a h."""
    th, p, v = x[:3], x[3:6], x[6:9]
    R = Exp3(th)
    return torch.cat([th + torch.linalg.solve(right_jacobian(th), w) * h, p + v * h + R @ a * (0.5 * h * h), v + R @ a * h])


def preintegrate_tangent(acc, gyro, h, bhat, aC, wC, iC):
    """-> dict(x [9] = theta, p, v; dt; Hb [9,6] = d x / d bias(acc, gyro); cov)."""
    k = acc.shape[0]

    def run(b):
        x = torch.zeros(9, dtype=DT)
        for s in range(k):
            x = _tangent_step(x, acc[s] - b[:3], gyro[s] - b[3:], h)
        return x
    x = run(bhat)
    Hb = torch.func.jacrev(run)(bhat)
    cov = torch.zeros(9, 9, dtype=DT)
    xs = torch.zeros(9, dtype=DT)
    for s in range(k):
        a, w = acc[s] - bhat[:3], gyro[s] - bhat[3:]
        A = torch.func.jacrev(_tangent_step, 0)(xs, a, w, h)
        B = torch.func.jacrev(_tangent_step, 1)(xs, a, w, h)
        C = torch.func.jacrev(_tangent_step, 2)(xs, a, w, h)
        cov = A @ cov @ A.T + B @ (aC / h) @ B.T + C @ (wC / h) @ C.T
        cov[3:6, 3:6] = cov[3:6, 3:6] + iC * h
        xs = _tangent_step(xs, a, w, h)
    return dict(x=x, dt=k * h, Hb=Hb, cov=cov)


def imu_factor(Ti, vi, Tj, vj, bias, pim, g, tangent):
    """gtsam/navigation/{ImuFactor,PreintegrationBase,NavState}.cpp, keys (X_i, V_i, X_j, V_j, B):
         bias-corrected delta    manifold: dR Exp(JRg dbg), dP + JPa dba + JPg dbg, dV + ...   (ManifoldPreintegration::biasCorrectedDelta)
                                 tangent : x + Hb (b - bhat), then dR = Exp(theta)              (TangentPreintegration::biasCorrectedDelta)
         predict                 R_i dR ;  p_i + v_i dt + g dt^2/2 + R_i dP ;  v_i + g dt + R_i dV      (NavState::correctPIM + retract)
         error                   NavState_j.localCoordinates(predicted)  -- frame j, predicted minus actual, order [rot, pos, vel]
       velocities are separate additive variables (ImuFactor keys), poses retract with T Exp(xi), bias is additive [acc, gyro]."""
    dt = pim["dt"]

    def resid(dxi, dvi, dxj, dvj, db):
        Ta, Tb = retract(Ti, dxi), retract(Tj, dxj)
        Ri, pi, Rj, pj = Ta[:3, :3], Ta[:3, 3], Tb[:3, :3], Tb[:3, 3]
        b = bias + db
        if tangent:
            xc = pim["x"] + pim["Hb"] @ (b - pim["bhat"])
            dR, dP, dV = Exp3(xc[:3]), xc[3:6], xc[6:9]
        else:
            dba, dbg = b[:3] - pim["bhat"][:3], b[3:] - pim["bhat"][3:]
            dR = pim["dR"] @ Exp3(pim["JRg"] @ dbg)
            dP = pim["dP"] + pim["JPa"] @ dba + pim["JPg"] @ dbg
            dV = pim["dV"] + pim["JVa"] @ dba + pim["JVg"] @ dbg
        Rp = Ri @ dR
        pp = pi + (vi + dvi) * dt + 0.5 * g * dt * dt + Ri @ dP
        vp = (vi + dvi) + g * dt + Ri @ dV
        return _nav_local(Rj, pj, vj + dvj, Rp, pp, vp)
    z6, z3 = torch.zeros(6, dtype=DT), torch.zeros(3, dtype=DT)
    args = (z6, z3, z6, z3, z6)
    e = resid(*args)
    Hs = [torch.func.jacrev(resid, i)(*args) for i in range(5)]
    return e, Hs
