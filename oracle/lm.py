"""TEST INFRASTRUCTURE ONLY -- CPU oracle, graph error / linearize / Levenberg-Marquardt.

Restates what `gtsam.LevenbergMarquardtOptimizer(graph, initial, LevenbergMarquardtParams())
.optimize()` does at /root/reference/batch.py:337 (all-default params), following upstream
gtsam 4.x (un-vendored, un-pinned; README.md:18):

  gtsam/nonlinear/NonlinearOptimizer.cpp        defaultOptimize / checkConvergence
  gtsam/nonlinear/LevenbergMarquardtOptimizer.cpp  iterate / tryLambda / buildDampedSystem
  gtsam/nonlinear/LevenbergMarquardtParams.h    defaults
  gtsam/nonlinear/NonlinearFactorGraph.cpp      error = sum 1/2 ||whitened r||^2, linearize
  gtsam/nonlinear/Values.cpp                    retract (per variable, x (+) delta)

GTSAM's linear solve is an exact sparse elimination (multifrontal Cholesky, COLAMD); here
the same damped normal equations (J^T J + lambda I) delta = J^T b are solved exactly with
SuperLU (scipy.sparse.linalg.splu), optionally after an exact landmark Schur complement --
the same linear system, so the same delta up to rounding.

PARITY UNPINNED: no GTSAM is installable here and the reference has no tests.

The input is the plain "problem" dict of numpy arrays that
visual_underwater_slam_b200.graph.NonlinearFactorGraph.to_problem() emits (keys, values,
per-type SoA factor tables with original insertion index) -- no product code is imported.
Column order of the linear system = GTSAM key order (b < l < v < x, Appendix A.8).
"""
import time
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from . import lie, factors as F

LM_DEFAULTS = dict(maxIterations=100, relativeErrorTol=1e-5, absoluteErrorTol=1e-5, errorTol=0.0,
                   lambdaInitial=1e-5, lambdaFactor=10.0, lambdaUpperBound=1e5, lambdaLowerBound=0.0,
                   minModelFidelity=1e-3, diagonalDamping=False, useFixedLambdaFactor=True)

FACTOR_TYPES = ('prior_pose', 'prior_vel', 'between', 'dvl', 'stereo', 'imu')


def split_pose(P):
    return P[:, :9].reshape(-1, 3, 3), P[:, 9:12]


class Layout:
    """Column offsets in GTSAM key order: b (6) < l (3) < v (3) < x (6)."""

    def __init__(self, prob):
        self.nb = len(prob['bias_keys'])
        self.nl = len(prob['lm_keys'])
        self.nv = len(prob['vel_keys'])
        self.nx = len(prob['pose_keys'])
        self.ob = 0
        self.ol = self.ob + 6 * self.nb
        self.ov = self.ol + 3 * self.nl
        self.ox = self.ov + 3 * self.nv
        self.n = self.ox + 6 * self.nx

    def cols(self, kind, idx):
        idx = np.asarray(idx, dtype=np.int64)
        if kind == 'b':
            return self.ob + 6 * idx, 6
        if kind == 'l':
            return self.ol + 3 * idx, 3
        if kind == 'v':
            return self.ov + 3 * idx, 3
        return self.ox + 6 * idx, 6


def values_of(prob):
    return dict(poses=prob['poses'].copy(), vels=prob['vels'].copy(),
                biases=prob['biases'].copy(), lms=prob['lms'].copy())


def build_options(prob):
    """The two compile-time switches of gtsam that change this path's arithmetic (prob['options'], written by
    graph.to_problem): GTSAM_TANGENT_PREINTEGRATION and GTSAM_SLOW_BUT_CORRECT_BETWEENFACTOR.  A problem without the
    entry is read as the gtsam 4.2 wheel build: tangent preintegration ON, slow-but-correct BetweenFactor OFF."""
    o = dict(tangent_preintegration=True, slow_but_correct_betweenfactor=False)
    o.update(prob.get('options') or {})
    return o


def eval_factors(prob, vals, ftype):
    """-> (r_whitened [n,m], [J_k], [(kind, idx)_k]) for one factor type at `vals`."""
    f = prob.get(ftype)
    if f is None or len(f['orig']) == 0:
        return None
    R, t = split_pose(vals['poses'])
    if ftype == 'prior_pose':
        Rm, tm = split_pose(f['meas'])
        r, J = F.prior_pose(R[f['x']], t[f['x']], Rm, tm, f['sqrt_info'])
        return r, J, [('x', f['x'])]
    if ftype == 'prior_vel':
        r, J = F.prior_vec(vals['vels'][f['v']], f['meas'], f['sqrt_info'])
        return r, J, [('v', f['v'])]
    if ftype == 'between':
        Rm, tm = split_pose(f['meas'])
        r, J = F.between(R[f['x1']], t[f['x1']], R[f['x2']], t[f['x2']], Rm, tm, f['sqrt_info'],
                         exact_jacobian=bool(build_options(prob)['slow_but_correct_betweenfactor']))
        return r, J, [('x', f['x1']), ('x', f['x2'])]
    if ftype == 'dvl':
        r, J = F.dvl(vals['vels'][f['v']], R[f['x']], f['meas'], f['sqrt_info'])
        return r, J, [('v', f['v']), ('x', f['x'])]
    if ftype == 'stereo':
        r, J = F.stereo(R[f['x']], t[f['x']], vals['lms'][f['l']], f['meas'], prob['calib'], f['sqrt_info'])
        return r, J, [('x', f['x']), ('l', f['l'])]
    if ftype == 'imu':
        r, J = F.imu(R[f['xi']], t[f['xi']], vals['vels'][f['vi']], R[f['xj']], t[f['xj']], vals['vels'][f['vj']],
                     vals['biases'][f['b']], f['pim'], f['sqrt_info'], prob['gravity'],
                     tangent=bool(build_options(prob)['tangent_preintegration']))
        return r, J, [('x', f['xi']), ('v', f['vi']), ('x', f['xj']), ('v', f['vj']), ('b', f['b'])]
    raise KeyError(ftype)


def factor_errors(prob, vals):
    """Per-factor 1/2||r||^2 in ORIGINAL insertion order (NonlinearFactorGraph index)."""
    nf = sum(len(prob[t]['orig']) for t in FACTOR_TYPES if t in prob)
    out = np.zeros(nf)
    for ft in FACTOR_TYPES:
        ev = eval_factors(prob, vals, ft)
        if ev is None:
            continue
        out[prob[ft]['orig']] = 0.5 * np.einsum('ni,ni->n', ev[0], ev[0])
    return out


def graph_error(prob, vals):
    """NonlinearFactorGraph::error: summed in insertion order like GTSAM's loop."""
    return float(np.sum(factor_errors(prob, vals)))


def linearize(prob, vals, lay=None):
    """-> (J csr [M,N] whitened, b [M] = -r). Row order: by type then factor (order is irrelevant to J^T J)."""
    lay = lay or Layout(prob)
    rows, cols, data, bs = [], [], [], []
    row0 = 0
    for ft in FACTOR_TYPES:
        ev = eval_factors(prob, vals, ft)
        if ev is None:
            continue
        r, Js, keys = ev
        n, m = r.shape
        base = row0 + m * np.arange(n)
        for Jk, (kind, idx) in zip(Js, keys):
            c0, d = lay.cols(kind, idx)
            rr = (base[:, None, None] + np.arange(m)[None, :, None]) + np.zeros((1, 1, d), dtype=np.int64)
            cc = (c0[:, None, None] + np.arange(d)[None, None, :]) + np.zeros((1, m, 1), dtype=np.int64)
            rows.append(rr.ravel())
            cols.append(cc.ravel())
            data.append(Jk.ravel())
        bs.append(-r.ravel())
        row0 += n * m
    J = sp.csr_matrix((np.concatenate(data), (np.concatenate(rows), np.concatenate(cols))), shape=(row0, lay.n))
    return J, np.concatenate(bs)


def retract(vals, delta, lay):
    """Values::retract(VectorValues): Pose3 -> T Exp(xi); vectors additive."""
    out = dict(vals)
    out['biases'] = vals['biases'] + delta[lay.ob:lay.ol].reshape(-1, 6)
    out['lms'] = vals['lms'] + delta[lay.ol:lay.ov].reshape(-1, 3)
    out['vels'] = vals['vels'] + delta[lay.ov:lay.ox].reshape(-1, 3)
    R, t = split_pose(vals['poses'])
    Rn, tn = lie.pose_retract(R, t, delta[lay.ox:].reshape(-1, 6))
    out['poses'] = np.concatenate([Rn.reshape(-1, 9), tn], axis=1)
    return out


def _splu_sym(A):
    """SuperLU in symmetric mode (minimum-degree on A+A^T, no pivoting): an exact sparse Cholesky-like solve."""
    return spla.splu(A.tocsc(), permc_spec='MMD_AT_PLUS_A', diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))


def _solve_banded_bordered(S, gc, lay, max_halfband=400):
    """Exact solve of the reduced camera system S dc = gc for a CHAIN graph by LAPACK's banded Cholesky (dpbtrf / dpbtrs
    through scipy.linalg): the unknowns are re-ordered node-major ([x_i (6), v_i (3)] per keyframe), which makes S banded
    with half-bandwidth (track length) x 9, plus the dense 6-wide border of the shared bias, eliminated by block
    elimination (6 extra right-hand sides).  This is the elimination order a fill-reducing ordering finds for a chain; it
    is what lets the CPU arm solve the 100 000-pose graph at all (SuperLU runs out of memory there).  Returns None when the
    graph is not a chain (loop closures put entries outside the band) -- the caller falls back to the general sparse solve.
    Column order of S / gc: the oracle's camera order [b | v | x]."""
    import scipy.linalg as sla
    nb, nv, nx = lay.nb, lay.nv, lay.nx
    if nb > 1 or nv not in (0, nx):
        return None
    D = 9 if nv else 6
    ncam = D * nx
    # position of oracle camera column c in the node-major order (bias last)
    perm = np.empty(6 * nb + 3 * nv + 6 * nx, dtype=np.int64)
    perm[:6 * nb] = ncam + np.arange(6 * nb)
    if nv:
        perm[6 * nb:6 * nb + 3 * nv] = (D * np.arange(nv)[:, None] + 6 + np.arange(3)[None, :]).ravel()
    perm[6 * nb + 3 * nv:] = (D * np.arange(nx)[:, None] + np.arange(6)[None, :]).ravel()
    C = S.tocoo()
    r, c = perm[C.row], perm[C.col]
    cam = (r < ncam) & (c < ncam)
    low = cam & (r >= c)
    hb = int((r[low] - c[low]).max()) if np.any(low) else 0
    if hb > max_halfband:
        return None
    ab = np.zeros((hb + 1, ncam))
    np.add.at(ab, (r[low] - c[low], c[low]), C.data[low])
    rhs = np.zeros((ncam, 1 + 6 * nb))
    g2 = np.empty(ncam + 6 * nb)
    g2[perm] = gc
    rhs[:, 0] = g2[:ncam]
    Hbb = np.zeros((6 * nb, 6 * nb))
    if nb:
        bor = (r < ncam) & (c >= ncam)
        np.add.at(rhs, (r[bor], 1 + c[bor] - ncam), C.data[bor])
        bb = (r >= ncam) & (c >= ncam)
        np.add.at(Hbb, (r[bb] - ncam, c[bb] - ncam), C.data[bb])
    cb = sla.cholesky_banded(ab, lower=True, overwrite_ab=True, check_finite=False)
    Z = sla.cho_solve_banded((cb, True), rhs, overwrite_b=True, check_finite=False)
    d2 = np.empty(ncam + 6 * nb)
    if nb:
        F = np.zeros((ncam, 6 * nb))
        np.add.at(F, (r[bor], c[bor] - ncam), C.data[bor])
        Sb = Hbb - F.T @ Z[:, 1:]
        db = np.linalg.solve(Sb, g2[ncam:] - F.T @ Z[:, 0])
        d2[:ncam] = Z[:, 0] - Z[:, 1:] @ db
        d2[ncam:] = db
    else:
        d2[:ncam] = Z[:, 0]
    return d2[perm]


def solve_damped(J, b, lam, lay, schur=True, banded=None):
    """Exact solve of (J^T J + lam I) d = J^T b.  schur=True eliminates landmarks first (same system).
    banded: solve the reduced camera system by banded Cholesky (_solve_banded_bordered) -- None = when it has more than
    200 000 unknowns (where SuperLU stops fitting); the parity tests check both routes against each other."""
    H = (J.T @ J).tocsc()
    g = J.T @ b
    n = lay.n
    H = H + lam * sp.identity(n, format='csc')
    if banded is None:
        banded = (n - 3 * lay.nl) > 200000
    if not schur or lay.nl == 0:
        if banded:
            d = _solve_banded_bordered(H, g, lay)
            if d is not None:
                return d
        return _splu_sym(H).solve(g)
    l0, l1 = lay.ol, lay.ov
    cam = np.concatenate([np.arange(0, l0), np.arange(l1, n)])
    lm = np.arange(l0, l1)
    Hc = H[cam][:, cam]
    E = H[cam][:, lm]
    C = H[lm][:, lm].tocsr()
    # C is block diagonal 3x3: invert blockwise
    nl = lay.nl
    Cd = np.zeros((nl, 3, 3))
    Cc = C.tocoo()
    Cd[Cc.row // 3, Cc.row % 3, Cc.col % 3] = Cc.data
    Ci = np.linalg.inv(Cd)
    ii = (3 * np.arange(nl)[:, None, None] + np.arange(3)[None, :, None]) + np.zeros((1, 1, 3), dtype=np.int64)
    jj = (3 * np.arange(nl)[:, None, None] + np.arange(3)[None, None, :]) + np.zeros((1, 3, 1), dtype=np.int64)
    Cinv = sp.csc_matrix((Ci.ravel(), (ii.ravel(), jj.ravel())), shape=(3 * nl, 3 * nl))
    ECi = (E @ Cinv).tocsc()
    S = (Hc - ECi @ E.T).tocsc()
    gc = g[cam] - ECi @ g[lm]
    dc = _solve_banded_bordered(S, gc, lay) if banded else None
    if dc is None:
        dc = _splu_sym(S).solve(gc)
    dl = Cinv @ (g[lm] - E.T @ dc)
    d = np.empty(n)
    d[cam] = dc
    d[lm] = dl
    return d


def lm_optimize(prob, params=None, schur=True, verbose=False):
    """GTSAM LM, verbatim control flow.  Returns (values, info)."""
    p = dict(LM_DEFAULTS)
    p.update(params or {})
    lay = Layout(prob)
    vals = values_of(prob)
    lam = p['lambdaInitial']
    err = graph_error(prob, vals)
    trace = dict(errors=[err], lambdas=[lam], tries=[], t_linearize=0.0, t_solve=0.0, t_error=0.0)
    iterations = 0
    info = dict(trace=trace)
    if err <= p['errorTol'] or p['maxIterations'] <= 0 or not np.isfinite(err):
        info.update(iterations=0, error=err, lam=lam)
        return vals, info
    while True:
        cur = err
        # ---- iterate(): linearize once, retry lambda
        t0 = time.perf_counter()
        J, b = linearize(prob, vals, lay)
        trace['t_linearize'] += time.perf_counter() - t0
        while True:
            t0 = time.perf_counter()
            try:
                delta = solve_damped(J, b, lam, lay, schur=schur)
                solved = bool(np.all(np.isfinite(delta)))
            except RuntimeError:
                solved = False
            trace['t_solve'] += time.perf_counter() - t0
            success = False
            stop = False
            new_err = np.inf
            if solved:
                old_lin = 0.5 * float(b @ b)
                rl = J @ delta - b
                new_lin = 0.5 * float(rl @ rl)
                lin_change = old_lin - new_lin
                if lin_change >= 0:
                    t0 = time.perf_counter()
                    new_vals = retract(vals, delta, lay)
                    new_err = graph_error(prob, new_vals)
                    trace['t_error'] += time.perf_counter() - t0
                    cost_change = err - new_err
                    if lin_change > np.finfo(float).eps * old_lin:
                        fidelity = cost_change / lin_change
                        success = fidelity > p['minModelFidelity']
                    if abs(cost_change) < p['relativeErrorTol'] * err:
                        stop = True
            trace['tries'].append(dict(lam=lam, solved=solved, success=success, new_err=float(new_err)))
            if verbose:
                print(f"  try lam={lam:.3e} solved={solved} new_err={new_err:.9e} success={success}")
            if success:
                vals = new_vals
                err = new_err
                lam = max(p['lambdaLowerBound'], lam / p['lambdaFactor'])
                iterations += 1
                break
            elif not stop:
                lam *= p['lambdaFactor']
                if lam >= p['lambdaUpperBound']:
                    break
            else:
                break
        trace['errors'].append(err)
        trace['lambdas'].append(lam)
        if verbose:
            print(f"iter {iterations} err={err:.12e} lam={lam:.3e}")
        # ---- checkConvergence
        if iterations >= p['maxIterations'] or not np.isfinite(cur):
            break
        if err <= p['errorTol']:
            break
        absdec = cur - err
        reldec = absdec / cur
        if (p['relativeErrorTol'] and reldec <= p['relativeErrorTol']) or absdec <= p['absoluteErrorTol']:
            break
    info.update(iterations=iterations, error=err, lam=lam)
    return vals, info


def marginal_covariance(prob, vals, queries):
    """gtsam::Marginals(graph, values).jointMarginalCovariance(keys).fullMatrix()  (gtsam/nonlinear/Marginals.cpp:
    linearize at `values`, eliminate, read the marginal off the Bayes tree) restated as the corresponding block of
    (J^T J)^-1, computed column by column with an exact sparse solve.  queries = [(kind, index)], kind in 'x','v','b','l';
    the tangent order of a Pose3 block is gtsam's (rotation, translation)."""
    lay = Layout(prob)
    J, _ = linearize(prob, vals, lay)
    lu = _splu_sym((J.T @ J).tocsc())
    cols = []
    for kind, idx in queries:
        c0, d = lay.cols(kind, [idx])
        cols.extend(range(int(c0[0]), int(c0[0]) + d))
    cols = np.asarray(cols)
    cov = np.empty((len(cols), len(cols)))
    for j, c in enumerate(cols):
        e = np.zeros(lay.n)
        e[c] = 1.0
        cov[:, j] = lu.solve(e)[cols]
    return 0.5 * (cov + cov.T)
