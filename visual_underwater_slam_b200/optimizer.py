"""gtsam.LevenbergMarquardtParams / gtsam.LevenbergMarquardtOptimizer facade (batch.py:337) and the
marshaller that hands the packed problem tables to the C-ABI (include/vus.h).

    results = LevenbergMarquardtOptimizer(graph, initial, LevenbergMarquardtParams()).optimize()

Host side only transposes the tables to component-major structure-of-arrays, passes pointers and
reads results back; PyTorch tensors are accepted as device-buffer carriers (`Session(..., device_tensors=True)`).
"""
import ctypes as C
import numpy as np
from . import _native, config
from ._native import LmParams, LmResult
from .values import Values

FACTOR_TYPES = ("prior_pose", "prior_vel", "between", "dvl", "stereo", "imu")
_SLOTS = {"prior_pose": ("x",), "prior_vel": ("v",), "between": ("x1", "x2"), "dvl": ("v", "x"),
          "stereo": ("x", "l"), "imu": ("xi", "vi", "xj", "vj", "b")}
_ROWS = {"prior_pose": 6, "prior_vel": 3, "between": 6, "dvl": 3, "stereo": 3, "imu": 9}
_COLS = {"prior_pose": 6, "prior_vel": 3, "between": 12, "dvl": 9, "stereo": 9, "imu": 24}
_MEAS = {"prior_pose": 12, "prior_vel": 3, "between": 12, "dvl": 3, "stereo": 3, "imu": 67}
_INFO = {"prior_pose": 6, "prior_vel": 3, "between": 6, "dvl": 3, "stereo": 3, "imu": 45}
_KINDS = (("pose", "pose_keys", "poses", 12), ("vel", "vel_keys", "vels", 3), ("bias", "bias_keys", "biases", 6),
          ("lm", "lm_keys", "lms", 3))


def _set(name):
    def f(self, v):
        setattr(self, name, v)
    return f


def _get(name):
    def f(self):
        return getattr(self, name)
    return f


class LevenbergMarquardtParams:
    """gtsam::LevenbergMarquardtParams with gtsam's setter/getter names and defaults."""

    def __init__(self):
        self.maxIterations = 100
        self.relativeErrorTol = 1e-5
        self.absoluteErrorTol = 1e-5
        self.errorTol = 0.0
        self.lambdaInitial = 1e-5
        self.lambdaFactor = 10.0
        self.lambdaUpperBound = 1e5
        self.lambdaLowerBound = 0.0
        self.minModelFidelity = 1e-3
        self.diagonalDamping = False
        self.useFixedLambdaFactor = True
        self.verbosityLM = "SILENT"
        # solver knobs of this implementation (gtsam: linearSolverType / ordering)
        self.pcgMaxIterations = 500
        self.pcgRelTol = 1e-12
        self.maxSupernode = 0
        self.profileKernels = False
        self.bandChunks = 0           # band factorization: 0 auto, n > 0 chunks, < 0 plain cyclic reduction (include/vus.h)

    setMaxIterations, getMaxIterations = _set("maxIterations"), _get("maxIterations")
    setRelativeErrorTol, getRelativeErrorTol = _set("relativeErrorTol"), _get("relativeErrorTol")
    setAbsoluteErrorTol, getAbsoluteErrorTol = _set("absoluteErrorTol"), _get("absoluteErrorTol")
    setErrorTol, getErrorTol = _set("errorTol"), _get("errorTol")
    setlambdaInitial, getlambdaInitial = _set("lambdaInitial"), _get("lambdaInitial")
    setlambdaFactor, getlambdaFactor = _set("lambdaFactor"), _get("lambdaFactor")
    setlambdaUpperBound, getlambdaUpperBound = _set("lambdaUpperBound"), _get("lambdaUpperBound")
    setlambdaLowerBound, getlambdaLowerBound = _set("lambdaLowerBound"), _get("lambdaLowerBound")

    def setDiagonalDamping(self, flag):
        if flag:
            raise NotImplementedError("diagonalDamping=True is not on the reference path (batch.py:337 uses defaults)")

    def setUseFixedLambdaFactor(self, flag):
        if not flag:
            raise NotImplementedError("useFixedLambdaFactor=False is not on the reference path")

    def setVerbosityLM(self, v):
        self.verbosityLM = v

    def to_c(self):
        p = LmParams()
        p.max_iterations = int(self.maxIterations)
        p.relative_error_tol = float(self.relativeErrorTol)
        p.absolute_error_tol = float(self.absoluteErrorTol)
        p.error_tol = float(self.errorTol)
        p.lambda_initial = float(self.lambdaInitial)
        p.lambda_factor = float(self.lambdaFactor)
        p.lambda_upper_bound = float(self.lambdaUpperBound)
        p.lambda_lower_bound = float(self.lambdaLowerBound)
        p.min_model_fidelity = float(self.minModelFidelity)
        p.pcg_max_iterations = int(self.pcgMaxIterations)
        p.pcg_rel_tol = float(self.pcgRelTol)
        p.max_supernode = int(self.maxSupernode)
        p.profile_kernels = int(bool(self.profileKernels))
        p.band_chunks = int(self.bandChunks)
        p.verbose = {"SILENT": 0, "SUMMARY": 1, "TERMINATION": 1, "LAMBDA": 1, "TRYLAMBDA": 1, "TRYCONFIG": 2,
                     "DAMPED": 2, "TRYDELTA": 2}.get(str(self.verbosityLM).upper(), 0)
        return p


def _rows(a, dim):
    """[n, dim] host table as a C-contiguous float64 array (no copy when it already is one): the library takes
    row-major tables as they are and transposes them on the device (VUS_MEM_*_ROWS)."""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1, dim))


class Session:
    """One vus_handle: uploads a packed problem (graph.to_problem), analyses it, exposes the C-ABI calls."""

    def __init__(self, prob, params=None, lib=None, device=0, device_tensors=False, partition=None, comm=None, components=None):
        self.lib = lib if lib is not None else _native.load()
        self.prob = prob
        self._h = C.c_void_p()
        rc = self.lib.vus_create(int(device), C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"vus_create failed ({rc}): no usable CUDA device {device}; this path has no CPU fallback")
        self._keep = []
        self.n = {}
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        try:
            self._upload(prob, device_tensors)
            self.set_params(params or LevenbergMarquardtParams())
            if partition is not None:                   # (n_owned_nodes, [owned factors per type]) -- parallel.py
                nf = (C.c_int64 * 6)(*[int(x) for x in partition[1]])
                self._check(self.lib.vus_set_partition(self._h, int(partition[0]), nf))
                if len(partition) > 2 and partition[2] is not None:   # (prev_local[16], next_local[16], rank, nranks): chain position
                    pl, nl_, rk, nr = partition[2]
                    pl = np.ascontiguousarray(pl, dtype=np.int64); nl_ = np.ascontiguousarray(nl_, dtype=np.int64)
                    self._check(self.lib.vus_set_partition_chain(self._h, len(pl), pl.ctypes.data_as(_native.c_i64_p),
                                                                 nl_.ctypes.data_as(_native.c_i64_p), int(rk), int(nr)))
            if comm is not None:                        # a _native.COMM_FN instance (kept alive by the caller and here)
                self._comm = comm
                self._check(self.lib.vus_set_comm(self._h, C.cast(comm, C.c_void_p), None))
            self.n_components = 1
            if components is not None:                  # node_start [ncomp + 1]: independent trajectories (parallel.solve_batched)
                ns = np.ascontiguousarray(components, dtype=np.int64)
                self.n_components = len(ns) - 1
                self._check(self.lib.vus_set_components(self._h, self.n_components, ns.ctypes.data_as(_native.c_i64_p)))
            self._check(self.lib.vus_analyze(self._h))
        except Exception:
            self.close()
            raise

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.vus_last_error(self._h).decode() or f"libvus error {rc}")

    def close(self):
        if self._h:
            self.lib.vus_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _table(self, arr, device_tensors):
        """-> (pointer, mem flag); keeps the carrier alive."""
        self.h2d_bytes += arr.nbytes
        if device_tensors:
            import torch
            t = torch.from_numpy(arr).cuda()
            self._keep.append(t)
            return C.c_void_p(t.data_ptr()), 3          # VUS_MEM_DEVICE_ROWS
        self._keep.append(arr)
        return arr.ctypes.data_as(C.c_void_p), 2        # VUS_MEM_HOST_ROWS

    def _upload(self, prob, device_tensors):
        lib = self.lib
        for kind, (name, kname, dname, dim) in enumerate(_KINDS):
            keys = np.ascontiguousarray(prob[kname], dtype=np.uint64)
            data = _rows(prob[dname], dim)
            self.n[name] = len(keys)
            ptr, mem = self._table(data, device_tensors)
            self._check(lib.vus_set_variables(self._h, kind, len(keys), keys.ctypes.data_as(_native.c_u64_p), ptr, mem))
        K = np.ascontiguousarray(prob["calib"], dtype=np.float64)
        g = np.ascontiguousarray(prob["gravity"], dtype=np.float64)
        self._check(lib.vus_set_calibration(self._h, K.ctypes.data_as(_native.c_double_p)))
        self._check(lib.vus_set_gravity(self._h, g.ctypes.data_as(_native.c_double_p)))
        opt = dict(config.gtsam_build())
        opt.update(prob.get("options") or {})
        self.options = opt
        self._check(lib.vus_set_gtsam_build(self._h, int(opt["tangent_preintegration"]), int(opt["slow_but_correct_betweenfactor"])))
        self.nf = {}
        for t, name in enumerate(FACTOR_TYPES):
            f = prob[name]
            n = len(f["orig"])
            self.nf[name] = n
            if n == 0:
                continue
            idx = np.ascontiguousarray(np.stack([f[s] for s in _SLOTS[name]], 0), dtype=np.int32)
            meas = _rows(f["meas"], _MEAS[name])
            info = _rows(f["sqrt_info"], _INFO[name])
            orig = np.ascontiguousarray(f["orig"], dtype=np.int64)
            pm, mem = self._table(meas, device_tensors)
            pi, _ = self._table(info, device_tensors)
            self.h2d_bytes += idx.nbytes
            self._check(lib.vus_add_factors(self._h, t, n, idx.ctypes.data_as(_native.c_i32_p), pm, pi,
                                            orig.ctypes.data_as(_native.c_i64_p), mem))
        self.n_factors = sum(self.nf.values())

    def comm_init(self, unique_id, rank, nranks):
        """Collectives issued by the library itself: NCCL communicator from a 128-byte id every rank received (vus_comm_init)."""
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        self._check(self.lib.vus_comm_init(self._h, buf, int(rank), int(nranks)))

    def set_halo(self, peers, send_lists, recv_ranges):
        """Halo exchange lists of a pose-range partition (vus_set_halo): peers [q], send_lists[q] = local indices of the owned
        nodes peer q needs, recv_ranges[q] = (offset, count) of peer q's nodes in this rank's halo."""
        peers = np.ascontiguousarray(peers, dtype=np.int32)
        ptr = np.zeros(len(peers) + 1, dtype=np.int64)
        for q, ix in enumerate(send_lists):
            ptr[q + 1] = ptr[q] + len(ix)
        idx = np.ascontiguousarray(np.concatenate([np.asarray(ix, dtype=np.int32) for ix in send_lists]) if len(peers) and ptr[-1] else np.zeros(0, np.int32))
        off = np.ascontiguousarray([r[0] for r in recv_ranges], dtype=np.int64)
        cnt = np.ascontiguousarray([r[1] for r in recv_ranges], dtype=np.int64)
        self._check(self.lib.vus_set_halo(self._h, len(peers), peers.ctypes.data_as(_native.c_i32_p), ptr.ctypes.data_as(_native.c_i64_p),
                                          idx.ctypes.data_as(_native.c_i32_p), off.ctypes.data_as(_native.c_i64_p), cnt.ctypes.data_as(_native.c_i64_p)))

    def set_params(self, params):
        p = params.to_c() if hasattr(params, "to_c") else params
        self._check(self.lib.vus_set_lm_params(self._h, C.byref(p)))

    # ------------------------------------------------------------------ C-ABI calls
    def layout(self):
        out = (C.c_int64 * 12)()
        self._check(self.lib.vus_get_layout(self._h, out))
        return dict(D=out[0], k=out[1], Ns=out[2], nrem=out[3], ndst=out[4], B=out[5], L=out[6], n_factors=out[7],
                    band_chunks=out[8], Ns_band=out[9], n_landmarks=out[10])

    def error(self, stream=None):
        v = C.c_double()
        self._check(self.lib.vus_error(self._h, stream, C.byref(v)))
        return v.value

    def factor_errors(self, stream=None):
        out = np.zeros(self.n_factors)
        self._check(self.lib.vus_factor_errors(self._h, stream, out.ctypes.data_as(_native.c_double_p)))
        return out

    def linearize(self, name, stream=None):
        """-> (r [n,m], J [n,m,cols]) whitened, node-ordered columns (include/vus.h)."""
        n, m, c = self.nf[name], _ROWS[name], _COLS[name]
        r = np.zeros((m, n))
        J = np.zeros((m * c, n))
        self._check(self.lib.vus_linearize(self._h, stream, FACTOR_TYPES.index(name), r.ctypes.data_as(_native.c_double_p),
                                           J.ctypes.data_as(_native.c_double_p)))
        return r.T.copy(), J.reshape(m, c, n).transpose(2, 0, 1).copy()

    def solve_step(self, lam, stream=None):
        dp = np.zeros((self.n["pose"], 6))
        dv = np.zeros((self.n["vel"], 3))
        db = np.zeros((max(self.n["bias"], 1), 6))
        dl = np.zeros((self.n["lm"], 3))
        its = C.c_int32()
        P = _native.c_double_p
        self._check(self.lib.vus_solve_step(self._h, stream, float(lam), dp.ctypes.data_as(P), dv.ctypes.data_as(P),
                                            db.ctypes.data_as(P), dl.ctypes.data_as(P), C.byref(its)))
        return dict(pose=dp, vel=dv, bias=db[:self.n["bias"]], lm=dl, pcg_iterations=its.value)

    def debug_band_solve(self, lam, rhs, stream=None):
        """Test hook: -> (SD [Ns,B,B], SU [Ns-1,B,B], x [nrhs, Ns*B], pivot_failed) for the damped band at the current values."""
        lay = self.layout()
        Ns, B = lay["Ns"], lay["B"]
        SD = np.zeros((Ns, B, B))
        SU = np.zeros((max(Ns - 1, 1), B, B))
        x = np.array(rhs, dtype=np.float64).reshape(-1, Ns * B).copy()
        P = _native.c_double_p
        rc = self.lib.vus_debug_band_solve(self._h, stream, float(lam), SD.ctypes.data_as(P), SU.ctypes.data_as(P),
                                           x.ctypes.data_as(P), x.shape[0])
        if rc < 0:
            self._check(rc)
        return SD, SU[:Ns - 1], x, bool(rc)

    def marginal_covariance(self, queries, stream=None):
        """Joint covariance of [(kind, index)] (kind in 'pose','vel','bias','lm') at the current values: dense [M, M]."""
        kinds = np.ascontiguousarray([("pose", "vel", "bias", "lm").index(k) for k, _ in queries], dtype=np.int32)
        idx = np.ascontiguousarray([i for _, i in queries], dtype=np.int32)
        M = int(sum((6, 3, 6, 3)[k] for k in kinds))
        out = np.zeros((M, M))
        self._check(self.lib.vus_marginal_covariance(self._h, stream, len(kinds), kinds.ctypes.data_as(_native.c_i32_p),
                                                     idx.ctypes.data_as(_native.c_i32_p), out.ctypes.data_as(_native.c_double_p)))
        return out

    def optimize(self, stream=None):
        res = LmResult()
        self._check(self.lib.vus_optimize(self._h, stream, C.byref(res)))
        return res.as_dict()

    def trace(self):
        """Try-by-try history of the last optimize(): list of dict(lam, solved, success, new_err, pcg_iterations)."""
        n = C.c_int32()
        self._check(self.lib.vus_get_trace(self._h, 0, None, C.byref(n)))
        arr = (_native.LmTry * max(n.value, 1))()
        self._check(self.lib.vus_get_trace(self._h, n.value, arr, C.byref(n)))
        return [dict(lam=t.lambda_, solved=bool(t.solved), success=bool(t.accepted), new_err=t.new_error,
                     pcg_iterations=t.pcg_iterations) for t in arr[:n.value]]

    def component_results(self):
        """Per-trajectory summary of the last batched optimize(): list of dicts (vus_component_result)."""
        arr = (_native.ComponentResult * self.n_components)()
        self._check(self.lib.vus_get_component_results(self._h, arr))
        return [{n: getattr(r, n) for n, _ in _native.ComponentResult._fields_} for r in arr]

    def save_values(self):
        self._check(self.lib.vus_save_values(self._h))

    def restore_values(self):
        self._check(self.lib.vus_restore_values(self._h))

    def time_linearize(self, reps=10, stream=None):
        v = C.c_double()
        self._check(self.lib.vus_time_linearize(self._h, stream, int(reps), C.byref(v)))
        return v.value

    def values(self):
        """Current values as host tables {name: [n, dim]}."""
        out = {}
        self.d2h_bytes = 0
        for kind, (name, kname, dname, dim) in enumerate(_KINDS):
            buf = np.empty((self.n[name], dim))
            self._check(self.lib.vus_get_variables(self._h, kind, buf.ctypes.data_as(C.c_void_p), 2))
            out[dname] = buf
            self.d2h_bytes += buf.nbytes
        return out


def _values_from(prob, tables):
    return Values.from_tables({"pose": (prob["pose_keys"], tables["poses"]), "vel": (prob["vel_keys"], tables["vels"]),
                               "bias": (prob["bias_keys"], tables["biases"]), "lm": (prob["lm_keys"], tables["lms"])})


def graph_error(graph, values, lib=None):
    s = Session(graph.to_problem(values), lib=lib)
    try:
        return s.error()
    finally:
        s.close()


class LevenbergMarquardtOptimizer:
    """gtsam.LevenbergMarquardtOptimizer(graph, initialValues, params) (batch.py:337)."""

    def __init__(self, graph, initialValues, params=None, lib=None):
        self._graph = graph
        self._params = params or LevenbergMarquardtParams()
        self._prob = graph.to_problem(initialValues)
        self._session = Session(self._prob, self._params, lib=lib)
        self._result = None
        self._values = initialValues
        self._error = None

    def optimize(self):
        self._result = self._session.optimize()
        self._values = _values_from(self._prob, self._session.values())
        self._error = self._result["final_error"]
        return self._values

    def optimizeSafely(self):
        return self.optimize()

    def values(self):
        return self._values

    def error(self):
        if self._error is None:
            self._error = self._session.error()
        return self._error

    def iterations(self):
        return 0 if self._result is None else self._result["iterations"]

    def lambda_(self):
        return self._params.lambdaInitial if self._result is None else self._result["final_lambda"]

    def stats(self):
        """Per-phase timings, PCG iterations, kernel launches and the try-by-try lambda / error history of the last optimize()."""
        d = dict(self._result or {})
        if self._result is not None:
            d["trace"] = self._session.trace()
        return d


def optimize_many(graphs, initial_values, params=None, lib=None, device=0):
    """N independent `LevenbergMarquardtOptimizer(graph_t, initial_t, params).optimize()` calls (batch.py:337 once per
    trajectory; BASELINE config 4) as ONE batched solve on the device: -> (list of Values, list of per-trajectory summaries
    with gtsam's iterations() / error() / lambda_()).  Every trajectory keeps its own LM path (include/vus.h,
    vus_set_components)."""
    from . import parallel
    graphs, initial_values = list(graphs), list(initial_values)
    if len(graphs) != len(initial_values) or not graphs:
        raise ValueError("optimize_many: one initial Values per graph, at least one graph")
    probs = [g.to_problem(v) for g, v in zip(graphs, initial_values)]
    res = parallel.solve_batched(probs, params or LevenbergMarquardtParams(), lib=lib, device=device, keep_values=True)
    out = []
    for p, r in zip(probs, res):
        out.append(_values_from(p, dict(poses=r["values"]["poses"], vels=r["values"]["vels"], biases=r["values"]["biases"],
                                        lms=r["values"]["lms"])))
    summaries = [dict(iterations=r["iterations"], error=r["final_error"], lambda_=r["final_lambda"],
                      inner_iterations=r["inner_iterations"], initial_error=r["initial_error"]) for r in res]
    return out, summaries


class JointMarginal:
    """gtsam.JointMarginal: the joint covariance (or information) of several variables, blocks in the order asked."""

    def __init__(self, keys, dims, full):
        self._keys = list(keys)
        self._off = np.concatenate([[0], np.cumsum(dims)]).astype(int)
        self._full = full

    def fullMatrix(self):
        return self._full

    def at(self, k1, k2):
        i, j = self._keys.index(int(k1)), self._keys.index(int(k2))
        return self._full[self._off[i]:self._off[i + 1], self._off[j]:self._off[j + 1]]


class Marginals:
    """gtsam.Marginals(graph, solution): marginal covariances of the optimised variables -- what a user of
    batch.py:337's result asks for next.  Linearizes at `solution`; every covariance column is one exact undamped solve
    on the device (vus_marginal_covariance)."""

    def __init__(self, graph, solution, lib=None):
        self._prob = graph.to_problem(solution)
        self._session = Session(self._prob, lib=lib)
        self._where = {}
        for name, kname in (("pose", "pose_keys"), ("vel", "vel_keys"), ("bias", "bias_keys"), ("lm", "lm_keys")):
            for i, k in enumerate(self._prob[kname]):
                self._where[int(k)] = (name, i)

    def _query(self, key):
        try:
            return self._where[int(key)]
        except KeyError:
            raise KeyError(f"Marginals: key {int(key)} is not in the solution") from None

    def marginalCovariance(self, key):
        return self._session.marginal_covariance([self._query(key)])

    def marginalInformation(self, key):
        return np.linalg.inv(self.marginalCovariance(key))

    def jointMarginalCovariance(self, keys):
        q = [self._query(k) for k in keys]
        dims = [dict(pose=6, vel=3, bias=6, lm=3)[k] for k, _ in q]
        return JointMarginal([int(k) for k in keys], dims, self._session.marginal_covariance(q))

    def jointMarginalInformation(self, keys):
        jm = self.jointMarginalCovariance(keys)
        return JointMarginal(jm._keys, np.diff(jm._off), np.linalg.inv(jm.fullMatrix()))


# ---------------------------------------------------------------------------------------------- front-end rows
def preintegrate_imu(acc, gyro, dt, params, bias_hat=None, lib=None, device=0, stream=None, tangent=None):
    """Device version of navigation.preintegrate_batch (the integrateMeasurement loop of batch.py:289-293):
    acc, gyro [n, k, 3], constant dt -> (pim [n, 67], sqrt_info_triu [n, 45]) ready for graph.add_imu_factors.
    tangent: preintegration variant (None = config.gtsam_build())."""
    lib = lib if lib is not None else _native.load()
    acc = np.ascontiguousarray(acc, dtype=np.float64)
    gyro = np.ascontiguousarray(gyro, dtype=np.float64)
    n, k, _ = acc.shape
    if gyro.shape != acc.shape:
        raise ValueError("preintegrate_imu: acc and gyro must both be [n, k, 3]")
    h = C.c_void_p()
    if lib.vus_create(int(device), C.byref(h)) != 0:
        raise RuntimeError("vus_create failed: no usable CUDA device; this path has no CPU fallback")
    try:
        tan = config.gtsam_build()["tangent_preintegration"] if tangent is None else bool(tangent)
        lib.vus_set_gtsam_build(h, int(tan), 0)
        pim = np.empty((n, 67))
        info = np.empty((n, 45))
        P = _native.c_double_p
        b = np.zeros(6) if bias_hat is None else np.ascontiguousarray(bias_hat, dtype=np.float64).reshape(6)
        cov = [np.ascontiguousarray(c, dtype=np.float64).reshape(9) for c in
               (params.accelerometerCovariance, params.gyroscopeCovariance, params.integrationCovariance)]
        rc = lib.vus_preintegrate_imu(h, stream, n, k, acc.ctypes.data_as(C.c_void_p), gyro.ctypes.data_as(C.c_void_p), float(dt),
                                      b.ctypes.data_as(P), cov[0].ctypes.data_as(P), cov[1].ctypes.data_as(P), cov[2].ctypes.data_as(P),
                                      pim.ctypes.data_as(C.c_void_p), info.ctypes.data_as(C.c_void_p), 2)
        if rc != 0:
            raise RuntimeError(lib.vus_last_error(h).decode() or f"libvus error {rc}")
        return pim, info
    finally:
        lib.vus_destroy(h)


def backproject_stereo(session, pose_idx, meas, stream=None):
    """Landmark initial values from stereo measurements (get_landmarks, batch.py:144-176) at the session's current
    poses and calibration: pose_idx [n] (indices into the pose table), meas [n, 3] (uL, uR, v) -> points [n, 3]."""
    idx = np.ascontiguousarray(pose_idx, dtype=np.int32)
    z = np.ascontiguousarray(meas, dtype=np.float64).reshape(len(idx), 3)
    out = np.empty((len(idx), 3))
    session._check(session.lib.vus_backproject_stereo(session._h, stream, len(idx), idx.ctypes.data_as(_native.c_i32_p),
                                                      z.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), 2))
    return out
