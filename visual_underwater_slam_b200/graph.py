"""gtsam.NonlinearFactorGraph facade + the packer that turns graph + Values into plain tables.

Reference call sites: batch.py:80, :272 (construction), :281-282 (.add), :291-292, :305
(.push_back), :338 (.saveGraph).  Factors keep their NonlinearFactorGraph insertion index
(`orig`) so factor indexing stays bit-identical to the reference order documented in SURVEY.md
3.1: [PriorPose3, PriorVector, then per pose: Imu_i, Dvl_i, Stereo_{i,*}].

`to_problem(values)` emits the "problem" dict of numpy arrays (keys sorted ascending per kind,
per-type structure-of-arrays factor tables holding int32 variable indices) that both the C-ABI
marshaller (optimizer.py) and the test oracle consume.  The bulk `add_*_factors` methods are
the vectorised construction API that bypasses the per-factor Python loop.
"""
import numpy as np
from . import config
from . import factors as F
from .values import Values
from .symbol import symbolChr, symbolIndex

FACTOR_TYPES = ("prior_pose", "prior_vel", "between", "dvl", "stereo", "imu")
_SLOTS = {"prior_pose": ("x",), "prior_vel": ("v",), "between": ("x1", "x2"), "dvl": ("v", "x"),
          "stereo": ("x", "l"), "imu": ("xi", "vi", "xj", "vj", "b")}
_SLOT_KIND = {"x": "pose", "x1": "pose", "x2": "pose", "xi": "pose", "xj": "pose",
              "v": "vel", "vi": "vel", "vj": "vel", "l": "lm", "b": "bias"}
_MEAS_DIM = {"prior_pose": 12, "prior_vel": 3, "between": 12, "dvl": 3, "stereo": 3, "imu": 67}
_INFO_DIM = {"prior_pose": 6, "prior_vel": 3, "between": 6, "dvl": 3, "stereo": 3, "imu": 45}


class NonlinearFactorGraph:
    def __init__(self):
        self._n = 0
        self._chunks = {t: [] for t in FACTOR_TYPES}   # list of dict(keys [n,slots] u64, meas, sqrt_info, orig)
        self._pending = {t: [] for t in FACTOR_TYPES}  # per-object adds: (keys, meas, sqrt_info, orig)
        self._custom = []
        self.calib = None
        self.gravity = None
        self.imu_tangent = None      # preintegration variant of the ImuFactors in this graph (all alike; config.py)

    # ------------------------------------------------------------------ gtsam API
    def add(self, factor):
        if isinstance(factor, NonlinearFactorGraph):          # push_back(graph): isam.py:341 hands whole graphs over
            self._merge(factor)
            return
        if isinstance(factor, F.CustomFactor):
            lowered = factor.as_dvl()
            if lowered is None:
                self._custom.append((self._n, factor))
                self._n += 1
                return
            factor = lowered
        t = factor.ftype
        if t not in self._pending:
            raise TypeError(f"unsupported factor type {type(factor).__name__}")
        if t == "stereo":
            self._set_once("calib", factor.K, "Cal3_S2Stereo")
        if t == "imu":
            self._set_once("gravity", factor.gravity, "n_gravity")
            self._set_variant(factor.tangent)
        meas = factor.pim if t == "imu" else factor.meas
        self._pending[t].append((factor.keys(), meas, factor.sqrt_info, self._n))
        self._n += 1

    push_back = add

    def _merge(self, other):
        """Append every factor of `other`, keeping its insertion order after this graph's factors."""
        if other._custom:
            for idx, f in other._custom:
                self._custom.append((self._n + idx, f))
        if other.calib is not None:
            self._set_once("calib", other.calib, "Cal3_S2Stereo")
        if other.gravity is not None:
            self._set_once("gravity", other.gravity, "n_gravity")
        if other.imu_tangent is not None:
            self._set_variant(other.imu_tangent)
        for t in FACTOR_TYPES:
            tab = other.table(t)
            if len(tab["orig"]) == 0:
                continue
            self._flush(t)
            self._chunks[t].append(dict(keys=tab["keys"].copy(), meas=tab["meas"].copy(), sqrt_info=tab["sqrt_info"].copy(),
                                        orig=tab["orig"] + self._n))
        self._n += other._n

    def resize(self, n):
        """gtsam's graph.resize(0) (isam.py clears its staging graph after every update)."""
        if n != 0:
            raise NotImplementedError("NonlinearFactorGraph.resize: only resize(0) is supported")
        self.__init__()

    def size(self):
        return self._n

    def nrFactors(self):
        return self._n

    def _set_once(self, name, value, what):
        cur = getattr(self, name)
        value = np.asarray(value, dtype=np.float64)
        if cur is None:
            setattr(self, name, value.copy())
        elif not np.array_equal(cur, value):
            raise NotImplementedError(f"all factors must share one {what} on the B200 path")

    def _set_variant(self, tangent):
        if self.imu_tangent is None:
            self.imu_tangent = bool(tangent)
        elif self.imu_tangent != bool(tangent):
            raise NotImplementedError("all ImuFactors of a graph must come from the same preintegration variant "
                                      "(tangent / manifold; a gtsam build has only one)")

    # ------------------------------------------------------------------ bulk construction API
    def _add_bulk(self, t, keys, meas, sqrt_info):
        keys = np.asarray(keys, dtype=np.uint64).reshape(-1, len(_SLOTS[t]))
        n = keys.shape[0]
        meas = np.asarray(meas, dtype=np.float64).reshape(n, _MEAS_DIM[t])
        sqrt_info = np.broadcast_to(np.asarray(sqrt_info, dtype=np.float64), (n, _INFO_DIM[t]))
        self._flush(t)
        self._chunks[t].append(dict(keys=keys.copy(), meas=meas.copy(), sqrt_info=sqrt_info.copy(),
                                    orig=np.arange(self._n, self._n + n, dtype=np.int64)))
        self._n += n

    def add_prior_pose_factors(self, keys, poses12, sqrt_info):
        self._add_bulk("prior_pose", keys, poses12, sqrt_info)

    def add_prior_vector_factors(self, keys, vecs, sqrt_info):
        self._add_bulk("prior_vel", keys, vecs, sqrt_info)

    def add_between_factors(self, keys1, keys2, poses12, sqrt_info):
        self._add_bulk("between", np.stack([keys1, keys2], 1), poses12, sqrt_info)

    def add_dvl_factors(self, vkeys, xkeys, meas, sqrt_info):
        self._add_bulk("dvl", np.stack([vkeys, xkeys], 1), meas, sqrt_info)

    def add_stereo_factors(self, xkeys, lkeys, meas, sqrt_info, K):
        self._set_once("calib", np.asarray(K.vector() if hasattr(K, "vector") else K), "Cal3_S2Stereo")
        self._add_bulk("stereo", np.stack([xkeys, lkeys], 1), meas, sqrt_info)

    def add_imu_factors(self, xi, vi, xj, vj, b, pim, sqrt_info_triu, n_gravity, tangent=None):
        """pim rows in the layout of navigation.py; tangent: the variant they were preintegrated with
        (None = config.gtsam_build()["tangent_preintegration"], what preintegrate_batch uses by default)."""
        self._set_once("gravity", n_gravity, "n_gravity")
        self._set_variant(config.gtsam_build()["tangent_preintegration"] if tangent is None else tangent)
        self._add_bulk("imu", np.stack([xi, vi, xj, vj, b], 1), pim, sqrt_info_triu)

    def set_insertion_order(self, ftype, orig):
        """Override the insertion indices of the LAST bulk chunk of `ftype` (used by generators that
        interleave types like batch.py:291-305 but add them type by type)."""
        self._chunks[ftype][-1]["orig"] = np.asarray(orig, dtype=np.int64).copy()

    def _flush(self, t):
        pend = self._pending[t]
        if not pend:
            return
        self._chunks[t].append(dict(
            keys=np.array([p[0] for p in pend], dtype=np.uint64).reshape(len(pend), -1),
            meas=np.array([p[1] for p in pend], dtype=np.float64).reshape(len(pend), -1),
            sqrt_info=np.array([p[2] for p in pend], dtype=np.float64).reshape(len(pend), -1),
            orig=np.array([p[3] for p in pend], dtype=np.int64)))
        self._pending[t] = []

    def table(self, t):
        self._flush(t)
        ch = self._chunks[t]
        if not ch:
            return dict(keys=np.zeros((0, len(_SLOTS[t])), dtype=np.uint64), meas=np.zeros((0, _MEAS_DIM[t])),
                        sqrt_info=np.zeros((0, _INFO_DIM[t])), orig=np.zeros(0, dtype=np.int64))
        if len(ch) > 1:
            ch = [dict(keys=np.concatenate([c["keys"] for c in ch]), meas=np.concatenate([c["meas"] for c in ch]),
                       sqrt_info=np.concatenate([c["sqrt_info"] for c in ch]),
                       orig=np.concatenate([c["orig"] for c in ch]))]
            self._chunks[t] = ch
        return ch[0]

    # ------------------------------------------------------------------ packer
    def to_problem(self, values):
        """graph + Values -> dict of numpy tables (see module docstring)."""
        if self._custom:
            idx, f = self._custom[0]
            raise RuntimeError(
                f"factor {idx} is a gtsam.CustomFactor with an arbitrary Python callback; a Python callback cannot "
                "run inside a CUDA kernel and this path has no CPU fallback. Use DvlVelocityFactor (the native "
                "form of batch.py:241-250) or a built-in factor type.")
        prob = {}
        keytab = {}
        for kind, kname, dname in (("pose", "pose_keys", "poses"), ("vel", "vel_keys", "vels"),
                                   ("bias", "bias_keys", "biases"), ("lm", "lm_keys", "lms")):
            k, d = values.table(kind)
            prob[kname], prob[dname] = k, d
            keytab[kind] = k
        prob["calib"] = self.calib if self.calib is not None else np.array([1.0, 1.0, 0.0, 0.0, 0.0, 1.0])
        prob["gravity"] = self.gravity if self.gravity is not None else np.array([0.0, 0.0, -9.81])
        for t in FACTOR_TYPES:
            tab = self.table(t)
            out = dict(meas=tab["meas"], sqrt_info=tab["sqrt_info"], orig=tab["orig"])
            if t == "imu":
                out["pim"] = tab["meas"]
            for s, slot in enumerate(_SLOTS[t]):
                kind = _SLOT_KIND[slot]
                ks = tab["keys"][:, s]
                pos = np.searchsorted(keytab[kind], ks)
                pos_c = np.minimum(pos, max(len(keytab[kind]) - 1, 0))
                ok = (pos < len(keytab[kind])) & (keytab[kind][pos_c] == ks) if len(keytab[kind]) else np.zeros(len(ks), bool)
                if not np.all(ok):
                    bad = int(ks[np.nonzero(~ok)[0][0]])
                    raise RuntimeError(f"Attempting to retrieve the key \"{symbolChr(bad)}{symbolIndex(bad)}\" "
                                       f"(slot {slot} of a {t} factor), which does not exist in the Values as a {kind}.")
                out[slot] = pos.astype(np.int32)
            prob[t] = out
        prob["n_factors"] = self._n
        build = config.gtsam_build()
        prob["options"] = dict(
            tangent_preintegration=build["tangent_preintegration"] if self.imu_tangent is None else self.imu_tangent,
            slow_but_correct_betweenfactor=build["slow_but_correct_betweenfactor"])
        return prob

    # ------------------------------------------------------------------ evaluation (CUDA path)
    def error(self, values):
        """NonlinearFactorGraph::error -- evaluated by the CUDA kernels (no CPU fallback)."""
        from .optimizer import graph_error
        return graph_error(self, values)

    def saveGraph(self, path, values=None):
        """Graphviz dump (batch.py:338): variables as circles, factors as points."""
        lines = ["graph {", "  size=\"5,5\";", ""]
        seen = {}
        for t in FACTOR_TYPES:
            for ks in self.table(t)["keys"]:
                for k in ks.tolist():
                    if k not in seen:
                        seen[k] = len(seen)
        for k in sorted(seen):
            lines.append(f"  var{k}[label=\"{symbolChr(k)}{symbolIndex(k)}\"];")
        lines.append("")
        fid = {}
        for t in FACTOR_TYPES:
            tab = self.table(t)
            for ks, o in zip(tab["keys"], tab["orig"]):
                fid[int(o)] = ks.tolist()
        for o in sorted(fid):
            ks = fid[o]
            lines.append(f"  factor{o}[label=\"\", shape=point];")
            for k in ks:
                lines.append(f"  var{k}--factor{o};")
        lines.append("}")
        with open(path, "w") as fh:
            fh.write("\n".join(lines) + "\n")
