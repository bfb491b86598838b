"""Multi-GPU host logic for the batch-LM path: one process per GPU, torch.distributed for the plumbing
(NCCL on the GPUs, gloo in the CPU tests).

The reference solves one graph per process (batch.py:337).  The path shards by *independent trajectories*
(BASELINE.json config 4: a batch of independent 500-pose graphs): trajectory t goes to rank t // ceil(n / world),
every rank solves its shard with its own handles on its own GPU, and the only communication is one all-gather of the
per-trajectory summary (final error, iterations, lambda) at the end -- there is NO data-path collective, so scaling is
weak and the N-rank result is bit-identical to the 1-rank result for every trajectory.

Inside a rank the trajectories of a shard are small (launch-latency bound), so several handles run concurrently from a
thread pool, each on its handle's own CUDA stream (ctypes releases the GIL during the C-ABI calls; one host thread per
handle, as include/vus.h requires; a stream must never be shared by two handles that run at the same time -- the
library captures its fixed launch sequences into CUDA graphs on it).

The second mode (BASELINE.json config 5) splits ONE pose graph by contiguous pose range: `partition_pose_graph` builds
every rank's local graph ([owned | halo] poses, owned + duplicated cut factors) and its halo send / receive lists,
`PartitionedSolver` drives one handle per rank and supplies the two collectives the C-ABI asks for (vus_set_comm): the
all-reduce of the PCG / LM scalars and the halo exchange of node vectors, both through torch.distributed.
"""
import ctypes as C
from concurrent.futures import ThreadPoolExecutor
import numpy as np
from .optimizer import Session, LevenbergMarquardtParams

SUMMARY_FIELDS = ("final_error", "iterations", "inner_iterations", "final_lambda", "initial_error")


def shard_range(n_items, rank, world):
    """Contiguous block partition: -> (first, last_exclusive) of `rank`; blocks differ in size by at most one."""
    base, extra = divmod(int(n_items), int(world))
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


CHAIN_SIDE = 16      # poses either side of a rank's range handed to vus_set_partition_chain (>= the widest supernode of a pose graph)


def pose_range(n_poses, rank, world):
    """Pose range of `rank` in a pose-range partition: shard_range in units of CHAIN_SIDE poses, so that every rank but the
    last owns a whole number of supernodes whatever width (1, 2, 4, 8 or 16 poses) the analysis finds -- the condition for
    tying the per-rank band factorizations together exactly (csrc/spike.cuh).  Short graphs fall back to shard_range."""
    nb = -(-int(n_poses) // CHAIN_SIDE)
    if nb < 2 * world:
        return shard_range(n_poses, rank, world)
    a, b = shard_range(nb, rank, world)
    return min(a * CHAIN_SIDE, n_poses), min(b * CHAIN_SIDE, n_poses)


def pose_owner(pose, n_poses, world):
    """Inverse of pose_range."""
    nb = -(-int(n_poses) // CHAIN_SIDE)
    if nb < 2 * world:
        return owner_of(pose, n_poses, world)
    return owner_of(pose // CHAIN_SIDE, nb, world)


def owner_of(item, n_items, world):
    """Inverse of shard_range."""
    base, extra = divmod(int(n_items), int(world))
    cut = extra * (base + 1)
    return item // (base + 1) if item < cut else extra + (item - cut) // max(base, 1)


def solve_local(problems, params=None, lib=None, device=0, threads=4, keep_values=True):
    """Solve a list of independent packed problems (graph.to_problem) on this rank's GPU.
    -> list of dict(summary fields..., values=tables or None), in input order."""
    params = params or LevenbergMarquardtParams()

    def work(arg):
        i, prob = arg
        s = Session(prob, params, lib=lib, device=device)
        try:
            res = s.optimize()                          # every handle runs on its own (library-owned) CUDA stream
            out = {k: res[k] for k in SUMMARY_FIELDS}
            out["values"] = s.values() if keep_values else None
            return out
        finally:
            s.close()

    items = list(enumerate(problems))
    if threads <= 1 or len(items) <= 1:
        return [work(it) for it in items]
    with ThreadPoolExecutor(max_workers=threads) as pool:
        return list(pool.map(work, items))


def concat_problems(problems):
    """Concatenate independent packed problems (same variable kinds, no landmarks) into ONE packed problem whose
    components are the inputs: pose / velocity rows of problem t follow those of problem t-1, its bias (if any) becomes
    bias t, factor indices are shifted accordingly and insertion indices continue across problems.
    -> (problem, node_start [n + 1]).  Keys are re-issued as X(i) / V(i) / B(t) over the concatenated index."""
    from .symbol import X, V, B
    if not problems:
        raise ValueError("concat_problems: empty batch")
    n_nodes = [len(p["pose_keys"]) for p in problems]
    node_start = np.concatenate([[0], np.cumsum(n_nodes)]).astype(np.int64)
    for p in problems:
        if len(p["lm_keys"]) or len(p["stereo"]["orig"]):
            raise NotImplementedError("concat_problems: stereo landmarks are not supported in a batch")
        if len(p["vel_keys"]) not in (0, len(p["pose_keys"])) or len(p["bias_keys"]) > 1:
            raise ValueError("concat_problems: every problem needs V(i) for every X(i) (or none) and at most one bias")
    has_vel = len(problems[0]["vel_keys"]) > 0
    has_bias = len(problems[0]["bias_keys"]) > 0
    if any((len(p["vel_keys"]) > 0) != has_vel or (len(p["bias_keys"]) > 0) != has_bias for p in problems):
        raise ValueError("concat_problems: the problems of a batch must hold the same variable kinds")
    n = int(node_start[-1])
    ar = np.arange(n, dtype=np.uint64)
    out = {"pose_keys": np.uint64(X(0)) + ar,
           "poses": np.concatenate([p["poses"] for p in problems], 0),
           "vel_keys": np.uint64(V(0)) + ar if has_vel else np.zeros(0, dtype=np.uint64),
           "vels": np.concatenate([p["vels"] for p in problems], 0) if has_vel else np.zeros((0, 3)),
           "bias_keys": np.uint64(B(0)) + np.arange(len(problems), dtype=np.uint64) if has_bias else np.zeros(0, dtype=np.uint64),
           "biases": np.concatenate([p["biases"] for p in problems], 0) if has_bias else np.zeros((0, 6)),
           "lm_keys": np.zeros(0, dtype=np.uint64), "lms": np.zeros((0, 3)),
           "calib": problems[0]["calib"], "gravity": problems[0]["gravity"], "options": problems[0].get("options")}
    node_slots = ("x", "x1", "x2", "xi", "xj", "v", "vi", "vj")
    f0 = 0
    tables = {}
    for t, p in enumerate(problems):
        if not np.array_equal(p["gravity"], out["gravity"]):
            raise NotImplementedError("concat_problems: all problems must share one n_gravity")
        if p.get("options") != out["options"]:
            raise NotImplementedError("concat_problems: all problems must come from the same gtsam build options (config.py)")
        for name in ("prior_pose", "prior_vel", "between", "dvl", "stereo", "imu"):
            f = p[name]
            cols = tables.setdefault(name, {})
            for k, v in f.items():
                v = np.asarray(v)
                if k in node_slots:
                    v = v.astype(np.int32) + np.int32(node_start[t])
                elif k == "b":
                    v = np.full(len(v), t, dtype=np.int32)
                elif k == "orig":
                    v = v.astype(np.int64) + f0
                cols.setdefault(k, []).append(v)
        f0 += int(p["n_factors"])
    for name, cols in tables.items():
        out[name] = {k: np.concatenate(v, 0) for k, v in cols.items()}
    out["n_factors"] = f0
    return out, node_start


def split_values(tables, node_start, has_bias=True):
    """Inverse of concat_problems for the value tables of a batched session: -> list of per-trajectory tables."""
    out = []
    for t in range(len(node_start) - 1):
        a, b = int(node_start[t]), int(node_start[t + 1])
        out.append(dict(poses=tables["poses"][a:b], vels=tables["vels"][a:b] if len(tables["vels"]) else tables["vels"],
                        biases=tables["biases"][t:t + 1] if has_bias and len(tables["biases"]) else tables["biases"][:0],
                        lms=tables["lms"][:0]))
    return out


def solve_batched(problems, params=None, lib=None, device=0, keep_values=True, stats=None):
    """Solve a list of independent packed problems as ONE block-diagonal system on this rank's GPU (vus_set_components):
    every trajectory keeps its own lambda / accept-reject / convergence path, as separate optimizers would.
    -> list of dict(summary fields..., values=tables or None), in input order (same shape as solve_local's result).
    `stats`: optional dict that receives the whole-batch vus_lm_result (rounds, PCG iterations, launches, per-class times)."""
    if len(problems) == 1:                               # a batch of one is the single-graph path (batch.py:337 as it stands)
        out = solve_local(problems, params, lib=lib, device=device, threads=1, keep_values=keep_values)
        if stats is not None:
            stats.update(inner_iterations=out[0]["inner_iterations"], iterations=out[0]["iterations"])
        return out
    prob, node_start = concat_problems(problems)
    s = Session(prob, params or LevenbergMarquardtParams(), lib=lib, device=device, components=node_start)
    try:
        total = s.optimize()
        per = s.component_results()
        vals = split_values(s.values(), node_start, has_bias=len(prob["bias_keys"]) > 0) if keep_values else [None] * len(per)
        out = []
        for r, v in zip(per, vals):
            d = {k: r[k] for k in SUMMARY_FIELDS}
            d["values"] = v
            out.append(d)
        if stats is not None:
            stats.update(total)
        return out
    finally:
        s.close()


def solve_sharded(make_problem, n_trajectories, params=None, lib=None, device=0, threads=4, group=None, keep_values=False,
                  batched=False):
    """Every rank builds and solves trajectories shard_range(n, rank, world) (make_problem(t) -> packed problem), then
    all ranks gather the [n, len(SUMMARY_FIELDS)] summary table.  -> (summary [n, F] float64, local results, (first, last)).
    batched=True solves a rank's shard as ONE block-diagonal system (solve_batched) instead of one handle per trajectory."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    first, last = shard_range(n_trajectories, rank, world)
    shard = [make_problem(t) for t in range(first, last)]
    if batched and shard:
        local = solve_batched(shard, params, lib, device, keep_values=keep_values)
    else:
        local = solve_local(shard, params, lib, device, threads, keep_values=keep_values)
    table = np.array([[r[k] for k in SUMMARY_FIELDS] for r in local], dtype=np.float64).reshape(last - first, len(SUMMARY_FIELDS))
    if world == 1:
        return table, local, (first, last)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", device) if backend == "nccl" else torch.device("cpu")
    # ragged all-gather: pad every shard to the largest one
    width = max(shard_range(n_trajectories, r, world)[1] - shard_range(n_trajectories, r, world)[0] for r in range(world))
    send = torch.zeros((width, len(SUMMARY_FIELDS)), dtype=torch.float64, device=dev)
    send[:last - first] = torch.from_numpy(table).to(dev)
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    parts = []
    for r in range(world):
        f, l = shard_range(n_trajectories, r, world)
        parts.append(recv[r][:l - f].cpu().numpy())
    return np.concatenate(parts, 0), local, (first, last)


# ======================================================================================================================
# one pose graph across ranks: contiguous pose ranges + halo exchange
# ======================================================================================================================
def partition_pose_graph(prob, world):
    """Split a packed pose-graph problem (poses, prior_pose, between only) into `world` local problems.

    Rank r owns poses pose_range(n, r, world).  Its local graph holds every factor touching an owned pose; a factor is
    OWNED by the rank of its lowest pose (its error is counted there), the other copy is a duplicate.  Local pose order
    is [owned (ascending) | halo (ascending global index)]; owned factors come first in every table.
    -> list over ranks of dict(prob, n_owned, nf_owned, owned (first, last), halo_global [nh], send {peer: local idx},
       recv {peer: (offset in halo, count)})."""
    for t in ("prior_vel", "dvl", "stereo", "imu"):
        if len(prob[t]["orig"]):
            raise NotImplementedError("pose-range partition supports pose graphs (PriorFactorPose3 / BetweenFactorPose3) only")
    n = len(prob["pose_keys"])
    x1, x2 = np.asarray(prob["between"]["x1"], np.int64), np.asarray(prob["between"]["x2"], np.int64)
    xp = np.asarray(prob["prior_pose"]["x"], np.int64)
    lo = np.minimum(x1, x2)
    parts = []
    for r in range(world):
        a, b = pose_range(n, r, world)
        in1, in2 = (x1 >= a) & (x1 < b), (x2 >= a) & (x2 < b)
        sel = np.nonzero(in1 | in2)[0]
        mine = (lo[sel] >= a) & (lo[sel] < b)
        sel = np.concatenate([sel[mine], sel[~mine]])                  # owned factors first (stable)
        n_own_b = int(mine.sum())
        ends = np.concatenate([x1[sel], x2[sel]])
        halo = np.unique(ends[(ends < a) | (ends >= b)])
        loc = np.full(n, -1, np.int64)
        loc[a:b] = np.arange(b - a)
        loc[halo] = (b - a) + np.arange(len(halo))
        selp = np.nonzero((xp >= a) & (xp < b))[0]
        nloc = (b - a) + len(halo)
        glob = np.concatenate([np.arange(a, b), halo])
        from .symbol import symbols
        lp = dict(prob)
        lp["pose_keys"] = symbols("x", np.arange(nloc))
        lp["poses"] = np.ascontiguousarray(prob["poses"][glob])
        bt = prob["between"]
        lp["between"] = dict(meas=np.ascontiguousarray(bt["meas"][sel]), sqrt_info=np.ascontiguousarray(bt["sqrt_info"][sel]),
                             orig=np.arange(len(selp), len(selp) + len(sel), dtype=np.int64),
                             x1=loc[x1[sel]].astype(np.int32), x2=loc[x2[sel]].astype(np.int32))
        pp = prob["prior_pose"]
        lp["prior_pose"] = dict(meas=np.ascontiguousarray(pp["meas"][selp]), sqrt_info=np.ascontiguousarray(pp["sqrt_info"][selp]),
                                orig=np.arange(len(selp), dtype=np.int64), x=loc[xp[selp]].astype(np.int32))
        lp["n_factors"] = len(selp) + len(sel)
        # the halo poses that continue the chain on either side (vus_set_partition_chain): local index or -1
        prev_local = [int(loc[a - 1 - j]) if a - 1 - j >= 0 else -1 for j in range(CHAIN_SIDE)]
        next_local = [int(loc[b + j]) if b + j < n else -1 for j in range(CHAIN_SIDE)]
        parts.append(dict(prob=lp, n_owned=b - a, nf_owned=[len(selp), 0, n_own_b, 0, 0, 0], owned=(a, b), halo_global=halo,
                          global_factor_index=dict(prior_pose=selp, between=sel), n_own_between=n_own_b,
                          chain=(prev_local, next_local, r, world)))
    for r, P in enumerate(parts):                                       # halo lists: who sends what to whom
        P["recv"], P["send"] = {}, {}
        owners = np.array([pose_owner(int(g), n, world) for g in P["halo_global"]], np.int64) if len(P["halo_global"]) else np.zeros(0, np.int64)
        for s_ in range(world):
            idx = np.nonzero(owners == s_)[0]
            if len(idx):
                P["recv"][s_] = (int(idx[0]), int(len(idx)))            # halo is sorted by global index -> contiguous per owner
    for r, P in enumerate(parts):
        for s_, Q in enumerate(parts):
            if s_ == r or r not in Q["recv"]:
                continue
            off, cnt = Q["recv"][r]
            P["send"][s_] = (Q["halo_global"][off:off + cnt] - P["owned"][0]).astype(np.int64)   # my local indices, in s_'s halo order
    return parts


class PartitionedSolver:
    """One rank of a pose graph split by pose range.  `part` is this rank's entry of partition_pose_graph(); the process
    group must already exist (nccl on GPUs, gloo for the CPU tests)."""

    def __init__(self, part, params=None, lib=None, device=0, group=None, stream_ordered=True, nccl_in_library=None, exact_band=True):
        import torch
        import torch.distributed as dist
        from . import _native
        self.torch, self.dist, self.group = torch, dist, group
        self.part = part
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.cuda = self.world > 1 and dist.get_backend(group) == "nccl" or (lib is None and torch.cuda.is_available())
        self.device = torch.device("cuda", device) if self.cuda else torch.device("cpu")
        self.D = 6
        self.n_owned = part["n_owned"]
        self.n_local = self.n_owned + len(part["halo_global"])
        self._send_idx = {p: torch.as_tensor(ix, device=self.device) for p, ix in part["send"].items()}
        self.comm_calls = {"allreduce": 0, "halo": 0}
        self._cb = _native.COMM_FN(self._comm)
        self._views = {}
        # NCCL: the library runs on a torch stream of ours and every collective is enqueued on that same stream, so neither
        # side waits on the host (vus_set_comm_mode); gloo (CPU tests): host-synchronous callbacks
        # On GPUs the library issues the collectives itself (vus_comm_init: NCCL on its own stream, inside its captured launch
        # sequences); torch.distributed only carries the 128-byte NCCL id to every rank.  Callbacks remain for gloo (CPU tests)
        # and as a fallback (nccl_in_library=False): there the library runs on a torch stream of ours and every collective is
        # enqueued on that same stream (vus_set_comm_mode).
        if nccl_in_library is None:
            nccl_in_library = self.cuda and self.world > 1 and dist.get_backend(group) == "nccl"
        self.nccl_in_library = bool(nccl_in_library) and self.world > 1
        self.stream = torch.cuda.Stream(self.device) if (self.cuda and self.world > 1 and stream_ordered and not self.nccl_in_library) else None
        self.session = Session(part["prob"], params, lib=lib, device=device,
                               partition=(self.n_owned, part["nf_owned"], part.get("chain") if (self.world > 1 and exact_band) else None),
                               comm=self._cb if (self.world > 1 and not self.nccl_in_library) else None)
        if self.nccl_in_library:
            uid = torch.zeros(128, dtype=torch.uint8)
            if self.rank == 0:
                buf = (C.c_char * 128)()
                self.session._check(self.session.lib.vus_nccl_unique_id(buf))
                uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
            uid = uid.to(self.device)
            dist.broadcast(uid, src=0, group=group)
            self.session.comm_init(uid.cpu().numpy().tobytes(), self.rank, self.world)
            peers = sorted(set(part["send"].keys()) | set(part["recv"].keys()))
            self.session.set_halo(peers, [part["send"].get(q, np.zeros(0, np.int64)) for q in peers],
                                  [part["recv"].get(q, (0, 0)) for q in peers])
        if self.stream is not None:
            self.session._check(self.session.lib.vus_set_comm_mode(self.session._h, 1))

    # ---- device memory handed over by the library as a torch tensor (no copy)
    def _view(self, ptr, count):
        key = (int(ptr), int(count))
        v = self._views.get(key)
        if v is None:
            v = self._views[key] = self._make_view(ptr, count)
        return v

    def _make_view(self, ptr, count):
        torch = self.torch
        if self.cuda:
            class _Arr:
                __cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}
            return torch.as_tensor(_Arr(), device=self.device)
        buf = (C.c_double * count).from_address(int(ptr))
        return torch.from_numpy(np.ctypeslib.as_array(buf))

    def _comm(self, ctx, op, buf, count):
        try:
            if self.stream is not None:
                with self.torch.cuda.stream(self.stream):
                    self._collective(op, buf, count)
            else:
                self._collective(op, buf, count)
                if self.cuda:
                    self.torch.cuda.synchronize(self.device)
            return 0
        except Exception as exc:                        # never let an exception cross the C boundary
            import traceback
            traceback.print_exc()
            self.error = exc
            return 1

    def _collective(self, op, buf, count):
        if True:
            from . import _native
            dist, torch = self.dist, self.torch
            if op == _native.COMM_ALLREDUCE_SUM:
                self.comm_calls["allreduce"] += 1
                dist.all_reduce(self._view(buf, int(count)), group=self.group)
            else:
                self.comm_calls["halo"] += 1
                D = int(count)
                vec = self._view(buf, self.n_local * D).view(self.n_local, D)
                ops, keep = [], []
                for peer, ix in self._send_idx.items():
                    out = vec.index_select(0, ix).contiguous()
                    keep.append(out)
                    ops.append(dist.P2POp(dist.isend, out, peer, group=self.group))
                for peer, (off, cnt) in self.part["recv"].items():
                    ops.append(dist.P2POp(dist.irecv, vec[self.n_owned + off:self.n_owned + off + cnt], peer, group=self.group))
                if ops:
                    for w in dist.batch_isend_irecv(ops):
                        w.wait()                        # NCCL: makes the current stream wait, not the host

    def optimize(self):
        return self.session.optimize(stream=self.stream.cuda_stream if self.stream is not None else None)

    def owned_poses(self):
        return self.session.values()["poses"][:self.n_owned]

    def gather_poses(self):
        """All ranks' owned poses, in global order, on every rank (result read-out, not part of the solve)."""
        mine = self.torch.from_numpy(np.ascontiguousarray(self.owned_poses()))
        if self.world == 1:
            return mine.numpy()
        sizes = [pose_range(self._n_global(), r, self.world) for r in range(self.world)]
        width = max(b - a for a, b in sizes)
        send = self.torch.zeros((width, 12), dtype=self.torch.float64, device=self.device)
        send[:self.n_owned] = mine.to(self.device)
        recv = [self.torch.empty_like(send) for _ in range(self.world)]
        self.dist.all_gather(recv, send, group=self.group)
        return np.concatenate([recv[r][:b - a].cpu().numpy() for r, (a, b) in enumerate(sizes)], 0)

    def _n_global(self):
        t = self.torch.tensor([self.n_owned], dtype=self.torch.int64, device=self.device)
        self.dist.all_reduce(t, group=self.group)
        return int(t.item())

    def close(self):
        self.session.close()
