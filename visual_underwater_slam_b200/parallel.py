"""Multi-GPU host logic for the batch-LM path: one process per GPU, torch.distributed for the plumbing
(NCCL on the GPUs, gloo in the CPU tests).

The reference solves one graph per process (batch.py:337).  The path shards by *independent trajectories*
(BASELINE.json config 4: a batch of independent 500-pose graphs): trajectory t goes to rank t // ceil(n / world),
every rank solves its shard with its own handles on its own GPU, and the only communication is one all-gather of the
per-trajectory summary (final error, iterations, lambda) at the end -- there is NO data-path collective, so scaling is
weak and the N-rank result is bit-identical to the 1-rank result for every trajectory.

Inside a rank the trajectories of a shard are small (launch-latency bound), so several handles run concurrently from a
thread pool, each on its own CUDA stream (ctypes releases the GIL during the C-ABI calls; one host thread per handle, as
include/vus.h requires).
"""
from concurrent.futures import ThreadPoolExecutor
import numpy as np
from .optimizer import Session, LevenbergMarquardtParams

SUMMARY_FIELDS = ("final_error", "iterations", "inner_iterations", "final_lambda", "initial_error")


def shard_range(n_items, rank, world):
    """Contiguous block partition: -> (first, last_exclusive) of `rank`; blocks differ in size by at most one."""
    base, extra = divmod(int(n_items), int(world))
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def owner_of(item, n_items, world):
    """Inverse of shard_range."""
    base, extra = divmod(int(n_items), int(world))
    cut = extra * (base + 1)
    return item // (base + 1) if item < cut else extra + (item - cut) // max(base, 1)


def solve_local(problems, params=None, lib=None, device=0, threads=4, streams=None, keep_values=True):
    """Solve a list of independent packed problems (graph.to_problem) on this rank's GPU.
    -> list of dict(summary fields..., values=tables or None), in input order."""
    params = params or LevenbergMarquardtParams()

    def work(arg):
        i, prob = arg
        stream = None
        if streams is not None:
            stream = streams[i % len(streams)]
        s = Session(prob, params, lib=lib, device=device)
        try:
            res = s.optimize(stream=stream)
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


def solve_sharded(make_problem, n_trajectories, params=None, lib=None, device=0, threads=4, group=None, keep_values=False):
    """Every rank builds and solves trajectories shard_range(n, rank, world) (make_problem(t) -> packed problem), then
    all ranks gather the [n, len(SUMMARY_FIELDS)] summary table.  -> (summary [n, F] float64, local results, (first, last))."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    first, last = shard_range(n_trajectories, rank, world)
    local = solve_local([make_problem(t) for t in range(first, last)], params, lib, device, threads, keep_values=keep_values)
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
