"""N > 1 path on CPU: world_size-2 gloo process group, trajectories sharded across ranks with no data-path
collective, the kernels through the test-only host emulation of csrc/ (tests/emu).  The 2-rank result must equal
the 1-rank result for every trajectory, bit for bit."""
import os
import subprocess
import sys
import numpy as np
import pytest
from visual_underwater_slam_b200 import parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from visual_underwater_slam_b200 import _native, parallel, synthetic
lib = _native.bind(os.path.join(sys.argv[1], "tests", "emu", "libvus_emu.so"))
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"])
if world > 1:
    dist.init_process_group("gloo", rank=rank, world_size=world)
n = int(sys.argv[2])
def make(t):
    d = synthetic.make_trajectory_graph(40, seed=4 + t, n_loops=2, loop_min_gap=10)
    return d["graph"].to_problem(d["initial"])
table, local, (first, last) = parallel.solve_sharded(make, n, lib=lib, threads=2, keep_values=True, batched=len(sys.argv) > 4)
if rank == 0:
    np.save(sys.argv[3], table)
np.save(sys.argv[3] + ".rank%d.npy" % rank, np.stack([r["values"]["poses"] for r in local]) if local else np.zeros((0,)))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
'''


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 4096):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
            for t in range(n):
                r = parallel.owner_of(t, n, world)
                assert spans[r][0] <= t < spans[r][1]


def _run(world, n, out, tmp_path, extra=()):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    port = 29500 + (os.getpid() % 500)
    for rank in range(world):
        env = dict(os.environ, WORLD_SIZE=str(world), RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT, str(n), str(out), *extra], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        o, _ = p.communicate(timeout=600)
        assert p.returncode == 0, o


def test_two_rank_gloo_equals_one_rank(tmp_path):
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "visual_underwater_slam_b200", "csrc"), "emu"], check=True)
    n = 5
    one, two = tmp_path / "one.npy", tmp_path / "two.npy"
    _run(1, n, one, tmp_path)
    _run(2, n, two, tmp_path)
    t1, t2 = np.load(one), np.load(two)
    assert t1.shape == (n, len(parallel.SUMMARY_FIELDS))
    assert np.array_equal(t1, t2)                       # same trajectories, same kernels, no data-path collective: bit-exact
    assert np.all(t1[:, 0] < t1[:, 4])                  # every trajectory's error went down
    p1 = np.load(str(one) + ".rank0.npy")
    p2 = np.concatenate([np.load(str(two) + ".rank%d.npy" % r) for r in range(2)], 0)
    assert np.array_equal(p1, p2)


def test_two_rank_gloo_batched_shards_match_one_handle_per_trajectory(tmp_path):
    """Every rank solves its shard as ONE block-diagonal system (vus_set_components); per trajectory the LM path and the
    result equal the one-handle-per-trajectory run (not bit for bit: a component may receive one more refinement pass
    when another component of its batch needs it)."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "visual_underwater_slam_b200", "csrc"), "emu"], check=True)
    n = 5
    one, two = tmp_path / "one.npy", tmp_path / "two.npy"
    _run(1, n, one, tmp_path)
    _run(2, n, two, tmp_path, extra=("batched",))
    t1, t2 = np.load(one), np.load(two)
    names = parallel.SUMMARY_FIELDS
    for f in ("iterations", "inner_iterations", "final_lambda"):
        assert np.array_equal(t1[:, names.index(f)], t2[:, names.index(f)]), f
    fe = names.index("final_error")
    assert np.all(np.abs(t1[:, fe] - t2[:, fe]) <= 1e-9 * t1[:, fe])
    p1 = np.load(str(one) + ".rank0.npy")
    p2 = np.concatenate([np.load(str(two) + ".rank%d.npy" % r) for r in range(2)], 0)
    assert np.abs(p1 - p2).max() < 1e-7


PART_WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from visual_underwater_slam_b200 import _native, parallel, synthetic
from visual_underwater_slam_b200.optimizer import Session
lib = _native.bind(os.path.join(sys.argv[1], "tests", "emu", "libvus_emu.so"))
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"])
dist.init_process_group("gloo", rank=rank, world_size=world)
d = synthetic.make_pose_graph(int(sys.argv[2]), seed=5, n_loops=int(sys.argv[3]))
prob = d["graph"].to_problem(d["initial"])
if len(sys.argv) > 5 and sys.argv[5] == "noskip":       # odometry + loop closures only: one pose per supernode (k = 1)
    n = int(sys.argv[2]); bt = prob["between"]
    keep = np.concatenate([np.arange(n - 1), np.arange(2 * n - 3, len(bt["orig"]))])
    prob["between"] = {k: np.ascontiguousarray(np.asarray(v)[keep]) for k, v in bt.items()}
    prob["between"]["orig"] = np.arange(1, 1 + len(keep), dtype=np.int64)
    prob["n_factors"] = 1 + len(keep)
part = parallel.partition_pose_graph(prob, world)[rank]
ps = parallel.PartitionedSolver(part, lib=lib)
res = ps.optimize()
poses = ps.gather_poses()
if rank == 0:
    np.save(sys.argv[4], poses)
    json.dump(dict(iterations=res["iterations"], tries=res["inner_iterations"], final_error=res["final_error"],
                   initial_error=res["initial_error"], pcg=res["pcg_iterations"], comm=ps.comm_calls), open(sys.argv[4] + ".json", "w"))
ps.close()
dist.barrier(); dist.destroy_process_group()
'''


def _run_part(world, n, loops, out, tmp_path, extra=()):
    script = tmp_path / "part_worker.py"
    script.write_text(PART_WORKER)
    port = 29500 + ((os.getpid() + 7 * world) % 500)
    procs = []
    for rank in range(world):
        env = dict(os.environ, WORLD_SIZE=str(world), RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT, str(n), str(loops), str(out), *extra], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        o, _ = p.communicate(timeout=900)
        assert p.returncode == 0, o


def test_partition_lists_are_consistent():
    from visual_underwater_slam_b200 import synthetic
    d = synthetic.make_pose_graph(300, seed=5, n_loops=40)
    prob = d["graph"].to_problem(d["initial"])
    for world in (2, 3, 4):
        parts = parallel.partition_pose_graph(prob, world)
        n_between = len(prob["between"]["orig"])
        owned = np.concatenate([P["global_factor_index"]["between"][:P["n_own_between"]] for P in parts])
        assert sorted(owned.tolist()) == list(range(n_between))          # every factor is owned exactly once
        assert sum(P["nf_owned"][0] for P in parts) == len(prob["prior_pose"]["orig"])
        for r, P in enumerate(parts):
            a, b = P["owned"]
            assert np.all((P["halo_global"] < a) | (P["halo_global"] >= b))
            for peer, ix in P["send"].items():                            # what I send is what the peer expects, in its halo order
                off, cnt = parts[peer]["recv"][r]
                assert np.array_equal(ix + a, parts[peer]["halo_global"][off:off + cnt])
            # local indices of every factor are in range and touch an owned pose
            bt = P["prob"]["between"]
            assert np.all((bt["x1"] < P["n_owned"]) | (bt["x2"] < P["n_owned"]))


def test_partitioned_two_rank_gloo_matches_one_rank(tmp_path):
    """BASELINE config 5 in miniature on CPU: a pose graph with loop closures split over 2 ranks (halo exchange +
    all-reduced PCG / LM scalars through gloo) reaches the 1-rank LM path and optimum."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "visual_underwater_slam_b200", "csrc"), "emu"], check=True)
    import json
    one, two = tmp_path / "p1.npy", tmp_path / "p2.npy"
    _run_part(1, 240, 30, one, tmp_path)
    _run_part(2, 240, 30, two, tmp_path)
    m1, m2 = json.load(open(str(one) + ".json")), json.load(open(str(two) + ".json"))
    assert m1["iterations"] == m2["iterations"] and m1["tries"] == m2["tries"]
    assert abs(m1["final_error"] - m2["final_error"]) <= 1e-6 * m1["final_error"]
    assert m2["final_error"] < 0.2 * m2["initial_error"]
    assert m2["comm"]["halo"] > 0 and m2["comm"]["allreduce"] > 0
    p1, p2 = np.load(one), np.load(two)
    assert p1.shape == p2.shape == (240, 12)
    assert np.abs(p1[:, 9:] - p2[:, 9:]).max() < 1e-6 and np.abs(p1[:, :9] - p2[:, :9]).max() < 1e-6


def test_partitioned_ranks_share_one_exact_band(tmp_path):
    """A long chain with few loop closures over 4 ranks: the per-rank band factorizations are tied together exactly
    (vus_set_partition_chain, csrc/spike.cuh), so the partitioned solve takes the one-rank LM path with (nearly) the one-rank
    PCG iteration count and ends on the same poses.  With block-Jacobi across ranks the same graph needs 4 x the PCG
    iterations, fails a damped solve and ends 1.6e-5 away (measured before the scheme existed)."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "visual_underwater_slam_b200", "csrc"), "emu"], check=True)
    import json
    one, four = tmp_path / "q1.npy", tmp_path / "q4.npy"
    _run_part(1, 1200, 3, one, tmp_path)
    _run_part(4, 1200, 3, four, tmp_path)
    m1, m4 = json.load(open(str(one) + ".json")), json.load(open(str(four) + ".json"))
    assert m1["iterations"] == m4["iterations"] and m1["tries"] == m4["tries"]
    assert m4["pcg"] <= 1.1 * m1["pcg"] + 8, (m1["pcg"], m4["pcg"])
    assert abs(m1["final_error"] - m4["final_error"]) <= 1e-9 * m1["final_error"]
    p1, p4 = np.load(one), np.load(four)
    assert np.abs(p1 - p4).max() < 1e-8


def test_pose_range_is_aligned_and_covers_everything():
    for n in (10, 100, 1200, 4001):
        for world in (1, 2, 3, 8):
            spans = [parallel.pose_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            if -(-n // parallel.CHAIN_SIDE) >= 2 * world:
                assert all(b % parallel.CHAIN_SIDE == 0 for _, b in spans[:-1])
            for t in range(0, n, 7):
                r = parallel.pose_owner(t, n, world)
                assert spans[r][0] <= t < spans[r][1]


def test_partitioned_exact_band_with_one_pose_per_supernode(tmp_path):
    """The same tie across ranks when the graph has no skip factors (k = 1: 6 x 6 interface blocks), on 3 ranks."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "visual_underwater_slam_b200", "csrc"), "emu"], check=True)
    import json
    one, three = tmp_path / "k1.npy", tmp_path / "k3.npy"
    _run_part(1, 960, 3, one, tmp_path, extra=("noskip",))
    _run_part(3, 960, 3, three, tmp_path, extra=("noskip",))
    m1, m3 = json.load(open(str(one) + ".json")), json.load(open(str(three) + ".json"))
    assert m1["iterations"] == m3["iterations"] and m1["tries"] == m3["tries"]
    assert m3["pcg"] <= 1.1 * m1["pcg"] + 8, (m1["pcg"], m3["pcg"])
    p1, p3 = np.load(one), np.load(three)
    assert np.abs(p1 - p3).max() < 1e-8
