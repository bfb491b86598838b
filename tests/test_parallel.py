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
table, local, (first, last) = parallel.solve_sharded(make, n, lib=lib, threads=2, keep_values=True)
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


def _run(world, n, out, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    port = 29500 + (os.getpid() % 500)
    for rank in range(world):
        env = dict(os.environ, WORLD_SIZE=str(world), RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT, str(n), str(out)], env=env,
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
