"""CPU (`-m "not gpu"`) checks of the host logic and of the kernel bodies through a sequential host emulation
(tests/emu/libvus_emu.so, built from the same csrc/ sources with -DVUS_EMU).  The emulation is test
infrastructure: the package never loads it.  The parity tests proper are the `-m gpu` ones (test_gpu_parity.py)."""
import os
import subprocess
import numpy as np
import pytest
from visual_underwater_slam_b200 import _native
import parity_common as pc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "visual_underwater_slam_b200", "csrc"), "emu"], check=True)
    return _native.bind(os.path.join(ROOT, "tests", "emu", "libvus_emu.so"))


def test_factor_parity_chain_loops(emu):
    _, prob = pc.make(80, n_loops=4)
    pc.check_factor_parity(emu, prob)


def test_factor_parity_stereo(emu):
    _, prob = pc.make(60, n_lm=120)
    pc.check_factor_parity(emu, prob)


def test_band_solve(emu):
    _, prob = pc.make(40)
    pc.check_band_solve(emu, prob, 1e-3, nrhs=6)
    _, prob = pc.make(30, n_lm=60)
    pc.check_band_solve(emu, prob, 1e-3, nrhs=2)


def test_solve_chain_is_exact(emu):
    _, prob = pc.make(70)
    its = pc.check_solve_parity(emu, prob, 1e-3, 1e-6)
    assert its <= 2          # band + bias border are factored exactly: PCG is only iterative refinement


def test_solve_stereo_schur(emu):
    _, prob = pc.make(60, n_lm=100)
    its = pc.check_solve_parity(emu, prob, 1e-2, 1e-6)
    assert its <= 3


def test_solve_loop_closures(emu):
    _, prob = pc.make(120, n_loops=5, loop_min_gap=30)
    pc.check_solve_parity(emu, prob, 1.0, 1e-6)


def test_lm_parity_chain(emu):
    _, prob = pc.make(120, n_loops=3, loop_min_gap=40)
    pc.check_lm_parity(emu, prob)


def test_lm_parity_stereo(emu):
    _, prob = pc.make(60, n_lm=100)
    pc.check_lm_parity(emu, prob)


def test_known_answer_noise_free(emu):
    """Noise-free graph: the optimum is the ground truth."""
    from visual_underwater_slam_b200.optimizer import Session
    d, prob = pc.make(50, n_lm=60, noise_scale=0.0)
    s = Session(prob, lib=emu)
    res = s.optimize()
    v = s.values()
    assert res["final_error"] < 1e-10
    assert np.abs(v["poses"] - d["truth"]["poses"]).max() < 1e-6
    assert np.abs(v["vels"] - d["truth"]["vels"]).max() < 1e-6
    s.close()


def test_tracks_longer_than_the_band_stay_exact(emu):
    """Landmark tracks that do not fit in the band (here: supernodes capped below the 10-pose track length) keep their
    Schur term implicit in the operator: the damped solve is still the exact one, only PCG works harder."""
    from visual_underwater_slam_b200.optimizer import LevenbergMarquardtParams, Session
    from oracle import lm
    _, prob = pc.make(60, n_lm=120)
    vals = lm.values_of(prob)
    lay = lm.Layout(prob)
    J, b = lm.linearize(prob, vals, lay)
    delta = lm.solve_damped(J, b, 1e-3, lay)
    p = LevenbergMarquardtParams()
    p.maxSupernode = 4
    p.pcgMaxIterations = 2000
    s = Session(prob, p, lib=emu)
    assert s.layout()["k"] < 9
    st = s.solve_step(1e-3)
    s.close()
    mine = np.concatenate([st["bias"].ravel(), st["lm"].ravel(), st["vel"].ravel(), st["pose"].ravel()])
    H = (J.T @ J).tocsr()
    g = J.T @ b
    assert np.linalg.norm(H @ mine + 1e-3 * mine - g) <= 1e-9 * np.linalg.norm(g)
    assert np.linalg.norm(mine - delta) <= 1e-5 * np.linalg.norm(delta)


def test_marginals_chain_and_bias(emu):
    _, prob = pc.make(50, n_loops=2, loop_min_gap=20)
    pc.check_marginals(emu, prob, [("pose", 0), ("pose", 37), ("vel", 12), ("bias", 0)])


def test_marginals_stereo_landmarks(emu):
    _, prob = pc.make(40, n_lm=60)
    pc.check_marginals(emu, prob, [("lm", 7), ("pose", 20), ("lm", 41), ("vel", 39)])


def test_incremental_resolve_matches_oracle_warm_start(emu):
    """ISAM2 facade (isam.py:341-342): after every update the estimate is the LM optimum of the accumulated
    graph started from the previous estimate -- checked against the oracle run from the same warm start."""
    import visual_underwater_slam_b200 as gtsam
    from oracle import lm
    d, _ = pc.make(60, n_lm=40, n_loops=2, loop_min_gap=20)
    isam = gtsam.ISAM2(lib=emu)
    acc_graph, acc_init = gtsam.NonlinearFactorGraph(), gtsam.Values()
    n_seen = 0
    for G, V in pc.split_for_incremental(d, [25, 45, 60]):
        warm = gtsam.Values(isam.calculateEstimate())
        warm.insert(V)
        acc_graph.push_back(G)
        res = isam.update(G, V)
        n_seen += G.size()
        assert isam.getFactorsUnsafe().size() == n_seen == acc_graph.size()
        prob = acc_graph.to_problem(warm)
        vals, info = lm.lm_optimize(prob)
        est = isam.calculateEstimate()
        assert abs(res.getErrorAfter() - info["error"]) <= 1e-6 * info["error"]
        assert np.abs(est.table("pose")[1] - vals["poses"]).max() < 1e-6
        assert np.abs(est.table("vel")[1] - vals["vels"]).max() < 1e-6
    assert n_seen == d["graph"].size()
    c = isam.marginalCovariance(gtsam.symbol_shorthand.X(59))
    assert c.shape == (6, 6) and np.all(np.linalg.eigvalsh(c) > 0)


def test_batched_trajectories_match_separate_optimizers(emu):
    """Ragged batch (different lengths, loop closures, LM paths with rejected tries) in one handle."""
    probs = [pc.make(40 + 7 * t, n_loops=2, loop_min_gap=15, seed=10 + t)[1] for t in range(3)]
    res = pc.check_batched_parity(emu, probs)
    assert len({r["inner_iterations"] for r in res}) > 1        # the trajectories really took different LM paths


def test_batched_pose_graphs_without_bias(emu):
    from visual_underwater_slam_b200 import synthetic
    probs = []
    for t in range(3):
        d = synthetic.make_pose_graph(60 + 10 * t, seed=20 + t, n_loops=4)
        probs.append(d["graph"].to_problem(d["initial"]))
    pc.check_batched_parity(emu, probs)


def test_batched_rejects_cross_component_factor(emu):
    from visual_underwater_slam_b200 import parallel
    from visual_underwater_slam_b200.optimizer import Session
    probs = [pc.make(20, seed=30 + t)[1] for t in range(2)]
    prob, node_start = parallel.concat_problems(probs)
    for slot in ("xj", "vj"):                                     # ImuFactor from keyframe 0 of trajectory 0 to keyframe 5 of trajectory 1
        prob["imu"][slot] = prob["imu"][slot].copy()
        prob["imu"][slot][0] = 25
    with pytest.raises(RuntimeError, match="connects two components"):
        Session(prob, lib=emu, components=node_start)


def test_batched_mixed_closure_counts_and_fallback(emu):
    """Capacitance solve of the loop closures: components with no / one / several closures in one batch (unused capacitance
    slots are identity rows), and the plain-PCG fallback when a component has more than 8 closures."""
    probs = [pc.make(60, n_loops=nl, loop_min_gap=12, seed=70 + t)[1] for t, nl in enumerate((0, 3, 1, 5))]
    pc.check_batched_parity(emu, probs)
    probs = [pc.make(90, n_loops=nl, loop_min_gap=10, seed=80 + t)[1] for t, nl in enumerate((10, 2))]
    pc.check_batched_parity(emu, probs)


def test_golden_rows_on_the_emulation(emu):
    """The frozen fixtures of the rows around the path (marginals, batched trajectories) through the host emulation."""
    import json
    from visual_underwater_slam_b200 import parallel
    from visual_underwater_slam_b200.optimizer import Session
    golden = os.path.join(ROOT, "tests", "golden")
    g = np.load(os.path.join(golden, "rows_marginals.npz"))
    meta = json.loads(str(g["meta"]))
    _, prob = pc.make(**meta["make"])
    s = Session(prob, lib=emu)
    cov = s.marginal_covariance([tuple(q) for q in meta["queries"]])
    s.close()
    scale = np.sqrt(np.outer(np.diag(g["cov"]), np.diag(g["cov"])))
    assert np.abs((cov - g["cov"]) / scale).max() <= 1e-4
    g = np.load(os.path.join(golden, "rows_batched.npz"))
    meta = json.loads(str(g["meta"]))
    res = parallel.solve_batched([pc.make(**mk)[1] for mk in meta["batch"]], lib=emu)
    for r, row in zip(res, g["summary"]):
        assert r["iterations"] == int(row[0]) and r["inner_iterations"] == int(row[1])
        assert abs(r["final_error"] - row[2]) <= 1e-6 * row[2]


def test_landmark_seen_twice_from_one_pose_stays_exact(emu):
    """A landmark observed twice from the same pose (two stereo factors on one (X, L) pair): the register-blocked Schur kernel
    expects one partner per pose, so such a landmark takes the implicit path of the long tracks -- still eliminated exactly."""
    _, prob = pc.make(40, n_lm=60, seed=21)
    st = {k: np.asarray(v).copy() for k, v in prob["stereo"].items()}
    dup = np.array([3, 4, 50, 51, 200])                              # duplicate a few observations (same pose, same landmark)
    for k in st:
        st[k] = np.concatenate([st[k], st[k][dup]], 0)
    st["meas"][-len(dup):] += 0.3                                     # a second, slightly different measurement
    st["orig"][-len(dup):] = prob["n_factors"] + np.arange(len(dup))
    prob = dict(prob)
    prob["stereo"] = st
    prob["n_factors"] = prob["n_factors"] + len(dup)
    pc.check_factor_parity(emu, prob)
    its = pc.check_solve_parity(emu, prob, 1e-2, 1e-6)
    assert its < 60
    pc.check_lm_parity(emu, prob)


def test_tracks_with_gaps_keep_parity(emu):
    """Ragged landmark tracks (poses missing inside a track, tracks of different lengths): the partner table of the Schur
    kernel must skip the gaps -- factor, solve and LM parity on a graph with 30 % of the observations removed."""
    _, prob = pc.make(50, n_lm=80, seed=31)
    st = {k: np.asarray(v) for k, v in prob["stereo"].items()}
    rng = np.random.default_rng(5)
    keep = rng.random(len(st["orig"])) > 0.3
    for l in range(len(prob["lm_keys"])):                           # every landmark keeps at least two observations
        rows = np.nonzero(st["l"] == l)[0]
        if keep[rows].sum() < 2:
            keep[rows[:2]] = True
    st = {k: v[keep].copy() for k, v in st.items()}
    removed = int((~keep).sum())
    prob = dict(prob)
    # insertion indices stay a permutation of 0 .. n_factors-1: renumber everything after dropping the removed factors
    order = np.sort(np.concatenate([np.asarray(prob[t]["orig"]) for t in ("prior_pose", "prior_vel", "between", "dvl", "imu")] + [st["orig"]]))
    remap = {int(o): i for i, o in enumerate(order)}
    for t in ("prior_pose", "prior_vel", "between", "dvl", "imu"):
        f = dict(prob[t])
        f["orig"] = np.array([remap[int(o)] for o in f["orig"]], dtype=np.int64)
        prob[t] = f
    st["orig"] = np.array([remap[int(o)] for o in st["orig"]], dtype=np.int64)
    prob["stereo"] = st
    prob["n_factors"] = prob["n_factors"] - removed
    pc.check_factor_parity(emu, prob)
    pc.check_solve_parity(emu, prob, 1e-2, 1e-6)
    pc.check_lm_parity(emu, prob)


def test_batch_of_one_is_the_single_graph_path(emu):
    probs = [pc.make(45, n_loops=2, loop_min_gap=12, seed=5)[1]]
    pc.check_batched_parity(emu, probs)


def test_marginals_never_return_an_unconverged_column(emu):
    """Tracks longer than the band leave most of the stereo information to PCG; the undamped unit-vector solves of
    gtsam.Marginals can then fail to converge.  The library must either return the oracle's covariance or refuse
    (VUS_ERR_STATE, 'did not converge') -- never hand back the unconverged column (found by tools/fuzz_single.py)."""
    from visual_underwater_slam_b200 import synthetic
    d = synthetic.make_trajectory_graph(68, seed=5892, n_landmarks=89, obs_per_landmark=13, n_loops=3, loop_min_gap=17, pixel_noise=1.0)
    prob = d["graph"].to_problem(d["initial"])
    answered = refused = 0
    for q in ([("pose", 10)], [("vel", 30)], [("bias", 0)], [("lm", 5)], [("pose", 0)], [("pose", 8)], [("lm", 0)]):
        try:
            pc.check_marginals(emu, prob, q, rtol=1e-5)      # an answer must be the oracle's
            answered += 1
        except RuntimeError as e:
            assert "did not converge" in str(e)              # ... or a refusal
            refused += 1
    assert answered + refused == 7


def check_bad_pivot_stays_in_its_component(lib):
    """A trajectory whose damped system loses positive definiteness to rounding (IMU information scaled to ~1e22 next to
    lambda = 1e-5) fails ITS tries and raises ITS lambda; the other trajectories of the batch take exactly the LM path they
    take without it (the fail flag is per component, a failed block is zeroed instead of leaking Inf / NaN)."""
    from visual_underwater_slam_b200 import parallel
    probs = [pc.make(40 + 7 * t, n_loops=2, loop_min_gap=15, seed=10 + t)[1] for t in range(3)]
    base = parallel.solve_batched(probs, lib=lib)
    bad = [dict(p) for p in probs]
    imu = dict(bad[1]["imu"])
    imu["sqrt_info"] = imu["sqrt_info"] * 1e8
    bad[1]["imu"] = imu
    st = {}
    res = parallel.solve_batched(bad, lib=lib, stats=st)
    assert st["solve_failures"] > 0 and res[1]["inner_iterations"] > res[1]["iterations"]
    for t in (0, 2):
        assert res[t]["iterations"] == base[t]["iterations"] and res[t]["inner_iterations"] == base[t]["inner_iterations"]
        assert abs(res[t]["final_error"] - base[t]["final_error"]) <= 1e-9 * base[t]["final_error"]
        assert np.abs(res[t]["values"]["poses"] - base[t]["values"]["poses"]).max() < 1e-9


def test_bad_pivot_stays_in_its_component(emu):
    check_bad_pivot_stays_in_its_component(emu)


def test_lm_parity_manifold_build_with_exact_between(emu):
    """The other gtsam build (GTSAM_TANGENT_PREINTEGRATION=OFF, GTSAM_SLOW_BUT_CORRECT_BETWEENFACTOR=ON) end to end."""
    import visual_underwater_slam_b200 as gtsam
    prev = gtsam.gtsam_build()
    gtsam.set_gtsam_build(tangent_preintegration=False, slow_but_correct_betweenfactor=True)
    try:
        _, prob = pc.make(90, n_lm=60, n_loops=3, loop_min_gap=30)
    finally:
        gtsam.set_gtsam_build(**prev)
    assert prob["options"] == dict(tangent_preintegration=False, slow_but_correct_betweenfactor=True)
    pc.check_factor_parity(emu, prob)
    pc.check_lm_parity(emu, prob)


def test_golden_configs_on_the_emulation(emu):
    """The reduced config-2 fixture through the host emulation (the full-size ones run on the GPU)."""
    pc.check_golden_config(emu, os.path.join(ROOT, "tests", "golden", "lm_c2s.npz"))


def test_damped_system_refresh_leaves_nothing_stale(emu):
    """form_system copies the whole base system once per graph and afterwards refreshes only the node / pair blocks and lets
    the Schur kernel assign its blocks: solving at lambda A, then B, then A again must reproduce the first answer bit for bit,
    and a fresh handle must give the same bits (stereo + loop closures + IMU chain: every kind of block)."""
    from visual_underwater_slam_b200.optimizer import Session
    _, prob = pc.make(90, n_lm=150, n_loops=3, loop_min_gap=30)
    s = Session(prob, lib=emu)
    a1 = s.solve_step(1e-3)
    s.solve_step(10.0)
    a2 = s.solve_step(1e-3)
    s.close()
    s = Session(prob, lib=emu)
    a3 = s.solve_step(1e-3)
    s.close()
    for k in ("pose", "vel", "lm", "bias"):
        assert np.array_equal(a1[k], a2[k]) and np.array_equal(a1[k], a3[k]), k


def _with_extra_between(prob, rows, swap):
    """Copy of a packed problem with between-factor rows `rows` appended again (swap[i]: with x1 / x2 exchanged)."""
    out = dict(prob)
    bt = prob["between"]
    rows = np.asarray(rows)
    x1, x2 = np.asarray(bt["x1"])[rows].copy(), np.asarray(bt["x2"])[rows].copy()
    sw = np.asarray(swap, bool)
    x1[sw], x2[sw] = np.asarray(bt["x2"])[rows][sw], np.asarray(bt["x1"])[rows][sw]
    n0 = int(prob["n_factors"])
    out["between"] = dict(meas=np.concatenate([bt["meas"], bt["meas"][rows]]), sqrt_info=np.concatenate([bt["sqrt_info"], bt["sqrt_info"][rows]]),
                          orig=np.concatenate([bt["orig"], n0 + np.arange(len(rows), dtype=np.int64)]),
                          x1=np.concatenate([bt["x1"], x1]).astype(np.int32), x2=np.concatenate([bt["x2"], x2]).astype(np.int32))
    out["n_factors"] = n0 + len(rows)
    return out


def test_several_factors_on_one_pose_pair_and_both_orientations(emu):
    """The pair gather (PairAsmBody) sums EVERY factor of a pose pair, whichever pose the factor names first: an odometry
    factor doubled, a skip factor doubled the other way round, an off-band loop closure tripled (one copy reversed)."""
    from visual_underwater_slam_b200 import synthetic
    d = synthetic.make_pose_graph(130, seed=8, n_loops=6)
    prob = d["graph"].to_problem(d["initial"])
    nb = len(prob["between"]["orig"])
    closure = nb - 1                                      # the generator appends the loop closures last
    prob2 = _with_extra_between(prob, [3, 129 + 7, closure, closure], [False, True, False, True])
    pc.check_factor_parity(emu, prob2)
    pc.check_solve_parity(emu, prob2, 1e-2, 1e-6)
    pc.check_lm_parity(emu, prob2)


def test_pose_without_chain_factor_of_its_own_keeps_a_defined_block(emu):
    """NodeAsmBody writes every entry of every node's block on every linearization (there is no zero-fill any more): a stereo
    graph re-linearized many times, with rejected tries in between, must keep matching the oracle."""
    _, prob = pc.make(50, n_lm=120, n_loops=2, loop_min_gap=20)
    pc.check_lm_parity(emu, prob)


def test_long_tracks_keep_the_oracle_path_with_a_strong_preconditioner(emu):
    """13 observations per landmark: longer than any band.  The operator keeps the exact implicit Schur term; the band
    preconditioner carries the track cut into segments that fit (every segment a landmark of its own), so a damped solve takes
    ~10 PCG iterations instead of 130-200, the LM path is the oracle's, and marginal queries are answered, not refused."""
    from visual_underwater_slam_b200 import synthetic
    from visual_underwater_slam_b200.optimizer import Session
    d = synthetic.make_trajectory_graph(74, seed=1, n_landmarks=60, obs_per_landmark=13, pixel_noise=1.0)
    prob = d["graph"].to_problem(d["initial"])
    s = Session(prob, lib=emu)
    assert s.layout()["k"] == 9
    st = s.solve_step(1e-5)
    s.close()
    assert st["pcg_iterations"] <= 20, st["pcg_iterations"]
    res, info = pc.check_lm_parity(emu, prob)
    assert res["pcg_iterations"] <= 20 * res["inner_iterations"]
    pc.check_marginals(emu, prob, [("pose", 40), ("lm", 7), ("vel", 12), ("bias", 0)], rtol=1e-5)
