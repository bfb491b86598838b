"""Host logic of the drop-in boundary (no GPU): the gtsam-named API batch.py uses, key / factor indexing (bit-exact,
SURVEY.md 8a-a1, A.8), error behaviour, and that the C-ABI library exports every symbol include/vus.h declares."""
import ctypes
import functools
import os
import re
import subprocess
import numpy as np
import pytest
import visual_underwater_slam_b200 as gtsam
from visual_underwater_slam_b200 import _native, synthetic
from visual_underwater_slam_b200.symbol import B, V, X, L, symbolChr, symbolIndex
from visual_underwater_slam_b200.optimizer import Session, LevenbergMarquardtParams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "visual_underwater_slam_b200", "csrc"), "emu"], check=True)
    return _native.bind(os.path.join(ROOT, "tests", "emu", "libvus_emu.so"))


def test_symbol_keys_are_gtsam_bit_layout():
    assert X(5) == (0x78 << 56) | 5 and V(0) == 0x76 << 56 and B(0) == 0x62 << 56 and L(7) == (0x6C << 56) | 7
    assert symbolChr(X(12)) == "x" and symbolIndex(X(12)) == 12
    assert B(10 ** 6) < L(0) < L(10 ** 6) < V(0) < V(10 ** 6) < X(0)          # std::map<Key> order: b < l < v < x
    assert gtsam.Symbol("x", 3).key() == X(3) and gtsam.Symbol(X(3)).index() == 3


def test_lm_params_defaults_are_gtsam_defaults():
    p = LevenbergMarquardtParams()
    assert (p.maxIterations, p.relativeErrorTol, p.absoluteErrorTol, p.errorTol) == (100, 1e-5, 1e-5, 0.0)
    assert (p.lambdaInitial, p.lambdaFactor, p.lambdaUpperBound, p.lambdaLowerBound) == (1e-5, 10.0, 1e5, 0.0)
    assert p.minModelFidelity == 1e-3 and p.diagonalDamping is False and p.useFixedLambdaFactor is True
    p.setMaxIterations(7)
    assert p.getMaxIterations() == 7
    with pytest.raises(NotImplementedError):
        p.setDiagonalDamping(True)


def velocity_error(measurement, this, values, jacobians):     # stand-in with the reference's signature (batch.py:196)
    raise AssertionError("a Python callback must never run on this path")


def _batch_py_graph(n=6, k_imu=4, seed=0):
    """The construction loop of batch.py:270-305, object by object."""
    rng = np.random.default_rng(seed)
    graph, initial = gtsam.NonlinearFactorGraph(), gtsam.Values()
    params = synthetic.reference_imu_params()
    pim = gtsam.PreintegratedImuMeasurements(params)
    K = gtsam.Cal3_S2Stereo(*synthetic.CALIB)
    pose_noise = gtsam.noiseModel.Diagonal.Sigmas(np.array([0.1, 0.1, 0.1, 0.3, 0.3, 0.3]))
    vel_noise = gtsam.noiseModel.Isotropic.Sigma(3, 0.1)
    dvl_noise = gtsam.noiseModel.Isotropic.Sigma(3, 0.1)
    cam_noise = gtsam.noiseModel.Isotropic.Sigma(3, 10.0)
    initial.insert(B(0), gtsam.imuBias.ConstantBias())                                   # batch.py:274
    expect = []
    seen = set()
    for i in range(n):
        pose = gtsam.Pose3(gtsam.Rot3.Rodrigues(0.0, 0.0, 0.01 * i), gtsam.Point3(0.1 * i, 0.0, -5.0))
        if i == 0:
            graph.add(gtsam.PriorFactorPose3(X(0), pose, pose_noise))                    # batch.py:281
            graph.add(gtsam.PriorFactorVector(V(0), np.zeros(3), vel_noise))             # batch.py:282
            expect += ["prior_pose", "prior_vel"]
            initial.insert(X(0), pose)
            initial.insert(V(0), np.zeros(3))
            continue
        initial.insert(X(i), pose)
        initial.insert(V(i), np.zeros(3))
        for _ in range(k_imu):
            pim.integrateMeasurement(np.array([0.0, 0.0, 9.81]) + 0.01 * rng.standard_normal(3), 0.01 * rng.standard_normal(3), 0.005)
        graph.push_back(gtsam.ImuFactor(X(i - 1), V(i - 1), X(i), V(i), B(0), pim))      # batch.py:291
        graph.push_back(gtsam.CustomFactor(dvl_noise, [V(i), X(i)], functools.partial(velocity_error, np.array([0.5, 0.0, 0.0]))))
        pim.resetIntegration()
        expect += ["imu", "dvl"]
        for lid in (i, i + 1):                                                           # two landmarks per frame, overlapping
            if lid not in seen:
                seen.add(lid)
                initial.insert(L(lid), gtsam.Point3(0.1 * lid, 0.2, -2.0))
            graph.push_back(gtsam.GenericStereoFactor3D(gtsam.StereoPoint2(1000.0, 960.0, 500.0), cam_noise, X(i), L(lid), K))
            expect.append("stereo")
    return graph, initial, expect


def test_factor_and_key_indexing_follows_batch_py():
    graph, initial, expect = _batch_py_graph()
    assert graph.size() == len(expect)
    prob = graph.to_problem(initial)
    order = [None] * graph.size()
    for t in ("prior_pose", "prior_vel", "between", "dvl", "stereo", "imu"):
        for o in prob[t]["orig"]:
            assert order[o] is None
            order[o] = t
    assert order == expect                                          # [PriorPose3, PriorVector, (Imu_i, Dvl_i, Stereo_i*)...]
    for name in ("bias_keys", "lm_keys", "vel_keys", "pose_keys"):  # ascending uint64 keys per kind
        k = prob[name]
        assert np.all(k[1:] > k[:-1])
    assert prob["dvl"]["v"].tolist() == prob["dvl"]["x"].tolist() == list(range(1, 6))      # keys [V(i), X(i)] (batch.py:247)
    assert prob["imu"]["xi"].tolist() == list(range(0, 5)) and prob["imu"]["xj"].tolist() == list(range(1, 6))
    assert np.allclose(prob["dvl"]["meas"], [[0.5, 0.0, 0.0]] * 5)
    assert prob["imu"]["pim"].shape == (5, 67) and prob["imu"]["sqrt_info"].shape == (5, 45)
    assert np.allclose(prob["imu"]["pim"][:, 15], 4 * 0.005)                                  # deltaTij of every factor


def test_values_api_and_errors():
    v = gtsam.Values()
    v.insert(X(0), gtsam.Pose3())
    v.insert(V(0), np.array([1.0, 2.0, 3.0]))
    assert v.exists(X(0)) and not v.exists(X(1)) and v.size() == 2
    assert np.allclose(v.atVector(V(0)), [1, 2, 3]) and v.atPose3(X(0)).x() == 0.0
    with pytest.raises(RuntimeError, match="already"):
        v.insert(X(0), gtsam.Pose3())
    with pytest.raises(RuntimeError, match="x7"):
        v.atPose3(X(7))


def test_missing_key_and_arbitrary_custom_factor_raise():
    graph, initial, _ = _batch_py_graph()
    graph.add(gtsam.PriorFactorPose3(X(99), gtsam.Pose3(), gtsam.noiseModel.Isotropic.Sigma(6, 1.0)))
    with pytest.raises(RuntimeError, match='"x99"'):
        graph.to_problem(initial)
    graph, initial, _ = _batch_py_graph()
    graph.add(gtsam.CustomFactor(gtsam.noiseModel.Isotropic.Sigma(3, 1.0), [X(0)], lambda this, values, H: np.zeros(3)))
    with pytest.raises(RuntimeError, match="CustomFactor"):
        graph.to_problem(initial)


def test_per_object_graph_solves_like_the_bulk_tables(emu):
    """The gtsam-style per-object construction and optimize() through the facade (kernel bodies via the host emulation)."""
    graph, initial, _ = _batch_py_graph(n=8)
    opt = gtsam.LevenbergMarquardtOptimizer(graph, initial, gtsam.LevenbergMarquardtParams(), lib=emu)
    e0 = opt.error()
    result = opt.optimize()
    assert opt.error() < e0 and opt.iterations() >= 1
    i = 0
    while result.exists(X(i)):                                      # constr3DPoints, batch.py:57-68
        p = result.atPose3(X(i))
        assert np.isfinite([p.x(), p.y(), p.z()]).all()
        i += 1
    assert i == 8


def test_save_graph_dot(tmp_path):
    graph, initial, expect = _batch_py_graph(n=3)
    path = tmp_path / "graph.dot"
    graph.saveGraph(str(path))                                       # batch.py:338
    txt = path.read_text()
    assert txt.startswith("graph {") and txt.count("shape=point") == len(expect) and 'label="x2"' in txt


def test_edge_cases(emu):
    # no poses at all
    empty = gtsam.NonlinearFactorGraph().to_problem(gtsam.Values())
    with pytest.raises(RuntimeError, match="no Pose3"):
        Session(empty, lib=emu)
    # a landmark nobody observes is an indeterminate system
    d = synthetic.make_trajectory_graph(12, seed=1, n_landmarks=4, obs_per_landmark=3)
    d["initial"].insert(L(1000), gtsam.Point3(0.0, 0.0, 0.0))
    with pytest.raises(RuntimeError, match="no stereo factor"):
        Session(d["graph"].to_problem(d["initial"]), lib=emu)
    # a single pose with only its prior: LM terminates immediately at zero error
    g, v = gtsam.NonlinearFactorGraph(), gtsam.Values()
    v.insert(X(0), gtsam.Pose3())
    g.add(gtsam.PriorFactorPose3(X(0), gtsam.Pose3(), gtsam.noiseModel.Isotropic.Sigma(6, 0.1)))
    s = Session(g.to_problem(v), lib=emu)
    res = s.optimize()
    assert res["iterations"] == 0 and res["final_error"] == 0.0
    s.close()
    # pose count that is not a multiple of the supernode size (padding nodes) and ragged tracks
    d = synthetic.make_trajectory_graph(23, seed=3, n_landmarks=9, obs_per_landmark=4)
    s = Session(d["graph"].to_problem(d["initial"]), lib=emu)
    lay = s.layout()
    assert lay["Ns"] * lay["k"] >= 23
    res = s.optimize()
    assert res["final_error"] < res["initial_error"] and res["solve_failures"] == 0
    s.close()


def test_c_abi_exports_every_declared_symbol(emu):
    header = open(os.path.join(ROOT, "include", "vus.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(vus_[a-z_0-9]+)\s*\(", header)) - {"vus_comm_fn"}
    assert declared, "no declarations parsed"
    assert declared == set(_native.EXPORTS), (declared ^ set(_native.EXPORTS))
    libs = [os.path.join(ROOT, "tests", "emu", "libvus_emu.so")]
    if os.path.exists(_native.LIB_PATH):
        libs.append(_native.LIB_PATH)                                # the sm_100a product library loads without a GPU
    for path in libs:
        lib = ctypes.CDLL(path)
        for name in declared:
            assert hasattr(lib, name), (path, name)


def test_marginals_facade_keys_and_errors(emu):
    """gtsam.Marginals facade: key lookup, joint blocks in the order asked, unknown keys raise."""
    d = synthetic.make_trajectory_graph(30, seed=3, n_landmarks=20, pixel_noise=1.0)
    m = gtsam.Marginals(d["graph"], d["initial"], lib=emu)
    c = m.marginalCovariance(X(10))
    assert c.shape == (6, 6) and np.all(np.linalg.eigvalsh(c) > 0)
    jm = m.jointMarginalCovariance([V(4), X(10), B(0)])
    assert jm.fullMatrix().shape == (15, 15)
    assert np.allclose(jm.at(X(10), X(10)), c, rtol=1e-8, atol=1e-16)
    assert np.array_equal(jm.at(V(4), B(0)), jm.at(B(0), V(4)).T)
    with pytest.raises(KeyError):
        m.marginalCovariance(X(999))


def test_marginals_and_batched_error_paths(emu):
    """Error behaviour of the rows added after the path: bad queries, unsupported combinations, malformed batches."""
    from visual_underwater_slam_b200 import parallel
    d = synthetic.make_pose_graph(30, seed=2, n_loops=2)
    prob = d["graph"].to_problem(d["initial"])
    s = Session(prob, lib=emu)
    with pytest.raises(RuntimeError, match="out of range"):
        s.marginal_covariance([("pose", 30)])
    with pytest.raises(RuntimeError, match="out of range"):
        s.marginal_covariance([("bias", 0)])                      # a pure pose graph has no bias variable
    with pytest.raises(RuntimeError, match="no batched optimize"):
        s.component_results()
    s.close()
    with pytest.raises(RuntimeError, match="node_start"):
        Session(prob, lib=emu, components=[1, 30])                # must start at 0
    with pytest.raises(RuntimeError, match="cover every pose"):
        Session(prob, lib=emu, components=[0, 10, 20])            # 30 poses
    two, node_start = parallel.concat_problems([prob, prob])
    sb = Session(two, lib=emu, components=node_start)
    with pytest.raises(RuntimeError, match="batched graph"):
        sb.marginal_covariance([("pose", 0)])
    with pytest.raises(RuntimeError, match="batched graph"):
        sb.solve_step(1e-3)
    sb.close()
    ds = synthetic.make_trajectory_graph(20, seed=3, n_landmarks=10, pixel_noise=1.0)
    with pytest.raises(NotImplementedError, match="stereo"):
        parallel.concat_problems([ds["graph"].to_problem(ds["initial"])] * 2)
    with pytest.raises(ValueError):
        parallel.concat_problems([])


def test_graph_and_values_merge_for_incremental_updates():
    """NonlinearFactorGraph.push_back(graph) keeps insertion order across graphs; Values.insert(Values) refuses duplicates."""
    d = synthetic.make_trajectory_graph(12, seed=1, pixel_noise=1.0)
    g = gtsam.NonlinearFactorGraph()
    g.push_back(d["graph"])
    g.push_back(d["graph"])
    n = d["graph"].size()
    assert g.size() == 2 * n
    orig = np.sort(np.concatenate([g.table(t)["orig"] for t in ("prior_pose", "prior_vel", "between", "dvl", "stereo", "imu")]))
    assert np.array_equal(orig, np.arange(2 * n))
    v = gtsam.Values()
    v.insert(d["initial"])
    assert v.size() == d["initial"].size()
    with pytest.raises(RuntimeError, match="already exists"):
        v.insert(d["initial"])
    g.resize(0)
    assert g.size() == 0


def test_optimize_many_is_n_separate_optimizers(emu):
    """gtsam-named entry for BASELINE config 4: N graphs + N initial Values in, N Values out, each equal to what its own
    LevenbergMarquardtOptimizer returns (keys of every trajectory are its own X(i) / V(i) / B(0))."""
    ds = [synthetic.make_trajectory_graph(40 + 8 * t, seed=50 + t, n_loops=2, loop_min_gap=12) for t in range(3)]
    vals, info = gtsam.optimize_many([d["graph"] for d in ds], [d["initial"] for d in ds], lib=emu)
    assert len(vals) == 3
    for d, v, s in zip(ds, vals, info):
        opt = gtsam.LevenbergMarquardtOptimizer(d["graph"], d["initial"], gtsam.LevenbergMarquardtParams(), lib=emu)
        ref = opt.optimize()
        assert s["iterations"] == opt.iterations() and abs(s["error"] - opt.error()) <= 1e-9 * opt.error()
        n = d["meta"]["n_poses"]
        assert v.exists(X(n - 1)) and not v.exists(X(n)) and v.exists(B(0))
        a, b = v.atPose3(X(n - 1)), ref.atPose3(X(n - 1))
        assert abs(a.x() - b.x()) + abs(a.y() - b.y()) + abs(a.z() - b.z()) < 1e-7
    with pytest.raises(ValueError):
        gtsam.optimize_many([], [])


def test_batch_py_import_list_resolves():
    """Every name batch.py imports from gtsam (batch.py:19-27) exists in the package, so its import block works with the
    module swapped; the ones it never uses on the path raise when constructed."""
    import importlib
    src = open("/root/reference/batch.py").read() if os.path.exists("/root/reference/batch.py") else None
    names = ["ISAM2", "BetweenFactorConstantBias", "Cal3_S2", "ConstantTwistScenario", "ImuFactor", "NonlinearFactorGraph",
             "PinholeCameraCal3_S2", "Point3", "Pose3", "PriorFactorConstantBias", "PriorFactorPose3", "PriorFactorVector", "Rot3",
             "Values", "PriorFactorPoint3", "NavState", "Cal3_S2Stereo", "StereoPoint2", "GenericStereoFactor3D"]
    if src is not None:                                   # the list above is batch.py's own (checked where the reference exists)
        m = re.search(r"from gtsam import \((.*?)\)", src, re.S)
        assert sorted(n.strip() for n in m.group(1).replace("\n", " ").split(",") if n.strip()) == sorted(names)
    for n in names:
        assert hasattr(gtsam, n), n
    sh = importlib.import_module("visual_underwater_slam_b200.symbol_shorthand") if False else gtsam.symbol_shorthand
    assert all(hasattr(sh, k) for k in "BVXL")
    plot = importlib.import_module("visual_underwater_slam_b200.utils").plot
    with pytest.raises(NotImplementedError):
        plot.plot_trajectory(1, None)
    with pytest.raises(NotImplementedError):
        gtsam.NavState()


def test_partition_chain_arguments_are_checked(emu):
    """vus_set_partition_chain rejects an impossible chain position instead of storing it."""
    import ctypes as C
    h = C.c_void_p()
    assert emu.vus_create(0, C.byref(h)) == 0
    try:
        side = (C.c_int64 * 16)(*([-1] * 16))
        assert emu.vus_set_partition_chain(h, 16, side, side, 0, 2) == 0
        assert emu.vus_set_partition_chain(h, 16, side, side, 2, 2) != 0            # rank >= nranks
        assert emu.vus_set_partition_chain(h, 16, None, side, 0, 2) != 0            # missing list
        assert b"vus_set_partition_chain" in emu.vus_last_error(h)
    finally:
        emu.vus_destroy(h)
