"""Runs LAST in the GPU suite (file name): the first GPU run of a path finished after the round's GPU budget was spent."""
import numpy as np
import pytest
import parity_common as pc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from visual_underwater_slam_b200 import _native
    return _native.load()


@pytest.mark.xfail(strict=False, reason="the segment preconditioner for tracks longer than the band was finished after the round's "
                   "GPU budget was spent: verified on the host emulation only (tests/test_emu.py), this is its first run on a GPU")
def test_long_tracks_on_the_gpu(lib):
    """13 observations per landmark (longer than any band): exact implicit Schur term in the operator, segment terms in the
    preconditioner; the oracle's LM path, few PCG iterations per solve, marginal queries answered."""
    from visual_underwater_slam_b200 import synthetic
    d = synthetic.make_trajectory_graph(74, seed=1, n_landmarks=60, obs_per_landmark=13, pixel_noise=1.0)
    prob = d["graph"].to_problem(d["initial"])
    res, info = pc.check_lm_parity(lib, prob)
    assert res["pcg_iterations"] <= 20 * res["inner_iterations"]
    pc.check_marginals(lib, prob, [("pose", 40), ("lm", 7), ("vel", 12), ("bias", 0)], rtol=1e-5)
