"""device IMU preintegration and stereo back-projection against the host restatements
(navigation.preintegrate_batch follows gtsam's Tangent- / ManifoldPreintegration; the back-projection follows StereoCamera)."""
import os
import subprocess
import numpy as np
import pytest
from visual_underwater_slam_b200 import _native, synthetic
from visual_underwater_slam_b200.navigation import preintegrate_batch
from visual_underwater_slam_b200.optimizer import preintegrate_imu, backproject_stereo, Session
import parity_common as pc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _imu(n, k, seed=0):
    rng = np.random.default_rng(seed)
    acc = np.array([0.1, -0.2, 9.81]) + 0.3 * rng.standard_normal((n, k, 3))
    gyro = 0.2 * rng.standard_normal((n, k, 3))
    return acc, gyro


def _check_preint(lib):
    params = synthetic.reference_imu_params()
    for k, bias, tangent in ((40, None, True), (7, np.array([0.02, -0.01, 0.015, 0.002, -0.001, 0.0015]), True),
                             (40, None, False), (7, np.array([0.02, -0.01, 0.015, 0.002, -0.001, 0.0015]), False)):
        acc, gyro = _imu(33, k, seed=k)
        pim_h, info_h, cov_h = preintegrate_batch(acc, gyro, synthetic.IMU_DT, params, bias, tangent=tangent)
        pim_d, info_d = preintegrate_imu(acc, gyro, synthetic.IMU_DT, params, bias, lib=lib, tangent=tangent)
        assert np.abs(pim_d - pim_h).max() <= 1e-12 * max(1.0, np.abs(pim_h).max())
        iu = np.triu_indices(9)
        Rh = np.zeros((33, 9, 9)); Rh[:, iu[0], iu[1]] = info_h
        Rd = np.zeros((33, 9, 9)); Rd[:, iu[0], iu[1]] = info_d
        Mh, Md = np.swapaxes(Rh, 1, 2) @ Rh, np.swapaxes(Rd, 1, 2) @ Rd          # information matrices
        assert np.abs(Md - Mh).max() <= 1e-8 * np.abs(Mh).max()
        assert np.abs(Md @ cov_h - np.eye(9)).max() < 1e-6                         # R^T R really is preintMeasCov^-1
        assert np.abs(info_d - info_h).max() <= 1e-7 * np.abs(info_h).max()


def _check_backproject(lib):
    d, prob = pc.make(30, n_lm=60, noise_scale=0.0)
    s = Session(prob, lib=lib)
    f = prob["stereo"]
    pts = backproject_stereo(s, f["x"], f["meas"])
    s.close()
    assert np.abs(pts - d["truth"]["lms"][f["l"]]).max() < 1e-8     # noise-free measurements back-project onto the landmark


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "visual_underwater_slam_b200", "csrc"), "emu"], check=True)
    return _native.bind(os.path.join(ROOT, "tests", "emu", "libvus_emu.so"))


def test_preintegration_emu(emu):
    _check_preint(emu)


def test_backprojection_emu(emu):
    _check_backproject(emu)


@pytest.mark.gpu
def test_preintegration_gpu():
    _check_preint(_native.load())


@pytest.mark.gpu
def test_backprojection_gpu():
    _check_backproject(_native.load())
