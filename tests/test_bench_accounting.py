"""bench.py's roofline accounting (no GPU): the algorithmic flop / byte figures DESIGN.md section 3 states, and the measured-traffic
look-up that feeds `roofline.traffic`."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def test_chunk_factor_flops_at_config3():
    fl, band = bench.chunk_factor_flops(11112, 81, 9, 148)
    assert abs(fl / 1e9 - 27.59) < 0.05                     # 27.6 GFLOP per chunked factorization
    assert abs(band / 1e9 - 7.29) < 0.05                    # sequential band Cholesky at the true half-bandwidth
    assert abs(bench.bcr_factor_flops(11112, 81) / 1e9 - 70.83) < 0.05   # cyclic reduction of the whole chain
    assert band < fl < bench.bcr_factor_flops(11112, 81)


def test_measured_traffic_lookup_matches_the_committed_capture():
    lay = {"Ns": 11112, "band_chunks": 148}
    per_try = bench.ncu_traffic_r2("bcr_factor", lay, {"launches": 57}, 1)
    assert 2.5e9 < per_try < 3.2e9                          # ChunkFactorBody: ~2.87 GB of DRAM traffic per launch
    per_solve = bench.ncu_traffic_r2("bcr_solve", lay, {"launches": 20}, 1)
    assert 3.5e9 < per_solve < 4.3e9                        # forward + backward chunk sweeps
    assert bench.ncu_traffic_r2("bcr_factor", {"Ns": 5000, "band_chunks": 148}, {"launches": 1}, 1) is None   # other layout: no capture
    assert bench.ncu_traffic_r2("linearize", lay, {"launches": 1}, 1) is None
