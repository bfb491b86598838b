import sys, numpy as np
sys.path.insert(0,'tests')
from visual_underwater_slam_b200 import _native
from visual_underwater_slam_b200.optimizer import Session, LevenbergMarquardtParams
import parity_common as pc
lib = _native.bind('tests/emu/libvus_emu.so')
_, prob = pc.make(120, n_loops=3, loop_min_gap=40)
p = LevenbergMarquardtParams(); p.verbosityLM='TRYDELTA'
s = Session(prob, p, lib=lib)
st = s.solve_step(1e-5)
