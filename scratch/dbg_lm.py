import sys, numpy as np
sys.path.insert(0,'tests')
from visual_underwater_slam_b200 import _native
from visual_underwater_slam_b200.optimizer import Session, LevenbergMarquardtParams
import parity_common as pc
from oracle import lm
lib = _native.bind('tests/emu/libvus_emu.so')
_, prob = pc.make(120, n_loops=3, loop_min_gap=40)
for tol in [1e-12, 1e-14]:
    p = LevenbergMarquardtParams(); p.pcgRelTol = tol
    s = Session(prob, p, lib=lib); res = s.optimize(); v = s.values()
    vals, info = lm.lm_optimize(prob)
    dt = v["poses"][:, 9:] - vals["poses"][:, 9:]
    print(tol, res['iterations'], info['iterations'], res['pcg_iterations'], 'err rel', abs(res['final_error']-info['error'])/info['error'], 'rmse', np.sqrt((dt**2).sum(1).mean()), 'max', np.abs(dt).max())
# oracle self-consistency: perturb the oracle's solve by using schur=False (different elimination order)
vals2, info2 = lm.lm_optimize(prob, schur=False)
import scipy.sparse.linalg as spla
print('oracle self', info2['iterations'], abs(info2['error']-info['error'])/info['error'])
# different ordering of the direct solver
old = lm._splu_sym
lm._splu_sym = lambda A: spla.splu(A.tocsc(), permc_spec='COLAMD', diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
vals3, info3 = lm.lm_optimize(prob)
dt = vals3["poses"][:, 9:] - vals["poses"][:, 9:]
print('oracle COLAMD vs MMD: iters', info3['iterations'], 'err rel', abs(info3['error']-info['error'])/info['error'], 'rmse', np.sqrt((dt**2).sum(1).mean()))
