import sys, numpy as np
sys.path.insert(0,'tests')
from visual_underwater_slam_b200 import _native
from visual_underwater_slam_b200.optimizer import Session, LevenbergMarquardtParams
import parity_common as pc
from oracle import lm
lib = _native.bind('tests/emu/libvus_emu.so')
_, prob = pc.make(120, n_loops=3, loop_min_gap=40)
lay = lm.Layout(prob)
vals = lm.values_of(prob)
lam = 1e-5
import copy
for it in range(4):
    p2 = dict(prob); p2['poses']=vals['poses']; p2['vels']=vals['vels']; p2['biases']=vals['biases']; p2['lms']=vals['lms']
    s = Session(p2, lib=lib)
    J,b = lm.linearize(prob, vals, lay)
    delta = lm.solve_damped(J,b,lam,lay)
    st = s.solve_step(lam)
    mine = np.concatenate([st["bias"].ravel(), st["lm"].ravel(), st["vel"].ravel(), st["pose"].ravel()])
    H = (J.T@J).tocsr(); g = J.T@b
    r1 = np.linalg.norm(H@mine+lam*mine-g)/np.linalg.norm(g); r2 = np.linalg.norm(H@delta+lam*delta-g)/np.linalg.norm(g)
    print(it, 'lam', lam, 'rel diff', np.linalg.norm(mine-delta)/np.linalg.norm(delta), 'res mine', r1, 'res oracle', r2, 'pcg', st['pcg_iterations'])
    nv_o = lm.retract(vals, delta, lay); nv_m = lm.retract(vals, mine, lay)
    print('   err after oracle step', lm.graph_error(prob, nv_o), 'after my step', lm.graph_error(prob, nv_m))
    vals = nv_o; lam/=10
    s.close()
