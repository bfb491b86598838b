import sys, time, numpy as np
from visual_underwater_slam_b200 import synthetic, _native
from visual_underwater_slam_b200.optimizer import Session, LevenbergMarquardtParams
from oracle import lm
lib = _native.bind('tests/emu/libvus_emu.so')
n_poses, n_lm, n_loops = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
d = synthetic.make_trajectory_graph(n_poses, seed=1, n_landmarks=n_lm, n_loops=n_loops, pixel_noise=1.0)
prob = d['graph'].to_problem(d['initial'])
s = Session(prob, lib=lib)
print('layout', s.layout())
vals = lm.values_of(prob)
print('error', s.error(), lm.graph_error(prob, vals))
fe = s.factor_errors(); fo = lm.factor_errors(prob, vals)
print('factor err max rel diff', np.max(np.abs(fe-fo)/(np.abs(fo)+1e-12)))
for name in ['prior_pose','prior_vel','between','dvl','stereo','imu']:
    ev = lm.eval_factors(prob, vals, name)
    if ev is None: continue
    r, J = s.linearize(name)
    ro, Js, _ = ev
    if name=='dvl': Jo = np.concatenate([Js[1], Js[0]], 2)
    else: Jo = np.concatenate(Js, 2)
    print(name, 'r', np.abs(r-ro).max()/ (np.abs(ro).max()+1e-300), 'J', np.abs(J-Jo).max()/np.abs(Jo).max())
lay = lm.Layout(prob)
Jm,b = lm.linearize(prob, vals, lay)
for lam in [1e-5, 1.0]:
    t=time.time(); delta = lm.solve_damped(Jm,b,lam,lay); t1=time.time()-t
    t=time.time(); st = s.solve_step(lam); t2=time.time()-t
    dx = delta[lay.ox:].reshape(-1,6); dv = delta[lay.ov:lay.ox].reshape(-1,3); db=delta[:6]; dl = delta[lay.ol:lay.ov].reshape(-1,3)
    rel = lambda a,b: np.linalg.norm(a-b)/(np.linalg.norm(b)+1e-300)
    print('lam',lam,'pcg',st['pcg_iterations'],'pose',rel(st['pose'],dx),'vel',rel(st['vel'],dv),'bias',rel(st['bias'][0],db),'lm',rel(st['lm'],dl) if n_lm else None, 'times', t1,t2)
