import sys, time, numpy as np
from visual_underwater_slam_b200 import synthetic, _native
from visual_underwater_slam_b200.optimizer import Session, LevenbergMarquardtParams
from oracle import lm
lib = _native.bind('tests/emu/libvus_emu.so')
n_poses, n_lm, n_loops = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
d = synthetic.make_trajectory_graph(n_poses, seed=1, n_landmarks=n_lm, n_loops=n_loops, pixel_noise=1.0)
prob = d['graph'].to_problem(d['initial'])
p = LevenbergMarquardtParams()
if len(sys.argv)>4: p.pcgRelTol=float(sys.argv[4])
s = Session(prob, p, lib=lib)
t=time.time(); res = s.optimize(); t_emu=time.time()-t
print(res)
t=time.time(); vals, info = lm.lm_optimize(prob); t_or=time.time()-t
print('oracle iters', info['iterations'], 'err', info['error'], 'time', t_or, 'emu time', t_emu)
print('rel err diff', abs(res['final_error']-info['error'])/info['error'], 'iters', res['iterations'], info['iterations'])
v = s.values()
print('pose t diff', np.abs(v['poses'][:,9:]-vals['poses'][:,9:]).max(), 'R diff', np.abs(v['poses'][:,:9]-vals['poses'][:,:9]).max())
print('oracle errors', [f"{e:.9e}" for e in info['trace']['errors']])
