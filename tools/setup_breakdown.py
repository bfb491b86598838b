import sys, time, numpy as np, torch, collections
sys.path.insert(0, '.')
from visual_underwater_slam_b200 import synthetic, _native
from visual_underwater_slam_b200.optimizer import Session, LevenbergMarquardtParams
d = synthetic.make_trajectory_graph(100000, seed=3, n_landmarks=200000, pixel_noise=1.0, drift_scale=0.1)
prob = d['graph'].to_problem(d['initial'])
def pin(x): return torch.from_numpy(np.ascontiguousarray(x)).pin_memory().numpy()
pinned = dict(prob)
for k in ("poses", "vels", "biases", "lms"): pinned[k] = pin(prob[k])
for k in ("prior_pose", "prior_vel", "between", "dvl", "stereo", "imu"):
    f = dict(prob[k]); f["meas"] = pin(f["meas"]); f["sqrt_info"] = pin(f["sqrt_info"]); pinned[k] = f
lib = _native.load()
acc = collections.OrderedDict()
class W:
    def __init__(self, lib): self._lib = lib
    def __getattr__(self, n):
        f = getattr(self._lib, n)
        def g(*a):
            t0 = time.perf_counter(); r = f(*a); torch.cuda.synchronize(); acc[n] = acc.get(n, 0) + time.perf_counter() - t0; return r
        return g
p = LevenbergMarquardtParams()
for i in range(5):
    acc.clear()
    torch.cuda.synchronize(); t0=time.perf_counter()
    s = Session(pinned, p, lib=W(lib)); t1=time.perf_counter()
    r = s.optimize(); t2=time.perf_counter()
    v = s.values(); t3=time.perf_counter()
    s.close(); t4=time.perf_counter()
    print('session %.1f optimize %.1f values %.1f close %.1f total %.1f ms' % tuple(1e3*x for x in (t1-t0, t2-t1, t3-t2, t4-t3, t4-t0)))
    print('   ', {k: round(1e3*v, 2) for k, v in acc.items()})
