import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from visual_underwater_slam_b200 import synthetic, _native, parallel
from visual_underwater_slam_b200.optimizer import Session, LevenbergMarquardtParams
T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 500
lib_path = sys.argv[3] if len(sys.argv) > 3 else None
lib = _native.bind(lib_path) if lib_path else _native.load()
t0 = time.time()
probs = []
for t in range(T):
    d = synthetic.make_trajectory_graph(n, seed=4 + t, n_loops=5, loop_min_gap=100 if n >= 300 else n // 3, pixel_noise=1.0)
    probs.append(d['graph'].to_problem(d['initial']))
print('generated %d trajectories x %d poses in %.1f s' % (T, n, time.time() - t0), flush=True)
for rep in range(2):
    t0 = time.perf_counter()
    prob, node_start = parallel.concat_problems(probs)
    t1 = time.perf_counter()
    s = Session(prob, LevenbergMarquardtParams(), lib=lib, components=node_start)
    t2 = time.perf_counter()
    tot = s.optimize()
    t3 = time.perf_counter()
    per = s.component_results()
    vals = s.values()
    t4 = time.perf_counter()
    s.close()
    its = [r['iterations'] for r in per]; tries = [r['inner_iterations'] for r in per]
    print('batched: concat %.1f ms, session %.1f ms, optimize %.1f ms (%d rounds, %d pcg its, %d launches), values %.1f ms -> %.3f ms/trajectory; iterations %d..%d tries %d..%d' % (
        1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), tot['inner_iterations'], tot['pcg_iterations'], tot['kernel_launches'], 1e3*(t4-t3), 1e3*(t3-t2)/T, min(its), max(its), min(tries), max(tries)), flush=True)
# the one-handle-per-trajectory path on a subset, for reference
m = min(T, 16)
for threads in (1, 8):
    t0 = time.perf_counter()
    loc = parallel.solve_local(probs[:m], lib=lib, threads=threads, keep_values=False)
    dt = time.perf_counter() - t0
    print('solve_local threads=%d: %.2f ms/trajectory' % (threads, 1e3 * dt / m), flush=True)
bad = 0
for r, l in zip(per[:m], loc):
    if r['iterations'] != l['iterations'] or r['inner_iterations'] != l['inner_iterations'] or abs(r['final_error'] - l['final_error']) > 1e-9 * l['final_error']:
        bad += 1
print('batched vs per-handle mismatches on the first %d: %d' % (m, bad))

pp = LevenbergMarquardtParams(); pp.profileKernels = True
s = Session(prob, pp, lib=lib, components=node_start)
tot = s.optimize()
print('profile (ms per class):', {k: round(v, 1) for k, v in tot['ms_class'].items() if v > 0})
print('launches per class:', {k: v for k, v in tot['launches_class'].items() if v > 0})
s.close()
