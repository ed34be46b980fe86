import argparse, os, sys, threading, time, subprocess
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import bench
from stencilstream_b200 import Grid, Params, StencilUpdate
ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='jacobi5'); ap.add_argument('--rows', type=int, default=16384); ap.add_argument('--cols', type=int, default=16384)
ap.add_argument('--iters', type=int, default=100); ap.add_argument('--fuse', type=int, default=1); ap.add_argument('--calls', type=int, default=3)
ap.add_argument('--profiling', type=int, default=1)
a = ap.parse_args()
params, halo, fill = bench.make_workload(a.workload, a.rows, a.cols)
grid = Grid(a.workload, a.rows, a.cols)
v = grid.accessor('write'); fill(v, 0, a.rows, a.rows); del v
grid.sync_to_device()
timer = bench.StreamTimer(0)
with bench.ClockSampler(0) as clocks:
  for i in range(a.calls):
    u = StencilUpdate(a.workload, Params(transition_function=params, halo_value=halo, n_iterations=a.iters, fused_iterations=a.fuse, profiling=bool(a.profiling)))
    t0=time.perf_counter(); timer.begin(); out = u(grid); t1=time.perf_counter(); ms = timer.end_ms()
    s = u.get_stats()
    print(f'k={s.fused_iterations} call {i}: total {ms:.3f} ms, host submit {1e3*(t1-t0):.2f} ms, sum of kernel events {s.kernel_runtime*1e3:.3f} ms over {s.n_launches} launches -> {s.kernel_runtime*1e6/max(s.n_launches,1):.1f} us/launch', flush=True)
print(clocks.summary(), 'samples', clocks.samples[:40])
