"""Run one configuration a few times (for ncu / quick timing)."""
import argparse, os, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import bench
from stencilstream_b200 import Grid, Params, StencilUpdate, workload_info
ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='jacobi5'); ap.add_argument('--rows', type=int, default=16384); ap.add_argument('--cols', type=int, default=16384)
ap.add_argument('--iters', type=int, default=8); ap.add_argument('--fuse', type=int, default=0); ap.add_argument('--calls', type=int, default=2)
ap.add_argument('--devices', default='', help='comma-separated CUDA ordinals: spread every call over these GPUs (Params.cuda_devices)')
a = ap.parse_args()
params, halo, fill = bench.make_workload(a.workload, a.rows, a.cols)
grid = Grid(a.workload, a.rows, a.cols)
v = grid.accessor('write'); fill(v, 0, a.rows, a.rows); del v
grid.sync_to_device()
timer = bench.StreamTimer(0)
devices = [int(d) for d in a.devices.split(',') if d] or None
u = StencilUpdate(a.workload, Params(transition_function=params, halo_value=halo, n_iterations=a.iters, fused_iterations=a.fuse, cuda_devices=devices))
for i in range(a.calls):
    timer.begin(); out = u(grid); ms = timer.end_ms()
    s = u.get_stats()
    print(f'{a.workload} slabs={s.n_slabs} k={s.fused_iterations} tile={s.tile_h}x{s.tile_w} call {i}: {ms:.3f} ms  {a.rows*a.cols*a.iters/ms/1e6:.1f} GCells/s', flush=True)
