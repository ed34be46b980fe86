"""Tuning sweep: GCell-updates/s for one workload over fusion depth / CTA shape / tile rows."""
import argparse, itertools, os, sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np
import bench
from stencilstream_b200 import Grid, Params, StencilUpdate, workload_info

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='jacobi5')
ap.add_argument('--rows', type=int, default=16384)
ap.add_argument('--cols', type=int, default=16384)
ap.add_argument('--iters', type=int, default=48)
ap.add_argument('--fuse', default='1,2,3,4,6,8')
ap.add_argument('--ctas', default='2')
ap.add_argument('--by', default='0')
ap.add_argument('--bx', default='0')
ap.add_argument('--tile_rows', default='0')
ap.add_argument('--tma', default='1')
ap.add_argument('--reps', type=int, default=3)
a = ap.parse_args()

params, halo, fill = bench.make_workload(a.workload, a.rows, a.cols)
info = workload_info(a.workload)
grid = Grid(a.workload, a.rows, a.cols)
v = grid.accessor('write'); fill(v, 0, a.rows, a.rows); del v
grid.sync_to_device()
timer = bench.StreamTimer(0)
ints = lambda s: [int(x) for x in s.split(',')]
print('workload rows cols | fuse ctas bx by trows tma | k tile smem | GCells/s  eff%%of8TB/s  ms/iter')
for fuse, ctas, bx, by, tr, tma in itertools.product(ints(a.fuse), ints(a.ctas), ints(a.bx), ints(a.by), ints(a.tile_rows), ints(a.tma)):
    os.environ.update(STST_CTAS_PER_SM=str(ctas), STST_BLOCK_X=str(bx), STST_BLOCK_Y=str(by), STST_TILE_ROWS=str(tr), STST_TMA=str(tma))
    iters = max(fuse, 1) * max(1, a.iters // max(fuse, 1))
    try:
        u = StencilUpdate(a.workload, Params(transition_function=params, halo_value=halo, n_iterations=iters, fused_iterations=fuse))
        out = u(grid); timer.sync()
        best = 1e30
        for _ in range(a.reps):
            timer.begin(); out = u(grid); ms = timer.end_ms(); best = min(best, ms)
        s = u.get_stats()
        g = a.rows * a.cols * iters / (best * 1e-3) / 1e9
        print(f'{a.workload} {a.rows} {a.cols} | {fuse} {ctas} {bx} {by} {tr} {tma} | k={s.fused_iterations} {s.tile_h}x{s.tile_w} b={s.block_x}x{s.block_y} smem={s.smem_bytes} | {g:8.1f} {100*g*info.bytes_per_cell_iteration/8000:6.1f}% {best/iters:7.4f}', flush=True)
        del out, u
    except Exception as e:
        print(f'{a.workload} | {fuse} {ctas} {bx} {by} {tr} {tma} | FAILED {str(e)[-100:]}', flush=True)
