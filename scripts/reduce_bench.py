"""Time Grid.max_abs (five convection norms) against the accessor route it replaces."""
import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np
import bench
from stencilstream_b200 import Grid
from stencilstream_b200.apps import convection_norm_extents
rows, cols = 4096, 8192
params, halo, fill = bench.make_workload('convection_pt', rows, cols)
g = Grid('convection_pt', rows, cols)
v = g.accessor('write'); fill(v, 0, rows, rows); v['Vx'][100, 100] = -7.5; del v
g.sync_to_device()
ext = convection_norm_extents(rows - 1, cols - 1)
for i in range(4):
    t0 = time.perf_counter(); n = g.max_abs(ext); t1 = time.perf_counter()
    print(f'max_abs x5: {(t1-t0)*1e3:.3f} ms  {5*8*rows*cols/(t1-t0)/1e9:.0f} GB/s algorithmic', n, flush=True)
T = None
for i in range(3):
    t0 = time.perf_counter(); T = g.field_to_numpy('T', out=T); t1 = time.perf_counter()
    print(f'field_to_numpy(T) pageable dst (call {i}): {(t1-t0)*1e3:.1f} ms {T.nbytes/(t1-t0)/1e9:.1f} GB/s')
t0 = time.perf_counter(); a = g.accessor('read'); t1 = time.perf_counter()
print(f'whole-grid accessor (what the reference route moves): {(t1-t0)*1e3:.1f} ms')
