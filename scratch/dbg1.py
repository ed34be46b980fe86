import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import cases, oracle
from stencilstream_b200 import Grid, Params, StencilUpdate
wl = sys.argv[1]; rows=int(sys.argv[2]); cols=int(sys.argv[3]); n=int(sys.argv[4])
fused = int(sys.argv[5]) if len(sys.argv)>5 else 0
p,h,cells = cases.make_case(wl, rows, cols, seed=1)
u = StencilUpdate(wl, Params(transition_function=p, halo_value=h, n_iterations=n, blocking=True, fused_iterations=fused))
try:
    got = u(Grid(wl, buffer=cells)).to_numpy()
except Exception as e:
    print(wl, 'TMA=%s'%os.environ.get('STST_TMA','1'), 'FAILED', str(e)[-80:]); sys.exit(0)
want = oracle.best().run(wl, p, h, cells, 0, n)
s = u.get_stats()
print(wl, 'TMA=%s'%os.environ.get('STST_TMA','1'), 'k=%d tile=%dx%d smem=%d'%(s.fused_iterations,s.tile_h,s.tile_w,s.smem_bytes), 'exact' if got.tobytes()==want.tobytes() else 'err=%g'%cases.rel_max_norm(got,want))
