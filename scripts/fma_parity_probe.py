"""How does the default (-fmad=true) GPU build compare with the CPU oracles, contracted and not?

For every workload: the GPU result of a seeded case against (a) the reference-built oracle compiled
with -ffp-contract=fast -mfma, (b) the C port compiled the same way, (c) the uncontracted oracle.
Prints bit-exactness and the relative max-norm; the answer decides which oracle the parity tests of
the default build use (tests/test_parity_gpu.py, tests/test_parity_fullsize_gpu.py).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

import cases
import oracle
from stencilstream_b200 import Grid, Params, StencilUpdate

oracle.set_threads()
checkers = {"ref_fma": oracle.reference(fma=True), "port_fma": oracle.port(fma=True),
            "ref_plain": oracle.best()}
for workload in ["jacobi5", "jacobi9", "jacobi_r2", "jacobi_r3", "hotspot", "fdtd", "convection_pt",
                 "convection_thermal"]:
    for shape, n in (((67, 93), 5), ((300, 420), 60), ((700, 900), 400)):
        if workload.startswith("convection") and n > 60:
            continue
        params, halo, cells = cases.make_case(workload, *shape, seed=4)
        update = StencilUpdate(workload, Params(transition_function=params, halo_value=halo,
                                                n_iterations=n, blocking=True))
        got = update(Grid(workload, buffer=cells)).to_numpy()
        line = f"{workload:20s} {shape[0]}x{shape[1]} n={n:4d} k={update.get_stats().fused_iterations}:"
        for name, checker in checkers.items():
            if checker is None:
                line += f"  {name}: n/a"
                continue
            want = checker.run(workload, params, halo, cells, 0, n)
            exact = got.tobytes() == want.tobytes()
            line += f"  {name}: {'EXACT' if exact else 'err %.2e' % cases.rel_max_norm(got, want)}"
        print(line, flush=True)
