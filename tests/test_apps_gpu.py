"""GPU tests of the application loops (stencilstream_b200/apps.py) against the same loops run with
the CPU oracle and host-side (numpy) glue — the way the reference's `main` functions do it
(examples/convection/convection.cpp:402-477, examples/fdtd/src/fdtd.cpp:218-252).
Strict (-fmad=false) build on the GPU side: everything must agree bit for bit, including the number
of pseudo-transient batches the convergence check lets through and the adaptive time step."""
import numpy as np
import pytest

from stencilstream_b200 import workloads as W
from cases import oracle_convection
from stencilstream_b200.apps import run_convection, run_fdtd

pytestmark = pytest.mark.gpu


def test_convection_application_loop_matches_oracle(oracle_best):
    config = W.convection_benchmark_config(res=48, n_iters=60, lx=1.0, ly=1.5)
    config.update(nt=4, nerr=10, nout=2, epsilon=3e-2, Ra=1e7)
    want_cells, want_steps, want_frames = oracle_convection(oracle_best, config)
    # the case exercises what it should: early and late exits of the convergence loop, and the
    # advective (velocity-dependent) time step
    assert sorted({s[1] for s in want_steps}) == [10, 20, 50, 60]
    assert all(s[4] < W.ConvectionExperiment(config).dt_diff for s in want_steps)
    frames = []
    grid, steps = run_convection(config, strict=True,
                                 on_frame=lambda it, T: frames.append((it, T.copy())))
    got_steps = [(s.it, s.iterations, s.errV, s.errP, s.dt) for s in steps]
    assert got_steps == want_steps
    assert grid.to_numpy().tobytes() == want_cells.tobytes()
    assert [it for it, _ in frames] == [it for it, _ in want_frames]
    for (_, got), (_, want) in zip(frames, want_frames):
        assert got.tobytes() == want.tobytes()


def test_fdtd_snapshot_loop_matches_oracle(oracle_best):
    """A shortened run of the default experiment (161 x 161 cells): 3 full snapshot intervals and a
    fourth that overshoots n_timesteps, exactly like the reference's loop does."""
    exp = W.FdtdExperiment(W.FDTD_DEFAULT)
    cells = exp.initial_grid()
    total, snap = 70, 20
    frames = []
    grid, simulation = run_fdtd(W.FDTD_DEFAULT, n_timesteps=total, n_snap_timesteps=snap,
                                strict=True,
                                on_frame=lambda f, i, v: frames.append((f, i, v.copy())))
    n_done = -(-total // snap) * snap
    want = oracle_best.run("fdtd", exp.kernel_params(), None, cells, 0, n_done)
    assert grid.to_numpy().tobytes() == want.tobytes()
    assert [(f, i) for f, i, _ in frames] == [("hz", 20), ("hz", 40), ("hz", 60), ("hz", 80),
                                               ("hz_sum", total)]
    mid = oracle_best.run("fdtd", exp.kernel_params(), None, cells, 0, 40)
    assert frames[1][2].tobytes() == np.ascontiguousarray(mid["hz"]).tobytes()
    assert frames[-1][2].tobytes() == np.ascontiguousarray(want["hz_sum"]).tobytes()
    assert simulation.get_n_processed_cells() == n_done * cells.size
