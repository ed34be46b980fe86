"""TEST INFRASTRUCTURE — check a window of a full-size result against the CPU oracle.

A window [wr0, wr1) x [wc0, wc1) of the grid after n iterations depends only on the input cells
within d = n * n_subiterations * radius of it (and on the grid border where that region crosses it).
`expected_window` cuts that region out of the input, runs the oracle on the crop with GLOBAL
coordinates (oracle_run_window2d: the reference cpu backend under a coordinate-shifting wrapper, or
the C port) and returns the window's cells — exactly what a whole-grid oracle run would hold there,
at (w + 2d)^2 instead of 16384^2 cells of work per iteration.
"""
from __future__ import annotations

import numpy as np

WORKLOAD_SHAPE = {  # (radius, n_subiterations); include/stst_workloads.h stst_workload_info
    "conway": (1, 1), "jacobi5": (1, 1), "jacobi9": (1, 1), "jacobi_r2": (2, 1), "jacobi_r3": (3, 1),
    "hotspot": (1, 1), "fdtd": (1, 2), "convection_pt": (1, 3), "convection_thermal": (1, 2),
    "kat": (1, 2), "kat_r2": (2, 2),
}


def crop_bounds(window, grid_shape, depth):
    """The window grown by `depth` cells on every side, clipped to the grid."""
    (wr0, wr1), (wc0, wc1) = window
    rows, cols = grid_shape
    return (max(wr0 - depth, 0), min(wr1 + depth, rows)), (max(wc0 - depth, 0), min(wc1 + depth, cols))


def expected_window(checker, workload, params, halo, cells_of, grid_shape, window, iteration_offset,
                    n_iterations):
    """Oracle cells of `window` = ((wr0, wr1), (wc0, wc1)) after `n_iterations`.
    `cells_of(r0, r1, c0, c1)` returns the INPUT cells of that region of the grid."""
    radius, n_sub = WORKLOAD_SHAPE[workload]
    depth = n_iterations * n_sub * radius
    (r0, r1), (c0, c1) = crop_bounds(window, grid_shape, depth)
    crop = np.ascontiguousarray(cells_of(r0, r1, c0, c1))
    out = checker.run_window2d(workload, params, halo, crop, r0, c0, grid_shape[0], grid_shape[1],
                               iteration_offset, n_iterations)
    (wr0, wr1), (wc0, wc1) = window
    return out[wr0 - r0:wr1 - r0, wc0 - c0:wc1 - c0]
