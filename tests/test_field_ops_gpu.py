"""GPU tests of the per-field device operations (Grid.max_abs, single-field transfers) and of the
staged transfer path for host memory that is not pinned.

The checker is numpy on the host image of the same cells: `max |v|` is exact arithmetic, so every
comparison is bit-exact. Index ranges follow the reference loop these operations replace
(examples/convection/convection.cpp:412-438).
"""
import numpy as np
import pytest

import cases
from stencilstream_b200 import Grid, Params, RangeError, StencilUpdate, _native
from stencilstream_b200.apps import convection_norm_extents

pytestmark = pytest.mark.gpu


def numpy_max_abs(cells, field, rows, cols):
    plane = cells[field] if cells.dtype.names else cells
    part = plane[:rows, :cols]
    return float(np.abs(part.astype(np.float64)).max()) if part.size else float("-inf")


@pytest.mark.parametrize("workload", ["jacobi5", "conway", "hotspot", "fdtd", "convection_pt", "kat"])
@pytest.mark.parametrize("shape", [(1, 1), (67, 93), (300, 1030)])
def test_max_abs_every_field_matches_numpy(workload, shape):
    cells = cases.random_cells(workload, shape, seed=11) if workload != "kat" \
        else cases.kat_input(*shape, 3)
    if cells.dtype.names and cells.dtype[0].kind == "f":
        cells[cells.dtype.names[0]] -= 0.5          # mixed signs
    grid = Grid(workload, buffer=cells)
    names = cells.dtype.names or (0,)
    rows, cols = shape
    extents = []
    for i, f in enumerate(names):                    # more than 8 for convection: two batches
        extents.append((f, rows, cols))
        extents.append((f, max(rows - 1 - i % 3, 0), max(cols - i % 5, 0)))
    got = grid.max_abs(extents)
    want = [numpy_max_abs(cells, f, r, c) for f, r, c in extents]
    assert got == want


def test_max_abs_empty_extent_and_clipping_and_nan():
    _, _, cells = cases.make_case("hotspot", 40, 50, seed=3)
    cells["temp"][7, 9] = np.nan                     # `abs(v) > max` never selects a NaN
    cells["temp"][8, 9] = -1e9
    grid = Grid("hotspot", buffer=cells)
    got = grid.max_abs([("temp", 0, 50), ("temp", 40, 0), ("temp", 10_000, 10_000), (1, 40, 50)])
    assert got[0] == float("-inf") and got[1] == float("-inf")
    assert got[2] == 1e9
    assert got[3] == numpy_max_abs(cells, "power", 40, 50)
    with pytest.raises(ValueError):
        grid.max_abs([(2, 1, 1)])                    # hotspot cells have two fields
    assert Grid("jacobi5", buffer=np.zeros((3, 3), np.float32)).max_abs([]) == []


def test_max_abs_follows_updates_and_host_writes(oracle_best):
    """The reduction sees what an accessor would see: results of an update that has not been
    downloaded, and host writes that have not been uploaded yet."""
    exp, cells = cases.convection_case(64, 80, seed=5)
    params = exp.pseudo_transient_params()
    grid = Grid("convection_pt", buffer=cells, strict=True)
    update = StencilUpdate("convection_pt", Params(transition_function=params, n_iterations=3,
                                                   blocking=True), strict=True)
    out = update(grid)
    want = oracle_best.run("convection_pt", params, None, cells, 0, 3)
    extents = convection_norm_extents(exp.nx, exp.ny)
    assert out.max_abs(extents) == [numpy_max_abs(want, f, r, c) for f, r, c in extents]
    view = out.accessor("read_write")
    view["Pt"][2, 3] = -123456.0
    del view
    assert out.max_abs([("Pt", exp.nx, exp.ny)]) == [123456.0]


@pytest.mark.parametrize("workload", ["hotspot", "fdtd", "convection_pt", "kat", "jacobi5"])
def test_single_field_transfers(workload):
    _, _, cells = cases.make_case(workload, 71, 130, seed=9)
    grid = Grid(workload, buffer=cells)
    names = cells.dtype.names or (0,)
    for f in names:
        plane = cells[f] if cells.dtype.names else cells
        assert grid.field_to_numpy(f).tobytes() == np.ascontiguousarray(plane).tobytes()
    last = names[-1]
    plane = np.ascontiguousarray(cells[last] if cells.dtype.names else cells)
    replaced = (plane * 2).astype(plane.dtype)
    grid.copy_field_from_buffer(last, replaced)
    back = grid.to_numpy()
    assert np.ascontiguousarray(back[last] if back.dtype.names else back).tobytes() == replaced.tobytes()
    if cells.dtype.names and len(names) > 1:         # the other fields are untouched
        assert back[names[0]].tobytes() == cells[names[0]].tobytes()
    with pytest.raises(RangeError):
        grid.copy_field_from_buffer(last, replaced[:-1])


@pytest.mark.parametrize("workload,shape", [
    ("jacobi5", (3000, 3001)),        # unsplit plane, padded pitch, 36 MB: wraps the 3-slot ring
    ("convection_pt", (900, 901)),    # 88-byte cells, 71 MB
    ("hotspot", (6100, 6000)),        # 293 MB: crosses the 256 MiB device-side chunk as well
])
def test_pageable_host_memory_round_trips_through_the_staged_pipeline(workload, shape):
    """copy_from_buffer / copy_to_buffer with ordinary numpy memory (stst_memcpy_2d_staged)."""
    cells = cases.random_cells(workload, shape, seed=4)
    pinned = _native.C.c_int(-1)
    _native.runtime_lib().stst_host_is_pinned(cells.ctypes.data_as(_native.C.c_void_p),
                                              _native.C.byref(pinned))
    assert pinned.value == 0
    grid = Grid(workload, buffer=cells)
    assert grid.to_numpy().tobytes() == cells.tobytes()
    # the accessor's (pinned, if the box allows) image is a third route to the same cells
    assert np.asarray(grid.accessor("read")).tobytes() == cells.tobytes()
