"""One `StencilUpdate` call spread over several devices of the box, from a single process.

`Params::cuda_devices` (C ABI: stst_update_params.cuda_devices; environment: STST_DEVICES) makes
`StencilUpdate::operator()` cut the grid into row slabs, run the generation loop on one
`SlabUpdate<F>` per entry — halo rows pushed between neighbours by the sweep kernel, ordering by
device flags — and gather the result into an ordinary grid (cuda/StencilUpdate.hpp, run_sharded).
The reference's updater is single-device (reference cuda/StencilUpdate.hpp:83, :124-127); what must
hold is that the result equals the single-device result, i.e. the oracle's, bit for bit in the
-fmad=false build. A device may be listed more than once, so the whole mechanism (peer copies,
pushes, flags, the two-launch pass) is exercised on a one-GPU box as well; with more GPUs present the
slabs are spread over them.
"""
import ctypes as C

import numpy as np
import pytest

import cases
from stencilstream_b200 import Grid, Params, StencilUpdate, _native

pytestmark = pytest.mark.gpu


def device_list(n_slabs):
    count = C.c_int(0)
    _native.runtime_lib().stst_device_count(C.byref(count))
    return [i % max(count.value, 1) for i in range(n_slabs)]


@pytest.mark.parametrize("n_slabs", [2, 3])
@pytest.mark.parametrize("workload,shape,offset,n,fused", [
    ("hotspot", (330, 700), 0, 23, 3),      # plane pass-through on slabs, n not a multiple of k
    ("kat", (97, 260), 5, 7, 2),            # self-checking functor: coordinates, iteration, tdv, halo
    ("fdtd", (150, 333), 0, 14, 2),         # hz_sum starts changing at iteration 10: collective repeat
    ("jacobi_r3", (200, 520), 0, 5, 2),     # ghost depth 6
    ("conway", (333, 640), 0, 9, 4),
    ("convection_pt", (120, 96), 0, 3, 1),
    ("jacobi5", (1000, 1030), 0, 40, 0),    # the planner's own depth
])
def test_sharded_call_equals_oracle(n_slabs, workload, shape, offset, n, fused, oracle_best):
    params, halo, cells = cases.make_case(workload, *shape, seed=31)
    if "kat" in workload:
        cells = cases.kat_input(*shape, offset)
    update = StencilUpdate(workload, Params(transition_function=params, halo_value=halo,
                                            iteration_offset=offset, n_iterations=n, blocking=True,
                                            fused_iterations=fused, cuda_devices=device_list(n_slabs)),
                           strict=True)
    grid = Grid(workload, buffer=cells, strict=True)
    out = update(grid)
    stats = update.get_stats()
    assert stats.n_slabs == n_slabs
    want = oracle_best.run(workload, params, halo, cells, offset, n)
    assert out.to_numpy().tobytes() == want.tobytes()
    assert grid.to_numpy().tobytes() == cells.tobytes()        # the source grid is never modified
    if workload == "fdtd":
        assert stats.speculation_redos >= 1
    # a second call reuses the slabs and resumes (iteration_offset is a live parameter)
    update.get_params().iteration_offset = offset + n
    update.get_params().n_iterations = 3
    again = update(out)
    want2 = oracle_best.run(workload, params, halo, want, offset + n, 3)
    assert again.to_numpy().tobytes() == want2.tobytes()
    # two launches per pass and slab (boundary strips + interior) where a slab is split
    assert update.get_stats().n_launches <= 2 * n_slabs * (n + 3 + 2)


def test_environment_variable_selects_the_devices(monkeypatch, oracle_best):
    params, halo, cells = cases.make_case("hotspot", 257, 300, seed=3)
    monkeypatch.setenv("STST_DEVICES", "0,0")
    update = StencilUpdate("hotspot", Params(transition_function=params, halo_value=halo,
                                             n_iterations=10, blocking=True), strict=True)
    out = update(Grid("hotspot", buffer=cells, strict=True))
    assert update.get_stats().n_slabs == 2
    assert out.to_numpy().tobytes() == oracle_best.run("hotspot", params, halo, cells, 0, 10).tobytes()
    monkeypatch.setenv("STST_DEVICES", "0-0")
    update = StencilUpdate("hotspot", Params(transition_function=params, halo_value=halo,
                                             n_iterations=10, blocking=True), strict=True)
    update(Grid("hotspot", buffer=cells, strict=True))
    assert update.get_stats().n_slabs == 1


def test_too_small_grids_fall_back_to_fewer_slabs(oracle_best):
    """A slab must own at least k * n_sub * radius rows: a 10-row grid cannot feed 8 slabs."""
    params, halo, cells = cases.make_case("jacobi5", 10, 64, seed=1)
    update = StencilUpdate("jacobi5", Params(transition_function=params, halo_value=halo,
                                             n_iterations=6, blocking=True, fused_iterations=3,
                                             cuda_devices=device_list(8)), strict=True)
    out = update(Grid("jacobi5", buffer=cells, strict=True))
    assert 1 <= update.get_stats().n_slabs <= 3
    assert out.to_numpy().tobytes() == oracle_best.run("jacobi5", params, halo, cells, 0, 6).tobytes()


def test_unknown_device_is_an_error():
    params, halo, cells = cases.make_case("jacobi5", 64, 64)
    update = StencilUpdate("jacobi5", Params(transition_function=params, halo_value=halo,
                                             n_iterations=2, cuda_devices=[0, 99]))
    with pytest.raises(ValueError):
        update(Grid("jacobi5", buffer=cells))


def test_cuda_device_must_match_the_grid():
    params, halo, cells = cases.make_case("jacobi5", 64, 64)
    update = StencilUpdate("jacobi5", Params(transition_function=params, halo_value=halo,
                                             n_iterations=2, cuda_device=7))
    with pytest.raises(ValueError):
        update(Grid("jacobi5", buffer=cells))


@pytest.mark.parametrize("workload,shape,n,fused", [("hotspot", (331, 700), 9, 3),
                                                    ("kat", (203, 260), 6, 2),
                                                    ("jacobi_r3", (257, 520), 5, 2)])
def test_sharded_call_with_cp_async_staging(monkeypatch, workload, shape, n, fused, oracle_best):
    """STST_TMA=0: tiles are staged with cp.async / plain loads, which — unlike a TMA box, whose
    tensor map ends with the slab — must not read rows beyond the slab's planes when the last tile row
    of a launch overshoots the row range (the tile height does not divide it)."""
    monkeypatch.setenv("STST_TMA", "0")
    params, halo, cells = cases.make_case(workload, *shape, seed=17)
    if "kat" in workload:
        cells = cases.kat_input(*shape, 0)
    update = StencilUpdate(workload, Params(transition_function=params, halo_value=halo,
                                            n_iterations=n, blocking=True, fused_iterations=fused,
                                            cuda_devices=device_list(3)), strict=True)
    out = update(Grid(workload, buffer=cells, strict=True))
    stats = update.get_stats()
    assert stats.n_slabs == 3 and stats.use_tma == 0
    assert out.to_numpy().tobytes() == oracle_best.run(workload, params, halo, cells, 0, n).tobytes()
