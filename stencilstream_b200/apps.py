"""Application loops around the generation loop — the callers on either side of the hot path.

Host-side mirrors of the two reference applications that do more than one `StencilUpdate` call:

* `run_convection`  — reference examples/convection/convection.cpp:402-477: batches of `nerr`
  pseudo-transient iterations until the velocity/pressure residuals fall below `epsilon`, then one
  thermal-solver iteration whose time step depends on the velocity maxima; repeated `nt` times.
* `run_fdtd`        — reference examples/fdtd/src/fdtd.cpp:218-252: the simulation in snapshot
  intervals with a live `iteration_offset`, one `hz` frame per interval, `hz_sum` at the end.

What differs from the reference is where the glue runs. There the five max-norms of a convergence
check are a host loop over a `GridAccessor`, which first migrates the whole 88-byte-per-cell grid
over PCIe (convection.cpp:412-438), and frames are written from a full-grid accessor
(fdtd.cpp:114-166, convection.cpp:460-477). Here norms are one device pass over the five planes
involved (`Grid.max_abs`, kernel `reduce_max_abs_kernel`) and a frame is a single-plane download
(`Grid.field_to_numpy`); the cells never leave HBM between updates. Every number is computed by
libstst_workloads.so; this module only sequences calls, as the reference's `main` functions do.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable

import numpy as np

from .api import Grid, Params, StencilUpdate
from .sharding import ShardedStencilUpdate
from . import workloads as W


@dataclass
class ConvectionStep:
    """What the reference prints per time step (convection.cpp:447-448) plus the chosen dt."""
    it: int
    iterations: int
    errV: float
    errP: float
    dt: float
    norms: dict = field(default_factory=dict)


def convection_norm_extents(nx: int, ny: int):
    """(field, rows, cols) of the five max-norms, index ranges as in convection.cpp:417-434 (the
    grid is (nx+1) x (ny+1) cells, x the row index)."""
    return [("ErrV", nx, ny + 1), ("ErrP", nx, ny), ("Vx", nx + 1, ny), ("Vy", nx, ny),
            ("Pt", nx, ny)]


def run_convection(config: dict, *, cells: np.ndarray | None = None, strict: bool | None = None,
                   device: int = -1, fused_iterations: int = 0,
                   on_frame: Callable[[int, np.ndarray], None] | None = None):
    """Run a whole convection experiment (`config`: the reference's experiment JSON as a dict).

    Returns (final Grid, [ConvectionStep, ...]). `on_frame(it, T)` receives the temperature field of
    the inner nx x ny cells every `nout` steps (what the reference writes to `<it>.csv`).
    """
    exp = W.ConvectionExperiment(config)
    nx, ny = exp.nx, exp.ny
    iter_max, nt = int(config["iterMax"]), int(config["nt"])
    nout, nerr, epsilon = int(config["nout"]), int(config["nerr"]), float(config["epsilon"])
    if cells is None:
        cells = exp.initial_grid()
    grid = Grid("convection_pt", buffer=cells, device=device, strict=strict)
    pseudo_transient = StencilUpdate(
        "convection_pt", Params(transition_function=exp.pseudo_transient_params(), halo_value=None,
                                n_iterations=nerr, blocking=True,
                                fused_iterations=fused_iterations), strict=strict)
    extents = convection_norm_extents(nx, ny)
    steps = []
    for it in range(1, nt + 1):
        errV = errP = 2 * epsilon
        norms = dict.fromkeys((e[0] for e in extents), float("-inf"))
        iterations = 0
        while iterations < iter_max and (errV > epsilon or errP > epsilon):
            grid = pseudo_transient(grid)
            norms = dict(zip((e[0] for e in extents), grid.max_abs(extents)))
            errV = norms["ErrV"] / (1e-12 + norms["Vy"])
            errP = norms["ErrP"] / (1e-12 + norms["Pt"])
            iterations += nerr
        with np.errstate(divide="ignore"):
            dt_adv = min(np.float64(exp.dx) / norms["Vx"], np.float64(exp.dy) / norms["Vy"]) / 2.1
        dt = float(min(exp.dt_diff, dt_adv))
        thermal = StencilUpdate(
            "convection_thermal", Params(transition_function=exp.thermal_params(dt), halo_value=None,
                                         n_iterations=1, blocking=True), strict=strict)
        grid = thermal(grid)
        steps.append(ConvectionStep(it, iterations, errV, errP, dt, norms))
        if on_frame is not None and it % nout == 0:
            on_frame(it, grid.field_to_numpy("T")[:nx, :ny])
    return grid, steps


def run_convection_sharded(config: dict, *, rank: int, world: int, comm=None, device: int = 0,
                           strict: bool | None = None, fused_iterations: int = 0,
                           slab_factory=None,
                           on_frame: Callable[[int, int, int, np.ndarray], None] | None = None):
    """`run_convection` on a row-sharded grid, one process per GPU (collective: every process of the
    group calls it with the same `config`).

    The pseudo-transient and the thermal update are two transition functions over the same cells, i.e.
    two slab objects per GPU; the cells move between them device to device (`load_from`). The five
    max-norms of a convergence check are reduced per slab on its GPU and combined with one
    all-reduce(MAX) — the only collective in the loop. `on_frame(it, row_lo, row_hi, T_rows)` receives
    this rank's rows of the temperature field every `nout` steps. Returns (the pseudo-transient
    ShardedStencilUpdate holding the final cells, [ConvectionStep, ...])."""
    exp = W.ConvectionExperiment(config)
    nx, ny = exp.nx, exp.ny
    rows, cols = exp.grid_shape
    iter_max, nt = int(config["iterMax"]), int(config["nt"])
    nout, nerr, epsilon = int(config["nout"]), int(config["nerr"]), float(config["epsilon"])
    common = dict(rank=rank, world=world, device=device, comm=comm, strict=strict,
                  slab_factory=slab_factory)
    pseudo_transient = ShardedStencilUpdate(
        "convection_pt", Params(transition_function=exp.pseudo_transient_params(), halo_value=None,
                                n_iterations=nerr, blocking=True,
                                fused_iterations=fused_iterations), rows, cols, **common)
    thermal = ShardedStencilUpdate(
        "convection_thermal", Params(transition_function=exp.thermal_params(0.0), halo_value=None,
                                     n_iterations=1, blocking=True), rows, cols, **common)
    lo, hi = pseudo_transient.row_lo, pseudo_transient.row_hi
    pseudo_transient.load(exp.initial_grid(lo, hi))
    extents = convection_norm_extents(nx, ny)
    steps = []
    for it in range(1, nt + 1):
        errV = errP = 2 * epsilon
        norms = dict.fromkeys((e[0] for e in extents), float("-inf"))
        iterations = 0
        while iterations < iter_max and (errV > epsilon or errP > epsilon):
            pseudo_transient()
            norms = dict(zip((e[0] for e in extents), pseudo_transient.max_abs(extents)))
            errV = norms["ErrV"] / (1e-12 + norms["Vy"])
            errP = norms["ErrP"] / (1e-12 + norms["Pt"])
            iterations += nerr
        with np.errstate(divide="ignore"):
            dt_adv = min(np.float64(exp.dx) / norms["Vx"], np.float64(exp.dy) / norms["Vy"]) / 2.1
        dt = float(min(exp.dt_diff, dt_adv))
        thermal.get_params().transition_function = exp.thermal_params(dt)
        thermal.load_from(pseudo_transient)
        thermal()
        pseudo_transient.load_from(thermal)
        steps.append(ConvectionStep(it, iterations, errV, errP, dt, norms))
        if on_frame is not None and it % nout == 0:
            T = pseudo_transient.field_to_numpy("T")
            on_frame(it, lo, hi, T[:max(0, min(hi, nx) - lo), :ny])
    thermal.close()
    return pseudo_transient, steps


def run_fdtd(config: dict, *, n_timesteps: int | None = None, n_snap_timesteps: int | None = None,
             cells: np.ndarray | None = None, strict: bool | None = None, device: int = -1,
             fused_iterations: int = 0,
             on_frame: Callable[[str, int, np.ndarray], None] | None = None):
    """Run an FDTD experiment (`config`: the reference's experiment JSON as a dict) in snapshot
    intervals. `on_frame(field, iteration, values)` receives the `hz` plane after every interval and
    `hz_sum` at the end (fdtd.cpp:233-250). `n_timesteps` / `n_snap_timesteps` override the
    experiment's values (tests shorten the run). Returns (final Grid, StencilUpdate)."""
    exp = W.FdtdExperiment(config)
    total = exp.n_timesteps() if n_timesteps is None else int(n_timesteps)
    snap = exp.n_snap_timesteps() if n_snap_timesteps is None else int(n_snap_timesteps)
    if cells is None:
        cells = exp.initial_grid()
    grid = Grid("fdtd", buffer=cells, device=device, strict=strict)
    simulation = StencilUpdate(
        "fdtd", Params(transition_function=exp.kernel_params(), halo_value=None, iteration_offset=0,
                       n_iterations=total, blocking=True, fused_iterations=fused_iterations),
        strict=strict)
    if snap:
        params = simulation.get_params()
        params.n_iterations = snap
        while params.iteration_offset < total:
            grid = simulation(grid)
            if on_frame is not None:
                on_frame("hz", params.iteration_offset + snap, grid.field_to_numpy("hz"))
            params.iteration_offset += snap
    else:
        grid = simulation(grid)
    if on_frame is not None:
        on_frame("hz_sum", total, grid.field_to_numpy("hz_sum"))
    return grid, simulation


def run_fdtd_sharded(config: dict, *, rank: int, world: int, device: int = 0, comm: Any = None,
                     n_timesteps: int | None = None, n_snap_timesteps: int | None = None,
                     strict: bool | None = None, fused_iterations: int = 0,
                     slab_factory: Callable[..., Any] | None = None, transport: str | None = None,
                     on_frame: Callable[[str, int, int, int, np.ndarray], None] | None = None):
    """`run_fdtd` on row slabs, one process per GPU (reference examples/fdtd/src/fdtd.cpp:218-252: the
    snapshot loop advances `iteration_offset` in steps of `n_snap_timesteps` and overshoots
    `n_timesteps` to the next multiple). `on_frame(field, iteration, row_lo, row_hi, values)` receives
    THIS rank's rows of `hz` after every interval and of `hz_sum` at the end — single-plane downloads,
    4 of the 32 bytes of a cell. Collective. Returns the ShardedStencilUpdate holding the final cells."""
    exp = W.FdtdExperiment(config)
    total = exp.n_timesteps() if n_timesteps is None else int(n_timesteps)
    snap = exp.n_snap_timesteps() if n_snap_timesteps is None else int(n_snap_timesteps)
    wh = exp.grid_wh()
    simulation = ShardedStencilUpdate(
        "fdtd", Params(transition_function=exp.kernel_params(), halo_value=None, iteration_offset=0,
                       n_iterations=snap or total, blocking=True, fused_iterations=fused_iterations),
        wh, wh, rank=rank, world=world, device=device, comm=comm, strict=strict,
        slab_factory=slab_factory, transport=transport)
    lo, hi = simulation.row_lo, simulation.row_hi
    simulation.load(exp.initial_grid()[lo:hi])
    params = simulation.get_params()
    if snap:
        while params.iteration_offset < total:
            simulation()
            if on_frame is not None:
                on_frame("hz", params.iteration_offset + snap, lo, hi, simulation.field_to_numpy("hz"))
            params.iteration_offset += snap
    else:
        simulation()
    if on_frame is not None:
        on_frame("hz_sum", total, lo, hi, simulation.field_to_numpy("hz_sum"))
    return simulation
