"""Synthetic inputs and transition-function parameters of the reference's example applications.

The reference generates its benchmark inputs inside the example `main()`s and Julia drivers; this
module restates exactly those recipes (citations per function) so that tests, `bench.py` and user
scripts can set up the same experiments without the reference tree. Floating-point derivations
mimic the C++ expression types step by step (float vs double promotions), because the derived
constants feed bit-exact parity checks; `tests/test_workload_setup.py` pins them against the
reference's own code where that is available.
"""
from __future__ import annotations

import math

import numpy as np

from ._native import (CELL_DTYPES, ConvectionPTParams, ConvectionThermalParams, FdtdParams,
                      HotspotParams, Jacobi5Params, Jacobi9Params, JacobiStarParams)

f32 = np.float32
f64 = np.float64


# ---- Conway -------------------------------------------------------------------------------------------

GOSPER_GUN = [
    "........................X...........",
    "......................X.X...........",
    "............XX......XX............XX",
    "...........X...X....XX............XX",
    "XX........X.....X...XX..............",
    "XX........X...X.XX....X.X...........",
    "..........X.....X.......X...........",
    "...........X...X....................",
    "............XX......................",
]


def conway_soup(rows: int, cols: int, seed: int = 42, density: float = 0.3) -> np.ndarray:
    """Bernoulli(`density`) soup with a Gosper glider gun stamped into every 64x64 block
    (Conway correctness input)."""
    rng = np.random.default_rng(seed)
    grid = rng.random((rows, cols)) < density
    gun = np.array([[ch == "X" for ch in line] for line in GOSPER_GUN], dtype=bool)
    for r0 in range(0, rows - gun.shape[0] - 1, 64):
        for c0 in range(0, cols - gun.shape[1] - 1, 64):
            grid[r0 + 1:r0 + 1 + gun.shape[0], c0 + 1:c0 + 1 + gun.shape[1]] = gun
    return grid.astype(np.bool_)


# ---- Jacobi -------------------------------------------------------------------------------------------

def jacobi_input(rows: int, cols: int) -> np.ndarray:
    """Centred unit square on a zero background (reference examples/jacobi/jacobi.cpp:111-124)."""
    r = np.arange(rows, dtype=np.float64)[:, None]
    c = np.arange(cols, dtype=np.float64)[None, :]
    inside = (r >= rows * 0.25) & (r < rows * 0.75) & (c >= cols * 0.25) & (c < cols * 0.75)
    return inside.astype(np.float32)


def jacobi5_params(coef=(0.2, 0.2, 0.2, 0.2, 0.2)) -> Jacobi5Params:
    """Five coefficients; the reference benchmark uses 0.2 each
    (examples/jacobi/scripts/benchmark.jl:44)."""
    p = Jacobi5Params()
    for i, v in enumerate(coef):
        p.coef[i] = float(f32(v))
    return p


def jacobi9_params(coef=None) -> Jacobi9Params:
    p = Jacobi9Params()
    coef = np.full((3, 3), 1.0 / 9.0) if coef is None else np.asarray(coef)
    for r in range(3):
        for c in range(3):
            p.coef[r][c] = float(f32(coef[r, c]))
    return p


def jacobi_star_params(radius: int) -> JacobiStarParams:
    """Equal weights 1/(4R+1) for the radius-R star stencil."""
    p = JacobiStarParams()
    w = float(f32(1.0 / (4 * radius + 1)))
    p.centre = w
    for i in range(3):
        p.arm[i] = w if i < radius else 0.0
    return p


# ---- HotSpot ------------------------------------------------------------------------------------------

def hotspot_input(rows: int, cols: int) -> np.ndarray:
    """temp = 30 everywhere; power = 0.5 in the centre half, 1-based inclusive bounds
    (reference examples/hotspot/data/input_gen.jl:3-15)."""
    cells = np.zeros((rows, cols), dtype=CELL_DTYPES["hotspot"])
    cells["temp"] = f32(30.0)
    r_lo, r_hi = rows // 4, (3 * rows) // 4          # 1-based, inclusive
    c_lo, c_hi = cols // 4, (3 * cols) // 4
    cells["power"][max(r_lo - 1, 0):r_hi, max(c_lo - 1, 0):c_hi] = f32(0.5)
    return cells


def hotspot_params(rows: int, cols: int) -> HotspotParams:
    """Physics constants exactly as reference examples/hotspot/hotspot.cpp:281-295 derives them."""
    MAX_PD, PRECISION, SPEC_HEAT_SI, K_SI, FACTOR_CHIP = f64(3.0e6), f64(0.001), f64(1.75e6), 100, f64(0.5)
    t_chip, chip_height, chip_width = f32(0.0005), f32(0.016), f32(0.016)

    grid_height = f32(chip_height / f32(rows))
    grid_width = f32(chip_width / f32(cols))

    Cap = f32(FACTOR_CHIP * SPEC_HEAT_SI * f64(t_chip) * f64(grid_height) * f64(grid_width))
    Rx = f32(f64(grid_width) / (f64(2.0) * f64(K_SI) * f64(t_chip) * f64(grid_height)))
    Ry = f32(f64(grid_height) / (f64(2.0) * f64(K_SI) * f64(t_chip) * f64(grid_width)))
    Rz = f32(t_chip / f32(f32(f32(K_SI) * grid_height) * grid_width))

    max_slope = f32(MAX_PD / (FACTOR_CHIP * f64(t_chip) * SPEC_HEAT_SI))
    step = f32(PRECISION / f64(max_slope) / f64(1000.0))

    p = HotspotParams()
    p.Rx_1 = float(f32(f32(1.0) / Rx))
    p.Ry_1 = float(f32(f32(1.0) / Ry))
    p.Rz_1 = float(f32(f32(1.0) / Rz))
    p.Cap_1 = float(f32(step / Cap))
    return p


# ---- FDTD ---------------------------------------------------------------------------------------------

FDTD_MAX_GRID = {  # reference examples/fdtd/experiments/max_grid.json
    "tau": 100e-15, "dx": 3.474e-10,
    "time": {"t_cutoff": 7.0, "t_detect": 14.0, "t_max": 1.5, "t_snap": 0.1},
    "source": {"frequency": 120e12, "phase": 3.0, "x": 0, "y": 0, "radius": 0},
    "cavity_rings": [{"radius": 800e-9, "mu_r": 11.56, "eps_r": 1.0, "sigma": 0.0}],
}

FDTD_DEFAULT = {  # reference examples/fdtd/experiments/default.json
    "tau": 100e-15, "dx": 10e-9,
    "time": {"t_cutoff": 7.0, "t_detect": 14.0, "t_max": 15.0, "t_snap": 0.1},
    "source": {"frequency": 120e12, "phase": 3.0, "x": 0, "y": 0, "radius": 0},
    "cavity_rings": [{"radius": 800e-9, "mu_r": 11.56, "eps_r": 1.0, "sigma": 0.0}],
}


class FdtdExperiment:
    """Derived quantities of an FDTD experiment (reference examples/fdtd/src/Parameters.hpp:224-262,
    examples/fdtd/src/Kernel.hpp:60-78, examples/fdtd/src/material/Material.hpp:24-78)."""

    c0 = f32(299792458.0)
    sqrt_2 = f32(1.4142135623730951)
    pi = f32(3.1415926535897932384626433)

    def __init__(self, config: dict):
        self.config = config
        self.tau = f32(config["tau"])
        self.dx = f32(config["dx"])
        t = config["time"]
        self.t_cutoff_factor, self.t_detect_factor = f32(t["t_cutoff"]), f32(t["t_detect"])
        self.t_max_factor = f32(t["t_max"])
        self.t_snap_factor = f32(t["t_snap"]) if "t_snap" in t else None
        s = config["source"]
        self.frequency, self.t_0_factor = f32(s["frequency"]), f32(s["phase"])
        self.source_x, self.source_y, self.source_radius = f32(s["x"]), f32(s["y"]), f32(s["radius"])
        self.rings = [(f32(r["radius"]), f32(r["mu_r"]), f32(r["eps_r"]), f32(r["sigma"]))
                      for r in config["cavity_rings"]]

    # Parameters.hpp:224-262
    def dt(self):
        return f32(f64(f32(self.dx / f32(self.c0 * self.sqrt_2))) * f64(0.99))

    def t_max(self):
        return f32(self.t_max_factor * self.tau)

    def n_timesteps(self) -> int:
        return int(np.ceil(f32(self.t_max() / self.dt())))

    def n_snap_timesteps(self):
        if self.t_snap_factor is None:
            return None
        return int(np.ceil(f32(f32(self.t_snap_factor * self.tau) / self.dt())))

    def grid_wh(self) -> int:
        outer = f32(0.0)
        for radius, *_ in self.rings:
            outer = f32(outer + radius)
        return int(np.ceil(f32(f32(f32(f32(2) * outer) / self.dx) + f32(2))))

    def source_rc(self):
        half = f32(self.grid_wh() // 2)
        return int(f32(half + f32(self.source_y / self.dx))), int(f32(half + f32(self.source_x / self.dx)))

    # Kernel.hpp:60-78
    def kernel_params(self) -> FdtdParams:
        p = FdtdParams()
        dt = self.dt()
        p.dt = float(dt)
        p.t_0 = float(f32(self.t_0_factor * self.tau))
        p.tau = float(self.tau)
        p.omega = float(f32(f64(2.0) * f64(self.pi) * f64(self.frequency)))
        p.cutoff_iteration = int(np.floor(f32(f32(self.t_cutoff_factor * self.tau) / dt)))
        p.detect_iteration = int(np.floor(f32(f32(self.t_detect_factor * self.tau) / dt)))
        srs = f32(self.source_radius / self.dx)
        p.source_radius_squared = float(f32(srs * srs))
        src_r, src_c = self.source_rc()
        sr, sc = f32(src_r), f32(src_c)
        p.source_r, p.source_c = float(sr), float(sc)
        bound = f32(self.source_radius / self.dx)
        bound = f32(bound * bound)
        bound = f32(bound - f32(f32(sc * sc) + f32(sr * sr)))
        p.source_distance_bound = float(bound)
        p.double_center_rc = float(f32(self.grid_wh()))
        return p

    # Material.hpp:24-78
    def _coefficients(self, mu_r, eps_r, sigma):
        dx, dt = self.dx, self.dt()
        mu_0 = f32(f64(4.0) * f64(self.pi) * f64(1.0e-7))
        eps_0 = f32(f64(1.0) / f64(f32(f32(self.c0 * self.c0) * mu_0)))
        one, two = f32(1), f32(2)
        sdt = f32(sigma * dt)
        ca = f32(f32(one - sdt) / f32(one + sdt))
        da = ca
        if np.isinf(eps_r):
            cb = f32(0.0)
        else:
            cb = f32(f32(dt / f32(f32(eps_0 * eps_r) * dx)) /
                     f32(one + f32(sdt / f32(f32(two * eps_0) * eps_r))))
        if np.isinf(mu_r):
            db = f32(0.0)
        else:
            db = f32(f32(dt / f32(f32(mu_0 * mu_r) * dx)) /
                     f32(one + f32(sdt / f32(f32(two * mu_0) * mu_r))))
        return ca, cb, da, db

    def initial_grid(self) -> np.ndarray:
        """Material map with zero fields (reference examples/fdtd/src/fdtd.cpp:193-216)."""
        wh = self.grid_wh()
        cells = np.zeros((wh, wh), dtype=CELL_DTYPES["fdtd"])
        idx = np.arange(wh, dtype=np.float32)
        half = f64(f32(wh)) / f64(2.0)
        a = (idx.astype(np.float64) - half).astype(np.float32)[:, None]
        b = (idx.astype(np.float64) - half).astype(np.float32)[None, :]
        with np.errstate(over="ignore"):
            distance = (self.dx * np.sqrt((a * a + b * b).astype(np.float32))).astype(np.float32)
        assigned = np.zeros((wh, wh), dtype=bool)
        radius = f32(0.0)
        for ring_radius, mu_r, eps_r, sigma in self.rings:
            radius = f32(radius + ring_radius)
            mask = (distance < radius) & ~assigned
            ca, cb, da, db = self._coefficients(mu_r, eps_r, sigma)
            for name, value in (("ca", ca), ("cb", cb), ("da", da), ("db", db)):
                cells[name][mask] = value
            assigned |= mask
        # cells outside every ring are MaterialCell::halo(): all zeros (CoefResolver.hpp:31-35)
        return cells


# ---- Mantle convection --------------------------------------------------------------------------------

def convection_benchmark_config(res: int, n_iters: int, lx: float = 1.0, ly: float = 1.0) -> dict:
    """The experiment dictionary of the reference's benchmark driver
    (examples/convection/scripts/benchmark.jl:159-178)."""
    return {"ly": ly, "lx": lx, "py": 1.0, "px": 1.0, "res": res, "eta0": 1.0, "DcT": 1.0,
            "deltaT": 1.0, "Ra": 1e7, "Pra": 1e3, "iterMax": n_iters, "nt": 1, "nout": 1,
            "nerr": n_iters, "epsilon": 1e-4, "dmp": 2}


class ConvectionExperiment:
    """Derived numerics of a convection experiment (reference examples/convection/convection.cpp:305-358)."""

    def __init__(self, config: dict):
        c = self.config = config
        self.lx, self.ly, self.px, self.py = float(c["lx"]), float(c["ly"]), float(c["px"]), float(c["py"])
        self.eta0, self.DcT, self.deltaT = float(c["eta0"]), float(c["DcT"]), float(c["deltaT"])
        self.Ra, self.Pra = float(c["Ra"]), float(c["Pra"])
        self.w = 1e-2 * self.ly
        self.roh0_g_alpha = self.Ra * self.eta0 * self.DcT / self.deltaT / math.pow(self.ly, 3)
        self.delta_eta_delta_T = 1e-10 / self.deltaT
        res = int(c["res"])
        self.nx = int(res * self.lx - 1)
        self.ny = int(res * self.ly - 1)
        self.nerr, self.dmp = int(c["nerr"]), float(c["dmp"])
        self.dx = self.lx / (self.nx - 1)
        self.dy = self.ly / (self.ny - 1)
        self.rho = 1.0 / self.Pra * self.eta0 / self.DcT
        self.dt_diff = 1.0 / 4.1 * math.pow(min(self.dx, self.dy), 2) / self.DcT
        self.delta_tau_iter = 1.0 / 6.1 * min(self.dx, self.dy) / math.sqrt(self.eta0 / self.rho)
        self.beta = 6.1 * math.pow(self.delta_tau_iter, 2) / math.pow(min(self.dx, self.dy), 2) / self.rho
        self.dampX = 1.0 - self.dmp / self.nx
        self.dampY = 1.0 - self.dmp / self.ny

    @property
    def grid_shape(self):
        return (self.nx + 1, self.ny + 1)

    def pseudo_transient_params(self) -> ConvectionPTParams:
        p = ConvectionPTParams()
        p.nx, p.ny = self.nx, self.ny
        p.roh0_g_alpha = self.roh0_g_alpha
        p.delta_eta_delta_T = self.delta_eta_delta_T
        p.eta0, p.deltaT = self.eta0, self.deltaT
        p.dx, p.dy = self.dx, self.dy
        p.delta_tau_iter, p.beta, p.rho = self.delta_tau_iter, self.beta, self.rho
        p.dampX, p.dampY, p.DcT = self.dampX, self.dampY, self.DcT
        return p

    def thermal_params(self, dt: float) -> ConvectionThermalParams:
        p = ConvectionThermalParams()
        p.nx, p.ny, p.dx, p.dy, p.dt, p.DcT = self.nx, self.ny, self.dx, self.dy, dt, self.DcT
        return p

    def initial_temperature(self, row_lo: int = 0, row_hi: int | None = None) -> np.ndarray:
        """The `T` field of the initial state (reference convection.cpp:380-397) for rows
        [row_lo, row_hi); every other field starts at zero."""
        nx, ny = self.nx, self.ny
        row_hi = nx + 1 if row_hi is None else row_hi
        x = np.arange(row_lo, row_hi, dtype=np.float64)[:, None]
        y = np.arange(ny + 1, dtype=np.float64)[None, :]
        # std::exp(-std::pow((x*dx - px)/w, 2) - std::pow((y*dy - py)/w, 2))
        blob = self.deltaT * np.exp(-np.power((x * self.dx - self.px) / self.w, 2)
                                    - np.power((y * self.dy - self.py) / self.w, 2))
        inside = (x < nx) & (y < ny)
        T = np.where(inside, blob, 0.0)
        T = np.where(y == ny - 1, -self.deltaT / 2.0, T)
        T = np.where(y == 0, self.deltaT / 2.0, T)
        return np.broadcast_to(T, (row_hi - row_lo, ny + 1))

    def initial_grid(self, row_lo: int = 0, row_hi: int | None = None) -> np.ndarray:
        """Initial temperature blob (reference convection.cpp:380-397); optionally only rows
        [row_lo, row_hi) of it (for slab-wise generation of very large grids)."""
        T = self.initial_temperature(row_lo, row_hi)
        cells = np.zeros(T.shape, dtype=CELL_DTYPES["convection_pt"])
        cells["T"] = T
        return cells
