"""TEST INFRASTRUCTURE — loader for the CPU parity oracles. Never imported by the product package.

Two interchangeable implementations of `oracle_run` (same C signature), each in two floating-point
flavours — `-ffp-contract=off` (every operation rounded; bit-identical to the `-fmad=false` GPU build)
and, `fma=True`, `-ffp-contract=fast -mfma` (a*b+c contracted like nvcc's default `-fmad=true` and like
the reference's own icpx build, whose default FP model contracts):

  kind "reference": oracle/_ref/liboracle_ref.so — the reference's own cpu backend and example
                    functors compiled in place from /root/reference (built only where that tree
                    exists; the prebuilt library travels with the repository snapshot).
  kind "port":      oracle/liboracle_port.so — the plain-C restatement in oracle/stencil_oracle.c,
                    buildable anywhere with gcc.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm use this module.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
PORT_LIB = HERE / "liboracle_port.so"
REF_LIB = HERE / "_ref" / "liboracle_ref.so"
PORT_FMA_LIB = HERE / "liboracle_port_fma.so"
REF_FMA_LIB = HERE / "_ref" / "liboracle_ref_fma.so"


class Oracle:
    def __init__(self, path: Path):
        self.path = path
        self.lib = C.CDLL(str(path))
        self.lib.oracle_run.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t]
        self.lib.oracle_run.restype = C.c_int
        self.lib.oracle_last_error.restype = C.c_char_p
        self.lib.oracle_kind.restype = C.c_char_p
        self.kind = self.lib.oracle_kind().decode()
        self.contracts = path.name.endswith("_fma.so")   # built with -ffp-contract=fast -mfma
        if hasattr(self.lib, "oracle_run_window"):
            self.lib.oracle_run_window.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                   C.c_void_p] + [C.c_size_t] * 6
            self.lib.oracle_run_window.restype = C.c_int
        self.lib.oracle_run_window2d.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_void_p] + [C.c_size_t] * 8
        self.lib.oracle_run_window2d.restype = C.c_int

    def run(self, workload: str, params, halo, cells: np.ndarray, iteration_offset: int,
            n_iterations: int) -> np.ndarray:
        """Advance `cells` (2-D array of the workload's cell dtype) by `n_iterations` iterations.
        `params`: ctypes struct (stst_<workload>_params); `halo`: one cell or None (all-zero)."""
        cells = np.ascontiguousarray(cells)
        out = np.empty_like(cells)
        halo_arr = None
        if halo is not None:
            halo_arr = np.zeros((), dtype=cells.dtype)
            halo_arr[()] = halo
        status = self.lib.oracle_run(
            workload.encode(), C.addressof(params) if params is not None else None,
            halo_arr.ctypes.data if halo_arr is not None else None,
            cells.ctypes.data, out.ctypes.data, cells.shape[0], cells.shape[1],
            int(iteration_offset), int(n_iterations))
        if status != 0:
            raise RuntimeError(f"oracle_run({workload}) failed: "
                               f"{self.lib.oracle_last_error().decode()}")
        return out


def _run_window(self, workload: str, params, halo, cells: np.ndarray, row0: int, global_rows: int,
                iteration_offset: int, n_iterations: int) -> np.ndarray:
    """`Oracle.run` on rows [row0, row0 + len(cells)) of a grid with `global_rows` rows (C port
    only): global coordinates, `halo` outside the window. See oracle_run_window in stencil_oracle.c."""
    cells = np.ascontiguousarray(cells)
    out = np.empty_like(cells)
    halo_arr = None
    if halo is not None:
        halo_arr = np.zeros((), dtype=cells.dtype)
        halo_arr[()] = halo
    status = self.lib.oracle_run_window(
        workload.encode(), C.addressof(params) if params is not None else None,
        halo_arr.ctypes.data if halo_arr is not None else None, cells.ctypes.data, out.ctypes.data,
        cells.shape[0], cells.shape[1], int(row0), int(global_rows), int(iteration_offset),
        int(n_iterations))
    if status != 0:
        raise RuntimeError(f"oracle_run_window({workload}) failed: "
                           f"{self.lib.oracle_last_error().decode()}")
    return out


Oracle.run_window = _run_window


def _run_window2d(self, workload: str, params, halo, cells: np.ndarray, row0: int, col0: int,
                  global_rows: int, global_cols: int, iteration_offset: int,
                  n_iterations: int) -> np.ndarray:
    """`Oracle.run` on the crop rows [row0, row0 + cells.shape[0]) x columns [col0, col0 +
    cells.shape[1]) of a `global_rows` x `global_cols` grid (both oracles): transition functions see
    global coordinates, `halo` beyond the crop. Exact further than n * n_sub * radius cells from every
    crop edge that is not a grid border."""
    cells = np.ascontiguousarray(cells)
    out = np.empty_like(cells)
    halo_arr = None
    if halo is not None:
        halo_arr = np.zeros((), dtype=cells.dtype)
        halo_arr[()] = halo
    status = self.lib.oracle_run_window2d(
        workload.encode(), C.addressof(params) if params is not None else None,
        halo_arr.ctypes.data if halo_arr is not None else None, cells.ctypes.data, out.ctypes.data,
        cells.shape[0], cells.shape[1], int(row0), int(col0), int(global_rows), int(global_cols),
        int(iteration_offset), int(n_iterations))
    if status != 0:
        raise RuntimeError(f"oracle_run_window2d({workload}) failed: "
                           f"{self.lib.oracle_last_error().decode()}")
    return out


Oracle.run_window2d = _run_window2d


def host_threads() -> int:
    """Host cores this process may run on."""
    import os
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def set_threads(n: int | None = None) -> int:
    """Make the oracles' OpenMP loops use `n` threads (default: every core this process may run on)
    whatever OMP_NUM_THREADS says — torchrun exports OMP_NUM_THREADS=1 to its workers, which once
    made a "32-core" CPU baseline run on one thread. Returns omp_get_max_threads() afterwards."""
    n = int(n or host_threads())
    gomp = C.CDLL("libgomp.so.1", mode=C.RTLD_GLOBAL)
    gomp.omp_set_num_threads(n)
    gomp.omp_get_max_threads.restype = C.c_int
    return int(gomp.omp_get_max_threads())


def cpu_has_fma() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " fma " in line + " "
    except OSError:
        pass
    return False


def build(verbose: bool = False) -> None:
    """Build whatever oracle can be built on this machine."""
    from . import recipes
    for fma in (False, True):
        recipes.build_oracle_port(verbose=verbose, fma=fma)
        recipes.build_oracle_ref(verbose=verbose, fma=fma)


def port(fma: bool = False) -> Oracle:
    lib = PORT_FMA_LIB if fma else PORT_LIB
    if fma and not cpu_has_fma():
        raise RuntimeError("this CPU has no FMA instructions: the contracting oracle cannot run here")
    if not lib.exists():
        build()
    return Oracle(lib)


def reference(fma: bool = False) -> Oracle | None:
    """The reference-built oracle, or None where it neither exists nor can be built (nor run)."""
    lib = REF_FMA_LIB if fma else REF_LIB
    if fma and not cpu_has_fma():
        return None
    if not lib.exists():
        try:
            build()
        except Exception:
            return None
    return Oracle(lib) if lib.exists() else None


def best(fma: bool = False) -> Oracle:
    """The reference-built oracle if available, else the C restatement."""
    return reference(fma) or port(fma)
