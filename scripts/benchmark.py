#!/usr/bin/env python
"""Benchmark driver with the experiments and output files of the reference's Julia drivers.

    python scripts/benchmark.py max_perf          <workload> [--out-dir DIR]
    python scripts/benchmark.py deep_grid_scaling <workload> [--out-dir DIR] [--target-runtime S]
                                                  [--min-wh 32] [--max-wh N]
    python scripts/benchmark.py ncu_metrics       <workload> [--rows R --cols C]

Counterparts (Julia is not part of this toolchain, so this is Python over the same C ABI the other
host code uses):

  max_perf           examples/*/scripts/benchmark.jl `max_perf_benchmark`, cuda branch
                     (hotspot: 16384 x 16384, 1000 iterations, best of 3 after a warm-up run,
                     examples/hotspot/scripts/benchmark.jl:103-107; jacobi: grid from `max_grid_wh`,
                     iterations for ~30 s of modelled run time, examples/jacobi/scripts/benchmark.jl:
                     100-126; fdtd: its benchmark experiment) -> `metrics.cuda.json` with the same
                     keys ("target", "measured" [cell-iterations/s], "FLOPS"), plus "model" /
                     "accuracy" from the roofline model below.
  deep_grid_scaling  `deep_grid_scaling_benchmark` (examples/hotspot/scripts/benchmark.jl:145-200):
                     square grids from `max_grid_wh` down to 32, divided by sqrt(2) per step,
                     iterations chosen so that the model predicts `--target-runtime` seconds
                     -> `scaling.cuda.csv` with the reference's columns
                     grid_wh,n_iters,runtime,measured_throughput,model_throughput; sizes already
                     in the file are skipped (resumable, as in the reference).
  ncu_metrics        the metric list of scripts/benchmark-common.jl:246-283, for use under
                     `ncu --csv --metrics ...` (prints the command; profiling is a separate run).

Model. The reference models its cuda backend as one HBM read + one write of the grid per sweep at
80 % of the A100's bandwidth, or 10 us of launch latency per sweep, whichever is larger
(scripts/benchmark-common.jl:148-158). The same model for this backend: one read + one write per
*fused pass* of k iterations at the measured B200 copy bandwidth (MEASURED_PEAKS.json), 10 us per
pass. "Walltime" is `StencilUpdate.get_walltime()` with `blocking=True`, exactly the figure the
reference's drivers scrape from the examples' output.
"""
from __future__ import annotations

import argparse
import csv
import json
import math
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

# OPERATIONS_PER_CELL / CELL_SIZE / target names of examples/*/scripts/benchmark.jl
OPERATIONS_PER_CELL = {"jacobi5": 9, "hotspot": 15, "fdtd": 8 + (6 + 4 + 2 + 2 + 2),
                       "convection_pt": (5 + 5 + 3 + 6 + 6 + 6) + (10 + 3 + 2 + 14 + 3 + 2) + 2}
TARGET_NAME = {"jacobi5": "Jacobi 5-point stencil, general coefficients, CUDA", "hotspot": "Hotspot, CUDA",
               "fdtd": "FDTD, CUDA", "convection_pt": "Convection, CUDA"}
SCHEDULING_LATENCY_PER_PASS = 0.5 / 50_000  # benchmark-common.jl:158


def hbm_bytes_per_second() -> float:
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        return float(json.loads(path.read_text())["hbm_gbs"]) * 1e9
    return 6650e9


def max_grid_wh(cell_size: int, memory_bytes: float, clip_to_base: float | None = None) -> int:
    """benchmark-common.jl:186-199: three grids must fit into device memory, and a grid must be
    indexable with signed 32-bit integers."""
    max_n_cells = min(memory_bytes / 3 / cell_size, 2 ** 31)
    wh = math.sqrt(max_n_cells)
    if clip_to_base is not None:
        wh = clip_to_base ** math.floor(math.log(wh, clip_to_base))
    return int(math.floor(wh))


def model_runtime(rows, cols, n_iters, cell_size, n_sub, k, bandwidth):
    passes = math.ceil(n_iters / max(k, 1))
    computation = passes * 2.0 * cell_size * rows * cols / bandwidth
    scheduling = passes * SCHEDULING_LATENCY_PER_PASS
    return max(computation, scheduling)


class Runner:
    """One workload on square (or given) grids through the public Grid / StencilUpdate API."""

    def __init__(self, workload: str):
        import bench
        from stencilstream_b200 import workload_info
        self.bench, self.workload = bench, workload
        self.info = workload_info(workload)

    def run(self, rows, cols, n_iters, n_samples=3, warmup=True):
        """(best walltime [s], fused iterations k) — minimum over `n_samples` runs, like the
        reference's drivers (hotspot/scripts/benchmark.jl:49-71)."""
        from stencilstream_b200 import Grid, Params, StencilUpdate
        params, halo, fill = self.bench.make_workload(self.workload, rows, cols)
        grid = Grid(self.workload, rows, cols)
        view = grid.accessor("write")
        fill(view, 0, rows, rows)
        del view
        grid.sync_to_device()
        best, k = float("inf"), 1
        for sample in range(n_samples + (1 if warmup else 0)):
            update = StencilUpdate(self.workload, Params(transition_function=params, halo_value=halo,
                                                         n_iterations=n_iters, blocking=True))
            out = update(grid)
            k = int(update.get_stats().fused_iterations)
            if sample > 0 or not warmup:
                best = min(best, update.get_walltime())
            del out, update
        return best, k


def max_perf(args):
    runner = Runner(args.workload)
    cell, n_sub = int(runner.info.cell_bytes), int(runner.info.n_subiterations)
    if args.workload == "fdtd":
        from stencilstream_b200 import workloads as W
        wh = W.FdtdExperiment(W.FDTD_MAX_GRID).grid_wh()
        rows = cols = wh
        n_iters = args.iterations or 1000
    else:
        rows = cols = args.rows or 16 * 2 ** 10
        n_iters = args.iterations or 1000
    print(f"Grid dimensions: {rows} x {cols}")
    print(f"Grid size: {rows * cols * cell * 2 ** -30} GB")
    print(f"No. of iterations: {n_iters}")
    runtime, k = runner.run(rows, cols, n_iters, n_samples=3)
    measured = rows * cols * n_iters / runtime
    model = rows * cols * n_iters / model_runtime(rows, cols, n_iters, cell, n_sub, k,
                                                  hbm_bytes_per_second())
    metrics = {"target": TARGET_NAME[args.workload], "measured": measured,
               "FLOPS": measured * OPERATIONS_PER_CELL[args.workload],
               "model": model, "accuracy": model / measured, "fused_iterations": k,
               "runtime": runtime}
    out = Path(args.out_dir) / "metrics.cuda.json"
    out.write_text(json.dumps(metrics, indent=1))
    print(json.dumps(metrics))


def deep_grid_scaling(args):
    runner = Runner(args.workload)
    if args.workload == "fdtd":
        raise SystemExit("the FDTD grid size follows from its experiment file; use max_perf")
    cell, n_sub = int(runner.info.cell_bytes), int(runner.info.n_subiterations)
    from stencilstream_b200 import _native
    import ctypes as C
    info = _native.DeviceInfo()
    _native.runtime_lib().stst_get_device_info(0, C.byref(info))
    wh = float(max_grid_wh(cell, float(info.total_mem), clip_to_base=math.sqrt(2)))
    if args.max_wh:
        wh = min(wh, float(args.max_wh))
    path = Path(args.out_dir) / "scaling.cuda.csv"
    columns = ["grid_wh", "n_iters", "runtime", "measured_throughput", "model_throughput"]
    rows_done = []
    if path.exists():
        with open(path) as f:
            rows_done = list(csv.DictReader(f))
    done = {int(r["grid_wh"]) for r in rows_done}
    bandwidth = hbm_bytes_per_second()
    first = True
    while round(wh) >= args.min_wh:
        true_wh = int(round(wh))
        wh /= math.sqrt(2)
        if true_wh in done:
            continue
        # iterations for `target_runtime` seconds according to the model of ONE iteration per pass
        # (the reference's proto_info for the cuda variant), then the run decides its own k
        per_iteration = model_runtime(true_wh, true_wh, 1, cell, n_sub, 1, bandwidth)
        n_iters = max(1, int(math.ceil(args.target_runtime / per_iteration)))
        runtime, k = runner.run(true_wh, true_wh, n_iters, n_samples=3, warmup=first)
        measured = true_wh * true_wh * n_iters / runtime
        model = true_wh * true_wh * n_iters / model_runtime(true_wh, true_wh, n_iters, cell, n_sub,
                                                            k, bandwidth)
        rows_done.append(dict(zip(columns, [true_wh, n_iters, runtime, measured, model])))
        with open(path, "w", newline="") as f:
            writer = csv.DictWriter(f, fieldnames=columns)
            writer.writeheader()
            writer.writerows(rows_done)
        print(f"{true_wh:6d}^2  {n_iters:8d} it  {runtime:8.4f} s  {measured / 1e9:9.2f} GCells/s "
              f"(model {model / 1e9:9.2f}, k={k})", flush=True)
        first = False


NCU_METRICS = [  # scripts/benchmark-common.jl:246-283
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__maximum_warps_per_active_cycle_pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
]


def ncu_metrics(args):
    rows, cols = args.rows or 16384, args.cols or args.rows or 16384
    print("ncu --csv --clock-control none -k regex:fused_sweep --metrics " + ",".join(NCU_METRICS) +
          f" python scripts/run_one.py --workload {args.workload} --rows {rows} --cols {cols} --iters 24")


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("experiment", choices=["max_perf", "deep_grid_scaling", "ncu_metrics"])
    ap.add_argument("workload", choices=sorted(OPERATIONS_PER_CELL))
    ap.add_argument("--out-dir", default=".")
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--cols", type=int, default=0)
    ap.add_argument("--iterations", type=int, default=0)
    ap.add_argument("--target-runtime", type=float, default=2.0,
                    help="seconds per grid size (the reference uses 30)")
    ap.add_argument("--min-wh", type=int, default=32)
    ap.add_argument("--max-wh", type=int, default=0)
    args = ap.parse_args()
    Path(args.out_dir).mkdir(parents=True, exist_ok=True)
    {"max_perf": max_perf, "deep_grid_scaling": deep_grid_scaling, "ncu_metrics": ncu_metrics}[
        args.experiment](args)


if __name__ == "__main__":
    main()
