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
  ncu_metrics        ncu_profile_command of scripts/benchmark-common.jl:246-283: profiles one launch of
                     the fused sweep kernel under ncu, scrapes the reference's ten figures (occupancy,
                     SM/DRAM/L1/L2 throughput, DRAM read/write volume, sectors per request) and merges
                     them into metrics.cuda.json (run it under gpurun; ncu replays the launch).

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


# ---- ncu scraping (reference scripts/benchmark-common.jl:246-283) ----------------------------------------
# The reference profiles its command twice (default sections, then MemoryWorkloadAnalysis_Tables with
# --print-details all), picks one launch and stores ten figures next to the throughput numbers of
# metrics.<variant>.json. Same figures here, from ONE profiled launch of the fused sweep kernel.

NCU_SECTIONS = ["SpeedOfLight", "Occupancy", "MemoryWorkloadAnalysis_Tables"]
NCU_RAW_METRICS = [
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
]
_UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def ncu_command(workload, rows, cols, iterations, launch_skip=1):
    cmd = ["ncu", "--csv", "--clock-control", "none", "-k", "regex:fused_sweep_kernel",
           "-s", str(launch_skip), "-c", "1", "--print-details", "all"]
    for section in NCU_SECTIONS:
        cmd += ["--section", section]
    cmd += ["--metrics", ",".join(NCU_RAW_METRICS)]
    cmd += [sys.executable, str(ROOT / "scripts" / "run_one.py"), "--workload", workload, "--rows",
            str(rows), "--cols", str(cols), "--iters", str(iterations), "--calls", "1"]
    return cmd


def parse_ncu_csv(text):
    """Rows of ncu's --csv output as dictionaries; everything before the header line (the profiled
    program's own output, ==PROF== lines, ==WARNING== lines) is skipped, like
    extract_ncu_profiling_data does (benchmark-common.jl:222-244)."""
    lines = [line for line in text.splitlines() if not line.startswith("==")]
    start = next((i for i, line in enumerate(lines) if line.startswith('"ID"')), None)
    if start is None:
        raise ValueError("no ncu CSV header in the output")
    return list(csv.DictReader(lines[start:]))


def _number(row):
    value = float(row["Metric Value"].replace(",", ""))
    return value * _UNIT_SCALE.get(row["Metric Unit"], 1.0)


def scrape_ncu_metrics(rows, launch_id=None):
    """The ten figures of ncu_profile_command (benchmark-common.jl:246-283) for one launch."""
    if launch_id is None:
        launch_id = rows[0]["ID"]
    mine = [r for r in rows if r["ID"] == str(launch_id)]

    def metric(name):
        for r in mine:
            if r["Metric Name"] == name:
                return _number(r)
        raise KeyError(f"ncu output has no metric {name!r} for launch {launch_id}")

    return {
        "achieved_occupancy": metric("Achieved Occupancy"),
        "compute_throughput": metric("Compute (SM) Throughput"),
        "dram_throughput": metric("DRAM Throughput"),
        "l1_throughput": metric("L1/TEX Cache Throughput"),
        "l2_throughput": metric("L2 Cache Throughput"),
        "theoretical_occupancy": metric("Theoretical Occupancy"),
        "read_volume": metric("dram__bytes_read.sum"),
        "write_volume": metric("dram__bytes_write.sum"),
        "sectors_per_load_request": metric("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")
        / max(metric("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"), 1.0),
        "sectors_per_store_request": metric("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum")
        / max(metric("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"), 1.0),
        "kernel": mine[0]["Kernel Name"] if mine else None,
    }


def ncu_metrics(args):
    """Profile one full-depth launch of the workload under ncu and merge the reference's ten
    profiling figures into <out-dir>/metrics.cuda.json (the file max_perf writes)."""
    import subprocess

    runner_info = Runner(args.workload).info
    if args.workload == "fdtd":
        from stencilstream_b200 import workloads as W
        rows = cols = W.FdtdExperiment(W.FDTD_MAX_GRID).grid_wh()
    else:
        rows, cols = args.rows or 16384, args.cols or args.rows or 16384
    cmd = ncu_command(args.workload, rows, cols, args.iterations or 48)
    print("+ " + " ".join(cmd), file=sys.stderr)
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if proc.returncode != 0:
        raise SystemExit(f"ncu failed ({proc.returncode}):\n{proc.stdout[-2000:]}\n{proc.stderr[-2000:]}")
    (Path(args.out_dir) / f"ncu_{args.workload}.csv").write_text(proc.stdout)
    figures = scrape_ncu_metrics(parse_ncu_csv(proc.stdout))
    figures["grid"] = [rows, cols]
    figures["cell_size"] = int(runner_info.cell_bytes)
    out = Path(args.out_dir) / "metrics.cuda.json"
    metrics = json.loads(out.read_text()) if out.exists() else {"target": TARGET_NAME[args.workload]}
    metrics.update(figures)
    out.write_text(json.dumps(metrics, indent=1))
    print(json.dumps(figures))


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
