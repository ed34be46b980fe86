"""The reference's FDTD `max_grid` experiment, complete: 4608 x 4608 cells, 184 911 time steps in 15
snapshot intervals of 12 328 steps (the loop overshoots to 184 920, examples/fdtd/src/fdtd.cpp:233-242;
experiments/max_grid.json), an `hz` frame fetched after every interval and `hz_sum` at the end.

    python scripts/fdtd_max_grid.py                       # one GPU
    torchrun --nproc-per-node 8 scripts/fdtd_max_grid.py  # row slabs, one process per GPU

Prints one JSON line: wall time of the whole loop (frames included), GCell-updates/s, per-interval
times and a checksum of every frame (so that the 1-GPU and the N-GPU run can be compared).
"""
import hashlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from stencilstream_b200 import workloads as W
from stencilstream_b200.apps import run_fdtd, run_fdtd_sharded


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    device = int(os.environ.get("LOCAL_RANK", "0"))
    strict = os.environ.get("STST_STRICT", "0") not in ("", "0")
    exp = W.FdtdExperiment(W.FDTD_MAX_GRID)
    total, snap, wh = exp.n_timesteps(), exp.n_snap_timesteps(), exp.grid_wh()
    stamps, digests = [], []
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(device)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", device))
        dist = dist_mod
        from stencilstream_b200.affinity import bind_to_gpu_numa_node
        bind_to_gpu_numa_node(device)

    t0 = time.perf_counter()
    if world == 1:
        def on_frame(field, iteration, values):
            stamps.append(time.perf_counter())
            digests.append((field, iteration, hashlib.sha256(values.tobytes()).hexdigest()[:16]))
        os.environ["STST_DEVICE"] = str(device)
        grid, simulation = run_fdtd(W.FDTD_MAX_GRID, strict=strict, on_frame=on_frame)
        launches = simulation.get_n_launches()
        k = int(simulation.get_stats().fused_iterations)
    else:
        def on_frame(field, iteration, lo, hi, values):
            stamps.append(time.perf_counter())
            digests.append((field, iteration, lo, hashlib.sha256(values.tobytes()).hexdigest()[:16]))
        simulation = run_fdtd_sharded(W.FDTD_MAX_GRID, rank=rank, world=world, device=device, comm=dist,
                                      strict=strict, on_frame=on_frame)
        launches = simulation.get_n_launches()
        k = int(simulation.info().fused_iterations)
    seconds = time.perf_counter() - t0
    if dist is not None:
        import torch
        t = torch.tensor([seconds], dtype=torch.float64, device=f"cuda:{device}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        seconds = float(t.item())
        everyone = [None] * world
        dist.all_gather_object(everyone, digests)
        digests = [d for per_rank in everyone for d in per_rank]
    if rank == 0:
        n_done = -(-total // snap) * snap
        frames = [s for s in stamps]
        print(json.dumps({
            "experiment": "examples/fdtd/experiments/max_grid.json", "grid": [wh, wh],
            "n_timesteps": total, "n_snap_timesteps": snap, "timesteps_computed": n_done,
            "n_gpus": world, "build": "-fmad=false" if strict else "default",
            "seconds_including_setup_and_frames": seconds,
            "gcell_updates_per_s": wh * wh * n_done / seconds / 1e9,
            "fused_iterations": k, "launches_rank0": int(launches),
            "interval_seconds": [round(b - a, 3) for a, b in zip([t0] + frames[:-1], frames)],
            "frame_digests": sorted(digests),
        }), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
