"""What the box's host links can move: pinned host <-> device copies on all ranks at once.

    torchrun --nproc-per-node N scripts/host_link_ceiling.py [--gib 1] [--bind]

Every rank pins `--gib` GiB, then all ranks copy it to their GPU and back simultaneously (barrier
before, max over ranks after), through the runtime's own copy call (stst_memcpy_2d_auto, the path the
cells take). Prints one JSON line: per-direction aggregate and per-GPU GB/s. This is the ceiling the
end-to-end figures of bench.py at N GPUs are to be judged against — the transfers of a step cannot go
faster than this, whatever the kernels do.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--bind", action="store_true", help="bind each rank to its GPU's NUMA node first")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    device = int(os.environ.get("LOCAL_RANK", "0"))
    placement = None
    if args.bind:
        from stencilstream_b200.affinity import bind_to_gpu_numa_node
        placement = bind_to_gpu_numa_node(device)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", device))
    from stencilstream_b200 import _native
    rt = _native.runtime_lib()
    n_bytes = int(args.gib * (1 << 30))
    host = C.c_void_p()
    assert rt.stst_malloc_host(n_bytes, C.byref(host)) == 0, rt.stst_last_error()
    C.memset(host, 1, n_bytes)
    stream = C.c_void_p()
    assert rt.stst_default_stream(device, C.byref(stream)) == 0
    dev = C.c_void_p()
    assert rt.stst_malloc(device, n_bytes, stream, C.byref(dev)) == 0
    row = 1 << 20

    def copy(direction):
        assert rt.stst_memcpy_2d_auto(dev, row, host, row, row, n_bytes // row, direction, device, stream) == 0
        assert rt.stst_stream_synchronize(stream) == 0

    def timed(direction):
        best = float("inf")
        for _ in range(args.reps):
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            copy(direction)
            seconds = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([seconds], dtype=torch.float64, device=f"cuda:{device}")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                seconds = float(t.item())
            best = min(best, seconds)
        return best

    copy(0)
    h2d, d2h = timed(0), timed(1)
    gathered = [placement]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, placement)
    if rank == 0:
        gb = n_bytes / 1e9
        print(json.dumps({
            "n_gpus": world, "gib_per_rank": args.gib, "bound_to_numa_node": bool(args.bind),
            "h2d_gbs_per_gpu": gb / h2d, "d2h_gbs_per_gpu": gb / d2h,
            "h2d_gbs_aggregate": world * gb / h2d, "d2h_gbs_aggregate": world * gb / d2h,
            "round_trip_ms_1gib": (h2d + d2h) * 1e3 / args.gib,
            "host_cores": len(os.sched_getaffinity(0)), "placement": gathered,
            "how": "all ranks copy simultaneously, pinned memory, stst_memcpy_2d_auto, best of "
                   f"{args.reps}, max over ranks",
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
