"""The multi-GPU partitioner's host logic under `gloo`, world_size 2 and 3, no GPU.

`ShardedStencilUpdate` (stencilstream_b200/sharding.py) runs with the host-memory slab double of
tests/fake_slab.py; the gathered result must equal the oracle's update of the whole grid — bit for
bit, since both sides are the same CPU arithmetic. What this pins: the row partition, neighbour
wiring through all_gather_object, the ghost depth k * n_subiterations * radius, tail passes,
global coordinates / iteration numbers / time-dependent values seen inside a slab.
"""
import os
import socket
import sys
import traceback
from pathlib import Path

import numpy as np
import pytest

import cases
from stencilstream_b200.sharding import partition_rows

ROOT = Path(__file__).resolve().parent.parent


def test_partition_rows_covers_everything_evenly():
    for rows in (1, 7, 64, 65, 1000, 16384):
        for count in (1, 2, 3, 4, 8):
            if count > rows:
                continue
            spans = [partition_rows(rows, count, i) for i in range(count)]
            assert spans[0][0] == 0 and spans[-1][1] == rows
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sorted(sizes, reverse=True) == sizes
    with pytest.raises(ValueError):
        partition_rows(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, workload, shape, offset, n, depth, failures, mode="host"):
    try:
        sys.path.insert(0, str(ROOT))
        sys.path.insert(0, str(ROOT / "tests"))
        import torch.distributed as dist
        import cases as cases_mod
        import oracle
        from fake_slab import HostSlab
        from stencilstream_b200 import Params
        from stencilstream_b200.sharding import ShardedStencilUpdate

        os.environ["OMP_NUM_THREADS"] = "1"
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                                world_size=world)
        checker = oracle.port()
        params, halo, cells = cases_mod.make_case(workload, *shape, seed=21)
        if "kat" in workload:
            cells = cases_mod.kat_input(*shape, offset)
        if mode in ("host", "host_violation"):
            extra = dict(slab_factory=lambda **kw: HostSlab(checker, depth, **kw))
        else:  # the real thing: sm_100a kernels, -fmad=false build, IPC-mapped neighbours
            from stencilstream_b200 import _native
            import ctypes as C
            count = C.c_int(0)
            _native.runtime_lib().stst_device_count(C.byref(count))
            extra = dict(device=rank % max(count.value, 1), strict=True,
                         overlap=(mode != "cuda-no-overlap"),
                         transport="nccl" if mode == "cuda-nccl" else "p2p")
        update = ShardedStencilUpdate(
            workload, Params(transition_function=params, halo_value=halo, iteration_offset=offset,
                             n_iterations=n, blocking=True, fused_iterations=depth),
            shape[0], shape[1], rank=rank, world=world, comm=dist, **extra)
        lo, hi = update.row_lo, update.row_hi
        update.load(cells[lo:hi])
        if mode == "host_violation" and rank == world - 1:
            update.slab.inject_violation_once = 0b10     # one slab reports a changed plane once
        update()
        if mode == "cuda-redo_expected":
            if update.n_speculation_redos < 1:
                failures.put(f"rank {rank}: no repeat although hz_sum changes at iteration 10")
            if update.info().passthrough_planes & 0b11111000 != 0b11110000:
                failures.put(f"rank {rank}: pass-through planes {update.info().passthrough_planes:#b}")
        if mode == "host_violation":
            # every slab went back to the saved generation and repeated the call
            if update.n_speculation_redos != 1 or update.slab.log.count("restore") != 1:
                failures.put(f"rank {rank}: redos {update.n_speculation_redos}, log {update.slab.log}")
        # a second call resumes where the first stopped (iteration_offset is a live parameter)
        update.get_params().iteration_offset += n
        update.get_params().n_iterations = 2
        update()
        mine = update.to_numpy()
        # per-field device operations: max-norms over the whole grid (collective) and a single-field
        # download of the owned rows
        names = mine.dtype.names or (0,)
        extents = [(f, shape[0] - 1 - (i % 2), shape[1] - (i % 3)) for i, f in enumerate(names)]
        norms = update.max_abs(extents)
        if mine.dtype.names:
            last = names[-1]
            if update.field_to_numpy(last).tobytes() != np.ascontiguousarray(mine[last]).tobytes():
                failures.put(f"{workload}: single-field download differs from the cells' field")
        gathered = [None] * world
        dist.all_gather_object(gathered, (lo, hi, mine.tobytes()))
        if rank == 0:
            want = checker.run(workload, params, halo, cells, offset, n + 2)
            got = np.empty_like(want)
            for glo, ghi, raw in gathered:
                got[glo:ghi] = np.frombuffer(raw, dtype=want.dtype).reshape(ghi - glo, shape[1])
            for (f, r, c), norm in zip(extents, norms):
                plane = want[f] if want.dtype.names else want
                expect_norm = float(np.abs(plane[:r, :c].astype(np.float64)).max())
                if norm != expect_norm:
                    failures.put(f"{workload}: max_abs({f}, {r}, {c}) = {norm!r}, oracle {expect_norm!r}")
            if got.tobytes() != want.tobytes():
                bad = np.argwhere(got.view(np.uint8).reshape(shape[0], -1)
                                  != want.view(np.uint8).reshape(shape[0], -1))
                failures.put(f"{workload}: sharded result differs from the whole-grid oracle, "
                             f"first at row {bad[0][0]}")
            if mode == "host_violation":
                dist.barrier()
                dist.destroy_process_group()
                return
            if mode != "host":
                if update.info().fused_iterations != depth:
                    failures.put(f"{workload}: fusion depth {update.info().fused_iterations}")
                dist.barrier()
                dist.destroy_process_group()
                return
            passes = [e for e in update.slab.log if isinstance(e, tuple)]
            expect = []
            it, rem = offset, n
            while rem > 0:
                expect.append(("pass", it, min(rem, depth)))
                it += min(rem, depth)
                rem -= min(rem, depth)
            expect.append(("pass", offset + n, 2 if depth >= 2 else 1))
            if depth < 2:
                expect.append(("pass", offset + n + 1, 1))
            if passes != expect:
                failures.put(f"{workload}: pass sequence {passes} != {expect}")
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        failures.put(f"rank {rank}: {traceback.format_exc()}")


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("workload,shape,offset,n,depth", [
    ("hotspot", (45, 37), 0, 7, 3),        # reads stencil.id / grid_range at the borders
    ("kat", (41, 19), 5, 5, 2),            # self-checking: coordinates, iteration, tdv, halo
    ("fdtd", (50, 33), 3, 4, 2),           # two sub-iterations + time-dependent value
    ("jacobi_r3", (60, 20), 0, 5, 2),      # radius 3: ghost depth 6
    ("conway", (33, 64), 0, 9, 4),
    ("convection_pt", (40, 24), 0, 3, 1),  # three sub-iterations, fp64
])
def test_sharded_update_equals_whole_grid(world, workload, shape, offset, n, depth, built):
    run_group(world, workload, shape, offset, n, depth, "host")


def test_reported_violation_makes_every_slab_repeat_the_call(built):
    """Plane pass-through on slabs: one slab reporting a violation makes ALL slabs restore the saved
    generation and repeat (one all-reduce per call); the result is still the whole-grid oracle's."""
    run_group(3, "hotspot", (45, 37), 0, 7, 3, "host_violation")


def run_group(world, workload, shape, offset, n, depth, mode):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    failures = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker,
                         args=(r, world, port, workload, shape, offset, n, depth, failures, mode))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    messages = []
    while not failures.empty():
        messages.append(failures.get())
    for p in procs:
        if p.is_alive():
            p.kill()
            messages.append("worker timed out")
        elif p.exitcode != 0:
            messages.append(f"worker exit code {p.exitcode}")
    assert not messages, "\n".join(messages)


def _sequence_worker(rank, world, port, failures):
    """Two sharded updates one after the other in the same process group (what bench.py's default
    invocation does for its `workloads`): close() is collective and leaves the group usable."""
    try:
        sys.path.insert(0, str(ROOT))
        sys.path.insert(0, str(ROOT / "tests"))
        import torch.distributed as dist
        import cases as cases_mod
        import oracle
        from fake_slab import HostSlab
        from stencilstream_b200 import Params
        from stencilstream_b200.sharding import ShardedStencilUpdate

        os.environ["OMP_NUM_THREADS"] = "1"
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                                world_size=world)
        checker = oracle.port()
        for workload, shape, n, depth in (("hotspot", (40, 30), 5, 2), ("jacobi5", (36, 28), 4, 3)):
            params, halo, cells = cases_mod.make_case(workload, *shape, seed=5)
            update = ShardedStencilUpdate(
                workload, Params(transition_function=params, halo_value=halo, n_iterations=n,
                                 blocking=True, fused_iterations=depth),
                shape[0], shape[1], rank=rank, world=world, comm=dist,
                slab_factory=lambda **kw: HostSlab(checker, depth, **kw))
            update.load(cells[update.row_lo:update.row_hi])
            update()
            mine = update.to_numpy()
            want = checker.run(workload, params, halo, cells, 0, n)
            if mine.tobytes() != np.ascontiguousarray(want[update.row_lo:update.row_hi]).tobytes():
                failures.put(f"rank {rank}: {workload} differs")
            update.close()
            update.close()          # idempotent
            if update.slab is not None:
                failures.put(f"rank {rank}: slab still there after close()")
        try:
            ShardedStencilUpdate("jacobi5", Params(n_iterations=1), 8, 8, rank=rank, world=world, comm=dist,
                                 slab_factory=lambda **kw: HostSlab(checker, 1, **kw), transport="carrier-pigeon")
            failures.put("an unknown transport was accepted")
        except ValueError:
            pass
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        failures.put(f"rank {rank}: {traceback.format_exc()}")


def test_sharded_updates_can_follow_one_another(built):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    failures = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sequence_worker, args=(r, 2, port, failures)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    messages = []
    while not failures.empty():
        messages.append(failures.get())
    for p in procs:
        if p.is_alive():
            p.kill()
            messages.append("worker timed out")
        elif p.exitcode != 0:
            messages.append(f"worker exit code {p.exitcode}")
    assert not messages, "\n".join(messages)
