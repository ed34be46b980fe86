"""The row-sharded convection application (`apps.run_convection_sharded`): two slab objects per rank
(pseudo-transient and thermal update over the same cells), device-to-device hand-over between them,
max-norms combined by all-reduce. Under gloo with the host-memory slab double (no GPU) and, marked
`gpu`, with real slabs (2 processes; they share the device on a one-GPU box). The result must equal
the oracle-driven single-process loop bit for bit."""
import os
import sys
import traceback
from pathlib import Path

import numpy as np
import pytest

from test_sharding_cpu import _free_port

ROOT = Path(__file__).resolve().parent.parent


def _config():
    from stencilstream_b200 import workloads as W
    config = W.convection_benchmark_config(res=48, n_iters=60, lx=1.0, ly=1.5)
    config.update(nt=3, nerr=10, nout=2, epsilon=3e-2, Ra=1e7)
    return config


def _worker(rank, world, port, mode, failures):
    try:
        sys.path.insert(0, str(ROOT))
        sys.path.insert(0, str(ROOT / "tests"))
        import torch.distributed as dist
        import cases as cases_mod
        import oracle
        from stencilstream_b200.apps import run_convection_sharded

        os.environ["OMP_NUM_THREADS"] = "1"
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                                world_size=world)
        checker = oracle.port()
        extra = {}
        if mode == "host":
            from fake_slab import HostSlab
            extra = dict(slab_factory=lambda **kw: HostSlab(checker, 1, **kw))
        else:
            import ctypes as C
            from stencilstream_b200 import _native
            count = C.c_int(0)
            _native.runtime_lib().stst_device_count(C.byref(count))
            extra = dict(device=rank % max(count.value, 1), strict=True)
        frames = []
        update, steps = run_convection_sharded(
            _config(), rank=rank, world=world, comm=dist,
            on_frame=lambda it, lo, hi, T: frames.append((it, lo, hi, T.copy())), **extra)
        mine = update.to_numpy()
        gathered = [None] * world
        dist.all_gather_object(gathered, (update.row_lo, update.row_hi, mine.tobytes(),
                                          [(f[0], f[1], f[2], f[3].tobytes()) for f in frames]))
        if rank == 0:
            want_cells, want_steps, want_frames = cases_mod.oracle_convection(checker, _config())
            got_steps = [(s.it, s.iterations, s.errV, s.errP, s.dt) for s in steps]
            if got_steps != want_steps:
                failures.put(f"steps differ: {got_steps} != {want_steps}")
            got = np.empty_like(want_cells)
            for lo, hi, raw, _ in gathered:
                got[lo:hi] = np.frombuffer(raw, dtype=want_cells.dtype).reshape(hi - lo, -1)
            if got.tobytes() != want_cells.tobytes():
                failures.put("sharded convection result differs from the oracle loop")
            for k, (it, T) in enumerate(want_frames):
                T_got = np.empty_like(T)
                for lo, hi, _, frs in gathered:
                    f_it, f_lo, f_hi, raw = frs[k]
                    n = max(0, min(f_hi, T.shape[0]) - f_lo)
                    if n:
                        T_got[f_lo:f_lo + n] = np.frombuffer(raw, dtype=T.dtype).reshape(n, -1)
                    if f_it != it:
                        failures.put(f"frame {k}: time step {f_it} != {it}")
                if T_got.tobytes() != T.tobytes():
                    failures.put(f"temperature frame of step {it} differs")
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        failures.put(f"rank {rank}: {traceback.format_exc()}")


def _run(world, mode):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    failures = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, failures)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
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


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_convection_application_under_gloo(world, built):
    _run(world, "host")


@pytest.mark.gpu
def test_sharded_convection_application_on_gpu(built):
    _run(2, "cuda")


def _fdtd_worker(rank, world, port, mode, failures):
    try:
        sys.path.insert(0, str(ROOT))
        sys.path.insert(0, str(ROOT / "tests"))
        import torch.distributed as dist
        import oracle
        from stencilstream_b200 import workloads as W
        from stencilstream_b200.apps import run_fdtd_sharded

        os.environ["OMP_NUM_THREADS"] = "1"
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                                world_size=world)
        checker = oracle.port()
        if mode == "host":
            from fake_slab import HostSlab
            extra = dict(slab_factory=lambda **kw: HostSlab(checker, 2, **kw), fused_iterations=2)
        else:
            import ctypes as C
            from stencilstream_b200 import _native
            count = C.c_int(0)
            _native.runtime_lib().stst_device_count(C.byref(count))
            extra = dict(device=rank % max(count.value, 1), strict=True)
        total, snap = 50, 20   # three snapshot intervals, the last one overshoots (fdtd.cpp:233-242)
        frames = []
        simulation = run_fdtd_sharded(
            W.FDTD_DEFAULT, rank=rank, world=world, comm=dist, n_timesteps=total,
            n_snap_timesteps=snap,
            on_frame=lambda f, i, lo, hi, v: frames.append((f, i, lo, hi, v.tobytes())), **extra)
        gathered = [None] * world
        dist.all_gather_object(gathered, (simulation.row_lo, simulation.row_hi,
                                          simulation.to_numpy().tobytes(), frames))
        if rank == 0:
            exp = W.FdtdExperiment(W.FDTD_DEFAULT)
            cells = exp.initial_grid()
            want = checker.run("fdtd", exp.kernel_params(), None, cells, 0, 60)
            got = np.empty_like(want)
            for lo, hi, raw, _ in gathered:
                got[lo:hi] = np.frombuffer(raw, dtype=want.dtype).reshape(hi - lo, -1)
            if got.tobytes() != want.tobytes():
                failures.put("sharded FDTD result differs from the oracle")
            labels = [(f, i) for f, i, *_ in gathered[0][3]]
            if labels != [("hz", 20), ("hz", 40), ("hz", 60), ("hz_sum", 50)]:
                failures.put(f"frames {labels}")
            mid = checker.run("fdtd", exp.kernel_params(), None, cells, 0, 40)
            hz = np.empty(mid.shape, dtype=np.float32)
            for lo, hi, _, frs in gathered:
                hz[lo:hi] = np.frombuffer(frs[1][4], dtype=np.float32).reshape(hi - lo, -1)
            if hz.tobytes() != np.ascontiguousarray(mid["hz"]).tobytes():
                failures.put("hz frame after 40 steps differs")
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        failures.put(f"rank {rank}: {traceback.format_exc()}")


def _run_fdtd(world, mode):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    failures = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fdtd_worker, args=(r, world, port, mode, failures)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
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


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_fdtd_snapshot_loop_under_gloo(world, built):
    _run_fdtd(world, "host")


@pytest.mark.gpu
def test_sharded_fdtd_snapshot_loop_on_gpu(built):
    _run_fdtd(2, "cuda")
