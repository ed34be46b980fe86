"""The multi-GPU partitioner on real hardware: `world` processes, one slab each, neighbours mapped
through CUDA IPC, halos pushed by the fused sweep kernel, ordering by stream-ordered flags.

The processes use device `rank % device_count`, so the whole protocol (IPC mapping, peer stores, flag
waits) is also exercised on a one-GPU box, where all slabs live on the same device. The gathered
result must equal the CPU oracle's update of the whole grid bit for bit (-fmad=false build).
"""
import pytest

from test_sharding_cpu import run_group

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["cuda", "cuda-no-overlap"])
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("workload,shape,offset,n,depth", [
    ("hotspot", (330, 700), 0, 7, 3),
    ("kat", (97, 260), 5, 5, 2),
    ("fdtd", (150, 333), 3, 4, 2),
    ("jacobi_r3", (200, 520), 0, 5, 2),
    ("conway", (333, 640), 0, 9, 4),
    ("convection_pt", (120, 96), 0, 3, 1),
    ("jacobi5", (1000, 1030), 0, 20, 4),
])
def test_sharded_update_equals_whole_grid_on_gpu(mode, world, workload, shape, offset, n, depth, built):
    run_group(world, workload, shape, offset, n, depth, mode)


@pytest.mark.parametrize("world", [2, 3])
def test_wrong_passthrough_guess_is_repeated_on_every_slab(world, built):
    """FDTD's `hz_sum` starts to change at detect_iteration = 10: the slabs' observing pass (iterations
    0-1) marks it as passing through, a later pass reports the change, all slabs restore the saved
    generation and repeat the call. `redo_expected` makes the workers assert that this happened and
    that the material coefficients still pass through afterwards."""
    run_group(world, "fdtd", (150, 333), 0, 14, 2, "cuda-redo_expected")


def _device_count():
    import ctypes as C
    from stencilstream_b200 import _native
    count = C.c_int(0)
    _native.runtime_lib().stst_device_count(C.byref(count))
    return count.value


@pytest.mark.parametrize("workload,shape,offset,n,depth", [
    ("hotspot", (330, 700), 0, 7, 3),
    ("fdtd", (150, 333), 3, 4, 2),
    ("jacobi5", (1000, 1030), 0, 20, 4),
])
def test_nccl_halo_transport_equals_whole_grid(workload, shape, offset, n, depth, built):
    """The portable halo route: one grouped ncclSend/ncclRecv exchange per pass through the runtime's
    own communicator (stst_nccl_comm_init_rank / stst_nccl_neighbor_exchange, include/stst_rt.h)
    instead of stores into IPC-mapped neighbour memory. NCCL refuses two ranks on one device, so this
    needs a box with at least two GPUs."""
    if _device_count() < 2:
        pytest.skip("NCCL needs one GPU per rank")
    run_group(2, workload, shape, offset, n, depth, "cuda-nccl")
