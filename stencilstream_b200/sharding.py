"""Row-sharded generation loop: one process per GPU, one slab of the grid per process.

Host-side mirror of the multi-GPU partitioner in
stencilstream_b200/include/StencilStream/cuda/internal/SlabUpdate.hpp (C ABI: the `stst_slab_*`
functions of include/stst_workloads.h). The reference's cuda backend is single-device
(reference StencilStream/cuda/StencilUpdate.hpp:83), so there is no reference interface to mirror for
the sharding itself; the object below keeps the shape of the reference's `StencilUpdate`
(`Params`, call operator advancing by `n_iterations`, `get_params()` live reference) and adds what a
slab needs: `load()` for the owned rows and `to_numpy()` to read them back.

`torch.distributed` is plumbing only: it moves the 64-byte CUDA IPC handle and the row range of every
slab to its two neighbours once, at construction. The data path has no collective — every fused
launch stores the boundary rows straight into the neighbours' ghost rows over NVLink and the slabs
order themselves with stream-ordered device flags.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Callable

import numpy as np

from . import _native
from .api import Params, StencilStreamError, _as_param_struct, _check


def partition_rows(rows: int, count: int, index: int) -> tuple[int, int]:
    """Rows [lo, hi) owned by shard `index` of `count`: as even as possible, the first
    `rows % count` shards one row larger (same rule as internal::partition_rows in SlabUpdate.hpp)."""
    if count <= 0 or not 0 <= index < count:
        raise ValueError("illegal shard index")
    base, extra = divmod(rows, count)
    lo = index * base + min(index, extra)
    return lo, lo + base + (1 if index < extra else 0)


class SlabInfo(C.Structure):
    _fields_ = [
        ("grid_rows", C.c_size_t), ("grid_cols", C.c_size_t), ("row_lo", C.c_size_t),
        ("row_hi", C.c_size_t), ("ghost_rows", C.c_size_t), ("device_bytes", C.c_size_t),
        ("n_launches", C.c_size_t), ("epoch", C.c_size_t), ("device", C.c_int),
        ("fused_iterations", C.c_uint), ("tile_h", C.c_uint), ("tile_w", C.c_uint),
        ("block_x", C.c_uint), ("block_y", C.c_uint), ("use_tma", C.c_uint), ("overlap", C.c_uint),
        ("smem_bytes", C.c_size_t), ("passthrough_planes", C.c_uint),
    ]


def _slab_lib(strict: bool | None = None):
    lib = _native.workloads_lib(strict)
    if not getattr(lib, "_slab_prototypes", False):
        vp = C.c_void_p
        lib.stst_slab_create.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t,
                                         C.c_int, C.c_uint, C.c_uint, C.c_int, C.POINTER(vp)]
        lib.stst_slab_destroy.argtypes = [vp]
        lib.stst_slab_get_info.argtypes = [vp, C.POINTER(SlabInfo)]
        lib.stst_slab_get_ipc_handle.argtypes = [vp, C.c_char_p]
        lib.stst_slab_attach_ipc.argtypes = [vp, C.c_int, C.c_char_p, C.c_size_t, C.c_size_t]
        lib.stst_slab_attach_local.argtypes = [vp, C.c_int, vp]
        lib.stst_slab_detach.argtypes = [vp]
        lib.stst_slab_use_nccl.argtypes = [vp, vp, C.c_int, C.c_int]
        lib.stst_slab_copy_from_host.argtypes = [vp, vp, C.c_size_t]
        lib.stst_slab_copy_to_host.argtypes = [vp, vp, C.c_size_t]
        lib.stst_slab_copy_rows_from_host.argtypes = [vp, C.c_size_t, C.c_size_t, vp, C.c_size_t]
        lib.stst_slab_copy_rows_to_host.argtypes = [vp, C.c_size_t, C.c_size_t, vp, C.c_size_t]
        lib.stst_slab_exchange_halos.argtypes = [vp]
        lib.stst_slab_update.argtypes = [vp, C.POINTER(_native.UpdateParams)]
        lib.stst_slab_synchronize.argtypes = [vp]
        lib.stst_slab_record_event.argtypes = [vp, vp]
        lib.stst_slab_copy_from_slab.argtypes = [vp, vp]
        lib.stst_slab_enable_speculation.argtypes = [vp, C.c_int, C.POINTER(C.c_int)]
        lib.stst_slab_backup.argtypes = [vp]
        lib.stst_slab_restore.argtypes = [vp]
        lib.stst_slab_take_violations.argtypes = [vp, C.POINTER(C.c_uint)]
        lib.stst_slab_drop_passthrough.argtypes = [vp, C.c_uint]
        lib.stst_slab_max_abs.argtypes = [vp, C.POINTER(_native.FieldExtent), C.c_size_t,
                                          C.POINTER(C.c_double)]
        lib.stst_slab_copy_field_rows_to_host.argtypes = [vp, C.c_size_t, C.c_size_t, C.c_size_t, vp,
                                                          C.c_size_t]
        lib._slab_prototypes = True
    return lib


class NativeSlab:
    """ctypes handle to one `stst_slab` (sm_100a kernels; fails loudly without a CUDA device)."""

    def __init__(self, workload: str, grid_rows: int, grid_cols: int, row_lo: int, row_hi: int,
                 device: int, fused_iterations: int = 0, tile_rows: int = 0, overlap: bool = True,
                 strict: bool | None = None):
        self.workload = workload
        self.dtype = _native.CELL_DTYPES[workload]
        self._lib = _slab_lib(strict)
        handle = C.c_void_p()
        _check(self._lib, self._lib.stst_slab_create(
            workload.encode(), grid_rows, grid_cols, row_lo, row_hi, device, fused_iterations,
            tile_rows, int(bool(overlap)), C.byref(handle)))
        self._handle = handle
        self._keepalive = None

    def close(self) -> None:
        handle = getattr(self, "_handle", None)
        if handle:
            self._lib.stst_slab_destroy(handle)
            self._handle = None

    __del__ = close

    def info(self) -> SlabInfo:
        info = SlabInfo()
        _check(self._lib, self._lib.stst_slab_get_info(self._handle, C.byref(info)))
        return info

    def ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        _check(self._lib, self._lib.stst_slab_get_ipc_handle(self._handle, buf))
        return buf.raw

    def attach_ipc(self, side: int, handle: bytes, row_lo: int, row_hi: int) -> None:
        _check(self._lib, self._lib.stst_slab_attach_ipc(self._handle, side, handle, row_lo, row_hi))

    def attach_local(self, side: int, peer: "NativeSlab") -> None:
        _check(self._lib, self._lib.stst_slab_attach_local(self._handle, side, peer._handle))

    def use_nccl(self, comm, up_rank: int, down_rank: int) -> None:
        """Halo transport through the NCCL communicator `comm` (see stst_slab_use_nccl)."""
        _check(self._lib, self._lib.stst_slab_use_nccl(self._handle, comm, up_rank, down_rank))

    def detach(self) -> None:
        """Wait for this slab, then forget and unmap both neighbours."""
        _check(self._lib, self._lib.stst_slab_detach(self._handle))

    def copy_from_host(self, cells: np.ndarray) -> None:
        arr = np.ascontiguousarray(cells, dtype=self.dtype)
        self._keepalive = arr  # the copy may still be in flight when this returns
        _check(self._lib, self._lib.stst_slab_copy_from_host(
            self._handle, arr.ctypes.data_as(C.c_void_p), arr.nbytes))

    def copy_to_host(self, out: np.ndarray) -> None:
        if out.dtype != self.dtype or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("copy_to_host needs a C-contiguous array of the workload's cell dtype")
        _check(self._lib, self._lib.stst_slab_copy_to_host(
            self._handle, out.ctypes.data_as(C.c_void_p), out.nbytes))

    def copy_rows_from_host(self, first_row: int, cells: np.ndarray) -> None:
        """Replace the owned rows [first_row, first_row + len(cells)) (slab-local indices)."""
        arr = np.ascontiguousarray(cells, dtype=self.dtype)
        _check(self._lib, self._lib.stst_slab_copy_rows_from_host(
            self._handle, first_row, arr.shape[0], arr.ctypes.data_as(C.c_void_p), arr.nbytes))
        self.synchronize()  # `arr` may be a temporary

    def copy_rows_to_host(self, first_row: int, out: np.ndarray) -> None:
        if out.dtype != self.dtype or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("copy_rows_to_host needs a C-contiguous array of the cell dtype")
        _check(self._lib, self._lib.stst_slab_copy_rows_to_host(
            self._handle, first_row, out.shape[0], out.ctypes.data_as(C.c_void_p), out.nbytes))

    def exchange_halos(self) -> None:
        _check(self._lib, self._lib.stst_slab_exchange_halos(self._handle))

    def copy_from_slab(self, source: "NativeSlab") -> None:
        """Replace the owned rows by those of `source` (same grid, rows and cell type, same device;
        device-to-device)."""
        _check(self._lib, self._lib.stst_slab_copy_from_slab(self._handle, source._handle))

    def max_abs(self, extents) -> list[float]:
        """[(field, rows, cols), ...] in GLOBAL grid coordinates -> this slab's share of each
        max-norm (-inf where it owns none of the rows)."""
        extents = list(extents)
        out = (C.c_double * max(len(extents), 1))()
        _check(self._lib, self._lib.stst_slab_max_abs(
            self._handle, _native.field_extents(self.workload, extents), len(extents), out))
        return [float(out[q]) for q in range(len(extents))]

    def field_rows_to_host(self, field, first_row: int, out: np.ndarray) -> None:
        """ONE field of the owned rows [first_row, first_row + len(out)) (slab-local indices)."""
        if out.dtype != _native.field_dtype(self.workload, field) or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("field_rows_to_host needs a C-contiguous array of the field's dtype")
        _check(self._lib, self._lib.stst_slab_copy_field_rows_to_host(
            self._handle, _native.field_index(self.workload, field), first_row, out.shape[0],
            out.ctypes.data_as(C.c_void_p), out.nbytes))

    def update(self, native_params) -> None:
        _check(self._lib, self._lib.stst_slab_update(self._handle, C.byref(native_params)))

    # -- speculative plane pass-through (include/stst_workloads.h) ------------------------------------
    def enable_speculation(self, on: bool = True) -> bool:
        enabled = C.c_int(0)
        _check(self._lib, self._lib.stst_slab_enable_speculation(self._handle, int(on),
                                                                 C.byref(enabled)))
        return bool(enabled.value)

    def backup(self) -> None:
        _check(self._lib, self._lib.stst_slab_backup(self._handle))

    def restore(self) -> None:
        _check(self._lib, self._lib.stst_slab_restore(self._handle))

    def take_violations(self) -> int:
        planes = C.c_uint(0)
        _check(self._lib, self._lib.stst_slab_take_violations(self._handle, C.byref(planes)))
        return int(planes.value)

    def drop_passthrough(self, planes: int) -> None:
        _check(self._lib, self._lib.stst_slab_drop_passthrough(self._handle, int(planes)))

    def synchronize(self) -> None:
        _check(self._lib, self._lib.stst_slab_synchronize(self._handle))

    def record_event(self, event) -> None:
        _check(self._lib, self._lib.stst_slab_record_event(self._handle, event))


class ShardedStencilUpdate:
    """The generation loop on this process's slab of a `grid_rows x grid_cols` grid.

    Every process of the group constructs one with the same arguments except `rank`/`device`
    (construction is collective: it exchanges slab handles), then issues the same sequence of
    `load()` and calls. `comm` is `torch.distributed` (or anything with `all_gather_object`);
    `slab_factory` builds the slab object (the product default is `NativeSlab`; the CPU test-suite
    injects a host-memory double to exercise this module without a GPU).
    """

    def __init__(self, workload: str, params: Params, grid_rows: int, grid_cols: int, *, rank: int,
                 world: int, device: int = 0, comm: Any = None, overlap: bool = True,
                 strict: bool | None = None, speculate: bool = True,
                 slab_factory: Callable[..., Any] | None = None, transport: str | None = None):
        if world > 1 and comm is None:
            raise ValueError("a process group is needed to exchange slab handles")
        self.workload, self.params = workload, params
        self._comm = comm
        self.grid_rows, self.grid_cols = int(grid_rows), int(grid_cols)
        self.rank, self.world, self.device = rank, world, device
        self.row_lo, self.row_hi = partition_rows(self.grid_rows, world, rank)
        self.dtype = _native.CELL_DTYPES[workload]
        factory = slab_factory or (lambda **kw: NativeSlab(strict=strict, **kw))

        def make(fused):
            return factory(workload=workload, grid_rows=self.grid_rows, grid_cols=self.grid_cols,
                           row_lo=self.row_lo, row_hi=self.row_hi, device=device,
                           fused_iterations=fused, tile_rows=int(params.tile_rows), overlap=overlap)

        self.slab = make(int(params.fused_iterations))
        if world > 1:
            # All slabs must fuse the same number of iterations (it fixes the ghost depth).
            depths = [None] * world
            comm.all_gather_object(depths, int(self.slab.info().fused_iterations))
            if len(set(depths)) != 1:
                self.slab.close()
                self.slab = make(min(depths))
            mine = (self.slab.ipc_handle(), self.row_lo, self.row_hi,
                    int(self.slab.info().fused_iterations))
            everyone = [None] * world
            comm.all_gather_object(everyone, mine)
            if len({e[3] for e in everyone}) != 1:
                raise StencilStreamError("slabs disagree on the fusion depth")
            import os
            self.transport = transport or os.environ.get("STST_HALO_TRANSPORT", "p2p")
            if self.transport == "nccl":
                # the portable route: one grouped ncclSend/ncclRecv exchange per pass, through a
                # communicator of the runtime's own (include/stst_rt.h); rank 0's id reaches the others
                # through the process group
                rt = _native.runtime_lib()
                uid = C.create_string_buffer(128)
                if rank == 0 and rt.stst_nccl_get_unique_id(uid) != 0:
                    raise StencilStreamError(rt.stst_last_error().decode())
                ids = [None] * world
                comm.all_gather_object(ids, uid.raw)
                rt.stst_set_device(device)
                self._nccl = C.c_void_p()
                if rt.stst_nccl_comm_init_rank(C.byref(self._nccl), world, ids[0], rank) != 0:
                    raise StencilStreamError(rt.stst_last_error().decode())
                self.slab.use_nccl(self._nccl, rank - 1 if rank > 0 else -1,
                                   rank + 1 if rank + 1 < world else -1)
            elif self.transport == "p2p":
                for side, other in ((0, rank - 1), (1, rank + 1)):
                    if 0 <= other < world:
                        handle, lo, hi, _ = everyone[other]
                        self.slab.attach_ipc(side, handle, lo, hi)
            else:
                raise ValueError(f"unknown halo transport {self.transport!r} (p2p or nccl)")
        self._keepalive = None
        # plane pass-through: every slab must agree on whether the protocol runs at all
        wants = bool(speculate) and hasattr(self.slab, "enable_speculation") and \
            self.slab.enable_speculation(True)
        if world > 1:
            votes = [None] * world
            comm.all_gather_object(votes, wants)
            wants = all(votes)
            if not wants and hasattr(self.slab, "enable_speculation"):
                self.slab.enable_speculation(False)
        self.speculating = wants
        self.n_speculation_redos = 0

    # -- data ---------------------------------------------------------------------------------------
    @property
    def owned_shape(self) -> tuple[int, int]:
        return (self.row_hi - self.row_lo, self.grid_cols)

    def load(self, cells: np.ndarray) -> None:
        """Replace the owned rows by `cells` (shape `owned_shape`) and publish the boundary rows to
        the neighbours' ghost rows. Collective."""
        if tuple(cells.shape) != self.owned_shape:
            from .api import RangeError
            raise RangeError("The target buffer has not the same size as the slab")
        self.slab.copy_from_host(cells)
        self.slab.exchange_halos()

    def load_chunks(self, generate, chunk_rows: int = 256) -> None:
        """`load()` for slabs that should not be staged on the host at once: `generate(lo, hi)` returns
        the cells of the GLOBAL rows [lo, hi), which are uploaded chunk by chunk. Collective."""
        for lo in range(self.row_lo, self.row_hi, chunk_rows):
            hi = min(lo + chunk_rows, self.row_hi)
            self.slab.copy_rows_from_host(lo - self.row_lo, generate(lo, hi))
        self.slab.exchange_halos()

    def load_from(self, other: "ShardedStencilUpdate") -> None:
        """`load()` from the slab of another updater over the same grid and cell type (e.g. the
        pseudo-transient and the thermal update of mantle convection): the owned rows move device to
        device, then the boundary rows are published. Collective."""
        if (other.grid_rows, other.grid_cols, other.row_lo, other.row_hi) != \
                (self.grid_rows, self.grid_cols, self.row_lo, self.row_hi):
            from .api import RangeError
            raise RangeError("The source slab has not the same rows and columns")
        self.slab.copy_from_slab(other.slab)
        self.slab.exchange_halos()

    def to_numpy(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.owned_shape, dtype=self.dtype)
        self.slab.copy_to_host(out)
        return out

    def max_abs(self, extents) -> list[float]:
        """Max-norms of single fields over the WHOLE grid: [(field, rows, cols), ...] ->
        max |cell.field| over the first rows x cols cells (global coordinates). Every slab reduces its
        own rows on its GPU; the per-slab values are combined with an all-reduce(MAX) over the process
        group — the only collective of the application loops (reference
        examples/convection/convection.cpp:412-438 computes these on the host). Collective."""
        extents = list(extents)
        local = self.slab.max_abs(extents)
        if self.world == 1 or not extents:
            return local
        import torch
        t = torch.tensor(local, dtype=torch.float64, device=self._collective_device())
        self._comm.all_reduce(t, op=self._comm.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    def field_to_numpy(self, field, out: np.ndarray | None = None) -> np.ndarray:
        """ONE field of the owned rows (single-plane download)."""
        if out is None:
            out = np.empty(self.owned_shape, dtype=_native.field_dtype(self.workload, field))
        self.slab.field_rows_to_host(field, 0, out)
        return out

    # -- the reference's StencilUpdate surface ----------------------------------------------------------
    def get_params(self) -> Params:
        return self.params

    def _native_params(self):
        p = self.params
        tf = _as_param_struct(self.workload, p.transition_function)
        native = _native.UpdateParams()
        native.transition_function = C.addressof(tf)
        native.transition_function_bytes = C.sizeof(tf)
        halo = None
        if p.halo_value is not None:
            halo = np.zeros((), dtype=self.dtype)
            halo[()] = p.halo_value
            native.halo_value = halo.ctypes.data
            native.halo_value_bytes = halo.nbytes
        native.iteration_offset = int(p.iteration_offset)
        native.n_iterations = int(p.n_iterations)
        native.blocking = int(bool(p.blocking))
        self._keepalive = (tf, halo)
        return native

    def __call__(self) -> "ShardedStencilUpdate":
        """Advance the slab by `params.n_iterations` iterations starting at `params.iteration_offset`
        (asynchronous unless `params.blocking` — or unless plane pass-through is active, whose
        verification ends every call with one all-reduce). Collective."""
        if not self.speculating:
            self.slab.update(self._native_params())
            return self
        # Speculative plane pass-through: the sweeps leave fields the transition function never
        # changes in place and verify that in every pass. If ANY slab saw such a field change, all
        # slabs return to the saved generation and repeat the call without trusting that field.
        self.slab.backup()
        while True:
            self.slab.update(self._native_params())
            still_useful = int(self.slab.info().passthrough_planes) != 0
            violated, anyone_useful = self._combine_or(self.slab.take_violations(), still_useful)
            if violated == 0:
                # once no slab has a plane left to pass through, the protocol (backup, all-reduce)
                # is dropped for good — on every rank at the same call, the decision is collective
                self.speculating = anyone_useful
                return self
            self.n_speculation_redos += 1
            self.slab.drop_passthrough(violated)
            self.slab.restore()

    def _collective_device(self):
        """Where all-reduce operands live: THIS slab's GPU under NCCL (never torch's current device,
        which a caller may not have set — two ranks on cuda:0 hang NCCL), host memory under gloo."""
        import torch
        if self._comm.get_backend() == "nccl":
            return torch.device("cuda", self.device)
        return torch.device("cpu")

    def _combine_or(self, mask: int, flag: bool) -> tuple[int, bool]:
        """(bitwise OR of `mask`, logical OR of `flag`) over all ranks."""
        if self.world == 1:
            return mask, flag
        import torch
        device = self._collective_device()
        # NCCL has no bitwise OR: one 0/1 entry per plane (and one for the flag), combined with MAX
        bits = torch.tensor([(mask >> i) & 1 for i in range(32)] + [int(flag)], dtype=torch.int32,
                            device=device)
        self._comm.all_reduce(bits, op=self._comm.ReduceOp.MAX)
        values = bits.cpu().tolist()
        return sum(int(b) << i for i, b in enumerate(values[:32])), bool(values[32])

    def synchronize(self) -> None:
        self.slab.synchronize()

    def get_n_launches(self) -> int:
        return int(self.slab.info().n_launches)

    def info(self) -> SlabInfo:
        return self.slab.info()

    def close(self) -> None:
        """Release the slab. Collective when the grid has several slabs: every slab first stops
        touching its neighbours' memory (barrier, detach, barrier), only then is any of them freed."""
        if self.slab is None:
            return
        self.slab.synchronize()
        detach = getattr(self.slab, "detach", None)
        if self.world > 1:
            self._comm.barrier()
        if detach is not None:
            detach()
        if self.world > 1:
            self._comm.barrier()
        self.slab.close()
        self.slab = None
        if getattr(self, "_nccl", None):
            _native.runtime_lib().stst_nccl_comm_destroy(self._nccl)
            self._nccl = None
