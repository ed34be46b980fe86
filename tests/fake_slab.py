"""TEST INFRASTRUCTURE — a host-memory stand-in for a GPU slab.

Implements the interface `stencilstream_b200.sharding.ShardedStencilUpdate` expects from a slab
(`NativeSlab`), with numpy arrays instead of HBM planes, `torch.distributed` point-to-point messages
instead of NVLink pushes and the CPU oracle instead of the sm_100a kernel. It exists so that the
partitioner's host logic — row partition, neighbour wiring, ghost depth k*n_sub*radius, pass
sequence with a shorter tail pass, collective call order — runs under the `gloo` backend with
world_size > 1 on machines without a GPU. It is never importable from the product package.
"""
from __future__ import annotations

import ctypes as C
import pickle
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist

from stencilstream_b200 import _native, workload_info


class HostSlab:
    def __init__(self, oracle, default_depth, *, workload, grid_rows, grid_cols, row_lo, row_hi,
                 device, fused_iterations, tile_rows, overlap):
        self.oracle, self.workload = oracle, workload
        self.dtype = _native.CELL_DTYPES[workload]
        self.grid_rows, self.grid_cols = grid_rows, grid_cols
        self.row_lo, self.row_hi = row_lo, row_hi
        info = workload_info(workload)
        self.k = int(fused_iterations) or int(default_depth)
        self.ghost = self.k * int(info.n_subiterations) * int(info.stencil_radius)
        self.neighbour = {}          # side -> (rank, row_lo, row_hi)
        self.ghost_rows = {0: None, 1: None}
        self.cells = np.zeros((row_hi - row_lo, grid_cols), dtype=self.dtype)
        self.n_launches = 0
        self.log = []

    # ---- interface of NativeSlab -----------------------------------------------------------------
    def info(self):
        return SimpleNamespace(fused_iterations=self.k, ghost_rows=self.ghost, row_lo=self.row_lo,
                               row_hi=self.row_hi, n_launches=self.n_launches,
                               passthrough_planes=1 if getattr(self, "speculation", False) else 0)

    def ipc_handle(self) -> bytes:
        return pickle.dumps(dist.get_rank()).ljust(64, b"\0")

    def attach_ipc(self, side, handle, row_lo, row_hi):
        rank = pickle.loads(handle.rstrip(b"\0"))
        assert (row_hi == self.row_lo) if side == 0 else (row_lo == self.row_hi), "not adjacent"
        assert row_hi - row_lo >= self.ghost
        self.neighbour[side] = (rank, row_lo, row_hi)

    def copy_from_host(self, cells):
        assert cells.shape == self.cells.shape
        self.cells = np.ascontiguousarray(cells, dtype=self.dtype).copy()
        self.log.append("upload")

    def copy_to_host(self, out):
        out[...] = self.cells

    def copy_from_slab(self, source):
        assert source.cells.shape == self.cells.shape and source.dtype == self.dtype
        self.cells = source.cells.copy()
        self.log.append("copy_from_slab")

    def max_abs(self, extents):
        out = []
        for field, rows, cols in extents:
            hi = min(int(rows), self.row_hi) - self.row_lo
            plane = self.cells[self.dtype.names[_native.field_index(self.workload, field)]] \
                if self.dtype.names else self.cells
            part = plane[:max(hi, 0), :int(cols)]
            out.append(float(np.abs(part.astype(np.float64)).max()) if part.size else float("-inf"))
        return out

    def field_rows_to_host(self, field, first_row, out):
        name = self.dtype.names[_native.field_index(self.workload, field)]
        out[...] = self.cells[name][first_row:first_row + out.shape[0]]

    def exchange_halos(self):
        self.log.append("exchange")
        self._exchange()

    def update(self, native):
        params = None
        if native.transition_function:
            params = _native.PARAM_TYPES[self.workload].from_address(native.transition_function)
        halo = None
        if native.halo_value:
            raw = C.string_at(native.halo_value, self.dtype.itemsize)
            halo = np.frombuffer(raw, dtype=self.dtype)[0]
        iteration, remaining = int(native.iteration_offset), int(native.n_iterations)
        while remaining > 0:
            n_gens = min(remaining, self.k)
            self._pass(params, halo, iteration, n_gens)
            iteration += n_gens
            remaining -= n_gens

    # plane pass-through: the host double copies whole cells; it takes part in the protocol (so that
    # ShardedStencilUpdate's backup / verify / repeat sequence runs under gloo) and can be told to
    # report a violation once, to exercise the repeat
    inject_violation_once = 0

    def enable_speculation(self, on=True):
        self.speculation = bool(on)
        return self.speculation

    def backup(self):
        self._saved = (self.cells.copy(), {k: (None if v is None else v.copy())
                                           for k, v in self.ghost_rows.items()})
        self.log.append("backup")

    def restore(self):
        self.cells = self._saved[0].copy()
        self.ghost_rows = {k: (None if v is None else v.copy()) for k, v in self._saved[1].items()}
        self.log.append("restore")

    def take_violations(self):
        mask, self.inject_violation_once = self.inject_violation_once, 0
        return mask

    def drop_passthrough(self, planes):
        self.log.append(("drop", planes))

    def synchronize(self):
        pass

    def close(self):
        pass

    # ---- internals -------------------------------------------------------------------------------------
    def _exchange(self):
        """Send my boundary rows, receive the neighbours' (what the GPU kernel pushes over NVLink)."""
        requests, inbox = [], {}
        for side, (rank, _, _) in self.neighbour.items():
            mine = self.cells[:self.ghost] if side == 0 else self.cells[-self.ghost:]
            out = torch.from_numpy(np.ascontiguousarray(mine).view(np.uint8).reshape(-1).copy())
            inbox[side] = torch.empty(self.ghost * self.grid_cols * self.dtype.itemsize,
                                      dtype=torch.uint8)
            requests.append(dist.isend(out, dst=rank))
            requests.append(dist.irecv(inbox[side], src=rank))
        for r in requests:
            r.wait()
        for side, buf in inbox.items():
            self.ghost_rows[side] = buf.numpy().view(self.dtype).reshape(self.ghost, self.grid_cols)

    def _pass(self, params, halo, iteration, n_gens):
        parts, row0 = [], self.row_lo
        if 0 in self.neighbour:
            parts.append(self.ghost_rows[0])
            row0 -= self.ghost
        parts.append(self.cells)
        if 1 in self.neighbour:
            parts.append(self.ghost_rows[1])
        window = np.ascontiguousarray(np.concatenate(parts, axis=0))
        out = self.oracle.run_window(self.workload, params, halo, window, row0, self.grid_rows,
                                     iteration, n_gens)
        first = self.row_lo - row0
        self.cells = np.ascontiguousarray(out[first:first + (self.row_hi - self.row_lo)])
        self.n_launches += 1
        self.log.append(("pass", iteration, n_gens))
        self._exchange()
