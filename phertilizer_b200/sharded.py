"""Cell-sharded observation store: one process per GPU, cells partitioned contiguously, all SNVs on every rank.

SURVEY.md section 8e: per-SNV tables (K1) are sums over cells, so every rank computes the partial table of its block of
cells (`ph_snv_partial`) and ONE sum all-reduce of the packed `[S0 | S1 | Nobs]` float64 table (NCCL over NVLink)
gives every rank the full table; the K1a decision (`ph_finish_linear` / `ph_finish_branching`) then runs redundantly
on every rank.  Per-cell tables (K2) need no exchange.  Block sums all-reduce a handful of scalars.

`torch.distributed` is plumbing only.  The `backend` indirection exists so the partition / exchange logic can be tested
with the gloo backend on a CPU-only box (tests/test_sharded_gloo.py) with a CPU stand-in for the two kernels.
"""
from __future__ import annotations

import numpy as np


def cell_block(n_cells: int, rank: int, world: int):
    """Contiguous block [lo, hi) of cells owned by `rank`."""
    return (n_cells * rank) // world, (n_cells * (rank + 1)) // world


class CudaBackend:
    """The product backend: device tensors + the C-ABI kernels."""

    def __init__(self, store, device):
        import torch

        self.torch = torch
        self.store = store
        self.device = device

    def empty(self, n, dtype):
        return self.torch.empty(n, dtype=dtype, device=self.device)

    def to_device(self, arr):
        return self.torch.from_numpy(np.ascontiguousarray(arr)).to(self.device)

    def stream(self):
        return self.torch.cuda.current_stream().cuda_stream

    def snv_partial(self, label_t, C, packed_t, list_t=None):
        self.store.snv_partial_dev(label_t, C, packed_t, list_t, stream=self.stream())

    def finish_linear(self, packed_t, k, lenA, lab_t, nobs_t):
        self.store.finish_linear_dev(packed_t, k, lenA, lab_t, nobs_t, stream=self.stream())

    def finish_branching(self, packed_t, k, lab_t, nobs_t):
        self.store.finish_branching_dev(packed_t, k, lab_t, nobs_t, stream=self.stream())


class ShardedStore:
    def __init__(self, backend, n_cells, n_snvs, lo, hi, group=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.b = backend
        self.n_cells, self.n_snvs = int(n_cells), int(n_snvs)
        self.lo, self.hi = int(lo), int(hi)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    @classmethod
    def from_coo(cls, n_cells, n_snvs, cell, snv, var, total, rank, world, device, alpha=0.001, max_copies=5, group=None):
        """Every rank passes the observations it has (the full COO or any superset of its block); rows outside
        [lo, hi) are dropped and the rest is re-indexed to local cell ids."""
        import torch

        from .store import ObservationStore

        lo, hi = cell_block(n_cells, rank, world)
        if isinstance(cell, np.ndarray):
            keep = (cell >= lo) & (cell < hi)
            cell, snv, var, total = cell[keep] - lo, snv[keep], var[keep], total[keep]
        else:
            keep = (cell >= lo) & (cell < hi)
            cell, snv, var, total = (cell[keep] - lo).contiguous(), snv[keep].contiguous(), var[keep].contiguous(), total[keep].contiguous()
        st = ObservationStore.from_coo(hi - lo, n_snvs, cell, snv, var, total, alpha=alpha, max_copies=max_copies,
                                       device=device if isinstance(device, int) else device.index)
        dev = torch.device("cuda", device) if isinstance(device, int) else device
        return cls(CudaBackend(st, dev), n_cells, n_snvs, lo, hi, group)

    # ------------------------------------------------------------------ helpers
    def local_labels(self, cell_label_global: np.ndarray):
        """Slice a global uint8 label vector to this rank's block, padded for the kernels' 4-byte staging reads."""
        loc = np.zeros(self.hi - self.lo + 16, np.uint8)
        loc[: self.hi - self.lo] = cell_label_global[self.lo:self.hi]
        return self.b.to_device(loc)

    def _allreduce(self, t):
        if self.world > 1:
            self.dist.all_reduce(t, group=self.group)
        return t

    # ------------------------------------------------------------------ K1 family on device tensors
    def assign_linear_dev(self, label_t, lenA, packed_t, lab_out_t, nobs_out_t, list_t=None):
        """label_t: this rank's local labels; lenA: GLOBAL |cellsA|; packed_t: float64 [3*k] scratch."""
        k = self.n_snvs if list_t is None else int(list_t.numel())
        self.b.snv_partial(label_t, 1, packed_t, list_t)
        self._allreduce(packed_t)
        self.b.finish_linear(packed_t, k, lenA, lab_out_t, nobs_out_t)

    def assign_branching_dev(self, label_t, packed_t, lab_out_t, nobs_out_t, list_t=None):
        k = self.n_snvs if list_t is None else int(list_t.numel())
        self.b.snv_partial(label_t, 2, packed_t, list_t)
        self._allreduce(packed_t)
        self.b.finish_branching(packed_t, k, lab_out_t, nobs_out_t)

    def snv_table_dev(self, label_t, C, packed_t, list_t=None):
        """Full [S0 | S1 | Nobs] table on every rank (float64, 3*k*C)."""
        self.b.snv_partial(label_t, C, packed_t, list_t)
        return self._allreduce(packed_t)

    # ------------------------------------------------------------------ host-array conveniences (same results on every rank)
    def assign_linear(self, cell_label_global, lenA, snv_list=None):
        t = self.torch
        lst = None if snv_list is None else self.b.to_device(np.asarray(snv_list, np.int32))
        k = self.n_snvs if snv_list is None else len(snv_list)
        packed = self.b.empty(3 * k, t.float64)
        lab = self.b.empty(k + 16, t.uint8)
        nobs = self.b.empty(k, t.int32)
        self.assign_linear_dev(self.local_labels(cell_label_global), lenA, packed, lab, nobs, lst)
        return lab[:k].cpu().numpy(), nobs.cpu().numpy()

    def assign_branching(self, cell_label_global, snv_list=None):
        t = self.torch
        lst = None if snv_list is None else self.b.to_device(np.asarray(snv_list, np.int32))
        k = self.n_snvs if snv_list is None else len(snv_list)
        packed = self.b.empty(6 * k, t.float64)
        lab = self.b.empty(k + 16, t.uint8)
        nobs = self.b.empty(2 * k, t.int32)
        self.assign_branching_dev(self.local_labels(cell_label_global), packed, lab, nobs, lst)
        return lab[:k].cpu().numpy(), nobs.cpu().numpy().reshape(k, 2)

    def snv_table(self, cell_label_global, C, snv_list=None):
        t = self.torch
        lst = None if snv_list is None else self.b.to_device(np.asarray(snv_list, np.int32))
        k = self.n_snvs if snv_list is None else len(snv_list)
        packed = self.b.empty(3 * k * C, t.float64)
        self.snv_table_dev(self.local_labels(cell_label_global), C, packed, lst)
        p = packed.cpu().numpy().reshape(3, k, C)
        return {"S0": p[0], "S1": p[1], "Nobs": p[2].astype(np.int32)}
