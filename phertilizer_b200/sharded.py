"""Cell-sharded observation store: one process per GPU, cells partitioned contiguously, all SNVs on every rank.

SURVEY.md section 8e: per-SNV tables (K1) are sums over cells.  Two transports combine the ranks' partial tables:

* `exchange="peer"` (default when it can be set up): the pass kernel stores every SNV's integer (cluster, code)
  histogram straight into the NVLink peer window of the rank owning the SNV, the owner decides and pushes label +
  counts to every rank (csrc/ph_peer.cu).  No NCCL on the data path; bit-identical to the single-GPU kernels.
* `exchange="nccl"`: every rank computes the partial `[S0 | S1 | Nobs]` float64 table (`ph_snv_partial`), ONE NCCL sum
  all-reduce, then the K1a decision (`ph_finish_linear` / `ph_finish_branching`) redundantly on every rank.

Per-cell tables (K2) need no exchange.  Block sums all-reduce a handful of scalars.

`torch.distributed` is plumbing only.  The `backend` indirection exists so the partition / exchange logic can be tested
with the gloo backend on a CPU-only box (tests/test_sharded_gloo.py) with a CPU stand-in for the two kernels.
"""
from __future__ import annotations

import numpy as np

from ._lib import PhError as _PhError


def cell_block(n_cells: int, rank: int, world: int):
    """Contiguous block [lo, hi) of cells owned by `rank`."""
    return (n_cells * rank) // world, (n_cells * (rank + 1)) // world


def global_pairs(var, total, group=None):
    """One (var,total) code table for all ranks: the local pair counts are gathered and merged, then ordered by global
    frequency.  The peer-memory exchange sums integer histograms indexed by code, so the codes must mean the same
    pair everywhere."""
    import torch.distributed as dist

    from .store import order_pairs, pair_counts

    pv, pt, cnt = pair_counts(var, total)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return order_pairs(pv, pt, cnt)
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, (pv.tolist(), pt.tolist(), cnt.tolist()), group=group)
    merged = {}
    for a, b, c in parts:
        for v, t, k in zip(a, b, c):
            merged[(v, t)] = merged.get((v, t), 0) + k
    keys = list(merged)
    return order_pairs(np.array([k[0] for k in keys], np.int64), np.array([k[1] for k in keys], np.int64),
                       np.array([merged[k] for k in keys], np.int64))


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


class PeerExchange:
    """NVLink peer-memory windows of the fused sharded K1 (include/phertilizer_b200.h, ph_peer_*).  torch.distributed
    only carries the 64-byte IPC handle of every rank's window once, at set-up."""

    def __init__(self, store, rank, world, group=None, total_snvs=0):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _lib

        L = _lib.lib()
        self._lib, self._L = _lib, L
        self._p = C.c_void_p()
        # Set-up is collective; a rank whose local step fails (unsupported shape, IPC not permitted in this container)
        # must still take part in every collective below, or the others would wait forever.  Everyone learns the outcome.
        nb = L.ph_peer_handle_bytes()
        blob, err = None, None
        try:
            _lib.check(L.ph_peer_create(store._h, int(rank), int(world), int(total_snvs), C.byref(self._p)))
            mine = C.create_string_buffer(nb)
            _lib.check(L.ph_peer_export(self._p, mine))
            blob = bytes(mine.raw)
        except _lib.PhError as e:
            err = str(e)
        blobs = [None] * world
        dist.all_gather_object(blobs, blob, group=group)
        if err is None and all(b is not None for b in blobs):
            try:
                _lib.check(L.ph_peer_connect(self._p, C.create_string_buffer(b"".join(blobs), nb * world)))
            except _lib.PhError as e:
                err = str(e)
        elif err is None:
            err = "a peer could not create its window"
        errs = [None] * world
        dist.all_gather_object(errs, err, group=group)  # also the barrier: every window is mapped before any kernel stores
        torch.cuda.synchronize()
        bad = [f"rank {r}: {e}" for r, e in enumerate(errs) if e is not None]
        if bad:
            self.close()
            raise _lib.PhError("peer-memory exchange unavailable (" + "; ".join(bad) + ")")

    def assign_linear(self, label_t, lenA, lab_out_t, nobs_out_t, list_t=None, k=0, stream=None):
        from .store import _ptr

        self._lib.check(self._L.ph_peer_assign_linear(self._p, _ptr(label_t), float(lenA), _ptr(list_t), int(k),
                                                      _ptr(lab_out_t), _ptr(nobs_out_t), stream))

    def assign_branching(self, label_t, lab_out_t, nobs_out_t, list_t=None, k=0, stream=None):
        from .store import _ptr

        self._lib.check(self._L.ph_peer_assign_branching(self._p, _ptr(label_t), _ptr(list_t), int(k), _ptr(lab_out_t),
                                                         _ptr(nobs_out_t), stream))

    def assign_linear_snv(self, label_t, lenA, snv_offset, lab_out_t, nobs_out_t, stream=None):
        from .store import _ptr

        self._lib.check(self._L.ph_peer_assign_linear_snv(self._p, _ptr(label_t), float(lenA), int(snv_offset),
                                                          _ptr(lab_out_t), _ptr(nobs_out_t), stream))

    def assign_branching_snv(self, label_t, snv_offset, lab_out_t, nobs_out_t, stream=None):
        from .store import _ptr

        self._lib.check(self._L.ph_peer_assign_branching_snv(self._p, _ptr(label_t), int(snv_offset), _ptr(lab_out_t),
                                                             _ptr(nobs_out_t), stream))

    def timed_out(self, stream=None) -> bool:
        if self._p is None:
            return True   # closed after a timeout
        return bool(self._L.ph_peer_timed_out(self._p, stream))

    def debug(self) -> dict:
        """What the first wait of an SNV-sharded step that gave up was waiting for (ph_peer_debug)."""
        import ctypes as C

        buf = (C.c_uint32 * 24)()
        if self._p is None:
            return {}
        self._lib.check(self._L.ph_peer_debug(self._p, buf))
        v = [int(x) for x in buf]
        names = ("timed_out", "kind", "source_rank", "slot", "step_wanted", "step_seen", "payload_seen", "own_step_counter")
        out = dict(zip(names, v[:8]))
        out["per_source(announced, arrived)"] = [(v[8 + 2 * r], v[9 + 2 * r]) for r in range(8)]
        return out

    def close(self):
        if self._p is not None:
            self._L.ph_peer_destroy(self._p)
            self._p = None


class ShardedStore:
    def __init__(self, backend, n_cells, n_snvs, lo, hi, group=None, exchange="auto"):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.b = backend
        self.n_cells, self.n_snvs = int(n_cells), int(n_snvs)
        self.lo, self.hi = int(lo), int(hi)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.peer = None
        self.exchange = "nccl"
        if exchange in ("auto", "peer") and self.world > 1 and isinstance(backend, CudaBackend):
            # PeerExchange raises on EVERY rank if any rank failed, so all ranks take the same branch
            try:
                self.peer = PeerExchange(backend.store, dist.get_rank(group), self.world, group)
                self.exchange = "peer"
            except Exception as e:  # PhError (unsupported shape, IPC not permitted) -> the NCCL transport
                if exchange == "peer":
                    raise
                print(f"[phertilizer_b200] {e}; using the NCCL all-reduce")

    @classmethod
    def from_coo(cls, n_cells, n_snvs, cell, snv, var, total, rank, world, device, alpha=0.001, max_copies=5, group=None,
                 exchange="auto"):
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
                                       device=device if isinstance(device, int) else device.index,
                                       pairs=global_pairs(var, total, group))
        dev = torch.device("cuda", device) if isinstance(device, int) else device
        return cls(CudaBackend(st, dev), n_cells, n_snvs, lo, hi, group, exchange)

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

    def _check_peer(self):
        """After the results have been copied back: a spin-wait that gave up (~3 s: a lost or stalled rank) means the
        window held stale data when it was summed / unpacked.  The outputs must not be used and the step counters of the
        ranks are no longer in step: the exchange is torn down.  Callers of the `*_dev` methods poll
        `self.peer.timed_out()` themselves (bench.py does, once per run)."""
        if self.peer is not None and self.peer.timed_out(self.b.stream()):
            self.peer.close()
            self.peer = None
            raise _PhError("peer-memory exchange timed out waiting for a rank (about 3 s): results discarded, exchange closed")

    # ------------------------------------------------------------------ K1 family on device tensors
    def assign_linear_dev(self, label_t, lenA, packed_t, lab_out_t, nobs_out_t, list_t=None):
        """label_t: this rank's local labels; lenA: GLOBAL |cellsA|; packed_t: float64 [3*k] scratch.
        Peer transport: the caller must poll `self.peer.timed_out()` before trusting the outputs (see _check_peer)."""
        k = self.n_snvs if list_t is None else int(list_t.numel())
        if self.peer is not None:
            return self.peer.assign_linear(label_t, lenA, lab_out_t, nobs_out_t, list_t, k, stream=self.b.stream())
        self.b.snv_partial(label_t, 1, packed_t, list_t)
        self._allreduce(packed_t)
        self.b.finish_linear(packed_t, k, lenA, lab_out_t, nobs_out_t)

    def assign_branching_dev(self, label_t, packed_t, lab_out_t, nobs_out_t, list_t=None):
        k = self.n_snvs if list_t is None else int(list_t.numel())
        if self.peer is not None:
            return self.peer.assign_branching(label_t, lab_out_t, nobs_out_t, list_t, k, stream=self.b.stream())
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
        out = lab[:k].cpu().numpy(), nobs.cpu().numpy()
        self._check_peer()
        return out

    def assign_branching(self, cell_label_global, snv_list=None):
        t = self.torch
        lst = None if snv_list is None else self.b.to_device(np.asarray(snv_list, np.int32))
        k = self.n_snvs if snv_list is None else len(snv_list)
        packed = self.b.empty(6 * k, t.float64)
        lab = self.b.empty(k + 16, t.uint8)
        nobs = self.b.empty(2 * k, t.int32)
        self.assign_branching_dev(self.local_labels(cell_label_global), packed, lab, nobs, lst)
        out = lab[:k].cpu().numpy(), nobs.cpu().numpy().reshape(k, 2)
        self._check_peer()
        return out

    def snv_table(self, cell_label_global, C, snv_list=None):
        t = self.torch
        lst = None if snv_list is None else self.b.to_device(np.asarray(snv_list, np.int32))
        k = self.n_snvs if snv_list is None else len(snv_list)
        packed = self.b.empty(3 * k * C, t.float64)
        self.snv_table_dev(self.local_labels(cell_label_global), C, packed, lst)
        p = packed.cpu().numpy().reshape(3, k, C)
        return {"S0": p[0], "S1": p[1], "Nobs": p[2].astype(np.int32)}

    # ------------------------------------------------------------------ the rest of the store API under cell sharding
    # (SURVEY 8e): K2 is local to the rank owning the cells, K3 and the tree log-likelihood all-reduce a handful of
    # scalars.  Only CudaBackend: these go through the rank's ObservationStore host-buffer calls.
    def cell_table(self, snv_label, Q, want=("S0", "S1", "Nobs", "Nvar")):
        """Per-cell table of ALL cells on every rank: each rank computes its block, the blocks are all-gathered."""
        t, d = self.torch, self.dist
        loc = self.b.store.cell_table(snv_label, Q, want=want)
        out = {}
        for k, a in loc.items():
            parts = [None] * self.world
            if self.world > 1:
                d.all_gather_object(parts, a, group=self.group)
            else:
                parts = [a]
            out[k] = np.concatenate(parts, axis=0)
        return out

    def _sum(self, arr):
        if self.world == 1:
            return arr
        t = self.torch.from_numpy(np.ascontiguousarray(arr).copy())
        if self.dist.get_backend(self.group) != "gloo":   # NCCL reduces device tensors
            t = t.to(self.b.device)
        self.dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()

    def block_sums(self, cell_label_global, C, snv_label, Q):
        """K3 over all ranks: (sum0, sum1, nobs, nvar) [Q, C]; counts exact, sums = sum of per-rank partials."""
        lab = np.zeros(self.hi - self.lo + 16, np.uint8)[: self.hi - self.lo]
        lab[:] = np.asarray(cell_label_global, np.uint8)[self.lo:self.hi]
        s0, s1, no, nv = self.b.store.block_sums(lab, C, snv_label, Q)
        return self._sum(s0), self._sum(s1), self._sum(no), self._sum(nv)

    def tree_loglik(self, cell_node_global, snv_node, present):
        """Variant log-likelihood per tree node (clonal_tree.py:932-944) over all ranks."""
        node = np.full(self.hi - self.lo + 16, 255, np.uint8)[: self.hi - self.lo]
        node[:] = np.asarray(cell_node_global, np.uint8)[self.lo:self.hi]
        return self._sum(self.b.store.tree_loglik(node, snv_node, present))


def snv_block(n_snvs: int, rank: int, world: int):
    """Contiguous block [lo, hi) of SNVs owned by `rank`."""
    return (n_snvs * rank) // world, (n_snvs * (rank + 1)) // world


class SnvShardedStore:
    """SNVs partitioned across ranks, every rank holds all cells of its SNV block.

    Per-SNV tables (K1) are independent across SNVs, so this sharding needs NO exchange of partial tables for the SNV
    reassignment: each rank runs the single-GPU fused kernel on its SNVs and the kernel epilogue stores label + counts
    into every rank's output window over NVLink peer memory (`ph_peer_assign_*_snv`).  What would need an exchange under
    this sharding is K2 (per-cell tables: a sum all-reduce of `n x Q` integer counts), the mirror image of the
    cell-sharded layout.  Results are those of the single-GPU kernel bit for bit (it IS that kernel, on a slice)."""

    def __init__(self, store, n_cells, n_snvs, lo, hi, device, group=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.store, self.device, self.group = store, device, group
        self.n_cells, self.n_snvs, self.lo, self.hi = int(n_cells), int(n_snvs), int(lo), int(hi)
        self.world = dist.get_world_size(group)
        self.peer = PeerExchange(store, dist.get_rank(group), self.world, group, total_snvs=self.n_snvs)

    @classmethod
    def from_coo(cls, n_cells, n_snvs, cell, snv, var, total, rank, world, device, alpha=0.001, max_copies=5, group=None):
        import torch

        from .store import ObservationStore

        lo, hi = snv_block(n_snvs, rank, world)
        keep = (snv >= lo) & (snv < hi)
        if isinstance(cell, np.ndarray):
            cell, snv, var, total = cell[keep], snv[keep] - lo, var[keep], total[keep]
        else:
            cell, snv, var, total = cell[keep].contiguous(), (snv[keep] - lo).contiguous(), var[keep].contiguous(), total[keep].contiguous()
        dev_index = device if isinstance(device, int) else device.index
        # the SNV-sharded step runs on the batched-4 copy of the stream like the single-GPU pass: its results travel as
        # self-describing words (SNV id + label + counts), so the shape order of a rank's lines needs no set-up exchange
        st = ObservationStore.from_coo(n_cells, hi - lo, cell, snv, var, total, alpha=alpha, max_copies=max_copies,
                                       device=dev_index, pairs=global_pairs(var, total, group))
        dev = torch.device("cuda", dev_index)
        try:
            return cls(st, n_cells, n_snvs, lo, hi, dev, group)
        except Exception:
            st.close()
            raise

    def _stream(self):
        return self.torch.cuda.current_stream().cuda_stream

    def assign_linear_dev(self, label_t, lenA, lab_out_t, nobs_out_t):
        """label_t: labels of ALL cells (uint8, device); outputs cover all n_snvs SNVs on every rank."""
        self.peer.assign_linear_snv(label_t, lenA, self.lo, lab_out_t, nobs_out_t, stream=self._stream())

    def assign_branching_dev(self, label_t, lab_out_t, nobs_out_t):
        self.peer.assign_branching_snv(label_t, self.lo, lab_out_t, nobs_out_t, stream=self._stream())

    def _labels_dev(self, cell_label):
        buf = np.zeros(self.n_cells + 16, np.uint8)
        buf[: self.n_cells] = cell_label
        return self.torch.from_numpy(buf).to(self.device)

    def assign_linear(self, cell_label, lenA):
        t = self.torch
        lab = t.empty(self.n_snvs + 16, dtype=t.uint8, device=self.device)
        nobs = t.empty(self.n_snvs, dtype=t.int32, device=self.device)
        self.assign_linear_dev(self._labels_dev(cell_label), lenA, lab, nobs)
        out = lab[: self.n_snvs].cpu().numpy(), nobs.cpu().numpy()
        self._check_peer()
        return out

    def assign_branching(self, cell_label):
        t = self.torch
        lab = t.empty(self.n_snvs + 16, dtype=t.uint8, device=self.device)
        nobs = t.empty(2 * self.n_snvs, dtype=t.int32, device=self.device)
        self.assign_branching_dev(self._labels_dev(cell_label), lab, nobs)
        out = lab[: self.n_snvs].cpu().numpy(), nobs.cpu().numpy().reshape(self.n_snvs, 2)
        self._check_peer()
        return out

    def _check_peer(self):
        """See ShardedStore._check_peer: a timed-out wait invalidates the step's outputs on this rank."""
        if self.peer.timed_out(self._stream()):
            raise _PhError(f"peer-memory exchange timed out waiting for a rank: results discarded ({self.peer.debug()})")

    def close(self):
        self.peer.close()
        self.store.close()
