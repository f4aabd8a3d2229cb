"""``CellShardedStore`` -- the ``ObservationStore`` API over cells sharded across GPUs.

One process per GPU; rank r holds the observations of a contiguous block of cells (all SNVs) in its own
``ObservationStore``.  Every method takes GLOBAL label vectors / index lists and returns the GLOBAL result on every rank,
so the tree-growing control flow (``Phertilizer``, ``Linear_split``, ``Branching_split``, ``ClonalTree``) runs unchanged and
redundantly on all ranks, on identical inputs -- the same RNG streams, the same decisions.

What makes the results independent of the number of GPUs:

* per-CELL quantities (K2 tables, a cell's tree log-likelihood, reassign_cell scores) are computed wholly by the rank that
  owns the cell; the blocks are assembled by position -- K2 tables with one NCCL all-gather of a device buffer each rank's
  pass writes its region of, everything else (and everything over gloo) by a sum all-reduce into a zero array (x + 0 is exact);
* the scalars that decide which split / tree wins -- block sums (compute_norm_likelihood, check_obs) and per-node tree
  log-likelihoods -- are fixed-order reductions of those assembled per-cell tables, run by the SAME kernels one GPU runs
  (``ph_line_reduce`` / ``ph_node_reduce``): bit-identical for any world size;
* per-SNV integer tables (Nobs, Nvar) are sums of integers;
* the fused SNV reassignment exchanges integer (cluster, code) histograms over NVLink peer memory (``sharded.ShardedStore``,
  exact), or falls back to one fp64 all-reduce (may differ on exact ties only) beyond its limits;
* per-SNV float sums over all cells (``snv_table`` S0/S1, ``snv_node_table``) are all-reduced partial sums: equal to one
  GPU's to rounding (1e-15 relative); columns that are structurally equal on every rank stay exactly equal.

torch.distributed is plumbing (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def _padded(a, n, fill=0):
    buf = np.full(n + 16, fill, dtype=np.uint8)
    buf[:n] = a
    return buf[:n]


class CellShardedStore:
    def __init__(self, local, n_cells, n_snvs, lo, hi, group=None, k1=None, dense=None):
        """local: this rank's store over cells [lo, hi); k1: the ``sharded.ShardedStore`` that runs the fused SNV
        reassignment across ranks (None: world 1); dense: a store-like object whose ``gauss_node_ll`` / ``set_embedding``
        see ALL cells (the embedding is replicated: K4 is n x D work and needs every cell's row for the node moments)."""
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.local, self.k1, self.dense = local, k1, dense
        self.n_cells, self.n_snvs = int(n_cells), int(n_snvs)
        self.lo, self.hi = int(lo), int(hi)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._device_reduce = self.world > 1 and dist.get_backend(group) != "gloo"
        self.pair_var = getattr(local, "pair_var", None)
        self.pair_total = getattr(local, "pair_total", None)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_coo(cls, n_cells, n_snvs, cell, snv, var, total, rank, world, device, alpha=0.001, max_copies=5, group=None,
                 exchange="auto"):
        from .sharded import ShardedStore
        from .store import ObservationStore

        k1 = ShardedStore.from_coo(n_cells, n_snvs, cell, snv, var, total, rank, world, device, alpha=alpha,
                                   max_copies=max_copies, group=group, exchange=exchange)
        dev = device if isinstance(device, int) else device.index
        # a handle without observations over all n cells: scratch + stream for the replicated dense kernels (K4)
        dense = ObservationStore.from_coo(n_cells, 1, np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32),
                                          np.zeros(0, np.int32), alpha=alpha, max_copies=max_copies, device=dev,
                                          pairs=(k1.b.store.pair_var, k1.b.store.pair_total))
        return cls(k1.b.store, n_cells, n_snvs, k1.lo, k1.hi, group, k1, dense)

    def close(self):
        if self.k1 is not None and self.k1.peer is not None:
            self.k1.peer.close()
        for s in (self.local, self.dense):
            if s is not None and hasattr(s, "close"):
                s.close()

    def info(self):
        inf = dict(self.local.info()) if hasattr(self.local, "info") else {}
        inf.update(world=self.world, rank=self.rank, cells_lo=self.lo, cells_hi=self.hi,
                   k1_exchange=getattr(self.k1, "exchange", None))
        return inf

    # ------------------------------------------------------------------ label helpers (global vectors)
    def cell_labels(self, *groups):
        from .store import _labels

        return _labels(self.n_cells, groups)

    def snv_labels(self, *groups):
        from .store import _labels

        return _labels(self.n_snvs, groups)

    def _loc(self, cell_vec, fill=0):
        """This rank's slice of a global per-cell uint8 vector, padded for the kernels' 4-byte staging reads."""
        return _padded(np.asarray(cell_vec, np.uint8)[self.lo:self.hi], self.hi - self.lo, fill)

    # ------------------------------------------------------------------ collectives on host arrays
    def _sum(self, arr):
        if self.world == 1:
            return arr
        t = self.torch.from_numpy(np.ascontiguousarray(arr).copy())
        if self._device_reduce:
            dev = self.torch.device("cuda", self.torch.cuda.current_device())
            t = t.to(dev)
        self.dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()

    def _assemble(self, block):
        """Rows of all ranks' blocks in cell order: zero array + own block, summed (exact: x + 0)."""
        if self.world == 1:
            return block
        full = np.zeros((self.n_cells,) + block.shape[1:], block.dtype)
        full[self.lo:self.hi] = block
        return self._sum(full)

    def _combine_tables(self, tables, rows_local):
        """ONE collective for a dict of per-line tables: they travel side by side in a float64 buffer (the integer counts
        are far below 2^53: exact) and come back with their own dtypes.  rows_local: the tables are this rank's block of
        cells (assembled by position) -- else per-SNV partial tables of all ranks (summed)."""
        if self.world == 1 or not tables:
            return tables
        keys = list(tables)
        cols = [np.asarray(tables[k]).reshape(len(tables[k]), -1) for k in keys]
        width = [c.shape[1] for c in cols]
        n_rows = self.n_cells if rows_local else cols[0].shape[0]
        buf = np.zeros((n_rows, sum(width)), np.float64)
        view = buf[self.lo:self.hi] if rows_local else buf
        at = 0
        for c, w in zip(cols, width):
            view[:, at:at + w] = c
            at += w
        buf = self._sum(buf)
        out, at = {}, 0
        for k, c, w in zip(keys, cols, width):
            part = buf[:, at:at + w]
            out[k] = part.astype(c.dtype) if c.dtype.kind in "iu" else np.ascontiguousarray(part)
            at += w
        return out

    _SIZES = {"S0": (8, np.float64), "S1": (8, np.float64), "Nobs": (4, np.int32), "Nvar": (4, np.int32)}

    def _cell_table_gathered(self, snv_label, Q, want):
        """K2 tables of ALL cells with the assembly on the device: every rank's pass writes its block straight into its
        region of one device buffer, ONE all-gather (NCCL, in place) completes it, ONE device-to-host copy brings it home.
        Same values as `_combine_tables` (rows are computed by their owner either way), a fifth of the host work."""
        torch, dist = self.torch, self.dist
        from .sharded import cell_block

        names = [k for k in ("S0", "S1", "Nobs", "Nvar") if k in want]
        blocks = [cell_block(self.n_cells, r, self.world) for r in range(self.world)]
        bmax = max(hi - lo for lo, hi in blocks)
        offs, at = {}, 0
        for k in names:   # the float64 tables first: every sub-array starts on a multiple of 8 bytes
            offs[k] = at
            at += bmax * Q * self._SIZES[k][0]
        region = (at + 15) & ~15
        cache = self.__dict__.setdefault("_gather_cache", {})
        key = (Q, tuple(names))
        if key not in cache:
            dev = torch.device("cuda", torch.cuda.current_device())
            cache[key] = (torch.zeros((self.world, region), dtype=torch.uint8, device=dev),
                          torch.empty((self.world, region), dtype=torch.uint8).pin_memory())
        if "_lab_dev" not in self.__dict__:
            dev = torch.device("cuda", torch.cuda.current_device())
            self._lab_pin = torch.zeros(self.n_snvs + 16, dtype=torch.uint8).pin_memory()
            self._lab_dev = torch.zeros(self.n_snvs + 16, dtype=torch.uint8, device=dev)
        full, host = cache[key]
        self._lab_pin.numpy()[:self.n_snvs] = np.asarray(snv_label, np.uint8)[:self.n_snvs]
        self._lab_dev.copy_(self._lab_pin, non_blocking=True)
        mine = full[self.rank]
        n_loc = self.hi - self.lo
        outs = {}
        for k in names:
            size, _ = self._SIZES[k]
            t = mine[offs[k]:offs[k] + n_loc * Q * size]
            outs[k] = t.view(torch.float64 if size == 8 else torch.int32)
        self.local.cell_table_dev(self._lab_dev, Q, S0=outs.get("S0"), S1=outs.get("S1"), Nobs=outs.get("Nobs"),
                                  Nvar=outs.get("Nvar"))   # legacy default stream = torch's current stream: ordered with NCCL
        dist.all_gather_into_tensor(full.view(-1), mine, group=self.group)
        host.copy_(full, non_blocking=False)
        h = host.numpy()
        res = {}
        for k in names:
            size, dt = self._SIZES[k]
            parts = [h[r, offs[k]:offs[k] + (hi - lo) * Q * size].view(dt).reshape(hi - lo, Q) for r, (lo, hi) in enumerate(blocks)]
            res[k] = np.concatenate(parts, axis=0)
        return res

    # ------------------------------------------------------------------ K1 family (per SNV, over cell clusters)
    def snv_table(self, cell_label, C_, snv_list=None, want=("S0", "S1", "Nobs", "Nvar")):
        t = self.local.snv_table(self._loc(cell_label), C_, snv_list, want=want)
        return self._combine_tables(t, rows_local=False)

    def assign_linear(self, cell_label, lenA, snv_list=None):
        if self.k1 is None or self.world == 1:
            return self.local.assign_linear(self._loc(cell_label), lenA, snv_list)
        return self.k1.assign_linear(np.asarray(cell_label, np.uint8), lenA, snv_list)

    def assign_branching(self, cell_label, snv_list=None):
        if self.k1 is None or self.world == 1:
            return self.local.assign_branching(self._loc(cell_label), snv_list)
        return self.k1.assign_branching(np.asarray(cell_label, np.uint8), snv_list)

    # ------------------------------------------------------------------ K2 family (per cell: owned by one rank)
    def cell_table(self, snv_label, Q, cell_list=None, want=("S0", "S1", "Nobs", "Nvar")):
        if self._device_reduce and self.world > 1:
            out = self._cell_table_gathered(snv_label, Q, want)
        else:
            out = self._combine_tables(self.local.cell_table(snv_label, Q, None, want=want), rows_local=True)
        if cell_list is not None:
            idx = np.asarray(cell_list, np.int64)
            out = {k: v[idx] for k, v in out.items()}
        return out

    def cell_node_table(self, snv_node, member, cell_list=None):
        if cell_list is None:
            return self._assemble(self.local.cell_node_table(snv_node, member, None))
        idx = np.asarray(cell_list, np.int64)
        K = np.asarray(member).shape[0]
        out = np.zeros((len(idx), K), np.float64)
        mine = (idx >= self.lo) & (idx < self.hi)
        if mine.any():
            out[mine] = self.local.cell_node_table(snv_node, member, (idx[mine] - self.lo).astype(np.int32))
        return self._sum(out)

    # ------------------------------------------------------------------ K3 and tree log-likelihood (deciding scalars)
    def block_sums(self, cell_label, C_, snv_label, Q):
        cl = np.asarray(cell_label, np.uint8)
        if Q > 3 or self.world == 1:
            s0, s1, no, nv = self.local.block_sums(self._loc(cl), C_, snv_label, Q)
            return self._sum(s0), self._sum(s1), self._sum(no), self._sum(nv)
        full = self.cell_table(snv_label, Q)
        return self.local.line_reduce(full["S0"], full["S1"], full["Nobs"], full["Nvar"], _padded(cl, self.n_cells), Q, C_)

    def tree_loglik(self, cell_node, snv_node, present, per_cell=False):
        cn = np.asarray(cell_node, np.uint8)
        if self.world == 1:
            return self.local.tree_loglik(self._loc(cn, 255), snv_node, present, per_cell=per_cell)
        _, pc = self.local.tree_loglik(self._loc(cn, 255), snv_node, present, per_cell=True)
        pc = self._assemble(pc)
        per_node = self.local.node_reduce(pc, _padded(cn, self.n_cells, 255), np.asarray(present).shape[0])
        return (per_node, pc) if per_cell else per_node

    def snv_node_table(self, cell_node, member, snv_list=None):
        return self._sum(self.local.snv_node_table(self._loc(cell_node, 255), member, snv_list))

    # ------------------------------------------------------------------ K4 (replicated embedding)
    def set_embedding(self, X):
        self.dense.set_embedding(X)

    def gauss_node_ll(self, X, cell_node, K, per_cell=False):
        return self.dense.gauss_node_ll(X, _padded(np.asarray(cell_node, np.uint8), self.n_cells, 255), K, per_cell=per_cell)
