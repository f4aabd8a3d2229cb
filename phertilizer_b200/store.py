"""``ObservationStore`` -- the device-resident replacement of the reference's dense ``Data`` matrices.

The reference keeps ``var, total, like0, like1_marg`` as dense ``n x m`` numpy arrays (phertilizer/data.py:8-20,
built at phertilizer/phertilizer.py:295-306) and reads them through ``M[np.ix_(cells, muts)].<reduce>``.  This class
owns one ``ph_handle`` (include/phertilizer_b200.h) on one GPU and exposes those reductions as methods taking the
same index arrays.  All compute happens in the CUDA library; numpy here only marshals labels and results.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .like_lut import encode_pairs, like_tables


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a.data_ptr()))  # torch tensor


def _labels(n: int, groups) -> np.ndarray:
    """index arrays (cluster 1, cluster 2, ...) -> uint8 label vector, 0 = not selected."""
    lab = np.zeros(n + 16, dtype=np.uint8)[:n]
    for c, idx in enumerate(groups, start=1):
        if idx is not None and len(idx):
            lab[np.asarray(idx, dtype=np.int64)] = c
    return lab


def _padded_u8(n: int, fill: int = 0) -> np.ndarray:
    buf = np.full(n + 16, fill, dtype=np.uint8)
    return buf[:n]


def torch_int32():
    import torch

    return torch.int32


def encode_pairs_device(var, total):
    """Device twin of like_lut.encode_pairs: distinct (var,total) pairs by descending frequency, int16 codes."""
    import torch

    width = int(var.max().item()) + 1 if var.numel() else 1
    key = total.to(torch.int64) * width + var.to(torch.int64)
    uniq, inv, cnt = torch.unique(key, return_inverse=True, return_counts=True)
    uniq_c, cnt_c = uniq.cpu().numpy(), cnt.cpu().numpy()
    if len(uniq_c) > 32767:
        raise _lib.PhError(f"{len(uniq_c)} distinct (var,total) pairs exceed the device code space")
    order = np.lexsort((uniq_c, -cnt_c))
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    code = torch.from_numpy(rank.astype(np.int16)).to(key.device)[inv].contiguous()
    u = uniq_c[order]
    return (u % width).astype(np.int64), (u // width).astype(np.int64), code


def pair_counts(var, total):
    """Distinct (var, total) pairs of a COO block with their counts (host arrays); inputs numpy or torch."""
    if isinstance(var, np.ndarray):
        pv, pt, cnt, _ = encode_pairs(var, total)
        return pv, pt, cnt
    import torch

    width = int(var.max().item()) + 1 if var.numel() else 1
    key = total.to(torch.int64) * width + var.to(torch.int64)
    uniq, cnt = torch.unique(key, return_counts=True)
    u = uniq.cpu().numpy()
    return (u % width).astype(np.int64), (u // width).astype(np.int64), cnt.cpu().numpy().astype(np.int64)


def order_pairs(pv, pt, cnt):
    """Code order: descending frequency, ties by ascending (total, var) -- like_lut.encode_pairs."""
    order = np.lexsort((pv, pt, -np.asarray(cnt)))
    return np.asarray(pv)[order], np.asarray(pt)[order]


def encode_with_table(var, total, pv, pt):
    """Codes of every observation in a GIVEN (var, total) table (all ranks of a sharded store share one table)."""
    width = int(max(int(pv.max()) if len(pv) else 0, int(var.max()) if len(var) else 0)) + 1
    tkey = np.asarray(pt, np.int64) * width + np.asarray(pv, np.int64)
    order = np.argsort(tkey)
    if isinstance(var, np.ndarray):
        key = np.asarray(total, np.int64) * width + np.asarray(var, np.int64)
        pos = np.searchsorted(tkey[order], key)
        pos = np.minimum(pos, len(order) - 1)
        if len(key) and not np.array_equal(tkey[order][pos], key):
            raise _lib.PhError("an observation's (var,total) pair is missing from the given table")
        return order[pos].astype(np.uint16)
    import torch

    key = total.to(torch.int64) * width + var.to(torch.int64)
    skey = torch.from_numpy(tkey[order]).to(key.device)
    pos = torch.searchsorted(skey, key).clamp_(max=len(order) - 1)
    if key.numel() and not bool((skey[pos] == key).all().item()):
        raise _lib.PhError("an observation's (var,total) pair is missing from the given table")
    return torch.from_numpy(order.astype(np.int16)).to(key.device)[pos].contiguous()


class ObservationStore:
    def __init__(self, handle, n_cells, n_snvs, pair_var, pair_total, L0, L1, device):
        self._h = handle
        self.n_cells = int(n_cells)
        self.n_snvs = int(n_snvs)
        self.pair_var = pair_var
        self.pair_total = pair_total
        self.L0 = L0
        self.L1 = L1
        self.device = device

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_coo(cls, n_cells, n_snvs, cell, snv, var, total, alpha=0.001, max_copies=5, device=0, pairs=None, codes=None):
        """cell/snv/var/total: numpy arrays (host) or torch CUDA tensors (stay on the device).
        pairs=(pair_var, pair_total): use this (var,total) table and code order instead of deriving it from the data
        (the ranks of a cell-sharded store must agree on the codes).
        codes: the observations' indices into `pairs`, already computed (int16 / uint16; var and total are then ignored)."""
        L = _lib.lib()
        on_dev = not isinstance(cell, np.ndarray)
        if codes is not None:
            if pairs is None:
                raise _lib.PhError("codes= needs the (var,total) table they index (pairs=)")
            pv, pt = (np.asarray(x, np.int64) for x in pairs)
            code = codes.contiguous() if on_dev else np.ascontiguousarray(codes, dtype=np.uint16)
            if on_dev:
                cell = cell.to(torch_int32()).contiguous()
                snv = snv.to(torch_int32()).contiguous()
                nnz = int(cell.numel())
            else:
                cell = np.ascontiguousarray(cell, dtype=np.int32)
                snv = np.ascontiguousarray(snv, dtype=np.int32)
                nnz = int(cell.shape[0])
        elif on_dev:
            if pairs is None:
                pv, pt, code = encode_pairs_device(var, total)
            else:
                pv, pt = (np.asarray(x, np.int64) for x in pairs)
                code = encode_with_table(var, total, pv, pt)
            cell = cell.to(torch_int32()).contiguous()
            snv = snv.to(torch_int32()).contiguous()
            nnz = int(cell.numel())
        else:
            if pairs is None:
                pv, pt, _, code32 = encode_pairs(var, total)
                code = code32.astype(np.uint16)
            else:
                pv, pt = (np.asarray(x, np.int64) for x in pairs)
                code = encode_with_table(np.asarray(var), np.asarray(total), pv, pt)
            cell = np.ascontiguousarray(cell, dtype=np.int32)
            snv = np.ascontiguousarray(snv, dtype=np.int32)
            nnz = int(cell.shape[0])
        if len(pv) > 65535:
            raise _lib.PhError(f"{len(pv)} distinct (var,total) pairs exceed the 16-bit code space")
        L0, L1 = like_tables(pv, pt, alpha, max_copies)
        L0 = np.ascontiguousarray(L0, dtype=np.float64)
        L1 = np.ascontiguousarray(L1, dtype=np.float64)
        vp = np.ascontiguousarray(pv > 0, dtype=np.uint8)
        h = C.c_void_p()
        _lib.check(L.ph_create(int(n_cells), int(n_snvs), nnz, _ptr(cell), _ptr(snv), _ptr(code), int(on_dev), len(pv),
                               _ptr(L0), _ptr(L1), _ptr(vp), int(device), C.byref(h)))
        return cls(h, n_cells, n_snvs, pv, pt, L0, L1, device)

    def close(self):
        if self._h is not None:
            _lib.lib().ph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> dict:
        inf = _lib.PhInfo()
        _lib.check(_lib.lib().ph_get_info(self._h, C.byref(inf)))
        return {k: getattr(inf, k) for k, _ in inf._fields_}

    def set_group(self, by_snv=0, by_cell=0):
        _lib.check(_lib.lib().ph_set_group(self._h, by_snv, by_cell))

    # ------------------------------------------------------------------ label helpers
    def cell_labels(self, *groups) -> np.ndarray:
        return _labels(self.n_cells, groups)

    def snv_labels(self, *groups) -> np.ndarray:
        return _labels(self.n_snvs, groups)

    @staticmethod
    def _list(idx):
        return None if idx is None else np.ascontiguousarray(idx, dtype=np.int32)

    # ------------------------------------------------------------------ host-buffer calls (C-ABI, on_device = 0)
    def _table(self, fn, n_major, label, C_, lst, want):
        lst = self._list(lst)
        k = n_major if lst is None else len(lst)
        out = {}
        S0 = np.empty((k, C_), np.float64) if "S0" in want else None
        S1 = np.empty((k, C_), np.float64) if "S1" in want else None
        No = np.empty((k, C_), np.int32) if "Nobs" in want else None
        Nv = np.empty((k, C_), np.int32) if "Nvar" in want else None
        if k:
            _lib.check(fn(self._h, _ptr(label), C_, _ptr(lst), k, _ptr(S0), _ptr(S1), _ptr(No), _ptr(Nv), 0, None))
        for name, arr in (("S0", S0), ("S1", S1), ("Nobs", No), ("Nvar", Nv)):
            if arr is not None:
                out[name] = arr
        return out

    def snv_table(self, cell_label, C_, snv_list=None, want=("S0", "S1", "Nobs", "Nvar")):
        """K1: per-SNV sums/counts over cell clusters 1..C_ -> dict of [len(snv_list), C_] arrays."""
        return self._table(_lib.lib().ph_snv_table, self.n_snvs, cell_label, C_, snv_list, want)

    def cell_table(self, snv_label, Q, cell_list=None, want=("S0", "S1", "Nobs", "Nvar")):
        """K2: per-cell sums/counts over SNV clusters 1..Q."""
        return self._table(_lib.lib().ph_cell_table, self.n_cells, snv_label, Q, cell_list, want)

    def assign_linear(self, cell_label, lenA, snv_list=None, out=None):
        """`out=(labels uint8 [k], nobs int32 [k])` reuses caller buffers; page-locked ones are written by the D2H copy
        directly (no staging memcpy)."""
        lst = self._list(snv_list)
        k = self.n_snvs if lst is None else len(lst)
        lab, nobs = out if out is not None else (_padded_u8(k), np.empty(k, np.int32))
        if k:
            _lib.check(_lib.lib().ph_assign_linear(self._h, _ptr(cell_label), float(lenA), _ptr(lst), k, _ptr(lab),
                                                   _ptr(nobs), 0, None))
        return lab, nobs

    def assign_branching(self, cell_label, snv_list=None, out=None):
        lst = self._list(snv_list)
        k = self.n_snvs if lst is None else len(lst)
        lab, nobs = out if out is not None else (_padded_u8(k), np.empty((k, 2), np.int32))
        if k:
            _lib.check(_lib.lib().ph_assign_branching(self._h, _ptr(cell_label), _ptr(lst), k, _ptr(lab), _ptr(nobs), 0,
                                                      None))
        return lab, nobs

    def block_sums(self, cell_label, C_, snv_label, Q):
        """K3 -> (sum0, sum1, nobs, nvar), each [Q, C_] (SNV cluster major)."""
        s0 = np.empty((Q, C_), np.float64)
        s1 = np.empty((Q, C_), np.float64)
        no = np.empty((Q, C_), np.int64)
        nv = np.empty((Q, C_), np.int64)
        _lib.check(_lib.lib().ph_block_sums(self._h, _ptr(cell_label), C_, _ptr(snv_label), Q, _ptr(s0), _ptr(s1),
                                            _ptr(no), _ptr(nv), 0, None))
        return s0, s1, no, nv

    def line_reduce(self, S0, S1, Nobs, Nvar, line_label, ncol, nlab):
        """Fixed-order sums of a per-line table [n_lines, ncol] over the lines of every label 1..nlab -> four [ncol, nlab]
        arrays: the reduction behind block_sums, on a caller-provided table (sharded_store.py assembles one)."""
        S0 = np.ascontiguousarray(S0, np.float64).reshape(-1, ncol)
        S1 = np.ascontiguousarray(S1, np.float64).reshape(-1, ncol)
        No = np.ascontiguousarray(Nobs, np.int32).reshape(-1, ncol)
        Nv = np.ascontiguousarray(Nvar, np.int32).reshape(-1, ncol)
        n_lines = S0.shape[0]
        s0 = np.empty((ncol, nlab), np.float64)
        s1 = np.empty((ncol, nlab), np.float64)
        no = np.empty((ncol, nlab), np.int64)
        nv = np.empty((ncol, nlab), np.int64)
        _lib.check(_lib.lib().ph_line_reduce(self._h, _ptr(S0), _ptr(S1), _ptr(No), _ptr(Nv), _ptr(line_label), n_lines, ncol,
                                             nlab, _ptr(s0), _ptr(s1), _ptr(no), _ptr(nv)))
        return s0, s1, no, nv

    def node_reduce(self, per_line, line_node, K):
        """per_node[k] = sum of per_line over the lines with line_node == k, in the fixed order tree_loglik uses."""
        v = np.ascontiguousarray(per_line, np.float64)
        out = np.empty(K, np.float64)
        _lib.check(_lib.lib().ph_node_reduce(self._h, _ptr(v), _ptr(line_node), len(v), K, _ptr(out)))
        return out

    def tree_loglik(self, cell_node, snv_node, present, per_cell=False):
        K = present.shape[0]
        present = np.ascontiguousarray(present, dtype=np.uint8)
        out = np.empty(K, np.float64)
        pc = np.empty(self.n_cells, np.float64) if per_cell else None
        _lib.check(_lib.lib().ph_tree_loglik(self._h, _ptr(cell_node), _ptr(snv_node), _ptr(present), K, _ptr(out),
                                             _ptr(pc), 0, None))
        return (out, pc) if per_cell else out

    def snv_node_table(self, cell_node, member, snv_list=None):
        return self._node_table(_lib.lib().ph_snv_node_table, self.n_snvs, cell_node, member, snv_list)

    def cell_node_table(self, snv_node, member, cell_list=None):
        return self._node_table(_lib.lib().ph_cell_node_table, self.n_cells, snv_node, member, cell_list)

    def _node_table(self, fn, n_major, minor_node, member, lst):
        member = np.ascontiguousarray(member, dtype=np.uint8)
        K = member.shape[0]
        lst = self._list(lst)
        k = n_major if lst is None else len(lst)
        T = np.empty((k, K), np.float64)
        if k:
            _lib.check(fn(self._h, _ptr(minor_node), _ptr(member), K, _ptr(lst), k, _ptr(T), 0, None))
        return T

    def set_embedding(self, X):
        """Keep the embedding on the device: `gauss_node_ll(X)` with the SAME array object then skips the upload."""
        Xc = np.ascontiguousarray(X, dtype=np.float64)
        _lib.check(_lib.lib().ph_set_embedding(self._h, _ptr(Xc), Xc.shape[1]))
        self._emb = X

    def gauss_node_ll(self, X, cell_node, K, per_cell=False):
        out = np.empty(K, np.float64)
        pc = np.empty(self.n_cells, np.float64) if per_cell else None
        if X is getattr(self, "_emb", None):
            _lib.check(_lib.lib().ph_gauss_node_ll(self._h, None, 0, _ptr(cell_node), K, _ptr(out), _ptr(pc), 0, None))
        else:
            X = np.ascontiguousarray(X, dtype=np.float64)
            _lib.check(_lib.lib().ph_gauss_node_ll(self._h, _ptr(X), X.shape[1], _ptr(cell_node), K, _ptr(out), _ptr(pc), 0,
                                                   None))
        return (out, pc) if per_cell else out

    def gauss_node_ll_dev(self, X_t, D, cell_node_t, K, per_node_t, per_cell_t=None, stream=None):
        """K4 on device tensors (X_t: float64 [n_cells, D] row-major); enqueued on `stream`."""
        _lib.check(_lib.lib().ph_gauss_node_ll(self._h, _ptr(X_t), int(D), _ptr(cell_node_t), int(K), _ptr(per_node_t),
                                               _ptr(per_cell_t), 1, stream))

    # ------------------------------------------------------------------ device-pointer calls (torch tensors, async)
    def snv_table_dev(self, cell_label_t, C_, S0=None, S1=None, Nobs=None, Nvar=None, snv_list_t=None, stream=None):
        k = self.n_snvs if snv_list_t is None else int(snv_list_t.numel())
        _lib.check(_lib.lib().ph_snv_table(self._h, _ptr(cell_label_t), C_, _ptr(snv_list_t), k, _ptr(S0), _ptr(S1),
                                           _ptr(Nobs), _ptr(Nvar), 1, stream))

    def cell_table_dev(self, snv_label_t, Q, S0=None, S1=None, Nobs=None, Nvar=None, cell_list_t=None, stream=None):
        k = self.n_cells if cell_list_t is None else int(cell_list_t.numel())
        _lib.check(_lib.lib().ph_cell_table(self._h, _ptr(snv_label_t), Q, _ptr(cell_list_t), k, _ptr(S0), _ptr(S1),
                                            _ptr(Nobs), _ptr(Nvar), 1, stream))

    def assign_linear_dev(self, cell_label_t, lenA, label_out_t, nobs_out_t, snv_list_t=None, stream=None):
        k = self.n_snvs if snv_list_t is None else int(snv_list_t.numel())
        _lib.check(_lib.lib().ph_assign_linear(self._h, _ptr(cell_label_t), float(lenA), _ptr(snv_list_t), k,
                                               _ptr(label_out_t), _ptr(nobs_out_t), 1, stream))

    def assign_branching_dev(self, cell_label_t, label_out_t, nobs_out_t, snv_list_t=None, stream=None):
        k = self.n_snvs if snv_list_t is None else int(snv_list_t.numel())
        _lib.check(_lib.lib().ph_assign_branching(self._h, _ptr(cell_label_t), _ptr(snv_list_t), k, _ptr(label_out_t),
                                                  _ptr(nobs_out_t), 1, stream))

    def snv_partial_dev(self, cell_label_t, C_, packed_t, snv_list_t=None, stream=None):
        """This rank's partial [S0|S1|Nobs] table (float64, 3*k*C) for one NCCL all-reduce (cell-sharded K1)."""
        k = self.n_snvs if snv_list_t is None else int(snv_list_t.numel())
        _lib.check(_lib.lib().ph_snv_partial(self._h, _ptr(cell_label_t), C_, _ptr(snv_list_t), k, _ptr(packed_t), 1, stream))

    @staticmethod
    def finish_linear_dev(packed_t, k, lenA, label_out_t, nobs_out_t, stream=None):
        _lib.check(_lib.lib().ph_finish_linear(_ptr(packed_t), k, float(lenA), _ptr(label_out_t), _ptr(nobs_out_t), stream))

    @staticmethod
    def finish_branching_dev(packed_t, k, label_out_t, nobs_out_t, stream=None):
        _lib.check(_lib.lib().ph_finish_branching(_ptr(packed_t), k, _ptr(label_out_t), _ptr(nobs_out_t), stream))

    def block_sums_dev(self, cell_label_t, C_, snv_label_t, Q, s0, s1, no, nv, stream=None):
        _lib.check(_lib.lib().ph_block_sums(self._h, _ptr(cell_label_t), C_, _ptr(snv_label_t), Q, _ptr(s0), _ptr(s1),
                                            _ptr(no), _ptr(nv), 1, stream))

    def tree_loglik_dev(self, cell_node_t, snv_node_t, present_t, K, per_node_t, per_cell_t=None, stream=None):
        _lib.check(_lib.lib().ph_tree_loglik(self._h, _ptr(cell_node_t), _ptr(snv_node_t), _ptr(present_t), K,
                                             _ptr(per_node_t), _ptr(per_cell_t), 1, stream))


def pairwise_dist(X, row_begin=0, row_end=None, device=0):
    """K5: squareform(pdist(X)) rows [row_begin, row_end) (phertilizer.py:286), computed on the GPU."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    n, D = X.shape
    row_end = n if row_end is None else row_end
    out = np.empty((row_end - row_begin, n), np.float64)
    _lib.check(_lib.lib().ph_pairwise_dist(_ptr(X), n, D, row_begin, row_end, _ptr(out), device, 0, None))
    return out


def pairwise_dist_gram(X, row_begin=0, row_end=None, device=0):
    """K5 on the FP64 tensor cores (ph_pairwise_dist_gram): rows [row_begin, row_end) of squareform(pdist(X)), every entry
    within 1e-9 relative of `pairwise_dist` (not bit-identical).  torch only holds the device buffers."""
    import torch

    dev = torch.device("cuda", device)
    Xd = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).to(dev)
    n, D = Xd.shape
    row_end = n if row_end is None else row_end
    out = torch.empty((row_end - row_begin, n), dtype=torch.float64, device=dev)
    sq = torch.empty(n, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().ph_pairwise_dist_gram(_ptr(Xd), n, D, row_begin, row_end, _ptr(out), _ptr(sq), device,
                                                    torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
    return out.cpu().numpy()
