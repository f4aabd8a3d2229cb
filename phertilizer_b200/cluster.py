"""Cell clustering of the tree operations on the GPU.

The reference builds a dense affinity between the cells of a seed -- `exp(-pdist(tmb)/kw)` times an optional kernel on
the copy-number (embedding) distance -- and cuts it in two with sklearn's `spectral_clustering(assign_labels=
"cluster_qr", eigen_solver="lobpcg")` (linear_split.py:277-390, branching_split.py:175-291, utils.py:131-150).  Here
the n x n embedding distance, the affinity and its normalised Laplacian live in HBM (`csrc/ph_affinity.cu`); the
eigensolver is still scipy's LOBPCG -- the reference's own third-party code, fed with the same initial block drawn from
the same `RandomState` -- but its operator is the GPU product `L @ X`.  Everything after the eigenvectors
(`_deterministic_vector_sign_flip`, `cluster_qr`) is sklearn's.

For sets of fewer than `SMALL` cells the (tiny) Laplacian is copied to the host and handed to the same scipy routines
as a dense array -- LOBPCG itself switches to a dense solver below 15 nodes, sklearn to `eigh` below 11.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

SMALL = 64


def _dptr(a):
    return a.ctypes.data_as(C.c_void_p)


class CopyDistance:
    """`squareform(pdist(embedding))` (phertilizer.py:286) resident on the GPU, with the affinity workspace."""

    def __init__(self, embedding, device=0):
        emb = np.ascontiguousarray(embedding, dtype=np.float64)
        self.n, self.D = emb.shape
        self._h = C.c_void_p()
        self._L = _lib.lib()
        _lib.check(self._L.ph_affinity_create(_dptr(emb), self.n, self.D, int(device), C.byref(self._h)))
        self._full = None
        self._k = 0

    shape = property(lambda self: (self.n, self.n))

    def close(self):
        if self._h is not None:
            self._L.ph_affinity_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ distance access
    def rows(self, idx) -> np.ndarray:
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        out = np.empty((len(idx), self.n), np.float64)
        if len(idx):
            _lib.check(self._L.ph_affinity_rows(self._h, _dptr(idx), len(idx), _dptr(out)))
        return out

    def nearest(self, rows, donor_mask):
        """For every cell of `rows`: the donor (donor_mask) at the smallest copy distance and the number of donors tied
        at that distance (GPU masked arg-min; replaces the full-row argsort of utils.py:100-112)."""
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        mask = np.ascontiguousarray(donor_mask, dtype=np.uint8)
        idx = np.empty(len(rows), np.int32)
        ties = np.empty(len(rows), np.int32)
        for lo in range(0, len(rows), self.n):
            part = rows[lo:lo + self.n]
            _lib.check(self._L.ph_affinity_nearest(self._h, _dptr(part), len(part), _dptr(mask), _dptr(idx[lo:]), _dptr(ties[lo:])))
        return idx, ties

    def full(self) -> np.ndarray:
        """The whole n x n matrix on the host (small inputs, tests, the < MIN_GPU_CELLS path)."""
        if self._full is None:
            out = np.empty((self.n, self.n), np.float64)
            _lib.check(self._L.ph_affinity_rows(self._h, None, self.n, _dptr(out)))
            self._full = out
        return self._full

    # ------------------------------------------------------------------ affinity + Laplacian + cut
    def build(self, cells, feats, use_copy_kernel, radius, strict):
        """Affinity of `cells` from features [k, F] -> (stats dict, dd) ; the Laplacian stays on the device."""
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        feats = np.ascontiguousarray(feats, dtype=np.float64).reshape(len(cells), -1)
        stats = np.zeros(8, np.float64)
        dd = np.empty(len(cells), np.float64)
        _lib.check(self._L.ph_affinity_build(self._h, _dptr(cells), len(cells), _dptr(feats), feats.shape[1],
                                             int(bool(use_copy_kernel)), float(radius), int(bool(strict)), _dptr(stats),
                                             _dptr(dd)))
        self._k = len(cells)
        names = ("kw", "kwc", "r", "has_nan", "max_d", "max_c", "q_lo", "q_hi")
        return dict(zip(names, stats.tolist())), dd

    def matmat(self, X) -> np.ndarray:
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X.reshape(-1, 1)
        Y = np.empty_like(X)
        _lib.check(self._L.ph_affinity_matmat(self._h, _dptr(X), X.shape[1], _dptr(Y)))
        return Y

    def laplacian(self) -> np.ndarray:
        out = np.empty((self._k, self._k), np.float64)
        _lib.check(self._L.ph_affinity_matrix(self._h, _dptr(out)))
        return out

    def min_cut(self, cells, feats, use_copy_kernel, radius, strict, random_state):
        """Two-way normalised cut of `cells` -> labels (0/1 per cell), or None when the affinity contains NaN (the
        reference then terminates the clustering).  Mirrors sklearn.cluster.spectral_clustering(n_clusters=2,
        assign_labels="cluster_qr", eigen_solver="lobpcg") -> _spectral_embedding(drop_first=False)."""
        from scipy.sparse.linalg import lobpcg
        from sklearn.cluster._spectral import cluster_qr
        from sklearn.utils.extmath import _deterministic_vector_sign_flip

        from scipy.linalg import eigh

        stats, dd = self.build(cells, feats, use_copy_kernel, radius, strict)
        if stats["has_nan"]:
            return None
        k, n_components = len(cells), 2
        if k < 5 * n_components + 1:
            # sklearn short-circuits LOBPCG for tiny graphs (manifold/_spectral_embedding.py, lobpcg branch)
            _, diffusion_map = eigh(self.laplacian(), check_finite=False)
        else:
            X = random_state.standard_normal(size=(k, n_components + 1))
            X[:, 0] = dd.ravel()
            A = self.laplacian() if k < SMALL else (lambda V: self.matmat(V))
            _, diffusion_map = lobpcg(A, X, tol=None, largest=False, maxiter=2000)
        embedding = diffusion_map.T[:n_components]
        embedding = embedding / dd          # recover u = D^-1/2 x
        embedding = _deterministic_vector_sign_flip(embedding)
        return cluster_qr(embedding[:n_components].T)


class ShardedCopyDistance(CopyDistance):
    """The cell clustering with the k x k affinity / Laplacian sharded by rows across the GPUs of a node.

    One process per GPU, every rank holds the (replicated) embedding and the rows [k r / world, k (r + 1) / world) of the
    current cell set's matrices: the affinity build -- O(k^2) exp / sqrt / div -- and every `L @ X` product of the
    eigensolver -- O(k^2) bytes -- cost 1 / world per rank.  The host side of the cut (scipy's LOBPCG iteration, sklearn's
    sign flip and cluster_qr) runs redundantly on every rank on identical inputs.  Every matrix entry, degree and product
    row is computed by one rank with the single-GPU kernels' arithmetic and the blocks are assembled exactly (maximum;
    sum into a zero array), so the cut is bit-identical to `CopyDistance.min_cut` for any number of ranks.

    Cell sets below `min_rows` cells, and radius < 1 (a quantile over all k^2 distances), take the single-GPU path on
    every rank (redundantly): the context keeps room for them.
    """

    def __init__(self, embedding, device=0, group=None, min_rows=4096):
        import torch.distributed as dist

        emb = np.ascontiguousarray(embedding, dtype=np.float64)
        self.n, self.D = emb.shape
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.min_rows = max(int(min_rows), SMALL, 2 * self.world)   # (below SMALL cells the cut reads the whole Laplacian)
        self._dist = dist
        self._device_reduce = dist.get_backend(group) != "gloo"
        # rows this context must hold: the largest row block, or a whole small cell set (k < min_rows: k * k entries)
        blk = (self.n + self.world - 1) // self.world + 1
        small = (min(self.min_rows, self.n) ** 2 + self.n - 1) // self.n
        self._h = C.c_void_p()
        self._L = _lib.lib()
        _lib.check(self._L.ph_affinity_create_rows(_dptr(emb), self.n, self.D, int(device), int(min(self.n, max(blk, small))),
                                                   C.byref(self._h)))
        self._full = None
        self._k = 0
        self._rows = None   # (lo, hi) while the context holds a row block
        self._F = 1

    # ------------------------------------------------------------------ collectives on host arrays
    def _reduce(self, arr, op):
        import torch

        t = torch.from_numpy(np.ascontiguousarray(arr).copy())
        if self._device_reduce:
            t = t.to(torch.device("cuda", torch.cuda.current_device()))
        self._dist.all_reduce(t, op=op, group=self.group)
        return t.cpu().numpy()

    def _assemble(self, block, k, lo, hi):
        full = np.zeros((k,) + block.shape[1:], np.float64)
        full[lo:hi] = block
        return self._reduce(full, self._dist.ReduceOp.SUM)

    def _sharded(self, k, radius):
        return self.world > 1 and k >= self.min_rows and radius >= 1.0

    # ------------------------------------------------------------------ affinity + Laplacian + product
    def build(self, cells, feats, use_copy_kernel, radius, strict):
        k = len(cells)
        if not self._sharded(k, radius):
            self._rows = None
            return super().build(cells, feats, use_copy_kernel, radius, strict)
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        feats = np.ascontiguousarray(feats, dtype=np.float64).reshape(k, -1)
        F = feats.shape[1]
        lo, hi = (k * self.rank) // self.world, (k * (self.rank + 1)) // self.world
        st3 = np.zeros(3, np.float64)
        _lib.check(self._L.ph_affinity_stats_rows(self._h, _dptr(cells), k, _dptr(feats), F, int(bool(use_copy_kernel)), lo, hi,
                                                  _dptr(st3)))
        st3 = self._reduce(st3, self._dist.ReduceOp.MAX)
        maxd, maxc, nan = float(st3[0]), float(st3[1]), bool(st3[2])
        kw, kwc = 0.2 * (maxd - 0.0), 0.2 * (maxc - 0.0)          # utils.py:75-84, min = 0 (the diagonal)
        stats = {"kw": kw, "kwc": kwc, "r": maxc, "has_nan": float(nan), "max_d": maxd, "max_c": maxc, "q_lo": maxc, "q_hi": maxc}
        self._k, self._rows, self._F = 0, None, F
        if nan:
            return stats, np.ones(k, np.float64)
        dd_blk = np.empty(max(hi - lo, 1), np.float64)
        has_nan = C.c_int(0)
        _lib.check(self._L.ph_affinity_build_rows(self._h, F, kw, kwc, maxc, int(bool(strict)), _dptr(dd_blk), C.byref(has_nan)))
        nan = bool(self._reduce(np.array([float(has_nan.value)]), self._dist.ReduceOp.MAX)[0])
        stats["has_nan"] = float(nan)
        if nan:
            return stats, np.ones(k, np.float64)
        dd = self._assemble(dd_blk[: hi - lo], k, lo, hi)
        _lib.check(self._L.ph_affinity_laplacian_rows(self._h, _dptr(dd)))
        self._k, self._rows = k, (lo, hi)
        return stats, dd

    def matmat(self, X):
        if self._rows is None:
            return super().matmat(X)
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X.reshape(-1, 1)
        lo, hi = self._rows
        Yb = np.empty((max(hi - lo, 1), X.shape[1]), np.float64)
        _lib.check(self._L.ph_affinity_matmat_rows(self._h, _dptr(X), X.shape[1], _dptr(Yb)))
        return self._assemble(Yb[: hi - lo], self._k, lo, hi)

    def laplacian(self):
        if self._rows is not None:
            raise _lib.PhError("the Laplacian is sharded by rows: no host copy")
        return super().laplacian()

    def nearest(self, rows, donor_mask):
        """Rows are independent: every rank takes a slice of the list, the slices are assembled."""
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        if self.world == 1 or len(rows) < 4 * self.world:
            return super().nearest(rows, donor_mask)
        lo, hi = (len(rows) * self.rank) // self.world, (len(rows) * (self.rank + 1)) // self.world
        idx, ties = super().nearest(rows[lo:hi], donor_mask)
        both = np.zeros((len(rows), 2), np.float64)      # indices and counts are small integers: exact in float64
        both[lo:hi, 0], both[lo:hi, 1] = idx, ties
        both = self._reduce(both, self._dist.ReduceOp.SUM)
        return both[:, 0].astype(np.int32), both[:, 1].astype(np.int32)
