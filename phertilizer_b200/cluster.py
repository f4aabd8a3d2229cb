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
# the LOBPCG iteration of the cut with its block vectors resident on the GPU (device_lobpcg below); PH_DEVICE_LOBPCG=0 keeps
# scipy's host iteration and only serves it `L @ X` products
DEVICE_LOBPCG = __import__("os").environ.get("PH_DEVICE_LOBPCG", "1") != "0"


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

    # ------------------------------------------------------------------ device-resident block vectors (k x 3 slots)
    def bv_set(self, slot, X):
        X = np.ascontiguousarray(X, dtype=np.float64)
        _lib.check(self._L.ph_affinity_bv_set(self._h, int(slot), _dptr(X)))

    def bv_get(self, slot) -> np.ndarray:
        out = np.empty((self._k, 3), np.float64)
        _lib.check(self._L.ph_affinity_bv_get(self._h, int(slot), _dptr(out)))
        return out

    def bv_apply(self, src, dst):
        """slot dst = L @ slot src, on the device."""
        _lib.check(self._L.ph_affinity_bv_apply(self._h, int(src), int(dst), None))

    def bv_gram(self, left, right) -> np.ndarray:
        """[3 len(left), 3 len(right)] matrix of inner products of the slots' columns (fixed-order reduction)."""
        l = np.ascontiguousarray(left, dtype=np.int32)
        r = np.ascontiguousarray(right, dtype=np.int32)
        out = np.empty((3 * len(l), 3 * len(r)), np.float64)
        _lib.check(self._L.ph_affinity_bv_gram(self._h, _dptr(l), len(l), _dptr(r), len(r), _dptr(out)))
        return out

    def bv_combine(self, dst, terms):
        """slot dst = sum over (slot, C) in terms of slot @ C, C 3 x 3 (dst may be one of the sources)."""
        buf = getattr(self, "_bv_buf", None)
        if buf is None:   # reused marshalling buffers: this is called a dozen times per LOBPCG iteration
            buf = self._bv_buf = (np.zeros(4, np.int32), np.zeros((4, 3, 3), np.float64))
            self._bv_ptr = (_dptr(buf[0]), _dptr(buf[1]))
        for t, (slot, Cm) in enumerate(terms):
            buf[0][t] = slot
            buf[1][t] = Cm
        rc = self._L.ph_affinity_bv_combine(self._h, int(dst), self._bv_ptr[0], self._bv_ptr[1], len(terms))
        if rc:
            _lib.check(rc)

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
            if k < SMALL:
                _, diffusion_map = lobpcg(self.laplacian(), X, tol=None, largest=False, maxiter=2000)
            elif DEVICE_LOBPCG:
                _, diffusion_map = device_lobpcg(self, X, maxiter=2000)
            else:
                _, diffusion_map = lobpcg(lambda V: self.matmat(V), X, tol=None, largest=False, maxiter=2000)
        embedding = diffusion_map.T[:n_components]
        embedding = embedding / dd          # recover u = D^-1/2 x
        embedding = _deterministic_vector_sign_flip(embedding)
        return cluster_qr(embedding[:n_components].T)


def device_lobpcg(ctx, X0, maxiter=2000, restart_control=20):
    """scipy.sparse.linalg.lobpcg(A, X0, tol=None, largest=False, maxiter=maxiter) for the Laplacian held by `ctx`, with the
    block vectors on the device.

    sklearn's spectral_embedding (the reference's normalizedMinCut, utils.py:131-150) hands the normalised Laplacian and a
    k x 3 start block to scipy's LOBPCG.  The method itself is a short recurrence on six k x 3 blocks (X, AX, residuals R, AR,
    directions P, AP) driven by 3 x 3 ... 9 x 9 Gram matrices: here the blocks never leave the GPU (`ctx.bv_*`: product,
    Gram, linear combination), the small dense algebra (Cholesky, inverse, symmetric eigenproblem) stays with the same scipy
    routines on the host, and the control flow -- active mask, restarts, implicit / explicit Gram matrices, the final
    Rayleigh-Ritz on the best iterate -- follows scipy's (no constraints, no preconditioner, standard problem).  Only the
    summation order of the k-long inner products differs from BLAS, as the order of the GPU product `L @ X` already did."""
    from scipy.linalg import LinAlgError, cholesky, eigh, inv

    k, m = X0.shape
    assert m == 3
    X, AX, R, AR, P, AP, BEST, T1, T2 = range(9)
    I3 = np.eye(3)
    tol = np.sqrt(np.finfo(np.float64).eps) * k

    def pad(M):
        out = np.zeros((3, 3))
        out[: M.shape[0], : M.shape[1]] = M
        return out

    def sym_blocks(diag, upper):
        """The symmetric block matrix np.block would build from the diagonal blocks and the blocks above them (row-major:
        (0,1), (0,2), (1,2)) -- by slice assignment: np.block's nesting checks cost more than the eigenproblem here."""
        sizes = [d.shape[0] for d in diag]
        at = np.concatenate([[0], np.cumsum(sizes)])
        M = np.empty((at[-1], at[-1]))
        u = 0
        for a in range(len(diag)):
            M[at[a]:at[a + 1], at[a]:at[a + 1]] = diag[a]
            for b in range(a + 1, len(diag)):
                M[at[a]:at[a + 1], at[b]:at[b + 1]] = upper[u]
                M[at[b]:at[b + 1], at[a]:at[a + 1]] = upper[u].T
                u += 1
        return M

    def orthonormalize(slot, ca):
        """In-place (B = I) orthonormalisation of the first ca columns by Cholesky; returns inv(chol) or None."""
        G = ctx.bv_gram([slot], [slot])[:ca, :ca]
        try:
            U = inv(cholesky(G, overwrite_a=True), overwrite_a=True)
        except LinAlgError:
            return None
        ctx.bv_combine(slot, [(slot, pad(U))])
        return U

    ctx.bv_set(X, X0)
    if orthonormalize(X, 3) is None:
        raise ValueError("Linearly dependent initial approximations")
    ctx.bv_apply(X, AX)
    lam, V = eigh(ctx.bv_gram([X], [AX]), check_finite=False)
    ii = np.argsort(lam)[:m]
    lam, V = lam[ii], np.asarray(V[:, ii])
    ctx.bv_combine(X, [(X, V)])
    ctx.bv_combine(AX, [(AX, V)])

    active = np.ones(m, dtype=bool)
    smallest = np.abs(np.finfo(np.float64).max)
    best_is_current = False
    it = -1
    restart, forced_restart, explicit = True, False, False
    have_p = False
    myeps = np.sqrt(np.finfo(np.float64).eps)
    while it < maxiter:
        it += 1
        ctx.bv_combine(R, [(AX, I3), (X, -np.diag(lam))])          # R = AX - X * lambda
        G0 = ctx.bv_gram([X, R], [R])                               # X^T R (used below) and R^T R in one reduction
        rn = np.sqrt(np.abs(np.diag(G0[3:])))
        rnorm = np.sum(np.abs(rn)) / m
        if rnorm < smallest:
            smallest = rnorm
            ctx.bv_combine(BEST, [(X, I3)])
        elif rnorm > 2 ** restart_control * smallest:
            forced_restart = True
            ctx.bv_apply(X, AX)
        active = active & (rn > tol)
        ca = int(active.sum())
        if ca == 0:
            break
        S = np.zeros((3, 3))
        S[np.nonzero(active)[0], np.arange(ca)] = 1.0                  # active columns first, zero columns behind
        ctx.bv_combine(R, [(R, S)])
        if it > 0:
            ctx.bv_combine(P, [(P, S)])
            ctx.bv_combine(AP, [(AP, S)])
        XR = G0[:3] @ S                                                # X^T (R S): the selection only permutes / drops columns
        ctx.bv_combine(R, [(R, I3), (X, -XR)])                        # orthogonal to X
        if orthonormalize(R, ca) is None:
            break
        ctx.bv_apply(R, AR)
        if it > 0:
            invR = orthonormalize(P, ca)
            if invR is not None:
                ctx.bv_combine(AP, [(AP, pad(invR))])
                restart = forced_restart
            else:
                restart = True
        if not (rn.max() > myeps and not explicit):
            explicit = True
        left = [X, R] if restart else [X, R, P]
        right = ([X, R, AX, AR] if restart else [X, R, P, AX, AR, AP])
        G = ctx.bv_gram(left, right)
        nl = len(left)

        def blk(a, b, rows, cols):
            return G[3 * a: 3 * a + rows, 3 * b: 3 * b + cols]

        iX, iR, iP = 0, 1, 2
        jAX, jAR, jAP = (2, 3, None) if restart else (3, 4, 5)
        gXAR = blk(iX, jAR, m, ca)
        gRAR = blk(iR, jAR, ca, ca)
        if explicit:
            gRAR = (gRAR + gRAR.T) / 2
            gXAX = blk(iX, jAX, m, m)
            gXAX = (gXAX + gXAX.T) / 2
            gXBX = blk(iX, iX, m, m)
            gRBR = blk(iR, iR, ca, ca)
            gXBR = blk(iX, iR, m, ca)
        else:
            gXAX = np.diag(lam)
            gXBX = np.eye(m)
            gRBR = np.eye(ca)
            gXBR = np.zeros((m, ca))
        lam_all = Vall = None
        if not restart:
            gXAP = blk(iX, jAP, m, ca)
            gRAP = blk(iR, jAP, ca, ca)
            gPAP = blk(iP, jAP, ca, ca)
            gXBP = blk(iX, iP, m, ca)
            gRBP = blk(iR, iP, ca, ca)
            if explicit:
                gPAP = (gPAP + gPAP.T) / 2
                gPBP = blk(iP, iP, ca, ca)
            else:
                gPBP = np.eye(ca)
            gramA = sym_blocks([gXAX, gRAR, gPAP], [gXAR, gXAP, gRAP])
            gramB = sym_blocks([gXBX, gRBR, gPBP], [gXBR, gXBP, gRBP])
            try:
                lam_all, Vall = eigh(gramA, gramB, check_finite=False)
            except LinAlgError:
                restart = True
        if restart:
            gramA = sym_blocks([gXAX, gRAR], [gXAR])
            gramB = sym_blocks([gXBX, gRBR], [gXBR])
            try:
                lam_all, Vall = eigh(gramA, gramB, check_finite=False)
            except LinAlgError:
                break
        ii = np.argsort(lam_all)[:m]
        lam = lam_all[ii]
        Vall = Vall[:, ii]
        eX, eR = Vall[:m], Vall[m: m + ca]
        if not restart:
            eP = Vall[m + ca:]
            ctx.bv_combine(T1, [(R, pad(eR)), (P, pad(eP))])
            ctx.bv_combine(T2, [(AR, pad(eR)), (AP, pad(eP))])
        else:
            ctx.bv_combine(T1, [(R, pad(eR))])
            ctx.bv_combine(T2, [(AR, pad(eR))])
        ctx.bv_combine(X, [(X, eX), (T1, I3)])
        ctx.bv_combine(AX, [(AX, eX), (T2, I3)])
        P, T1 = T1, P                                                   # the new directions; their old slots become scratch
        AP, T2 = T2, AP
        have_p = True

    ctx.bv_combine(R, [(AX, I3), (X, -np.diag(lam))])
    rn = np.sqrt(np.abs(np.diag(ctx.bv_gram([R], [R]))))
    if np.sum(np.abs(rn)) / m < smallest:
        ctx.bv_combine(BEST, [(X, I3)])
    # final "exact" Rayleigh-Ritz on the best iterate
    ctx.bv_apply(BEST, AX)
    G = ctx.bv_gram([BEST], [BEST, AX])
    gXBX, gXAX = G[:, :3], G[:, 3:]
    gXAX = (gXAX + gXAX.T) / 2
    gXBX = (gXBX + gXBX.T) / 2
    lam, V = eigh(gXAX, gXBX, check_finite=False)
    ii = np.argsort(lam)[:m]
    lam, V = lam[ii], np.asarray(V[:, ii])
    ctx.bv_combine(BEST, [(BEST, V)])
    return lam, ctx.bv_get(BEST)


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

    def bv_apply(self, src, dst):
        if self._rows is None:
            return super().bv_apply(src, dst)
        lo, hi = self._rows
        Yb = np.empty((max(hi - lo, 1), 3), np.float64)
        _lib.check(self._L.ph_affinity_bv_apply(self._h, int(src), int(dst), _dptr(Yb)))
        self.bv_set(dst, self._assemble(Yb[: hi - lo], self._k, lo, hi))

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
