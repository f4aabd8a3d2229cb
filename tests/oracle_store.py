"""TEST DOUBLE: an `ObservationStore` look-alike backed by the dense CPU oracle, so the host-side control flow of
phertilizer_b200 (splits, tree growing, post-processing) can be exercised on a box without a GPU.  Never used by the
product; it exists only under tests/."""
import numpy as np

from oracle import oracle_np as O


def _groups(lab, C):
    return [np.nonzero(lab == c)[0] for c in range(1, C + 1)]


class OracleStore:
    def __init__(self, n, m, cell, snv, var, total, alpha=0.001, max_copies=5):
        self.n_cells, self.n_snvs = int(n), int(m)
        self.D = O.DenseData(n, m, cell, snv, var, total, alpha, max_copies)

    @classmethod
    def from_coo(cls, n, m, cell, snv, var, total, alpha=0.001, max_copies=5, device=0):
        return cls(n, m, cell, snv, var, total, alpha, max_copies)

    def close(self):
        pass

    def cell_labels(self, *groups):
        lab = np.zeros(self.n_cells, np.uint8)
        for c, idx in enumerate(groups, 1):
            if idx is not None and len(idx):
                lab[np.asarray(idx, np.int64)] = c
        return lab

    def snv_labels(self, *groups):
        lab = np.zeros(self.n_snvs, np.uint8)
        for c, idx in enumerate(groups, 1):
            if idx is not None and len(idx):
                lab[np.asarray(idx, np.int64)] = c
        return lab

    def snv_table(self, cell_label, C, snv_list=None, want=("S0", "S1", "Nobs", "Nvar")):
        muts = np.arange(self.n_snvs) if snv_list is None else np.asarray(snv_list, np.int64)
        S0, S1, No, Nv = O.snv_table(self.D, _groups(cell_label, C), muts)
        return {"S0": S0, "S1": S1, "Nobs": No.astype(np.int32), "Nvar": Nv.astype(np.int32)}

    def cell_table(self, snv_label, Q, cell_list=None, want=("S0", "S1", "Nobs", "Nvar")):
        cells = np.arange(self.n_cells) if cell_list is None else np.asarray(cell_list, np.int64)
        A0, A1, No, Nv = O.cell_table(self.D, cells, _groups(snv_label, Q))
        return {"S0": A0, "S1": A1, "Nobs": No.astype(np.int32), "Nvar": Nv.astype(np.int32)}

    def assign_linear(self, cell_label, lenA, snv_list=None):
        muts = np.arange(self.n_snvs) if snv_list is None else np.asarray(snv_list, np.int64)
        cellsA = np.nonzero(cell_label == 1)[0]
        S0, S1, No, _ = O.snv_table(self.D, [cellsA], muts)
        lab = np.where(S0[:, 0] / lenA >= S1[:, 0] / lenA, 2, 1).astype(np.uint8)
        return lab, No[:, 0].astype(np.int32)

    def assign_branching(self, cell_label, snv_list=None):
        muts = np.arange(self.n_snvs) if snv_list is None else np.asarray(snv_list, np.int64)
        cA, cB = np.nonzero(cell_label == 1)[0], np.nonzero(cell_label == 2)[0]
        a, b, c = O.branching_candidates(self.D, cA, cB, muts)
        mx = np.maximum(a, np.maximum(b, c))
        lab = np.where(a == mx, 1, np.where(b == mx, 2, 3)).astype(np.uint8)
        _, _, No, _ = O.snv_table(self.D, [cA, cB], muts)
        return lab, No.astype(np.int32)

    def block_sums(self, cell_label, C, snv_label, Q):
        return O.block_sums(self.D, _groups(cell_label, C), _groups(snv_label, Q))

    def tree_loglik(self, cell_node, snv_node, present, per_cell=False):
        K = present.shape[0]
        out = np.zeros(K)
        pc = np.zeros(self.n_cells)
        for k in range(K):
            cells = np.nonzero(cell_node == k)[0]
            if len(cells) == 0:
                continue
            on = (snv_node < K) & (present[k][np.minimum(snv_node, K - 1)] != 0)
            pm = np.nonzero(on)[0]
            out[k] = O.variant_likelihood_by_node(self.D, cells, pm)
            pc[cells] = self.D.like1[np.ix_(cells, pm)].sum(axis=1) + self.D.like0[np.ix_(cells, np.nonzero(~on)[0])].sum(axis=1)
        return (out, pc) if per_cell else out

    def line_reduce(self, S0, S1, Nobs, Nvar, line_label, ncol, nlab):
        """CPU stand-in of ObservationStore.line_reduce (ph_line_reduce): [ncol, nlab] sums over the lines of every label."""
        lab = np.asarray(line_label)
        tabs = [np.asarray(t).reshape(-1, ncol) for t in (S0, S1, Nobs, Nvar)]
        outs = [np.zeros((ncol, nlab), np.float64), np.zeros((ncol, nlab), np.float64), np.zeros((ncol, nlab), np.int64),
                np.zeros((ncol, nlab), np.int64)]
        for c in range(1, nlab + 1):
            rows = lab[: tabs[0].shape[0]] == c
            for t, o in zip(tabs, outs):
                o[:, c - 1] = t[rows].sum(axis=0)
        return tuple(outs)

    def node_reduce(self, per_line, line_node, K):
        v, lab = np.asarray(per_line, np.float64), np.asarray(line_node)[: len(per_line)]
        return np.array([v[lab == k].sum() for k in range(K)])

    def set_embedding(self, X):
        pass

    def gauss_node_ll(self, X, cell_node, K, per_cell=False):
        out = np.zeros(K)
        for k in range(K):
            cells = np.nonzero(cell_node == k)[0]
            if len(cells):
                out[k] = O.gauss_node_loglik(np.asarray(X, float), cells)[0]
        return out

    def snv_node_table(self, cell_node, member, snv_list=None):
        muts = np.arange(self.n_snvs) if snv_list is None else np.asarray(snv_list, np.int64)
        K = member.shape[0]
        T = np.zeros((len(muts), K))
        for k in range(K):
            y = member[k][np.minimum(cell_node, K - 1)].astype(float) * (cell_node < K)
            T[:, k] = y @ self.D.like1[:, muts] + (1 - y) @ self.D.like0[:, muts]
        return T

    def cell_node_table(self, snv_node, member, cell_list=None):
        cells = np.arange(self.n_cells) if cell_list is None else np.asarray(cell_list, np.int64)
        K = member.shape[0]
        T = np.zeros((len(cells), K))
        for k in range(K):
            y = member[k][np.minimum(snv_node, K - 1)].astype(float) * (snv_node < K)
            T[:, k] = self.D.like1[cells, :] @ y + self.D.like0[cells, :] @ (1 - y)
        return T


class HostCopyDistance:
    """TEST DOUBLE of phertilizer_b200.cluster.CopyDistance: the reference's own host arithmetic (scipy pdist, numpy
    exp / quantile, sklearn spectral_clustering -- linear_split.py:308-335, branching_split.py:215-232, utils.py:131-150)."""

    def __init__(self, embedding, device=0):
        self._full = O.pairwise_dist(np.asarray(embedding, np.float64))
        self.n = self._full.shape[0]
        self.shape = self._full.shape

    def full(self):
        return self._full

    def rows(self, idx):
        return self._full[np.asarray(idx, np.int64), :]

    def affinity(self, cells, feats, use_copy_kernel, radius, strict):
        from scipy.spatial.distance import pdist, squareform

        feats = np.asarray(feats, np.float64).reshape(len(cells), -1)
        d_mat = squareform(pdist(feats, metric="euclidean"))
        kernel = np.exp(-1 * d_mat / (0.2 * (np.max(d_mat) - np.min(d_mat))))
        c_mat = self._full[np.ix_(cells, cells)]
        r = np.quantile(c_mat, radius)
        if use_copy_kernel:
            copy_kernel = np.exp(-1 * c_mat / (0.2 * (np.max(c_mat) - np.min(c_mat))))
            copy_kernel[(c_mat >= r) if strict else (c_mat > r)] = 0
            kernel = np.multiply(kernel, copy_kernel)
        return kernel

    def min_cut(self, cells, feats, use_copy_kernel, radius, strict, random_state):
        from sklearn.cluster import spectral_clustering

        W = self.affinity(cells, feats, use_copy_kernel, radius, strict)
        if np.any(np.isnan(W)):
            return None
        return spectral_clustering(W, n_clusters=2, random_state=random_state, assign_labels="cluster_qr",
                                   eigen_solver="lobpcg")


def patch_phertilizer(monkeypatch):
    """Route phertilizer_b200.phertilizer's data layer to the CPU test doubles."""
    import phertilizer_b200.phertilizer as P

    monkeypatch.setattr(P, "ObservationStore", OracleStore)
    monkeypatch.setattr(P, "CopyDistance", HostCopyDistance)
