"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, dense semantics) of the reference's likelihood hot path.

Every function restates one family of ``M[np.ix_(cells, muts)].<reduce>`` expressions of elkebir-group/phertilizer
and cites the lines it follows.  It is the checker for the CUDA path in ``tests/`` and ``__graft_entry__.smoke()``;
nothing under ``phertilizer_b200/`` may import it.  The restatement is pinned against the reference itself:
``tests/golden/make_goldens.py`` runs the unmodified reference (via ``oracle/ref_shim.py``) on the bundled example
and on synthetic inputs and stores per-call inputs/outputs; ``tests/test_oracle_vs_golden.py`` replays them here.

Dense on purpose: like the reference it materialises ``n x m`` arrays, so it is only usable at the small shapes the
reference itself can hold.  ``oracle/oracle_sparse.c`` restates the same sums over sparse CSR/CSC data for the large
shapes and for the CPU baseline of ``bench.py``.
"""
from __future__ import annotations

import numpy as np
import scipy.special
from scipy.spatial.distance import pdist, squareform

np.seterr(invalid="ignore", divide="ignore")


# ---------------------------------------------------------------------------------------------- a-1 / a-2
def _ipow(a, b):
    """float64 ** int64 the way numba lowers it (numba/cpython/numbers.py int_power_impl): LSB-first
    square-and-multiply, not libm pow."""
    r = 1.0
    while b != 0:
        if b & 1:
            r *= a
        b >>= 1
        a *= a
    return r


def like0_like1(var, total, alpha=0.001, max_copies=5, min_copies=1):
    """phertilizer.py:29-40 (binom_pdf), 52-67, 70-100, 102-119, 121-151, 295."""
    var = np.asarray(var, np.int64)
    total = np.asarray(total, np.int64)
    coeff = scipy.special.comb(total, var)
    vafs_all = np.array([i / c for c in np.arange(min_copies, max_copies + 1, 1) for i in np.arange(1, c + 1, 1)])
    vafs = np.unique(vafs_all)
    probs = np.array([np.count_nonzero(vafs_all == v) / len(vafs_all) for v in vafs])
    prime = np.array([v * (1 - alpha) + (1 - v) * (alpha / 3) for v in vafs])
    l0 = np.empty(len(var))
    l1 = np.empty(len(var))
    for i, (k, n, c) in enumerate(zip(var.tolist(), total.tolist(), coeff.tolist())):
        l0[i] = c * _ipow(alpha, k) * _ipow(1 - alpha, n - k)
        p = 0.0
        for v, w in zip(prime.tolist(), probs.tolist()):
            p += c * _ipow(v, k) * _ipow(1 - v, n - k) * w
        l1[i] = p
    return np.log(l0), np.log(l1)


class DenseData:
    """The four dense matrices of the reference's ``Data`` (data.py:8-20; phertilizer.py:153-161, 295-306)."""

    def __init__(self, n, m, cell, snv, var, total, alpha=0.001, max_copies=5):
        cell = np.asarray(cell, np.int64)
        snv = np.asarray(snv, np.int64)
        var = np.asarray(var, np.int64)
        total = np.asarray(total, np.int64)
        key = total * (var.max() + 1 if len(var) else 1) + var
        uniq, inv = np.unique(key, return_inverse=True)
        w = (var.max() + 1) if len(var) else 1
        u0, u1 = like0_like1(uniq % w, uniq // w, alpha, max_copies)
        self.n, self.m = n, m
        self.var = np.zeros((n, m), np.int64)
        self.total = np.zeros((n, m), np.int64)
        self.like0 = np.zeros((n, m))
        self.like1 = np.zeros((n, m))
        self.var[cell, snv] = var
        self.total[cell, snv] = total
        self.like0[cell, snv] = u0[inv]
        self.like1[cell, snv] = u1[inv]


def _ix(M, cells, muts):
    return M[np.ix_(np.asarray(cells, np.int64), np.asarray(muts, np.int64))]


# ---------------------------------------------------------------------------------------------- K1 family
def snv_table(D: DenseData, cell_groups, muts):
    """Per-SNV S0,S1,Nobs,Nvar for each cell group: linear_split.py:186-188, 224, 229-230;
    branching_split.py:316-317, 332-337."""
    C = len(cell_groups)
    k = len(muts)
    S0 = np.zeros((k, C)); S1 = np.zeros((k, C))
    No = np.zeros((k, C), np.int64); Nv = np.zeros((k, C), np.int64)
    for c, cells in enumerate(cell_groups):
        if len(cells) == 0:
            continue
        S0[:, c] = _ix(D.like0, cells, muts).sum(axis=0)
        S1[:, c] = _ix(D.like1, cells, muts).sum(axis=0)
        No[:, c] = np.count_nonzero(_ix(D.total, cells, muts), axis=0)
        Nv[:, c] = np.count_nonzero(_ix(D.var, cells, muts), axis=0)
    return S0, S1, No, Nv


def linear_mut_assignment(D: DenseData, cellsA, muts, nobs_per_cluster=3):
    """linear_split.py:203-237."""
    muts = np.asarray(muts)
    num_obs = np.count_nonzero(_ix(D.total, cellsA, muts), axis=0)
    nobs = np.quantile(num_obs, 0.75)
    if nobs < nobs_per_cluster:
        return muts, None
    like0_array = _ix(D.like0, cellsA, muts).mean(axis=0)
    like1_array = _ix(D.like1, cellsA, muts).mean(axis=0)
    mask_pred = np.array(like0_array >= like1_array)
    mutsB = muts[mask_pred]
    mutsA = np.setdiff1d(muts, mutsB)
    return mutsA, mutsB


def branching_mut_assignment(D: DenseData, cellsA, cellsB, muts, nobs_per_cluster=3):
    """branching_split.py:293-348."""
    muts = np.asarray(muts)
    num_obsA = np.count_nonzero(_ix(D.total, cellsA, muts), axis=0)
    num_obsB = np.count_nonzero(_ix(D.total, cellsB, muts), axis=0)
    if np.quantile(num_obsA, 0.75) < nobs_per_cluster or np.quantile(num_obsB, 0.75) < nobs_per_cluster:
        return np.empty(0, int), np.empty(0, int), muts
    candA = _ix(D.like1, cellsA, muts).sum(axis=0) + _ix(D.like0, cellsB, muts).sum(axis=0)
    candB = _ix(D.like0, cellsA, muts).sum(axis=0) + _ix(D.like1, cellsB, muts).sum(axis=0)
    candC = _ix(D.like1, cellsA, muts).sum(axis=0) + _ix(D.like1, cellsB, muts).sum(axis=0)
    mx = np.maximum(candA, np.maximum(candB, candC))
    isA = candA == mx
    isB = (~isA) & (candB == mx)
    return muts[isA], muts[isB], muts[~isA & ~isB]


def branching_candidates(D: DenseData, cellsA, cellsB, muts):
    """branching_split.py:332-337: the three candidate log-likelihoods per SNV."""
    candA = _ix(D.like1, cellsA, muts).sum(axis=0) + _ix(D.like0, cellsB, muts).sum(axis=0)
    candB = _ix(D.like0, cellsA, muts).sum(axis=0) + _ix(D.like1, cellsB, muts).sum(axis=0)
    candC = _ix(D.like1, cellsA, muts).sum(axis=0) + _ix(D.like1, cellsB, muts).sum(axis=0)
    return candA, candB, candC


def weighted_init_vaf(D: DenseData, cells, muts):
    """linear_split.py:186-191: binary VAF per SNV and its normalisation."""
    bc = np.count_nonzero(_ix(D.var, cells, muts), axis=0)
    bt = np.count_nonzero(_ix(D.total, cells, muts), axis=0)
    vaf = bc / bt
    return vaf, vaf / np.sum(vaf)


# ---------------------------------------------------------------------------------------------- K2 family
def cell_table(D: DenseData, cells, mut_groups):
    Q = len(mut_groups)
    k = len(cells)
    A0 = np.zeros((k, Q)); A1 = np.zeros((k, Q))
    No = np.zeros((k, Q), np.int64); Nv = np.zeros((k, Q), np.int64)
    for q, muts in enumerate(mut_groups):
        if len(muts) == 0:
            continue
        A0[:, q] = _ix(D.like0, cells, muts).sum(axis=1)
        A1[:, q] = _ix(D.like1, cells, muts).sum(axis=1)
        No[:, q] = np.count_nonzero(_ix(D.total, cells, muts), axis=1)
        Nv[:, q] = np.count_nonzero(_ix(D.var, cells, muts), axis=1)
    return A0, A1, No, Nv


def calc_tmb(D: DenseData, cells, muts):
    """branching_split.py:147-173 / linear_split.py:301-310."""
    v = np.count_nonzero(_ix(D.var, cells, muts), axis=1).reshape(-1, 1)
    t = np.count_nonzero(_ix(D.total, cells, muts), axis=1).reshape(-1, 1)
    return v / t


def check_reads(D: DenseData, cells, muts, axis, nobs_per_cluster=3):
    """linear_split.py:393-396 / branching_split.py:428-431."""
    nobs = np.count_nonzero(_ix(D.total, cells, muts), axis=axis)
    return np.median(nobs) >= nobs_per_cluster


def linear_check_metrics(D: DenseData, all_cells, ca, cb, ma, mb, nobs_per_cluster=3):
    """linear_split.py:398-425."""
    f1 = np.nanmedian(np.count_nonzero(_ix(D.var, ca, mb), axis=1) / np.count_nonzero(_ix(D.total, ca, mb), axis=1))
    f2 = np.nanmedian(np.count_nonzero(_ix(D.var, all_cells, ma), axis=1) / np.count_nonzero(_ix(D.total, all_cells, ma), axis=1))
    feats = [check_reads(D, c, m, a, nobs_per_cluster) for a in [0, 1] for c in [ca, cb] for m in [ma, mb]]
    return f1, f2, all(feats)


def branching_check_metrics(D: DenseData, ca, cb, ma, mb, nobs_per_cluster=3):
    """branching_split.py:393-426 (feat2 duplicates feat1; feat3 is overwritten by all(feats))."""
    f1 = np.nanmedian(np.count_nonzero(_ix(D.var, ca, mb), axis=1) / np.count_nonzero(_ix(D.total, ca, mb), axis=1))
    feats = [check_reads(D, c, m, a, nobs_per_cluster) for a in [0, 1] for c in [ca, cb] for m in [ma, mb]]
    return f1, f1, all(feats), feats


# ---------------------------------------------------------------------------------------------- K3 family
def check_obs(D: DenseData, cells, muts):
    """utils.py:218-220; clonal_tree.py:1144-1146."""
    return int(np.count_nonzero(_ix(D.total, cells, muts)))


def block_sums(D: DenseData, cell_groups, mut_groups):
    Q, C = len(mut_groups), len(cell_groups)
    s0 = np.zeros((Q, C)); s1 = np.zeros((Q, C))
    no = np.zeros((Q, C), np.int64); nv = np.zeros((Q, C), np.int64)
    for q, muts in enumerate(mut_groups):
        for c, cells in enumerate(cell_groups):
            if len(muts) == 0 or len(cells) == 0:
                continue
            s0[q, c] = _ix(D.like0, cells, muts).sum()
            s1[q, c] = _ix(D.like1, cells, muts).sum()
            no[q, c] = np.count_nonzero(_ix(D.total, cells, muts))
            nv[q, c] = np.count_nonzero(_ix(D.var, cells, muts))
    return s0, s1, no, nv


def linear_norm_likelihood(D: DenseData, all_cells, cellsA, mutsA, mutsB):
    """linear_split.py:523-556."""
    total_like = _ix(D.like0, cellsA, mutsB).sum() + _ix(D.like1, all_cells, mutsA).sum()
    total = np.count_nonzero(_ix(D.like0, cellsA, mutsB)) + np.count_nonzero(_ix(D.like1, all_cells, mutsA))
    return total_like / total


def branching_norm_likelihood(D: DenseData, all_cells, cellsA, cellsB, mutsA, mutsB, mutsC):
    """branching_split.py:434-468."""
    mutsC = np.asarray(mutsC).astype(int)
    total_like = _ix(D.like1, all_cells, mutsC).sum() + _ix(D.like0, cellsA, mutsB).sum() + _ix(D.like0, cellsB, mutsA).sum()
    total = (np.count_nonzero(_ix(D.like0, all_cells, mutsC)) + np.count_nonzero(_ix(D.like0, cellsA, mutsB))
             + np.count_nonzero(_ix(D.like0, cellsB, mutsA)))
    return total_like / total


def seed_strip(D: DenseData, cells, muts):
    """clonal_tree.py:1135-1142 (two stages, order matters)."""
    cells = np.asarray(cells); muts = np.asarray(muts)
    vs = _ix(D.var, cells, muts).sum(axis=0)
    muts = np.setdiff1d(muts, muts[vs == 0])
    vc = _ix(D.var, cells, muts).sum(axis=1)
    cells = np.setdiff1d(cells, cells[vc == 0])
    return cells, muts


# ---------------------------------------------------------------------------------------------- tree level
def variant_likelihood_by_node(D: DenseData, cells, present_muts):
    """clonal_tree.py:932-944 with presence_by_node 343-358."""
    cells = np.asarray(cells, np.int64)
    y = np.zeros((len(cells), D.m), dtype=int)
    if len(cells):
        y[:, np.asarray(present_muts, np.int64)] = 1
    return np.multiply(1 - y, D.like0[cells, :]).sum() + np.multiply(y, D.like1[cells, :]).sum()


def snv_node_scores(D: DenseData, s, c_dict):
    """clonal_tree.py:651-661: per-node likelihood of SNV s, c_dict[node] = 0/1 vector over cells."""
    return {k: np.dot(y, D.like1[:, s]) + np.dot(1 - y, D.like0[:, s]) for k, y in c_dict.items()}


def cell_node_scores(D: DenseData, c, y_dict):
    """clonal_tree.py:614-635: per-node likelihood of cell c, y_dict[node] = 0/1 vector over SNVs."""
    return {k: np.dot(y, D.like1[c, :]) + np.dot(1 - y, D.like0[c, :]) for k, y in y_dict.items()}


def find_bad_snvs(D: DenseData, clade_cells, muts, q):
    """clonal_tree.py:573-600."""
    cells = np.asarray(clade_cells).astype(int)
    muts = np.asarray(muts).astype(int)
    tot = np.count_nonzero(_ix(D.total, cells, muts), axis=0).reshape(-1)
    na = muts[tot == 0]
    muts = np.setdiff1d(muts, na).astype(int)
    tot = np.count_nonzero(_ix(D.total, cells, muts), axis=0).reshape(-1)
    mc = np.count_nonzero(_ix(D.var, cells, muts), axis=0).reshape(-1)
    vaf = mc / tot
    bad = muts[vaf <= np.quantile(vaf, q)] if len(vaf) else np.empty(0, int)
    return np.union1d(bad, na)


# ---------------------------------------------------------------------------------------------- K4 / K5
def gauss_node_loglik(X, cells):
    """clonal_tree.py:395-407 with scipy.stats.multivariate_normal(allow_singular=True) restated for a DIAGONAL
    covariance (scipy/stats/_multivariate.py: _eigvalsh_to_eps, _pinv_1d, _PSD, _logpdf, frozen logpdf support rule).
    Returns (sum, per-cell)."""
    X = np.asarray(X, float)
    x = X[np.asarray(cells, np.int64), :]
    mu = x.mean(axis=0)
    var = np.var(x, axis=0)
    eps = 1e6 * np.finfo("d").eps * np.max(np.abs(var))
    keep = var > eps
    rank = int(keep.sum())
    log_pdet = np.sum(np.log(var[keep]))
    dev = x - mu
    maha = np.sum(np.square(dev[:, keep] * np.sqrt(1.0 / var[keep])), axis=-1)
    out = -0.5 * (rank * np.log(2 * np.pi) + log_pdet + maha)
    if rank < x.shape[1]:
        resid = np.linalg.norm(dev[:, ~keep], axis=-1)
        out[~(resid < 1e3 * eps)] = -np.inf
    return out.sum(), out


def pairwise_dist(X):
    """phertilizer.py:286."""
    return squareform(pdist(np.asarray(X, float), metric="euclidean"))
