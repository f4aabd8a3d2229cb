"""Branching tree operation (a seed splits into two sibling clones under a common parent).

Follows the reference's `Branching_split` (phertilizer/branching_split.py:88-488); reductions on the GPU store:

  calc_tmb          (147-173)  -> ObservationStore.cell_table (both SNV sets in one K2 pass)
  mut_assignment    (293-348)  -> ObservationStore.assign_branching (K1 + 3-way K1a; the reference's Python loop over
                                   all SNVs is the fused epilogue)
  check_metrics     (393-431)  -> K1 / K2 count tables (incl. the reference's quirks: feat2 == feat1, feat3 = all(feats))
  compute_norm_likelihood (434-468) -> ObservationStore.block_sums (K3)
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from .clonal_tree import BranchingTree
from .clonal_tree_list import ClonalTreeList
from .utils import impute_mut_features, setdiff1d, split_by_labels

np.seterr(invalid="ignore", divide="ignore")


class Branching_split:
    low_cmb: float = 0.075   # unused class attributes of the reference (branching_split.py:117-118)
    high_cmb: float = 0.15

    def __init__(self, data, seed, rng, params):
        self.data = data
        self.store = data.store
        self.cells = seed.cells
        self.muts = seed.muts
        self.rng = rng
        self.np_rng = np.random.RandomState(params.seed)
        self.starts = params.starts
        self.iterations = params.iterations
        self.use_copy_kernel = params.use_copy_kernel
        self.radius = params.radius
        self.params = params

    def random_init(self, arr):
        """Uniform three-way split (branching_split.py:121-145)."""
        draws = self.rng.random(arr.shape[0])
        arrA = arr[draws < 1 / 3]
        arrB = arr[draws > 2 / 3]
        return arrA, arrB, setdiff1d(arr, np.concatenate([arrA, arrB]))

    def calc_tmb(self, cells, muts):
        t = self.store.cell_table(self.store.snv_labels(muts), 1, cells, want=("Nobs", "Nvar"))
        return (t["Nvar"][:, 0] / t["Nobs"][:, 0]).reshape(-1, 1)

    def create_affinity(self, mutsA, mutsB, cells):
        """Features of the affinity (branching_split.py:175-215): per-cell burden over mutsA and mutsB, imputed."""
        if len(mutsA) == 0 or len(mutsB) == 0:
            return None
        t = self.store.cell_table(self.store.snv_labels(mutsA, mutsB), 2, cells, want=("Nobs", "Nvar"))
        mut_features = t["Nvar"] / t["Nobs"]          # columns: tmb over mutsA, tmb over mutsB
        mut_features = impute_mut_features(self.data.copy_distance, mut_features, cells)
        if np.any(np.isnan(mut_features)):
            return None
        return mut_features

    def cluster_cells(self, mutsA, mutsB, cells):
        mut_features = self.create_affinity(mutsA, mutsB, cells)
        if mut_features is None:
            return cells, np.empty(shape=0, dtype=int)
        if len(cells) == 1:                     # utils.py:133-135
            cells1, cells2 = cells, np.empty(shape=0)
        else:
            # branching kernel: copy-distance entries >= r are cut (branching_split.py:228)
            labels = self.data.copy_distance.min_cut(cells, mut_features, self.use_copy_kernel, self.radius, True, self.np_rng)
            if labels is None:
                print("Terminating cell clustering due to NANs in affinity matrix")
                return cells, np.empty(shape=0, dtype=int)
            cells1, cells2 = split_by_labels(cells, labels)
        feats = pd.DataFrame(mut_features, cells)
        avg = feats[feats.index.isin(cells1)].mean(axis=0)
        # the cluster whose mutsA burden exceeds its mutsB burden becomes cellsA (branching_split.py:275-286)
        return (cells1, cells2) if avg[0] > avg[1] else (cells2, cells1)

    def mut_assignment(self, cellsA, cellsB):
        label, nobs = self.store.assign_branching(self.store.cell_labels(cellsA, cellsB), self.muts)
        lim = self.params.nobs_per_cluster
        if np.quantile(nobs[:, 0], 0.75) < lim or np.quantile(nobs[:, 1], 0.75) < lim:
            return np.empty(shape=0, dtype=int), np.empty(shape=0, dtype=int), self.muts
        muts = np.asarray(self.muts)
        # the reference appends to python lists and wraps them in np.array: dtype float64 when a list stays empty
        parts = (muts[label == k] for k in (1, 2, 3))
        return tuple(part if len(part) else np.array([]) for part in parts)

    def check_metrics(self, ca, cb, ma, mb, mc):
        st = self.store
        ca_t = st.cell_table(st.snv_labels(ma, mb), 2, ca, want=("Nobs", "Nvar"))
        cb_t = st.cell_table(st.snv_labels(ma, mb), 2, cb, want=("Nobs",))
        feat1 = np.nanmedian(ca_t["Nvar"][:, 1] / ca_t["Nobs"][:, 1])
        feat2 = feat1                                      # branching_split.py:400-404
        both = np.concatenate([np.asarray(ma, np.int64), np.asarray(mb, np.int64)])
        s_t = st.snv_table(st.cell_labels(ca, cb), 2, both, want=("Nobs",))["Nobs"]
        na = len(ma)
        lim = self.params.nobs_per_cluster
        feats = []
        for axis in (0, 1):
            for ci, ctab in ((0, ca_t), (1, cb_t)):
                for mi in (0, 1):
                    nobs = (s_t[:na, ci] if mi == 0 else s_t[na:, ci]) if axis == 0 else ctab["Nobs"][:, mi]
                    feats.append(np.median(nobs) >= lim)
        return feat1, feat2, all(feats), feats             # feat3 is overwritten by all(feats) (414-426)

    def compute_norm_likelihood(self, cellsA, cellsB, mutsA, mutsB, mutsC):
        st = self.store
        mutsC = np.asarray(mutsC).astype(int)
        rest = setdiff1d(self.cells, np.concatenate([cellsA, cellsB]))
        s0, s1, nobs, _ = st.block_sums(st.cell_labels(cellsA, cellsB, rest), 3, st.snv_labels(mutsA, mutsB, mutsC), 3)
        total_like = ((s1[2, 0] + s1[2, 1]) + s1[2, 2]) + s0[1, 0] + s0[0, 1]
        total = ((nobs[2, 0] + nobs[2, 1]) + nobs[2, 2]) + nobs[1, 0] + nobs[0, 1]
        return total_like / total

    def run(self):
        cells, muts = self.cells, self.muts
        mutsA, mutsB, mutsC = self.random_init(muts)
        oldA = np.empty(shape=0, dtype=int)
        cellsA = cellsB = np.empty(shape=0, dtype=int)   # the reference raises UnboundLocalError here when an
        for _ in range(self.iterations):                # initial third is empty; treated as "no tree" instead
            if mutsB.shape[0] == 0 or mutsA.shape[0] == 0:
                break
            cellsA, cellsB = self.cluster_cells(mutsA, mutsB, cells)
            if len(cellsB) == 0 or len(cellsA) == 0 or np.array_equal(np.sort(cellsA), np.sort(oldA)):
                break
            oldA = cellsA
            mutsA, mutsB, mutsC = self.mut_assignment(cellsA, cellsB)
        if len(cellsA) > 0 and len(cellsB) > 0 and len(mutsA) > 0 and len(mutsB) > 0:
            norm_like = self.compute_norm_likelihood(cellsA, cellsB, mutsA, mutsB, mutsC)
            cand = BranchingTree(cellsA, cellsB, mutsA, mutsB, mutsC)
            f1, f2, f3, f4 = self.check_metrics(cellsA, cellsB, mutsA, mutsB, mutsC)
            p = self.params
            if f1 <= p.low_cmb and f2 <= p.low_cmb and f3 >= p.high_cmb and f4:
                if norm_like > self.best_norm_like:
                    self.best_norm_like = norm_like
                    self.best_tree = cand

    def sprout(self):
        self.cand_trees = ClonalTreeList()
        self.best_norm_like = -np.inf
        self.best_tree = None
        for _ in range(self.starts):
            self.run()
            self.use_copy_kernel = not self.use_copy_kernel
        return self.best_tree
