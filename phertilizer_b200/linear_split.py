"""Linear tree operation (parent -> child split of a seed).

Control flow, RNG draws and acceptance rules follow the reference's `Linear_split` (phertilizer/linear_split.py:100-787,
live methods only); every reduction over the observation matrices goes to the GPU store:

  random_weighted_init_array (163-201)  -> per-SNV Nvar/Nobs            : ObservationStore.snv_table      (K1)
  mut_assignment            (203-237)  -> fused q75 inputs + argmax    : ObservationStore.assign_linear  (K1+K1a)
  create_affinity features  (301-310)  -> per-cell Nvar/Nobs over mutsB : ObservationStore.cell_table     (K2)
  check_metrics / check_reads (393-425)-> K1 / K2 count tables
  compute_norm_likelihood   (523-556)  -> block sums                    : ObservationStore.block_sums     (K3)

The cell clustering (308-390): the affinity and its normalised Laplacian are built on the GPU from the mutation-burden
features (`cluster.CopyDistance.min_cut`); the eigenvectors come from scipy's LOBPCG as in the reference.
"""
from __future__ import annotations

import logging

import numpy as np
import pandas as pd

from .clonal_tree import LinearTree
from .clonal_tree_list import ClonalTreeList
from .utils import check_obs, impute_mut_features, setdiff1d, split_by_labels

np.seterr(invalid="ignore", divide="ignore")


class Linear_split:
    def __init__(self, data, seed, rng, params):
        self.rng = rng
        self.np_rng = np.random.RandomState(params.seed)
        self.data = data
        self.store = data.store
        self.cells = seed.cells
        self.muts = seed.muts
        self.params = params
        self.radius = params.radius
        self.minobs = params.minobs
        self.use_copy_kernel = params.use_copy_kernel
        self.iterations = params.iterations
        self.starts = params.starts
        self.copy_distance_matrix = data.copy_distance

    # ------------------------------------------------------------------ initialisation
    def random_init_array(self, array, p):
        draws = self.rng.random(len(array))
        return array[draws <= p], array[draws > p]

    def random_weighted_init_array(self, cells, muts, p):
        t = self.store.snv_table(self.store.cell_labels(cells), 1, muts, want=("Nobs", "Nvar"))
        vaf = t["Nvar"][:, 0] / t["Nobs"][:, 0]
        vaf_prob = vaf / np.sum(vaf)
        num_mutsA = int(np.floor(p * muts.shape[0]))
        if num_mutsA < np.count_nonzero(vaf_prob) and not np.any(np.isnan(vaf_prob)):
            mutsA = self.rng.choice(muts, size=num_mutsA, p=vaf_prob, replace=False)
            return mutsA, setdiff1d(muts, mutsA)
        return self.random_init_array(muts, p)

    # ------------------------------------------------------------------ SNV step
    def mut_assignment(self, cellsA, muts):
        label, nobs = self.store.assign_linear(self.store.cell_labels(cellsA), len(cellsA), muts)
        if np.quantile(nobs, 0.75) < self.params.nobs_per_cluster:
            return muts, None
        mutsB = muts[label == 2]
        return setdiff1d(muts, mutsB), mutsB

    # ------------------------------------------------------------------ cell step
    def create_affinity(self, cells, mutsB=None):
        """Per-cell mutation burden over mutsB (linear_split.py:301-310) -> the features of the affinity; the affinity
        matrix itself (308-335) is formed on the GPU by `cluster_cells`."""
        if mutsB is None:
            return None, None
        t = self.store.cell_table(self.store.snv_labels(mutsB), 1, cells, want=("Nobs", "Nvar"))
        mb_mut_count = t["Nvar"][:, 0].astype(np.int64)
        tmb = (mb_mut_count / t["Nobs"][:, 0]).reshape(-1, 1)
        if np.any(np.isnan(tmb)):
            tmb = impute_mut_features(self.copy_distance_matrix, tmb, cells)
        return tmb, pd.Series(mb_mut_count.reshape(-1), index=cells)

    def cluster_cells(self, cells, mutsB):
        tmb, counts = self.create_affinity(cells, mutsB)
        if tmb is None:
            return cells, np.empty(shape=0, dtype=int), None
        stats = {}
        if len(cells) == 1:                     # utils.py:133-135
            cells1, cells2 = cells, np.empty(shape=0)
        else:
            # linear kernel: copy-distance entries > r are cut (linear_split.py:329)
            labels = self.copy_distance_matrix.min_cut(cells, tmb, self.use_copy_kernel, self.radius, False, self.np_rng)
            if labels is None:
                print("Terminating cell clustering due to NANs in affinity matrix")
                return cells, np.empty(shape=0, dtype=int), None
            cells1, cells2 = split_by_labels(cells, labels)
            stats.update(jump_perc=0.08, first_gap=0.03, largest_gap=0.5, spectral_num_k=2)
        mean1, mean2 = counts[cells1].mean(), counts[cells2].mean()
        if np.isnan(mean1) or np.isnan(mean2):
            print("warning mut count mean is Nan")
        if mean1 < mean2:
            cellsA, cellsB = cells1, cells2
        else:
            cellsA, cellsB = cells2, cells1
        logging.info(f"Mb Count a: {min(mean1, mean2)} Mb Count b: {max(mean1, mean2)}")
        stats["norm_count_diff"] = abs(mean1 - mean2)
        stats["ca_mb_count"] = min(mean1, mean2)
        return cellsA, cellsB, stats

    # ------------------------------------------------------------------ acceptance metrics
    def check_metrics(self, ca, cb, ma, mb):
        st = self.store
        ca_t = st.cell_table(st.snv_labels(ma, mb), 2, ca, want=("Nobs", "Nvar"))
        cb_t = st.cell_table(st.snv_labels(ma, mb), 2, cb, want=("Nobs",))
        feat1 = np.nanmedian(ca_t["Nvar"][:, 1] / ca_t["Nobs"][:, 1])
        all_t = st.cell_table(st.snv_labels(ma), 1, self.cells, want=("Nobs", "Nvar"))
        feat2 = np.nanmedian(all_t["Nvar"][:, 0] / all_t["Nobs"][:, 0])
        # per-SNV observation counts over ca / cb (axis 0) and per-cell counts over ma / mb (axis 1)
        both = np.concatenate([ma, mb])
        s_t = st.snv_table(st.cell_labels(ca, cb), 2, both, want=("Nobs",))["Nobs"]
        na = len(ma)
        lim = self.params.nobs_per_cluster
        feats = []
        for axis in (0, 1):
            for ci, ctab in ((0, ca_t), (1, cb_t)):
                for mi in (0, 1):
                    if axis == 0:
                        nobs = s_t[:na, ci] if mi == 0 else s_t[na:, ci]
                    else:
                        nobs = ctab["Nobs"][:, mi]
                    feats.append(np.median(nobs) >= lim)
        return feat1, feat2, all(feats)

    def compute_norm_likelihood(self, cellsA, cellsB, mutsA, mutsB):
        """[sum like0(cellsA x mutsB) + sum like1(seed cells x mutsA)] / observed entries of the same blocks."""
        st = self.store
        rest = setdiff1d(self.cells, cellsA)
        s0, s1, nobs, _ = st.block_sums(st.cell_labels(cellsA, rest), 2, st.snv_labels(mutsA, mutsB), 2)
        total_like = s0[1, 0] + (s1[0, 0] + s1[0, 1])
        total = nobs[1, 0] + (nobs[0, 0] + nobs[0, 1])
        return total_like / total

    @staticmethod
    def best_norm_like(tree_list, norm_like_list):
        if tree_list.size() == 0:
            return None
        return tree_list.index_tree(np.argmax(np.array(norm_like_list)))

    # ------------------------------------------------------------------ main loop
    def run(self, cells, muts, p, parent_norm=-np.inf):
        if check_obs(cells, muts, self.store) < self.minobs:
            print("not enough observations, terminating")
            return
        candidates = ClonalTreeList()
        norms = []
        for _ in range(self.starts):
            mutsA, mutsB = self.random_weighted_init_array(cells, muts, p)
            oldA = cells
            for _j in range(self.iterations):
                cellsA, cellsB, _stats = self.cluster_cells(cells, mutsB)
                if len(cellsB) == 0 or np.array_equal(np.sort(cellsA), np.sort(oldA)):
                    break
                oldA = cellsA
                mutsA, mutsB = self.mut_assignment(cellsA, muts)
            if len(cellsA) > 0 and len(cellsB) > 0:
                cellsB_tree = setdiff1d(self.cells, cellsA)
                mutsB_tree = setdiff1d(self.muts, mutsA)
                f1, f2, f3 = self.check_metrics(cellsA, cellsB_tree, mutsA, mutsB_tree)
                if f1 <= self.params.low_cmb and f2 >= self.params.high_cmb and f3:
                    candidates.insert(LinearTree(cellsA, cellsB_tree, mutsA, mutsB_tree))
                    norms.append(self.compute_norm_likelihood(cellsA, cellsB, mutsA, mutsB_tree))
            self.use_copy_kernel = not self.use_copy_kernel
        best = self.best_norm_like(candidates, norms)
        if best is not None:
            like_norm = self.compute_norm_likelihood(best.get_tip_cells(0), best.get_tip_cells(1), best.get_tip_muts(0),
                                                     best.get_tip_muts(1))
            print(f"like norm {like_norm}: parent norm {parent_norm}")
            if like_norm > parent_norm:
                self.cand_splits.insert(best)
                self.norm_like_list.append(like_norm)
                self.run(best.get_tip_cells(0), best.get_tip_muts(0), p, parent_norm=like_norm)

    def sprout(self):
        self.norm_like_list = []
        self.cand_splits = ClonalTreeList()
        self.run(self.cells, self.muts, p=0.5)
        if self.cand_splits.size() == 0:
            return self.cand_splits, None
        return self.cand_splits, self.cand_splits.index_tree(np.argmax(np.array(self.norm_like_list)))
