"""`Phertilizer`: grows a clonal tree from ultra-low-coverage single-cell DNA sequencing data.

Drop-in for the reference's `phertilizer.phertilizer.Phertilizer` (phertilizer/phertilizer.py:163-745): same constructor
and `phertilize` arguments, same return values, same RNG streams and the same planting / growing control flow on the
host.  What changes is the data layer: instead of densifying four n x m matrices (lines 295-306) the observations are
handed to the CUDA library once (`ObservationStore.from_coo`) and every reduction over them runs on the B200.
"""
from __future__ import annotations

from collections import deque
from copy import deepcopy

import numpy as np
import pandas as pd

from . import branching_split as bs
from . import linear_split as ls
from .clonal_tree import IdentityTree, Seed
from .clonal_tree_list import ClonalTreeList
from .data import Data
from .params import Params
from .cluster import CopyDistance
from .store import ObservationStore
from .utils import check_obs, power_calc, setdiff1d


def _factorize(col):
    """(codes, uniques) with codes[i] the index of col[i] in uniques (any bijection: callers re-rank the uniques).
    Small non-negative integer ids -- the usual case -- take a counting pass instead of pandas' hash table."""
    if getattr(col.dtype, "kind", "O") in "iu" and len(col):  # (string columns: no object-array detour, straight to pandas)
        a = col.to_numpy()
        lo, hi = int(a.min()), int(a.max())
        if lo >= 0 and hi < (1 << 27):
            seen = np.bincount(a, minlength=hi + 1) > 0
            rank = np.cumsum(seen, dtype=np.int32)
            rank -= 1
            return rank[a], np.nonzero(seen)[0].astype(a.dtype)
    return pd.factorize(col, sort=False)


def index_observations(variant_count_data):
    """DataFrame[chr, mutation_id, cell_id, base, var, total] -> (cell labels, SNV labels, cell index, SNV index, var,
    total) of the OBSERVED entries (phertilizer.py:235-253, 295-306).

    The label sets come from ALL rows, as in the reference.  Rows with total == 0 are not observations there: its dense
    matrices are `unstack(fill_value=0)` and "observed" is `count_nonzero(total)` everywhere (phertilizer.py:153-161,
    linear_split.py:186-188), so such a row is indistinguishable from a missing one -- they are dropped here, after the
    labels are fixed.  A (cell, mutation) pair listed twice makes the reference's `unstack` raise; so does this."""
    # labels -> dense indices: sorted unique cell ids and "chr_mutation" strings (phertilizer.py:235-253).
    # Same labels and order as the reference; the per-observation string concatenation and label lookups (tens of
    # seconds at 2.4e7 observations) are done on the distinct (chr, mutation_id) pairs / cell ids only.
    vc = variant_count_data
    # Every per-observation pass below is one streaming numpy operation on 32-bit indices (the store takes int32).
    cell_codes, cell_uniques = _factorize(vc["cell_id"])
    cell_labels = np.sort(np.asarray(cell_uniques))
    cell_rank = pd.Series(np.arange(len(cell_labels)), index=cell_labels)[np.asarray(cell_uniques)].to_numpy()
    cell_idx = cell_rank.astype(np.int32)[cell_codes]
    del cell_codes
    chr_codes, chr_uniques = _factorize(vc["chr"])
    mid_codes, mid_uniques = _factorize(vc["mutation_id"])
    M = max(len(mid_uniques), 1)
    n_pair_space = len(chr_uniques) * M
    wide = np.int64 if n_pair_space >= (1 << 31) else np.int32
    pair = chr_codes.astype(wide)
    pair *= wide(M)
    pair += mid_codes.astype(wide, copy=False)
    del chr_codes, mid_codes
    counting = n_pair_space <= (1 << 27)
    if counting:
        # counting pass instead of sorting all observations
        pair_u = np.nonzero(np.bincount(pair, minlength=n_pair_space))[0]
    else:
        pair_u, pair_inv = np.unique(pair, return_inverse=True)
    pair_str = (pd.Series(np.asarray(chr_uniques)[pair_u // M]).astype("str") + "_"
                + pd.Series(np.asarray(mid_uniques)[pair_u % M]).astype(str)).to_numpy()
    mut_labels = np.sort(pd.unique(pair_str))
    mut_rank = pd.Series(np.arange(len(mut_labels)), index=mut_labels)[pair_str].to_numpy().astype(np.int32)
    if counting:
        lut = np.zeros(n_pair_space, dtype=np.int32)   # pair -> index of its "chr_mutation" label, one gather
        lut[pair_u] = mut_rank
        mut_idx = lut[pair]
    else:
        mut_idx = mut_rank[pair_inv]
    del pair
    var = vc["var"].to_numpy()
    total = vc["total"].to_numpy()
    if len(total) and bool((total <= 0).any()):
        if bool(((total <= 0) & (var > 0)).any()):
            raise ValueError("variant_count_data has rows with var > 0 and total == 0")
        keep = total > 0
        cell_idx, mut_idx, var, total = cell_idx[keep], mut_idx[keep], var[keep], total[keep]
    # duplicates: the store rejects them on the device (ph_create sorts the (SNV, cell) keys anyway)
    return cell_labels, mut_labels, cell_idx, mut_idx, var, total


def prepare_inputs(variant_count_data, bin_count_data):
    """Everything `Phertilizer.__init__` derives from the two input frames on the host: index_observations() plus the binned
    read counts in cell-index order (phertilizer.py:259-263)."""
    cell_labels, mut_labels, cell_idx, mut_idx, var, total = index_observations(variant_count_data)
    cell_series = pd.Series(data=np.arange(cell_labels.shape[0], dtype="int"), index=cell_labels)
    bc = bin_count_data.copy()
    bc["cell_index"] = cell_series[bc["cell"]].values
    bc = bc.sort_values(by=["cell_index"]).drop(["cell", "cell_index"], axis=1).to_numpy()
    return cell_labels, mut_labels, cell_idx, mut_idx, var, total, bc


def _embed(bin_counts: np.ndarray, dim_reduce: bool) -> np.ndarray:
    """Row-normalise + UMAP (phertilizer.py:267-279) or the raw bins (284-285).  UMAP is third-party and stays so."""
    if not dim_reduce:
        return bin_counts
    try:
        import umap
    except ImportError as e:
        raise ImportError("umap-learn is not installed: pass dim_reduce=False (--no-umap) or a precomputed embedding") from e
    norm = bin_counts / (bin_counts.sum(axis=1).reshape(-1, 1))
    reducer = umap.UMAP(n_neighbors=40, min_dist=0, n_components=2, spread=1, metric="manhattan", random_state=55)
    return reducer.fit_transform(norm)


class Phertilizer:
    def __init__(self, variant_count_data, bin_count_data, alpha=0.001, max_copies=5, mode="var", dim_reduce=True,
                 device=None, device_group=None, _indexed=None):
        """`device_group`: a torch.distributed process group (or True for the default group) whose ranks each drive one
        GPU.  The observations are then sharded by cells across the ranks (`sharded_store.CellShardedStore`)
        and EVERY rank must construct the object and make the same calls: the tree-growing control flow runs redundantly
        on all ranks, on identical data and RNG streams, and returns the same tree everywhere."""
        use_read_depth = mode != "var"
        self.mapping_list = None
        self.explored_seeds = None
        import os

        self.device_group = None
        world = 1
        if device_group is not None:
            import torch.distributed as dist

            self.device_group = None if device_group is True else device_group
            world = dist.get_world_size(self.device_group)
            if device is None:
                device = int(os.environ.get("PHERTILIZER_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        device = int(os.environ.get("PHERTILIZER_DEVICE", "0")) if device is None else device

        # `_indexed` (fanout.py): the label -> index pass and the cell-ordered bin matrix were done once by the parent of
        # the `--runs` workers and arrive as plain arrays (memory-mapped), instead of every worker repeating them
        if _indexed is not None:
            cell_labels, mut_labels, cell_idx, mut_idx, var, total, bc = _indexed
        else:
            cell_labels, mut_labels, cell_idx, mut_idx, var, total, bc = prepare_inputs(variant_count_data, bin_count_data)
        self.cells = np.arange(cell_labels.shape[0], dtype="int")
        self.muts = np.arange(mut_labels.shape[0], dtype="int")
        cell_lookup = pd.Series(data=cell_labels, index=self.cells)
        mut_lookup = pd.Series(data=mut_labels, index=self.muts)
        embedding = _embed(bc, dim_reduce)
        if world > 1:
            from .cluster import ShardedCopyDistance

            copy_distance = ShardedCopyDistance(np.asarray(embedding, dtype=np.float64), device=device, group=self.device_group,
                                                min_rows=int(os.environ.get("PH_SHARD_MIN_ROWS", "4096")))
        else:
            copy_distance = CopyDistance(np.asarray(embedding, dtype=np.float64), device=device)

        # per-observation likelihoods become a (var,total) -> (like0, like1) table inside the store (295-306)
        if world > 1:
            import torch.distributed as dist

            from .sharded_store import CellShardedStore

            store = CellShardedStore.from_coo(len(self.cells), len(self.muts), cell_idx, mut_idx, var, total,
                                              dist.get_rank(self.device_group), world, device, alpha=alpha,
                                              max_copies=max_copies, group=self.device_group)
        else:
            store = ObservationStore.from_coo(len(self.cells), len(self.muts), cell_idx, mut_idx, var, total, alpha=alpha,
                                              max_copies=max_copies, device=device)
        if hasattr(store, "set_embedding"):
            store.set_embedding(embedding)       # tree scoring (K4) reads it hundreds of times
        self.data = Data(cell_lookup, mut_lookup, store, embedding, copy_distance, use_read_depth)

    # ------------------------------------------------------------------ small accessors
    def save_embedding(self, fname):
        cells = self.data.cell_lookup.sort_index().values.reshape(-1, 1)
        pd.DataFrame(np.hstack([cells, self.data.read_depth]), columns=["cell", "V1", "V2"]).to_csv(fname, index=False)

    def get_id_mappings(self):
        return self.data.cell_lookup, self.data.mut_lookup

    # ------------------------------------------------------------------ main entry
    def phertilize(self, max_iterations=10, starts=3, seed=1026, radius=0.8, low_cmb=0.05, high_cmb=0.15,
                   nobs_per_cluster=3.0, gamma=0.95, d=7, use_copy_kernel=False, post_process=False):
        """Enumerate clonal trees and return (best tree, ClonalTreeList of all trees, Series of normalised
        log-likelihoods) -- phertilizer.py:350-460."""
        Nmax = check_obs(self.cells, self.muts, self.data.store)
        minobs = power_calc(d, gamma, 6, Nmax)
        self.rng = np.random.default_rng(seed)
        self.params = Params(starts, max_iterations, radius, minobs, seed, post_process, use_copy_kernel, low_cmb,
                             high_cmb, nobs_per_cluster)
        self.init_seed = Seed(self.cells, self.muts, key=0)
        seed_list = deque([self.init_seed])
        self.mapping_list = {}
        self.explored_seeds = {}
        print(f"\nPhertilizing a tree with n: {len(self.cells)} cells and m: {len(self.muts)} SNVs")
        print("\nStarting planting phase....")
        self.planting(seed_list)
        print("\nPlanting phase complete!")
        print("\nStarting grow phase....")
        pre_proc_trees = self.grow()
        print("\nGrowing phase complete!")
        print("\nIdentifying the maximum likelihood clonal tree....")
        optimal_tree, loglikelihoods = pre_proc_trees.find_best_tree(self.data)
        if optimal_tree is None:
            optimal_tree = IdentityTree(self.cells, self.muts)
        if post_process:
            print("Post-processing...")
            optimal_tree.post_process(self.data)
        return optimal_tree, pre_proc_trees, loglikelihoods

    # ------------------------------------------------------------------ planting: elementary operations on every seed
    def planting(self, seed_list):
        """LIFO over seeds; keys advance only for seeds with enough observations (phertilizer.py:463-535)."""
        key = 0
        while len(seed_list) > 0:
            curr = seed_list.pop()
            curr.set_key(key)
            print(curr)
            self.mapping_list[key] = []
            self.explored_seeds[key] = deepcopy(curr)
            if curr.has_linear():
                operations = [self.sprout_branching]
                self.mapping_list[key].append(curr.linear_tree)
            elif curr.has_branching():
                operations = [self.sprout_linear]
                self.mapping_list[key].append(curr.branching_tree)
            else:
                operations = [self.sprout_linear, self.sprout_branching]
            curr.strip(self.data.store)          # mutates the seed AFTER it was memoised (reference quirk 8)
            if curr.count_obs(self.data.store) < self.params.minobs:
                print("Insufficient number of observations, seed is terminal")
                continue
            for op in operations:
                tree, new_seeds = op(curr)
                if tree is not None:
                    print("\n")
                    print(tree)
                    self.mapping_list[key].append(tree)
                    seed_list += new_seeds
                else:
                    print("No inferred elementary tree.")
            print(f"Number of subproblems remaining: {len(seed_list)}")
            key += 1

    def lookup_seed_key(self, seed):
        for s in self.explored_seeds:
            if self.explored_seeds[s] == seed:
                return s

    # ------------------------------------------------------------------ growing: enumerate trees from the memo
    def grow(self, max_trees=np.inf):
        """phertilizer.py:558-614."""
        all_trees = ClonalTreeList()
        frontier = ClonalTreeList()
        for t in self.mapping_list[0]:
            all_trees.insert(t)
            if type(t) is not IdentityTree:
                frontier.insert(t)
            while frontier.has_trees():
                if all_trees.size() > max_trees:
                    break
                curr = frontier.pop_tree()
                for leaf in curr.get_leaves():
                    seed = curr.get_seed_by_node(leaf)
                    if seed is None:
                        continue
                    seed_key = self.lookup_seed_key(seed)
                    if seed_key is None:
                        continue
                    for sub in self.mapping_list[seed_key]:
                        if type(sub) is IdentityTree:
                            continue
                        merged = deepcopy(curr)
                        merged.merge(sub, leaf)
                        if not all_trees.contains(merged):
                            all_trees.insert(merged)
                        if not frontier.contains(merged):
                            frontier.insert(merged)
        return all_trees

    # ------------------------------------------------------------------ elementary operations
    def sprout_branching(self, seed):
        print("\nSprouting a branching tree from seed:")
        print(seed)
        tree = bs.Branching_split(self.data, seed, self.rng, self.params).sprout()
        return tree, (tree.get_seeds() if tree is not None else [])

    def sprout_linear(self, seed):
        """Linear operation; a chain of recursive splits is unrolled into nested seeds (phertilizer.py:655-724)."""
        print("\nSprouting a linear tree from seed:")
        print(seed)
        tree_list, best = ls.Linear_split(self.data, seed, self.rng, self.params).sprout()
        if best is None:
            return None, []
        seed_list = best.get_seeds()
        seed.set_linear(best)
        idx = tree_list.size() - 2
        if idx >= 0:
            ca_cells, ma_muts = best.get_tip_cells(0), best.get_tip_muts(0)
            curr_seed = seed_list[0]
            while idx >= 0:
                nxt = tree_list.index_tree(idx)
                cA, mA = nxt.get_tip_cells(0), nxt.get_tip_muts(0)
                nxt.cell_mapping[0] = {0: setdiff1d(cA, ca_cells)}
                nxt.mut_mapping[0] = setdiff1d(mA, ma_muts)
                curr_seed.set_linear(nxt)
                ca_cells, ma_muts = cA, mA
                next_seeds = nxt.get_seeds()
                seed_list += next_seeds
                curr_seed = next_seeds[0]
                idx -= 1
        return best, seed_list

    def sprout_identity(self, seed):
        return IdentityTree(seed.cells, seed.muts, None)
