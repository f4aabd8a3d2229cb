"""Clonal trees with cell attachments and SNV placements, scored on the GPU.

Same public surface as the reference's `phertilizer/clonal_tree.py` (attributes `tree`, `cell_mapping`, `mut_mapping`,
`mut_loss_mapping`, `event_mapping`, `loglikelihood`, `variant_likelihood`, `bin_count_likelihood`,
`norm_loglikelihood`, `node_likelihood`; classes `ClonalTree`, `LinearTree`, `BranchingTree`, `IdentityTree`, `Seed`).
The likelihood and reassignment arithmetic that the reference evaluates on dense matrices
(clonal_tree.py:395-407, 573-729, 858-944, 1135-1146) is delegated to the `ObservationStore` kernels.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import networkx as nx
import numpy as np

from .utils import (setdiff1d, dict_to_dataframe, generate_cell_dataframe, generate_mut_dataframe, get_next_label, pickle_save)

_NOT_IN_TREE = 255


def _empty():
    return np.empty(shape=0, dtype=int)


class ClonalTree:
    def __init__(self, tree, cell_mapping, mut_mapping, mut_loss_mapping=None, event_mapping=None, key=0):
        self.tree: nx.DiGraph = tree
        self.cell_mapping = cell_mapping
        self.mut_mapping = mut_mapping
        self.mut_loss_mapping = {} if mut_loss_mapping is None else mut_loss_mapping
        # CNA events are never inferred on the live path of the reference (phertilizer.py:289); kept for compatibility
        self.event_mapping = {} if not event_mapping else event_mapping
        self.cna_genotype_mode = bool(event_mapping)
        self.key = key
        self.loglikelihood = self.variant_likelihood = self.bin_count_likelihood = None
        self.states = ["gain", "loss", "neutral"]
        self.use_rd = True
        self.norm_loglikelihood = -10000

    # ------------------------------------------------------------------ identity / printing
    def __eq__(self, other):
        """Trees are "equal" when isomorphic with matching cell counts per node (clonal_tree.py:187-192)."""
        return nx.is_isomorphic(self.tree, other.tree, node_match=lambda a, b: a["ncells"] == b["ncells"])

    __hash__ = None

    def __str__(self):
        lines = [f"Clonal Tree {self.key} ", f"{list(self.tree.edges())} ", f"Loss Nodes:{len(self.mut_loss_mapping)}"]
        for n in self.tree.nodes():
            lost = len(self.mut_loss_mapping[n]) if n in self.mut_loss_mapping else 0
            lines.append(f"Node: {n} Cells: {len(self.get_tip_cells(n))} Muts: {len(self.get_tip_muts(n))} LostMuts: {lost}")
        return "\n".join(lines) + f"\nTotal Cells: {len(self.get_all_cells())} Total Muts: {len(self.get_all_muts())}"

    def set_key(self, key):
        self.key = key

    def has_loss(self):
        return len(self.mut_loss_mapping) > 0

    # ------------------------------------------------------------------ accessors
    def get_tip_cells(self, t):
        if t in self.cell_mapping:
            return np.concatenate([self.cell_mapping[t][k] for k in self.cell_mapping[t]])
        return _empty()

    def get_tip_muts(self, t):
        return self.mut_mapping[t] if t in self.mut_mapping else _empty()

    def get_leaves(self):
        return [n for n in list(self.tree.nodes()) if self.tree.out_degree(n) == 0]

    def get_all_cells(self):
        return np.concatenate([self.cell_mapping[i][k] for i in self.cell_mapping for k in self.cell_mapping[i]])

    def get_all_muts(self):
        return np.sort(np.concatenate([self.mut_mapping[i] for i in self.mut_mapping]))

    def get_cell_clusters(self):
        clusters = np.zeros(len(self.get_all_cells()), dtype=int)
        for cluster, cells in self.cell_mapping.items():
            if len(cells[0]) > 0:
                clusters[cells[0]] = cluster
        return clusters

    def get_mut_clusters(self, n_mut=None):
        if n_mut is None:
            clusters = np.zeros(len(self.get_all_muts()), dtype=int)
        else:
            clusters = np.full(shape=n_mut, fill_value=-1)
        for cluster, muts in self.mut_mapping.items():
            if len(muts) > 0:
                clusters[muts] = cluster
        return clusters

    def find_root(self):
        for n in list(self.tree.nodes()):
            if self.tree.in_degree(n) == 0:
                return n

    def get_seed_by_node(self, node):
        if node in list(self.tree.nodes()):
            return Seed(self.get_tip_cells(node), self.get_tip_muts(node))
        return None

    def _path_to(self, node):
        return nx.shortest_path(self.tree, self.find_root(), node)

    def get_ancestral_muts(self, node):
        """SNVs gained on the root->parent(node) path minus losses (clonal_tree.py:979-992)."""
        path = self._path_to(node)[:-1]
        if not path:
            return _empty()
        present = np.concatenate([self.mut_mapping[p] for p in path])
        lost = [self.mut_loss_mapping[p] for p in path if p in self.mut_loss_mapping]
        if lost:
            present = np.setdiff1d(present, np.concatenate(lost))
        return present

    def get_cell_cluster(self, c):
        for node in self.cell_mapping:
            if c in self.cell_mapping[node][0]:
                return node

    def get_snv_cluster(self, s):
        for node in self.mut_mapping:
            if s in self.mut_mapping[node]:
                return node

    # ------------------------------------------------------------------ topology edits
    def relabel(self):
        """Relabel nodes 0..K-1 in pre-order and carry the mappings along (clonal_tree.py:238-266)."""
        order = list(nx.dfs_preorder_nodes(self.tree, self.find_root()))
        new_of = {old: i for i, old in enumerate(order)}
        self.tree = nx.relabel_nodes(self.tree, new_of, copy=True)
        for name in ("cell_mapping", "mut_mapping", "mut_loss_mapping", "event_mapping"):
            old = getattr(self, name)
            setattr(self, name, {i: old[o] for i, o in enumerate(order) if o in old})
        self.labels = order

    def merge(self, tree, l):
        """Replace leaf `l` by `tree` (root of `tree` takes the leaf's place) and relabel (clonal_tree.py:269-309)."""
        if type(tree) is not IdentityTree:
            parent = list(self.tree.predecessors(l))[0]
            self.tree.remove_node(l)
            self.cell_mapping.pop(l)
            self.mut_mapping.pop(l)
            self.mut_loss_mapping.pop(l, None)
            if self.cna_genotype_mode:
                self.event_mapping.pop(l, None)
            top = get_next_label(self.tree)
            self.tree.add_node(top, ncells=len(tree.get_tip_cells(0)))
            self.tree.add_edge(parent, top)
            self.mut_mapping[top] = tree.mut_mapping[0]
            self.cell_mapping[top] = tree.cell_mapping[0]
            if 0 in tree.mut_loss_mapping:
                self.mut_loss_mapping[top] = tree.mut_loss_mapping[0]
            if 0 in tree.event_mapping:
                self.event_mapping[top] = tree.event_mapping[0]
            for child in tree.tree.successors(0):
                lab = top + child
                self.tree.add_node(lab, ncells=len(tree.get_tip_cells(child)))
                self.tree.add_edge(top, lab)
                self.mut_mapping[lab] = tree.mut_mapping[child]
                self.cell_mapping[lab] = tree.cell_mapping[child]
                if child in tree.mut_loss_mapping:
                    self.mut_loss_mapping[lab] = tree.mut_loss_mapping[child]
                if child in tree.event_mapping:
                    self.event_mapping[lab] = tree.event_mapping[child]
        self.relabel()

    # ------------------------------------------------------------------ genotypes (host, small)
    def _present_muts(self, node):
        present = np.concatenate([self.get_ancestral_muts(node).astype(int), self.mut_mapping[node]])
        if node in self.mut_loss_mapping:
            present = np.setdiff1d(present, self.mut_loss_mapping[node])
        return present.astype(int)

    def snv_genotypes(self, m=None):
        """node -> 0/1 vector over SNVs (clonal_tree.py:311-327)."""
        if m is None:
            m = len(self.get_all_muts())
        out = {}
        for node in self.mut_mapping:
            y = np.zeros(m, dtype=int)
            y[self._present_muts(node)] = 1
            out[node] = y
        return out

    def cell_cluster_genotypes(self, n=None):
        """node -> 0/1 vector over cells of the node's clade (clonal_tree.py:329-341)."""
        if n is None:
            n = len(self.get_all_cells())
        out = {}
        for node in self.cell_mapping:
            c = np.zeros(n, dtype=int)
            clade = list(nx.dfs_preorder_nodes(self.tree, node))
            c[np.concatenate([self.cell_mapping[i][0] for i in clade])] = 1
            out[node] = c
        return out

    def presence_by_node(self, m, cells, node):
        presence = np.zeros(shape=(len(cells), m), dtype=int)
        if len(cells):
            presence[:, self._present_muts(node)] = 1
        return presence

    # ------------------------------------------------------------------ device label vectors
    def _node_arrays(self, data):
        """(K, cell_node[n], snv_node[m], present[K,K]) for the scoring kernels; nodes must be labelled 0..K-1 < 64."""
        nodes = list(self.tree.nodes())
        K = int(max(nodes)) + 1
        if K > 64:
            raise ValueError(f"tree with node label {K - 1} exceeds the 64 nodes the scoring kernel supports")
        cell_node = np.full(data.n_cells + 16, _NOT_IN_TREE, np.uint8)[: data.n_cells]
        snv_node = np.full(data.n_snvs + 16, _NOT_IN_TREE, np.uint8)[: data.n_snvs]
        for k in self.cell_mapping:
            cells = self.get_tip_cells(k)
            if len(cells):
                cell_node[cells.astype(np.int64)] = k
        present = np.zeros((K, K), np.uint8)
        placed = 0
        for k in self.mut_mapping:
            muts = np.asarray(self.mut_mapping[k]).astype(np.int64)
            placed += len(muts)
            if len(muts):
                snv_node[muts] = k
        # one node per SNV and no losses <=> presence is exactly "SNV's node is on the root->k path"
        simple = (not self.mut_loss_mapping) and placed == int(np.count_nonzero(snv_node != _NOT_IN_TREE))
        for k in nodes:
            for q in self._path_to(k):
                present[k, q] = 1
        return K, cell_node, snv_node, present, simple

    # ------------------------------------------------------------------ likelihood
    def compute_likelihood(self, data):
        """Tree log-likelihood = sum over nodes of variant + read-depth terms (clonal_tree.py:858-910); the per-node
        variant sums (932-944) and the Gaussian embedding term (395-407) each come from one kernel call."""
        self.use_rd = data.use_read_depth
        K, cell_node, snv_node, present, simple = self._node_arrays(data)
        if simple:
            variant = data.store.tree_loglik(cell_node, snv_node, present)
        else:  # an SNV placed on several nodes / losses: score node by node with explicit presence sets
            variant = np.zeros(K)
            for k in self.cell_mapping:
                cells = self.get_tip_cells(k)
                if len(cells) == 0:
                    continue
                cn = np.full(data.n_cells + 16, _NOT_IN_TREE, np.uint8)[: data.n_cells]
                sn = np.full(data.n_snvs + 16, _NOT_IN_TREE, np.uint8)[: data.n_snvs]
                cn[cells.astype(np.int64)] = 0
                sn[self._present_muts(k)] = 0
                variant[k] = data.store.tree_loglik(cn, sn, np.ones((1, 1), np.uint8))[0]
        bins = data.store.gauss_node_ll(data.read_depth, cell_node, K) if self.use_rd else np.zeros(K)
        totals = {"total": 0, "variant": 0, "bin": 0}
        self.node_likelihood = {}
        for n in self.cell_mapping:
            if len(self.get_tip_cells(n)) > 0:
                v, b = float(variant[n]), float(bins[n]) if self.use_rd else 0
                self.node_likelihood[n] = {"variant": v, "total": v + b, "bin": b}
        for n in self.node_likelihood:
            for k in totals:
                totals[k] += self.node_likelihood[n][k]
        self.loglikelihood_dict = totals
        self.loglikelihood = totals["total"]
        self.variant_likelihood = totals["variant"]
        self.bin_count_likelihood = totals["bin"]
        self.norm_loglikelihood = self.loglikelihood / (len(self.get_all_cells()) * len(self.get_all_muts()))
        return self.loglikelihood

    def get_loglikelihood(self):
        return self.loglikelihood, self.variant_likelihood, self.bin_count_likelihood

    # ------------------------------------------------------------------ post-processing
    def find_bad_snvs(self, data, node, q):
        """SNVs of `node` with binary VAF over the clade's cells <= the q-quantile, plus unobserved ones
        (clonal_tree.py:573-600)."""
        clade = list(nx.dfs_preorder_nodes(self.tree, node))
        cells = np.concatenate([self.cell_mapping[c][0] for c in clade]).astype(int)
        muts = np.asarray(self.mut_mapping[node]).astype(int)
        if len(muts) == 0:
            return _empty()
        t = data.store.snv_table(data.store.cell_labels(cells), 1, muts, want=("Nobs", "Nvar"))
        tot, var = t["Nobs"][:, 0], t["Nvar"][:, 0]
        na = muts[tot == 0]
        seen = tot > 0
        order = np.argsort(muts[seen], kind="stable")  # the reference continues on setdiff1d(muts, na): sorted
        kept = muts[seen][order]
        vaf = (var[seen] / tot[seen])[order]
        bad = kept[vaf <= np.quantile(vaf, q)] if len(vaf) > 0 else _empty()
        return np.union1d(bad, na)

    def _move_snv(self, s, scores_by_node):
        """Strict-> scan in mapping order, then move s (clonal_tree.py:647-666)."""
        clust = self.get_snv_cluster(s)
        if clust is None:
            clust = -1
        best_like, best = -np.inf, None
        for node, val in scores_by_node:
            if val > best_like:
                best_like, best = val, node
        if best != clust:
            if clust >= 0:
                self.mut_mapping[clust] = np.setdiff1d(self.mut_mapping[clust], np.array(s))
            self.mut_mapping[best] = np.union1d(self.mut_mapping[best], np.array(s))

    def _snv_scores(self, data, snvs, c_dict):
        """scores[i][k] = dot(c_k, like1[:, s_i]) + dot(1-c_k, like0[:, s_i]) for all nodes k (one kernel call)."""
        nodes = list(c_dict.keys())
        pats, inv = np.unique(np.array([c_dict[k] for k in nodes]).T, axis=0, return_inverse=True)
        P, K = len(pats), len(nodes)
        KK = max(P, K)
        if KK > 64:
            raise ValueError("more than 64 distinct clades")
        member = np.zeros((KK, KK), np.uint8)
        member[:K, :P] = pats.T
        cell_node = np.zeros(data.n_cells + 16, np.uint8)[: data.n_cells]
        cell_node[:] = inv.reshape(-1)
        T = data.store.snv_node_table(cell_node, member, np.asarray(snvs, np.int32))
        return nodes, T[:, :K]

    def reassign_snv(self, s, c_dict, data):
        nodes, T = self._snv_scores(data, [s], c_dict)
        self._move_snv(s, zip(nodes, T[0]))

    def reassign_cell(self, c, y_dict, data, _scores=None):
        """clonal_tree.py:612-645: best node among those with <= 1 child, strict >, starting from the current node."""
        clust = self.get_cell_cluster(c)
        nodes = list(y_dict.keys())
        if _scores is None:
            _scores = self._cell_scores(data, [c], y_dict)[1][0]
        score = dict(zip(nodes, _scores))
        if clust in y_dict:
            best_like = score[clust]
            best = clust
        else:
            clust, best, best_like = -1, None, -np.inf
        for node in nodes:
            if len(list(self.tree.successors(node))) <= 1 and score[node] > best_like:
                best_like, best = score[node], node
        if best != clust:
            if clust >= 0:
                self.cell_mapping[clust][0] = np.setdiff1d(self.cell_mapping[clust][0], np.array(c))
            self.cell_mapping[best][0] = np.union1d(self.cell_mapping[best][0], np.array(c))

    def _cell_scores(self, data, cells, y_dict):
        nodes = list(y_dict.keys())
        pats, inv = np.unique(np.array([y_dict[k] for k in nodes]).T, axis=0, return_inverse=True)
        P, K = len(pats), len(nodes)
        KK = max(P, K)
        if KK > 64:
            raise ValueError("more than 64 distinct SNV placements")
        member = np.zeros((KK, KK), np.uint8)
        member[:K, :P] = pats.T
        snv_node = np.zeros(data.n_snvs + 16, np.uint8)[: data.n_snvs]
        snv_node[:] = inv.reshape(-1)
        T = data.store.cell_node_table(snv_node, member, np.asarray(cells, np.int32))
        return nodes, T[:, :K]

    def find_missing_cells(self, data):
        return np.setdiff1d(data.cell_lookup.index.to_numpy(), self.get_all_cells())

    def find_missing_snvs(self, data):
        return np.setdiff1d(data.mut_lookup.index.to_numpy(), self.get_all_muts())

    def post_process(self, data, q=0.1, iterations=1, place_missing=True):
        """Place missing cells/SNVs, then move each non-root node's low-VAF SNVs to their best node
        (clonal_tree.py:680-729).  The sequential semantics are kept; the per-SNV x node score tables are batched
        (they do not depend on the SNV placements being edited)."""
        ncells, nmuts = data.cell_lookup.shape[0], data.mut_lookup.shape[0]
        nodes = list(nx.dfs_postorder_nodes(self.tree, source=self.find_root()))
        prev = self.compute_likelihood(data)
        if place_missing:
            missing_snvs = self.find_missing_snvs(data)
            missing_cells = self.find_missing_cells(data)
            if len(missing_cells):
                y_dict = self.snv_genotypes(m=nmuts)
                _, scores = self._cell_scores(data, missing_cells, y_dict)
                for c, row in zip(missing_cells, scores):
                    self.reassign_cell(c, y_dict, data, _scores=row)
            if len(missing_snvs):
                c_dict = self.cell_cluster_genotypes(n=ncells)
                nds, T = self._snv_scores(data, missing_snvs, c_dict)
                for s, row in zip(missing_snvs, T):
                    self._move_snv(s, zip(nds, row))
        loglikelihood = self.compute_likelihood(data)
        for i in range(iterations):
            c_dict = self.cell_cluster_genotypes(n=ncells)
            for n in nodes:
                if n != self.find_root():
                    bad = self.find_bad_snvs(data, n, q)
                    if len(bad):
                        nds, T = self._snv_scores(data, bad, c_dict)
                        for s, row in zip(bad, T):
                            self._move_snv(s, zip(nds, row))
                loglikelihood = self.compute_likelihood(data)
                print(f"iteration {i} node: {n} variant {self.variant_likelihood} loglike: {loglikelihood}")
                if prev == loglikelihood:
                    break
                prev = loglikelihood
        return loglikelihood

    # ------------------------------------------------------------------ outputs
    def generate_results(self, cell_lookup, mut_lookup):
        return (generate_cell_dataframe(self.cell_mapping, cell_lookup), generate_mut_dataframe(self.mut_mapping, mut_lookup),
                generate_mut_dataframe(self.mut_loss_mapping, mut_lookup), dict_to_dataframe(self.event_mapping))

    def save_results(self, cell_lookup, mut_lookup, pcell_fname, pmut_fname, ploss_fname, pevents_fname):
        pcell, pmut, ploss, pevents = self.generate_results(cell_lookup, mut_lookup)
        pcell.to_csv(pcell_fname, index=False)
        pmut.to_csv(pmut_fname, index=False)
        ploss.to_csv(ploss_fname, index=False)
        pevents.to_csv(pevents_fname)

    def save(self, path):
        pickle_save(self, path)

    def save_text(self, path):
        """Edge list + leaves in the reference's text format (clonal_tree.py:423-435)."""
        leafs = [n for n in self.tree.nodes if len(list(self.tree.successors(n))) == 0]
        with open(path, "w+") as f:
            f.write(f"{len(list(self.tree.edges))} #edges\n")
            for u, v in list(self.tree.edges):
                f.write(f"{u} {v}\n")
            f.write(f"{len(leafs)} #leaves\n")
            for l in leafs:
                f.write(f"{l}\n")

    def _to_agraph(self):
        try:
            import pygraphviz as pgv
        except ImportError as e:  # output cosmetics only; out of scope of the GPU path
            raise ImportError("pygraphviz is required to draw trees") from e
        G = pgv.AGraph(strict=False, directed=True)
        like = f"\\nLL: {self.loglikelihood:.1f}" if self.loglikelihood is not None else ""
        for n in self.tree.nodes():
            G.add_node(n, label=f"{n}\\nCells: {len(self.get_tip_cells(n))}{like if n == self.find_root() else ''}")
        for u, v in self.tree.edges():
            G.add_edge(u, v, label=f"+{len(self.get_tip_muts(v))} SNVs")
        return G

    def tree_png(self, fname, chrom_map_file: str = None):
        G = self._to_agraph()
        G.layout("dot")
        G.draw(fname)

    def tree_dot(self, fname, chrom_map_file: str = None):
        self._to_agraph().write(fname)


class LinearTree(ClonalTree):
    """parent (node 0: cellsA, mutsA) -> child (node 1: cellsB, mutsB)  (clonal_tree.py:1011-1039)."""

    def __init__(self, cellsA, cellsB, mutsA, mutsB, eA=None, eB=None, key=None):
        t = nx.DiGraph()
        t.add_node(0, ncells=len(cellsA))
        t.add_node(1, ncells=len(cellsB))
        t.add_edge(0, 1)
        em = {k: e for k, e in ((0, eA), (1, eB)) if e is not None}
        super().__init__(t, {0: {0: cellsA}, 1: {0: cellsB}}, {0: mutsA, 1: mutsB}, {}, em)

    def get_seeds(self):
        return [Seed(self.get_tip_cells(1), self.get_tip_muts(1))]


class BranchingTree(ClonalTree):
    """empty root (node 0: mutsC) with children 1 (cellsA, mutsA) and 2 (cellsB, mutsB)  (clonal_tree.py:1042-1078)."""

    def __init__(self, cellsA, cellsB, mutsA, mutsB, mutsC, eA=None, eB=None, eC=None, key=None):
        t = nx.DiGraph()
        t.add_node(0, ncells=0)
        t.add_node(1, ncells=len(cellsA))
        t.add_node(2, ncells=len(cellsB))
        t.add_edge(0, 1)
        t.add_edge(0, 2)
        em = {1: eA, 2: eB} if eA is not None and eB is not None else {}
        super().__init__(t, {0: {0: _empty()}, 1: {0: cellsA}, 2: {0: cellsB}}, {0: mutsC, 1: mutsA, 2: mutsB}, {}, em)

    def get_seeds(self):
        return [Seed(self.get_tip_cells(l), self.mut_mapping[l]) for l in (1, 2)]


class IdentityTree(ClonalTree):
    def __init__(self, cells, muts, events=None):
        t = nx.DiGraph()
        t.add_node(0, ncells=len(cells))
        super().__init__(t, {0: {0: cells}}, {0: muts}, {}, {0: events} if events is not None else {})

    def get_seeds(self):
        return []


@dataclass
class Seed:
    """A leaf (cells, SNVs) still to be split (clonal_tree.py:1100-1158)."""

    cells: np.ndarray
    muts: np.ndarray
    ancestral_muts: np.ndarray = None
    key: int = None
    linear_tree: LinearTree = None
    branching_tree: BranchingTree = None
    tree_list: list = field(default_factory=list)

    def __str__(self):
        return f"Cells: {len(self.cells)} Muts: {len(self.muts)}"

    def __eq__(self, other):
        return type(other) is type(self) and np.array_equal(self.cells, other.cells) and np.array_equal(self.muts, other.muts)

    __hash__ = None

    def set_key(self, key):
        self.key = key

    def get_key(self):
        return self.key

    def strip(self, store):
        """Drop SNVs without a variant read among the seed's cells, then cells without a variant read among the
        remaining SNVs (two stages, clonal_tree.py:1135-1142; sum(var)==0 <=> no entry with var>0)."""
        muts = np.asarray(self.muts)
        cells = np.asarray(self.cells)
        if len(muts) and len(cells):
            nv = store.snv_table(store.cell_labels(cells), 1, muts, want=("Nvar",))["Nvar"][:, 0]
            self.muts = setdiff1d(muts, muts[nv == 0])
        else:
            self.muts = setdiff1d(muts, muts)
        if len(self.muts) and len(cells):
            nc = store.cell_table(store.snv_labels(self.muts), 1, cells, want=("Nvar",))["Nvar"][:, 0]
            self.cells = setdiff1d(cells, cells[nc == 0])
        else:
            self.cells = setdiff1d(cells, cells)

    def count_obs(self, store):
        from .utils import check_obs

        return check_obs(self.cells, self.muts, store)

    def set_linear(self, ct):
        self.tree_list.append(ct)
        self.linear_tree = ct

    def has_linear(self):
        return self.linear_tree is not None

    def set_branching(self, ct):
        self.tree_list.append(ct)
        self.branching_tree = ct

    def has_branching(self):
        return self.branching_tree is not None
