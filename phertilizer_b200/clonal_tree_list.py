"""A queue of candidate clonal trees (mirrors the reference's `ClonalTreeList`, phertilizer/clonal_tree_list.py)."""
from collections import deque

import numpy as np
from pandas import Series


class ClonalTreeList:
    def __init__(self):
        self.counter = 0
        self.stack = deque()

    def insert(self, tree):
        tree.set_key(self.counter)
        self.counter += 1
        self.stack.append(tree)

    def index_tree(self, index):
        return self.stack[index] if index < self.size() else None

    def pop_tree(self):
        return self.stack.popleft()

    def get_tree(self, key):
        for t in self.stack:
            if t.key == key:
                return t

    def has_trees(self):
        return len(self.stack) > 0

    def get_total_count(self):
        return self.counter

    def size(self):
        return len(self.stack)

    def get_all_trees(self):
        return self.stack

    @staticmethod
    def _signature(tree):
        """Invariant of the reference's tree equality (isomorphism with matching `ncells`, clonal_tree.py:187-192): the
        sorted (ncells, out-degree, in-degree) triples of the nodes.  Different signatures => not equal, so the
        quadratic number of `nx.is_isomorphic` calls of `grow` (phertilizer.py:558-614) shrinks to the real candidates."""
        g = tree.tree
        return tuple(sorted((d.get("ncells"), g.out_degree(v), g.in_degree(v)) for v, d in g.nodes(data=True)))

    def _cached_signature(self, tree):
        # trees are not edited once they sit in a list; a relabel/merge replaces `tree.tree`, which invalidates the entry
        ent = getattr(tree, "_sig_cache", None)
        if ent is None or ent[0] is not tree.tree or ent[1] != tree.tree.number_of_nodes():
            ent = (tree.tree, tree.tree.number_of_nodes(), self._signature(tree))
            tree._sig_cache = ent
        return ent[2]

    def contains(self, obj):
        sig = self._signature(obj)
        return any(self._cached_signature(t) == sig and t == obj for t in self.stack)

    def find_best_tree(self, data):
        """Score every candidate on the GPU and keep the first maximum of the normalised log-likelihood
        (clonal_tree_list.py:196-234; strict > keeps the earliest tree on ties)."""
        n = self.size()
        print(f"Finding the maximum likelihood tree out of {n} trees....")
        values = np.zeros(shape=n)
        best_like, best_tree, keys = -np.inf, None, []
        for i, tree in enumerate(self.stack):
            keys.append(tree.key)
            tree.compute_likelihood(data)
            values[i] = tree.norm_loglikelihood
            print(f"Tree: {tree.key} Log Likelihood: {values[i]}")
            if values[i] > best_like:
                best_like, best_tree = values[i], tree
        return best_tree, Series(values, index=keys)

    def save_the_trees(self, path):
        for tree in self.stack:
            loglike = np.round(tree.loglikelihood)
            tag = ("+" if loglike >= 0 else "-") + str(loglike)
            tree.tree_png(f"{path}/tree{tree.key}_{tag}.png")
