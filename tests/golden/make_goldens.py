"""Generate the golden vectors under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE (build container only).

    python tests/golden/make_goldens.py            # needs /root/reference; writes *.npz + manifest json

What is recorded (SURVEY.md section 4 "per-call golden capture"):
  * inputs  : the example's observations as COO in the reference's internal indexing (after run.py's dropped first
              row and the np.sort(unique) label mapping), the binned read counts, and the reference's per-observation
              like0/like1 values for every distinct (var,total) pair;
  * kat     : like0/like1 from the reference's numba functions for all (k,n), n<=60, three (alpha, max_copies) settings;
  * calls   : (arguments, result) of every hot-path method while the reference runs
              `-s 3 -j 10 --post_process --no-umap` on the example (config 1), `-s 2 -j 6 --post_process` on a small
              synthetic data set (T0) and `-s 2 -j 5 --post_process` on a 2000 x 3000 synthetic (T1);
  * e2e     : candidate-tree likelihoods, final tree, cell/SNV clusters.
The GPU box has no /root/reference: tests only read the files written here.
"""
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import pandas as pd
import scipy.special

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_shim import example_paths, load_reference  # noqa: E402

R = load_reference()
PH, LS, BS, CT, UT = R["phertilizer"], R["linear_split"], R["branching_split"], R["clonal_tree"], R["utils"]


class Recorder:
    def __init__(self):
        self.arrays = {}
        self.manifest = []

    def add(self, kind, **fields):
        idx = len(self.manifest)
        entry = {"kind": kind, "fields": {}}
        for k, v in fields.items():
            if v is None:
                entry["fields"][k] = None
            elif isinstance(v, (bool, np.bool_)):
                entry["fields"][k] = {"py": bool(v)}
            elif isinstance(v, (int, np.integer)):
                entry["fields"][k] = {"py": int(v)}
            elif isinstance(v, (float, np.floating)):
                entry["fields"][k] = {"f64": repr(float(v))}
            else:
                a = np.asarray(v)
                if a.dtype.kind in "iu":
                    a = a.astype(np.int32)
                elif a.dtype.kind == "b":
                    a = a.astype(np.uint8)
                key = f"r{idx}_{k}"
                self.arrays[key] = a
                entry["fields"][k] = {"npz": key}
        self.manifest.append(entry)

    def save(self, stem):
        np.savez_compressed(os.path.join(HERE, stem + ".npz"), **self.arrays)
        with open(os.path.join(HERE, stem + ".json"), "w") as f:
            json.dump(self.manifest, f)


REC = None


def tree_fields(t):
    nodes = sorted(t.cell_mapping.keys())
    f = {"edges": np.array(list(t.tree.edges()), dtype=np.int32).reshape(-1, 2), "nodes": np.array(nodes)}
    for k in nodes:
        f[f"cells_{k}"] = t.get_tip_cells(k)
        f[f"muts_{k}"] = np.asarray(t.get_tip_muts(k)).astype(int)
    return f


def install_hooks():
    def wrap(cls, name, fn):
        orig = getattr(cls, name)

        def inner(self, *a, **k):
            return fn(orig, self, *a, **k)

        setattr(cls, name, inner)

    def ls_mut(orig, self, cellsA, muts):
        out = orig(self, cellsA, muts)
        REC.add("linear_mut_assignment", cellsA=cellsA, muts=muts, nobs_per_cluster=float(self.params.nobs_per_cluster),
                mutsA=out[0], mutsB=out[1])
        return out

    def bs_mut(orig, self, cellsA, cellsB):
        out = orig(self, cellsA, cellsB)
        REC.add("branching_mut_assignment", cellsA=cellsA, cellsB=cellsB, muts=self.muts,
                nobs_per_cluster=float(self.params.nobs_per_cluster), mutsA=out[0], mutsB=out[1], mutsC=out[2])
        return out

    def ls_init(orig, self, cells, muts, p):
        bc = np.count_nonzero(self.var[np.ix_(cells, muts)], axis=0)
        bt = np.count_nonzero(self.total[np.ix_(cells, muts)], axis=0)
        vaf = bc / bt
        REC.add("weighted_init_vaf", cells=cells, muts=muts, vaf=vaf, vaf_prob=vaf / np.sum(vaf))
        return orig(self, cells, muts, p)

    def bs_tmb(orig, self, cells, muts):
        out = orig(self, cells, muts)
        REC.add("calc_tmb", cells=cells, muts=muts, tmb=out)
        return out

    def ls_aff(orig, self, cells, mutsB=None):
        out = orig(self, cells, mutsB)
        if mutsB is not None:
            v = np.count_nonzero(self.var[np.ix_(cells, mutsB)], axis=1)
            t = np.count_nonzero(self.total[np.ix_(cells, mutsB)], axis=1)
            REC.add("linear_affinity_features", cells=cells, mutsB=mutsB, nvar=v, nobs=t)
        return out

    def ls_metrics(orig, self, ca, cb, ma, mb):
        out = orig(self, ca, cb, ma, mb)
        REC.add("linear_check_metrics", all_cells=self.cells, ca=ca, cb=cb, ma=ma, mb=mb,
                nobs_per_cluster=float(self.params.nobs_per_cluster), f1=out[0], f2=out[1], f3=bool(out[2]))
        return out

    def bs_metrics(orig, self, ca, cb, ma, mb, mc):
        out = orig(self, ca, cb, ma, mb, mc)
        REC.add("branching_check_metrics", ca=ca, cb=cb, ma=ma, mb=mb, mc=mc,
                nobs_per_cluster=float(self.params.nobs_per_cluster), f1=out[0], f2=out[1], f3=bool(out[2]),
                feats=np.array(out[3]))
        return out

    def ls_norm(orig, self, cellsA, cellsB, mutsA, mutsB):
        out = orig(self, cellsA, cellsB, mutsA, mutsB)
        REC.add("linear_norm_likelihood", all_cells=self.cells, cellsA=cellsA, mutsA=mutsA, mutsB=mutsB, value=out)
        return out

    def bs_norm(orig, self, cellsA, cellsB, mutsA, mutsB, mutsC):
        out = orig(self, cellsA, cellsB, mutsA, mutsB, mutsC)
        REC.add("branching_norm_likelihood", all_cells=self.cells, cellsA=cellsA, cellsB=cellsB, mutsA=mutsA, mutsB=mutsB,
                mutsC=mutsC, value=out)
        return out

    def seed_strip(orig, self, var):
        c0, m0 = self.cells, self.muts
        orig(self, var)
        REC.add("seed_strip", cells=c0, muts=m0, cells_out=self.cells, muts_out=self.muts)

    def seed_count(orig, self, total):
        out = orig(self, total)
        REC.add("count_obs", cells=self.cells, muts=self.muts, value=int(out))
        return out

    def ct_var(orig, self, node, like0, like1, bin_mapping=None):
        out = orig(self, node, like0, like1, bin_mapping)
        cells = self.get_tip_cells(node)
        present = np.concatenate([self.get_ancestral_muts(node).astype(int), self.mut_mapping[node]]).astype(int)
        REC.add("variant_likelihood_by_node", cells=cells, present_muts=present, value=out)
        return out

    def ct_rd(orig, self, n, read_depth, cells=None):
        out = orig(self, n, read_depth, cells)
        c = self.get_tip_cells(n) if cells is None else cells
        if len(c):
            REC.add("read_depth_likelihood_by_node", cells=c, value=out[0], per_cell=out[1].to_numpy())
        return out

    def ct_like(orig, self, data):
        out = orig(self, data)
        REC.add("tree_likelihood", loglikelihood=self.loglikelihood, variant=self.variant_likelihood,
                bin=self.bin_count_likelihood, norm=self.norm_loglikelihood, **tree_fields(self))
        return out

    def ct_reassign_snv(orig, self, s, c_dict, data):
        nodes = list(c_dict.keys())
        scores = [np.dot(c_dict[k], data.like1_marg[:, s]) + np.dot(1 - c_dict[k], data.like0[:, s]) for k in nodes]
        before = self.get_snv_cluster(s)
        orig(self, s, c_dict, data)
        REC.add("reassign_snv", s=int(s), nodes=np.array(nodes), cdict=np.array([c_dict[k] for k in nodes]).astype(np.uint8),
                scores=np.array(scores), before=-1 if before is None else int(before), after=int(self.get_snv_cluster(s)))

    def ct_reassign_cell(orig, self, c, y_dict, data):
        nodes = list(y_dict.keys())
        scores = [np.dot(y_dict[k], data.like1_marg[c, :]) + np.dot(1 - y_dict[k], data.like0[c, :]) for k in nodes]
        before = self.get_cell_cluster(c)
        orig(self, c, y_dict, data)
        REC.add("reassign_cell", c=int(c), nodes=np.array(nodes), ydict=np.array([y_dict[k] for k in nodes]).astype(np.uint8),
                scores=np.array(scores), before=-1 if before is None else int(before), after=int(self.get_cell_cluster(c)),
                succ=np.array([len(list(self.tree.successors(k))) for k in nodes]))

    def ct_bad(orig, self, data, node, q):
        import networkx as nx
        clade = list(nx.dfs_preorder_nodes(self.tree, node))
        cells = np.concatenate([self.cell_mapping[c][0] for c in clade])
        muts = self.mut_mapping[node]
        out = orig(self, data, node, q)
        REC.add("find_bad_snvs", clade_cells=cells, muts=muts, q=float(q), bad=out)
        return out

    def ct_post(orig, self, data, *a, **k):
        REC.add("post_process_begin", **tree_fields(self))
        out = orig(self, data, *a, **k)
        REC.add("post_process_end", loglikelihood=out, variant=self.variant_likelihood, norm=self.norm_loglikelihood,
                **tree_fields(self))
        return out

    wrap(LS.Linear_split, "mut_assignment", ls_mut)
    wrap(BS.Branching_split, "mut_assignment", bs_mut)
    wrap(LS.Linear_split, "random_weighted_init_array", ls_init)
    wrap(BS.Branching_split, "calc_tmb", bs_tmb)
    wrap(LS.Linear_split, "create_affinity", ls_aff)
    wrap(LS.Linear_split, "check_metrics", ls_metrics)
    wrap(BS.Branching_split, "check_metrics", bs_metrics)
    wrap(LS.Linear_split, "compute_norm_likelihood", ls_norm)
    wrap(BS.Branching_split, "compute_norm_likelihood", bs_norm)
    wrap(CT.Seed, "strip", seed_strip)
    wrap(CT.Seed, "count_obs", seed_count)
    wrap(CT.ClonalTree, "compute_variant_likelihood_by_node_without_events", ct_var)
    wrap(CT.ClonalTree, "read_depth_likelihood_by_node", ct_rd)
    wrap(CT.ClonalTree, "compute_likelihood", ct_like)
    wrap(CT.ClonalTree, "reassign_snv", ct_reassign_snv)
    wrap(CT.ClonalTree, "reassign_cell", ct_reassign_cell)
    wrap(CT.ClonalTree, "find_bad_snvs", ct_bad)
    wrap(CT.ClonalTree, "post_process", ct_post)
    # ClonalTree.__init__ binds compute_variant_likelihood to the bound method at construction: it resolves through
    # the class attribute, so wrapping the class before any tree is built is enough.


def run_reference(variant_df, bin_df, stem, starts, iterations, seed, inputs_extra=None, **phert_kw):
    """Runs the reference end to end, records calls, saves <stem>_inputs/_calls."""
    global REC
    ph = PH.Phertilizer(variant_df.copy(), bin_df.copy(), alpha=0.001, max_copies=5, mode="rd", dim_reduce=False)
    d = ph.data
    cell, snv = np.nonzero(d.total)
    inputs = dict(n=np.int64(d.total.shape[0]), m=np.int64(d.total.shape[1]), cell=cell.astype(np.int32), snv=snv.astype(np.int32),
                  var=d.var[cell, snv].astype(np.int32), total=d.total[cell, snv].astype(np.int32),
                  like0=d.like0[cell, snv], like1=d.like1_marg[cell, snv], read_depth=np.asarray(d.read_depth),
                  cell_labels=d.cell_lookup.to_numpy().astype(str), mut_labels=d.mut_lookup.to_numpy().astype(str))
    if np.asarray(d.read_depth).dtype.kind == "i":
        inputs["read_depth"] = np.asarray(d.read_depth).astype(np.int32)
    if inputs_extra:
        inputs.update(inputs_extra)
    np.savez_compressed(os.path.join(HERE, stem + "_inputs.npz"), **inputs)

    REC = Recorder()
    REC.add("check_obs", cells=ph.cells, muts=ph.muts, value=int(UT.check_obs(ph.cells, ph.muts, d.total)))
    buf = io.StringIO()
    with redirect_stdout(buf):
        tree, tree_list, loglikes = ph.phertilize(max_iterations=iterations, starts=starts, seed=seed, radius=1,
                                                  gamma=0.95, d=10, use_copy_kernel=True, post_process=True,
                                                  low_cmb=0.05, high_cmb=0.15, nobs_per_cluster=3, **phert_kw)
    REC.add("final", minobs=int(ph.params.minobs), loglikes=loglikes.to_numpy(), loglike_keys=loglikes.index.to_numpy(),
            norm=tree.norm_loglikelihood, n_trees=tree_list.size(), **tree_fields(tree))
    pcell, pmut, _, _ = tree.generate_results(*ph.get_id_mappings())
    REC.add("final_frames", cell_cluster=pcell["cluster"].to_numpy(), cell_id=pcell["cell"].to_numpy().astype(str),
            mut_cluster=pmut["cluster"].to_numpy(), mut_id=pmut["mutation"].to_numpy().astype(str))
    for i, t in enumerate(tree_list.get_all_trees()):
        REC.add("candidate_tree", index=i, key=int(t.key), norm=t.norm_loglikelihood, **tree_fields(t))
    REC.save(stem + "_calls")
    open(os.path.join(HERE, stem + "_stdout.txt"), "w").write(buf.getvalue())
    kinds = {}
    for e in REC.manifest:
        kinds[e["kind"]] = kinds.get(e["kind"], 0) + 1
    print(stem, "records:", kinds)
    print(stem, "final edges", list(tree.tree.edges()), "norm", tree.norm_loglikelihood)


def kat():
    ks, ns = [], []
    for n in range(1, 61):
        for k in range(0, n + 1):
            ks.append(k)
            ns.append(n)
    ks, ns = np.array(ks), np.array(ns)
    out = {"var": ks.astype(np.int32), "total": ns.astype(np.int32)}
    settings = [(0.001, 5), (0.01, 3), (0.0005, 8)]
    for i, (alpha, mc) in enumerate(settings):
        coeff = scipy.special.comb(ns, ks)
        df = pd.DataFrame({"var": ks, "total": ns})
        out[f"like0_{i}"] = PH.compute_like0(df, alpha, coeff).to_numpy()
        out[f"like1_{i}"] = PH.compute_like1(df, 1, mc, alpha, coeff).to_numpy()
    out["settings"] = np.array(settings)
    np.savez_compressed(os.path.join(HERE, "kat_like.npz"), **out)
    print("kat:", len(ks), "pairs x", len(settings), "settings")


def main():
    import numba
    import scipy
    import sklearn
    import networkx

    versions = dict(numpy=np.__version__, pandas=pd.__version__, scipy=scipy.__version__, sklearn=sklearn.__version__,
                    numba=numba.__version__, networkx=networkx.__version__)
    json.dump(versions, open(os.path.join(HERE, "versions.json"), "w"), indent=1)
    kat()
    install_hooks()

    # ---- config 1: the bundled example exactly as run.py reads it (run.py:16-27; first line dropped)
    vfile, bfile = example_paths()
    col_names = ["chr", "mutation_id", "cell_id", "base", "var", "total", "copies"]
    variant = pd.read_table(vfile, sep="\t", header=None, names=col_names, skiprows=[0])
    variant.drop("copies", inplace=True, axis=1)
    bins = pd.read_csv(bfile)
    run_reference(variant, bins, "example", starts=3, iterations=10, seed=100 * 99059 + 0)

    # ---- small synthetic (T0): D=2 embedding, exercises the singular / low-dimensional Gaussian path
    from phertilizer_b200.synth import make_config, to_dataframes

    s = make_config("T0")
    vdf, bdf = to_dataframes(s)
    run_reference(vdf, bdf, "synth_t0", starts=2, iterations=6, seed=4242,
                  inputs_extra=dict(cell_clone=s.cell_clone.numpy(), snv_node=s.snv_node.numpy(), parent=s.parent))

    # ---- mid-size synthetic (T1, 2000 x 3000, 4 clones): an 8-node tree, exercises reassign_cell and deeper recursion
    s = make_config("T1")
    vdf, bdf = to_dataframes(s)
    run_reference(vdf, bdf, "synth_t1", starts=2, iterations=5, seed=777,
                  inputs_extra=dict(cell_clone=s.cell_clone.numpy(), snv_node=s.snv_node.numpy(), parent=s.parent))


if __name__ == "__main__":
    main()
