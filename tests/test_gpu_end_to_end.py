"""End-to-end drop-in check: `phertilizer_b200.Phertilizer` (GPU store, host control flow) reproduces the reference's
recorded run -- same seed, `--no-umap`, `-s 3 -j 10 --post_process` on the bundled example (config 1) and `-s 2 -j 6`
on the small synthetic.  Assignments and tree must be identical, log-likelihoods within 1e-9 relative."""
import io
from contextlib import redirect_stdout

import numpy as np
import pandas as pd
import pytest

from tests.golden_util import load_golden, rel_close, tree_from_record

pytestmark = pytest.mark.gpu


def frames_from_golden(g):
    i = g.inputs
    mut_labels = i["mut_labels"]
    cell_ids = i["cell_labels"].astype(np.int64)
    chrom = np.array([s.rsplit("_", 1)[0] for s in mut_labels])
    mid = np.array([s.rsplit("_", 1)[1] for s in mut_labels])
    vc = pd.DataFrame({"chr": chrom[i["snv"]], "mutation_id": mid[i["snv"]], "cell_id": cell_ids[i["cell"]], "base": "A",
                       "var": i["var"].astype(np.int64), "total": i["total"].astype(np.int64)})
    rd = i["read_depth"]
    bc = pd.DataFrame(rd, columns=[f"b{k}" for k in range(rd.shape[1])])
    bc.insert(0, "cell", cell_ids)
    # shuffle the rows: the constructor must restore cell-index order (phertilizer.py:259-263)
    return vc.sample(frac=1.0, random_state=3).reset_index(drop=True), bc.sample(frac=1.0, random_state=4).reset_index(drop=True)


@pytest.mark.parametrize("stem,starts,iters,seed", [("synth_t0", 2, 6, 4242), ("example", 3, 10, 100 * 99059), ("synth_t1", 2, 5, 777)])
def test_phertilize_matches_reference_run(stem, starts, iters, seed):
    from phertilizer_b200.phertilizer import Phertilizer

    g = load_golden(stem)
    vc, bc = frames_from_golden(g)
    ph = Phertilizer(vc, bc, alpha=0.001, max_copies=5, mode="rd", dim_reduce=False)
    assert list(ph.data.mut_lookup.to_numpy().astype(str)) == list(g.inputs["mut_labels"])
    buf = io.StringIO()
    with redirect_stdout(buf):
        tree, tree_list, loglikes = ph.phertilize(max_iterations=iters, starts=starts, seed=seed, radius=1, gamma=0.95, d=10,
                                                  use_copy_kernel=True, post_process=True, low_cmb=0.05, high_cmb=0.15,
                                                  nobs_per_cluster=3)
    fin = g.of("final")[0]
    assert ph.params.minobs == fin.minobs
    assert tree_list.size() == fin.n_trees
    assert rel_close(loglikes.to_numpy(), fin.loglikes, 1e-9), (loglikes.to_numpy(), fin.loglikes)
    assert list(loglikes.index) == list(fin.loglike_keys)
    # candidate trees: identical topology, cells and SNVs per node
    for rec, t in zip(g.of("candidate_tree"), tree_list.get_all_trees()):
        nodes, edges, cells, muts = tree_from_record(rec)
        assert sorted(t.tree.edges()) == sorted(edges)
        for k in nodes:
            assert np.array_equal(t.get_tip_cells(k), cells[k]), (rec.index, k)
            assert np.array_equal(np.asarray(t.get_tip_muts(k)).astype(np.int64), muts[k]), (rec.index, k)
    # final (post-processed) tree
    nodes, edges, cells, muts = tree_from_record(fin)
    assert sorted(tree.tree.edges()) == sorted(edges)
    for k in nodes:
        assert np.array_equal(tree.get_tip_cells(k), cells[k])
        assert np.array_equal(np.asarray(tree.get_tip_muts(k)).astype(np.int64), muts[k])
    assert rel_close(tree.norm_loglikelihood, fin.norm, 1e-9)
    pe = g.of("post_process_end")[0]
    assert rel_close(tree.variant_likelihood, pe.variant, 1e-9) and rel_close(tree.loglikelihood, pe.loglikelihood, 1e-9)
    # output frames (generate_results)
    ff = g.of("final_frames")[0]
    pcell, pmut, _, _ = tree.generate_results(*ph.get_id_mappings())
    assert np.array_equal(pcell["cluster"].to_numpy(), ff.cell_cluster) and list(pcell["cell"].astype(str)) == list(ff.cell_id)
    assert np.array_equal(pmut["cluster"].to_numpy(), ff.mut_cluster) and list(pmut["mutation"].astype(str)) == list(ff.mut_id)


def test_pairwise_distance_bitwise_vs_scipy():
    """copy_distance feeds quantiles and exp() kernels of the cell clustering; report exactness vs scipy."""
    from scipy.spatial.distance import pdist, squareform

    from phertilizer_b200.store import pairwise_dist

    g = load_golden("example")
    X = g.inputs["read_depth"].astype(np.float64)
    got = pairwise_dist(X)
    ref = squareform(pdist(X, metric="euclidean"))
    assert rel_close(got, ref, 1e-12)
    frac_exact = float(np.mean(got == ref))
    print("pairwise distance bit-exact fraction:", frac_exact)
    assert frac_exact > 0.5


def test_runs_fanout_equals_sequential_loop():
    """`--runs` spread over worker processes (two workers; on a 1-GPU box both share device 0) must return, run by run,
    what the reference's sequential loop (run.py:55-74) returns, and pick the same winner."""
    import torch

    from phertilizer_b200.fanout import best_run, run_many
    from phertilizer_b200.phertilizer import Phertilizer

    g = load_golden("synth_t0")
    vc, bc = frames_from_golden(g)
    ctor = dict(alpha=0.001, max_copies=5, mode="rd", dim_reduce=False)
    kws = [dict(max_iterations=4, starts=2, seed=100 * 7 + i, radius=1, gamma=0.95, d=10, use_copy_kernel=True,
                post_process=False) for i in range(3)]
    devices = [0, 1] if torch.cuda.device_count() > 1 else [0, 0]
    results, (cell_lookup, mut_lookup) = run_many(vc, bc, ctor, kws, devices)
    ph = Phertilizer(vc, bc, **ctor)
    assert list(cell_lookup) == list(ph.get_id_mappings()[0]) and list(mut_lookup) == list(ph.get_id_mappings()[1])
    seq = []
    for kw in kws:
        with redirect_stdout(io.StringIO()):
            seq.append(ph.phertilize(**kw))
    for (tree, tl, ll, _), (t2, tl2, ll2) in zip(results, seq):
        assert tree.norm_loglikelihood == t2.norm_loglikelihood and tl.size() == tl2.size()
        assert sorted(tree.tree.edges()) == sorted(t2.tree.edges())
        for k in tree.tree.nodes():
            assert np.array_equal(tree.get_tip_cells(k), t2.get_tip_cells(k))
            assert np.array_equal(tree.get_tip_muts(k), t2.get_tip_muts(k))
        assert np.array_equal(ll.to_numpy(), ll2.to_numpy())
    best_seq, bl = None, -np.inf
    for i, r in enumerate(seq):
        if r[0].norm_loglikelihood > bl:
            bl, best_seq = r[0].norm_loglikelihood, i
    assert best_run(results) == best_seq


@pytest.mark.parametrize("case", ["mid", "c2"])
def test_mid_size_run_matches_reference(case):
    """4000 x 20000 (3.9 M observations) and BASELINE config 2 itself, 10000 x 50000 (24 M): the unmodified reference's
    end-to-end result (tests/golden/make_golden_mid.py, dense path) vs phertilizer_b200 on the GPU -- same candidate trees,
    same cell and SNV assignments, log-likelihoods to 1e-9."""
    import json
    import os

    from phertilizer_b200.phertilizer import Phertilizer
    from tests.golden.make_golden_mid import CASES, make_inputs

    RUN = CASES[case][1]
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    if not os.path.exists(os.path.join(here, f"{case}_e2e.npz")):
        pytest.skip(f"golden '{case}' not generated (tests/golden/make_golden_mid.py {case})")
    meta = json.load(open(os.path.join(here, f"{case}_e2e.json")))
    gold = np.load(os.path.join(here, f"{case}_e2e.npz"))
    s, (vc, bc), digest = make_inputs(case)
    if digest != meta["inputs_sha256"]:
        pytest.skip("the seeded generator produced different inputs on this machine (torch CPU RNG): golden not comparable")
    ph = Phertilizer(vc, bc, alpha=0.001, max_copies=5, mode="rd", dim_reduce=False)
    with redirect_stdout(io.StringIO()):
        tree, tree_list, loglikes = ph.phertilize(**RUN)
    assert ph.params.minobs == meta["minobs"] and tree_list.size() == meta["n_trees"]
    assert rel_close(loglikes.to_numpy(), gold["loglikes"], 1e-9)

    def same(prefix, t):
        assert sorted(int(v) for v in t.tree.nodes()) == list(gold[prefix + "nodes"])
        assert sorted((int(a), int(b)) for a, b in t.tree.edges()) == [tuple(e) for e in gold[prefix + "edges"]]
        for k in gold[prefix + "nodes"]:
            assert np.array_equal(np.asarray(t.get_tip_cells(int(k)), np.int64), gold[f"{prefix}cells_{k}"])
            assert np.array_equal(np.asarray(t.get_tip_muts(int(k))).astype(np.int64), gold[f"{prefix}muts_{k}"])

    for i, t in enumerate(tree_list.get_all_trees()):
        same(f"cand{i}_", t)
    same("final_", tree)
    assert rel_close(tree.norm_loglikelihood, float(meta["norm"]), 1e-9)
    assert rel_close(tree.variant_likelihood, float(meta["variant_likelihood"]), 1e-9)
