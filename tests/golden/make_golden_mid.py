"""Mid-size end-to-end golden: RUN THE UNMODIFIED REFERENCE (build container only) on a seeded synthetic data set two
orders of magnitude larger than the per-call goldens -- 4000 cells x 20000 SNVs at 0.05x coverage, 5 clones (a 2/5-scale
BASELINE config 2; the dense reference needs 4 x 640 MB here, 4 x 4 GB + hours of O(n^2 log n) imputation sorts on the
full config 2).  Only results are stored (the inputs are regenerated from the seed; their SHA-256 is recorded so that the
test can tell "different data" from "different answer"):

    python tests/golden/make_golden_mid.py [mid|c2]     # -> tests/golden/<case>_e2e.npz + <case>_e2e.json
"""
import hashlib
import io
import json
import os
import sys
import time
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CASES = {
    # 2/5-scale config 2: minutes for the dense reference
    "mid": (dict(n=4000, m=20000, cov=0.05, K=5, D=2, seed=2026),
            dict(max_iterations=4, starts=2, seed=31337, radius=1, gamma=0.95, d=10, use_copy_kernel=True, post_process=True,
                 low_cmb=0.05, high_cmb=0.15, nobs_per_cluster=3)),
    # BASELINE.json configs[1] itself (synth.CONFIGS["C2"]): 4 x 4 GB dense matrices, about an hour for the reference
    "c2": (dict(n=10_000, m=50_000, cov=0.05, K=5, D=2, seed=1028),
           dict(max_iterations=3, starts=1, seed=9905900, radius=1, gamma=0.95, d=10, use_copy_kernel=True, post_process=True,
                low_cmb=0.05, high_cmb=0.15, nobs_per_cluster=3)),
}
CFG, RUN = CASES["mid"]


def make_inputs(case="mid"):
    from phertilizer_b200.synth import make_synthetic, to_dataframes

    s = make_synthetic(device="cpu", **CASES[case][0])
    h = hashlib.sha256()
    for t in (s.cell, s.snv, s.var, s.total):
        h.update(t.numpy().tobytes())
    h.update(s.embedding.numpy().tobytes())
    return s, to_dataframes(s), h.hexdigest()


def tree_arrays(prefix, tree, out):
    nodes = sorted(int(v) for v in tree.tree.nodes())
    out[prefix + "nodes"] = np.array(nodes, np.int32)
    out[prefix + "edges"] = np.array(sorted((int(a), int(b)) for a, b in tree.tree.edges()), np.int32).reshape(-1, 2)
    for k in nodes:
        out[f"{prefix}cells_{k}"] = np.asarray(tree.get_tip_cells(k), np.int32)
        out[f"{prefix}muts_{k}"] = np.asarray(tree.get_tip_muts(k)).astype(np.int32)


if __name__ == "__main__":
    from oracle.ref_shim import load_reference

    case = sys.argv[1] if len(sys.argv) > 1 else "mid"
    CFG, RUN = CASES[case]
    R = load_reference()
    s, (vc, bc), digest = make_inputs(case)
    t0 = time.time()
    ph = R["phertilizer"].Phertilizer(vc, bc, alpha=0.001, max_copies=5, mode="rd", dim_reduce=False)
    t_init = time.time() - t0
    buf = io.StringIO()
    t0 = time.time()
    with redirect_stdout(buf):
        tree, tree_list, loglikes = ph.phertilize(**RUN)
    t_run = time.time() - t0
    out = {}
    tree_arrays("final_", tree, out)
    for i, t in enumerate(tree_list.get_all_trees()):
        tree_arrays(f"cand{i}_", t, out)
    out["loglikes"] = loglikes.to_numpy()
    np.savez_compressed(os.path.join(HERE, f"{case}_e2e.npz"), **out)
    meta = dict(cfg=CFG, run=RUN, inputs_sha256=digest, nnz=int(s.nnz), n_trees=int(tree_list.size()), minobs=int(ph.params.minobs),
                norm=repr(float(tree.norm_loglikelihood)), loglikelihood=repr(float(tree.loglikelihood)),
                variant_likelihood=repr(float(tree.variant_likelihood)),
                reference_init_s=round(t_init, 1), reference_phertilize_s=round(t_run, 1), reference_threads=1)
    json.dump(meta, open(os.path.join(HERE, f"{case}_e2e.json"), "w"), indent=1)
    print(json.dumps(meta))
