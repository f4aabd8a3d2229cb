"""End-to-end tree wall time on synthetic data (BASELINE.json metric, second half): Phertilizer(...).phertilize(...)
through the public API, with a per-stage breakdown (GPU store calls vs the host-side cell clustering).

  python scripts/e2e_tree.py C2 [--n N --m M] [--starts S --iterations J] [--post]
Prints one JSON line.  The dense reference cannot be timed beside it at these shapes unless --reference is given
(only sensible for tiny shapes; needs /root/reference, so never on the GPU box).
"""
import argparse, io, json, sys, time
from collections import defaultdict
from contextlib import redirect_stdout

import numpy as np

sys.path.insert(0, ".")

ap = argparse.ArgumentParser()
ap.add_argument("config", nargs="?", default="C2")
ap.add_argument("--n", type=int)
ap.add_argument("--m", type=int)
ap.add_argument("--starts", type=int, default=3)
ap.add_argument("--iterations", type=int, default=10)
ap.add_argument("--seed", type=int, default=9905900)
ap.add_argument("--post", action="store_true")
ap.add_argument("--quiet", action="store_true")
args = ap.parse_args()

from phertilizer_b200 import branching_split as BS, linear_split as LS, utils as U
from phertilizer_b200.clonal_tree import ClonalTree
from phertilizer_b200.phertilizer import Phertilizer
from phertilizer_b200.store import ObservationStore
from phertilizer_b200.synth import CONFIGS, make_config, to_dataframes

over = {}
if args.n: over["n"] = args.n
if args.m: over["m"] = args.m
t0 = time.time()
import torch
s = make_config(args.config, device="cuda" if torch.cuda.is_available() else "cpu", **over)
vc, bc = to_dataframes(s)
t_data = time.time() - t0

acc = defaultdict(float); cnt = defaultdict(int)
def timed(obj, name, label):
    f = getattr(obj, name)
    def w(*a, **k):
        t = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            acc[label] += time.perf_counter() - t; cnt[label] += 1
    setattr(obj, name, w)

for name in ("snv_table", "cell_table", "assign_linear", "assign_branching", "block_sums", "tree_loglik", "snv_node_table",
             "cell_node_table", "gauss_node_ll"):
    timed(ObservationStore, name, "gpu:" + name)
# the modules bind these helpers by name: patch them where they are used
from phertilizer_b200.cluster import CopyDistance
for mod in (LS, BS):
    timed(mod, "impute_mut_features", "host:impute_mut_features")
timed(CopyDistance, "min_cut", "cluster:min_cut(total: GPU affinity + Laplacian, LOBPCG with GPU matmat)")
timed(CopyDistance, "build", "gpu:affinity_build")
timed(CopyDistance, "matmat", "gpu:laplacian_matmat")
timed(LS.Linear_split, "cluster_cells", "split:cluster_cells(total)")
timed(BS.Branching_split, "cluster_cells", "split:cluster_cells(total)")
timed(ClonalTree, "compute_likelihood", "tree:compute_likelihood(total)")
timed(ClonalTree, "post_process", "tree:post_process(total)")

t0 = time.time()
ph = Phertilizer(vc, bc, mode="rd", dim_reduce=False)
t_init = time.time() - t0
buf = io.StringIO()
t0 = time.time()
with redirect_stdout(buf):
    tree, trees, ll = ph.phertilize(max_iterations=args.iterations, starts=args.starts, seed=args.seed, radius=1, gamma=0.95, d=10,
                                    use_copy_kernel=True, post_process=args.post)
t_run = time.time() - t0
if not args.quiet:
    sys.stderr.write(buf.getvalue()[-1500:])
gpu = sum(v for k, v in acc.items() if k.startswith("gpu:") and "affinity" not in k and "matmat" not in k)
out = {"config": args.config, "n": s.n, "m": s.m, "nnz": s.nnz, "starts": args.starts, "iterations": args.iterations,
       "post_process": args.post, "synth_s": round(t_data, 2), "init_s": round(t_init, 2), "phertilize_s": round(t_run, 2),
       "gpu_store_calls_s": round(gpu, 3), "gpu_store_calls": int(sum(c for k, c in cnt.items() if k.startswith("gpu:"))),
       "stages_s": {k: [round(v, 3), cnt[k]] for k, v in sorted(acc.items(), key=lambda kv: -kv[1])},
       "tree_nodes": int(tree.tree.number_of_nodes()), "n_trees": int(trees.size()),
       "norm_loglikelihood": float(tree.norm_loglikelihood)}
try:  # agreement of the inferred cell clusters with the generator's clones (sanity, not parity)
    from sklearn.metrics import adjusted_rand_score

    pred = np.full(s.n, -1)
    for node in tree.tree.nodes():
        pred[np.asarray(tree.get_tip_cells(node), dtype=np.int64)] = node
    out["cell_ari_vs_truth"] = round(float(adjusted_rand_score(s.cell_clone.cpu().numpy(), pred)), 4)
    out["true_clones"] = int(len(s.parent))
except Exception as e:  # pragma: no cover
    out["cell_ari_vs_truth"] = f"n/a ({e})"
print(json.dumps(out))
