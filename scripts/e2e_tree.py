"""End-to-end tree wall time on synthetic data (BASELINE.json metric, second half): Phertilizer(...).phertilize(...)
through the public API, with a per-stage breakdown (GPU store calls vs the host-side cell clustering).

  python scripts/e2e_tree.py C2 [--n N --m M] [--starts S --iterations J] [--post]
  torchrun --nproc-per-node N scripts/e2e_tree.py C3 ...        cells sharded over N GPUs (Phertilizer(device_group=True))
  torchrun --nproc-per-node N scripts/e2e_tree.py T1 --one-gpu  every rank on cuda:0, gloo rendez-vous (tests)
Prints one JSON line (rank 0) with a `tree_signature` that must not depend on the number of GPUs.
"""
import argparse, hashlib, io, json, os, sys, time
from collections import defaultdict
from contextlib import redirect_stdout

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("config", nargs="?", default="C2")
ap.add_argument("--n", type=int)
ap.add_argument("--m", type=int)
ap.add_argument("--starts", type=int, default=3)
ap.add_argument("--iterations", type=int, default=10)
ap.add_argument("--seed", type=int, default=9905900)
ap.add_argument("--post", action="store_true")
ap.add_argument("--quiet", action="store_true")
ap.add_argument("--one-gpu", action="store_true", help="multi-rank run with every rank on cuda:0 (gloo)")
ap.add_argument("--out", help="also write the JSON line to this file (rank 0)")
args = ap.parse_args()

import torch

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = 0 if args.one_gpu else int(os.environ.get("LOCAL_RANK", "0"))
if torch.cuda.is_available():
    torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist

    if args.one_gpu:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

from phertilizer_b200 import branching_split as BS, linear_split as LS, utils as U
from phertilizer_b200.clonal_tree import ClonalTree
from phertilizer_b200.phertilizer import Phertilizer
from phertilizer_b200.sharded_store import CellShardedStore
from phertilizer_b200.store import ObservationStore
from phertilizer_b200.synth import CONFIGS, make_config, to_dataframes

over = {}
if args.n: over["n"] = args.n
if args.m: over["m"] = args.m
t0 = time.time()
s = make_config(args.config, device=f"cuda:{local}" if torch.cuda.is_available() else "cpu", **over)
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

store_cls = CellShardedStore if world > 1 else ObservationStore
for name in ("snv_table", "cell_table", "assign_linear", "assign_branching", "block_sums", "tree_loglik", "snv_node_table",
             "cell_node_table", "gauss_node_ll"):
    timed(store_cls, name, "gpu:" + name)
# the modules bind these helpers by name: patch them where they are used
from phertilizer_b200.cluster import CopyDistance
for mod in (LS, BS):
    timed(mod, "impute_mut_features", "host:impute_mut_features")
from phertilizer_b200.cluster import ShardedCopyDistance
cd_cls = ShardedCopyDistance if world > 1 else CopyDistance
timed(CopyDistance, "min_cut", "cluster:min_cut(total: GPU affinity + Laplacian, LOBPCG)")
timed(cd_cls, "build", "gpu:affinity_build")
timed(cd_cls, "matmat", "gpu:laplacian_matmat(host operands)")
timed(cd_cls, "bv_apply", "gpu:laplacian_product(device-resident block)")
timed(CopyDistance, "bv_gram", "gpu:lobpcg_gram")
timed(CopyDistance, "bv_combine", "gpu:lobpcg_combine")
timed(LS.Linear_split, "cluster_cells", "split:cluster_cells(total)")
timed(BS.Branching_split, "cluster_cells", "split:cluster_cells(total)")
timed(ClonalTree, "compute_likelihood", "tree:compute_likelihood(total)")
timed(ClonalTree, "post_process", "tree:post_process(total)")

t0 = time.time()
ph = Phertilizer(vc, bc, mode="rd", dim_reduce=False, device=local, device_group=True if world > 1 else None)
if world > 1:
    dist.barrier()
t_init = time.time() - t0
buf = io.StringIO()
t0 = time.time()
with redirect_stdout(buf):
    tree, trees, ll = ph.phertilize(max_iterations=args.iterations, starts=args.starts, seed=args.seed, radius=1, gamma=0.95, d=10,
                                    use_copy_kernel=True, post_process=args.post)
if world > 1:
    dist.barrier()
t_run = time.time() - t0
if not args.quiet and rank == 0:
    sys.stderr.write(buf.getvalue()[-1500:])


def signature(tree, ll):
    """What must be identical for every number of GPUs: topology, every cell / SNV assignment, the likelihoods' digits."""
    cell_node = np.full(s.n, -1, np.int64)
    snv_node = np.full(s.m, -1, np.int64)
    for node in tree.tree.nodes():
        cell_node[np.asarray(tree.get_tip_cells(node), dtype=np.int64)] = node
        snv_node[np.asarray(tree.get_tip_muts(node), dtype=np.int64)] = node
    return {"edges": sorted([int(a), int(b)] for a, b in tree.tree.edges()),
            "cells_per_node": {int(k): int((cell_node == k).sum()) for k in sorted(tree.tree.nodes())},
            "snvs_per_node": {int(k): int((snv_node == k).sum()) for k in sorted(tree.tree.nodes())},
            "assignments_sha1": hashlib.sha1(cell_node.tobytes() + snv_node.tobytes()).hexdigest(),
            "norm_loglikelihood": repr(float(tree.norm_loglikelihood)),
            "candidate_loglikelihoods_sha1": hashlib.sha1(np.asarray(ll.to_numpy(), np.float64).tobytes()).hexdigest(),
            "n_candidates": int(len(ll))}


gpu = sum(v for k, v in acc.items() if k.startswith("gpu:") and "affinity" not in k and "matmat" not in k)
out = {"config": args.config, "n": s.n, "m": s.m, "nnz": s.nnz, "gpus": world, "starts": args.starts, "iterations": args.iterations,
       "post_process": args.post, "synth_s": round(t_data, 2), "init_s": round(t_init, 2), "phertilize_s": round(t_run, 2),
       "gpu_store_calls_s": round(gpu, 3), "gpu_store_calls": int(sum(c for k, c in cnt.items() if k.startswith("gpu:"))),
       "stages_s": {k: [round(v, 3), cnt[k]] for k, v in sorted(acc.items(), key=lambda kv: -kv[1])},
       "tree_nodes": int(tree.tree.number_of_nodes()), "n_trees": int(trees.size()),
       "norm_loglikelihood": float(tree.norm_loglikelihood), "tree_signature": signature(tree, ll)}
if world > 1:
    out["k1_exchange"] = ph.data.store.info().get("k1_exchange")
    sigs = [None] * world
    dist.all_gather_object(sigs, json.dumps(out["tree_signature"], sort_keys=True))
    out["all_ranks_agree"] = len(set(sigs)) == 1
try:  # agreement of the inferred cell clusters with the generator's clones (sanity, not parity)
    from sklearn.metrics import adjusted_rand_score

    pred = np.full(s.n, -1)
    for node in tree.tree.nodes():
        pred[np.asarray(tree.get_tip_cells(node), dtype=np.int64)] = node
    out["cell_ari_vs_truth"] = round(float(adjusted_rand_score(s.cell_clone.cpu().numpy(), pred)), 4)
    out["true_clones"] = int(len(s.parent))
except Exception as e:  # pragma: no cover
    out["cell_ari_vs_truth"] = f"n/a ({e})"
if rank == 0:
    print(json.dumps(out))
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        open(args.out, "w").write(json.dumps(out) + "\n")
if world > 1:
    dist.destroy_process_group()
