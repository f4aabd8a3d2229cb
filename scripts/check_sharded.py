"""torchrun --nproc-per-node N scripts/check_sharded.py [--one-gpu] : the sharded K1 steps equal the single-GPU path.

Exercises, on N ranks: the SNV-sharded step (one fused kernel per rank, results pushed as self-validating words over
peer memory), the cell-sharded peer step (integer histograms over peer memory) and -- with the NCCL backend -- the
all-reduce transport, on linear and branching reassignment, plus K2 / K3 / tree log-likelihood under cell sharding.
Peer transports must be BIT-IDENTICAL to the single-GPU call; the fp64 all-reduce may differ on exact ties only.

--one-gpu: every rank uses cuda:0 and the rendez-vous runs on gloo.  CUDA IPC maps a window of another process on the
SAME device just as well, and the kernels do not assume that all ranks are co-resident (the GPU time-slices the
processes), so the driver's single-GPU box covers the multi-rank exchange with world > 1."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phertilizer_b200.sharded import ShardedStore, SnvShardedStore
from phertilizer_b200.store import ObservationStore
from phertilizer_b200.synth import make_config
from tests.golden_util import load_golden

ap = argparse.ArgumentParser()
ap.add_argument("--one-gpu", action="store_true")
ap.add_argument("--quick", action="store_true", help="fewer datasets / trials (the time-sliced one-GPU mode is slow)")
args = ap.parse_args()

if args.one_gpu:   # the ranks time-slice ONE GPU: a rank may wait several scheduler slices for its peer's kernel to run
    os.environ.setdefault("PH_PEER_TIMEOUT_MS", "20000")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = 0 if args.one_gpu else int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
if args.one_gpu:
    dist.init_process_group("gloo")
else:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nccl = dist.get_backend() == "nccl"


def datasets():
    for stem in (("synth_t0",) if args.quick else ("example", "synth_t0")):
        g = load_golden(stem); i = g.inputs
        yield stem, g.n, g.m, i["cell"], i["snv"], i["var"], i["total"]
    s = make_config("T1", with_embedding=False)
    yield "T1", s.n, s.m, s.cell.numpy(), s.snv.numpy(), s.var.numpy(), s.total.numpy()
    if not args.quick:
        # > 65520 cells: two chunks of the label image; m not a multiple of 8 * world
        s = make_config("C2", with_embedding=False, n=70000, m=3001)
        yield "C2w", s.n, s.m, s.cell.numpy(), s.snv.numpy(), s.var.numpy(), s.total.numpy()


def labels(rng, n):
    lab = np.zeros(n + 16, np.uint8)[:n]; lab[:] = rng.integers(0, 3, n)
    labA = np.zeros(n + 16, np.uint8)[:n]; labA[:] = (lab == 1)
    return lab, labA


ok = True
used = set()
trials = 2 if args.quick else 4
for stem, n, m, cell, snv, var, total in datasets():
    full = ObservationStore.from_coo(n, m, cell, snv, var, total, device=local)
    for exchange in (("peer", "nccl") if nccl else ("peer",)):
        if stem == "C2w" and exchange == "peer" and (n + world - 1) // world > 65535:
            continue
        sh = ShardedStore.from_coo(n, m, cell, snv, var, total, rank, world, local, exchange="auto" if exchange == "peer" else "nccl")
        used.add(sh.exchange)
        exact = sh.exchange == "peer"
        rng = np.random.default_rng(11)
        for trial in range(trials):
            lab, labA = labels(rng, n)
            sub = None if trial % 2 == 0 else np.sort(rng.choice(m, m // 3, replace=False)).astype(np.int32)
            l1, n1 = full.assign_linear(labA, float(labA.sum()), sub)
            l2, n2 = sh.assign_linear(labA, float(labA.sum()), sub)
            good = np.array_equal(l1, l2) and np.array_equal(n1, n2)
            if not good: print(f"[rank {rank}] linear mismatch", stem, exchange, trial, int((l1 != l2).sum()), int((n1 != n2).sum()))
            ok &= good
            b1, m1 = full.assign_branching(lab, sub)
            b2, m2 = sh.assign_branching(lab, sub)
            good = np.array_equal(m1, m2)
            if not good: print(f"[rank {rank}] branching nobs mismatch", stem, exchange, trial)
            ok &= good
            diff = np.nonzero(b1 != b2)[0]
            if len(diff) and exact:
                print(f"[rank {rank}] peer exchange: {len(diff)} branching labels differ", stem, trial)
                ok = False
            elif len(diff):
                t1 = full.snv_table(lab, 2, sub)
                S0, S1 = t1["S0"], t1["S1"]
                cand = np.stack([S1[:, 0] + S0[:, 1], S0[:, 0] + S1[:, 1], S1[:, 0] + S1[:, 1]])
                for d in diff:
                    pair = cand[[b1[d] - 1, b2[d] - 1], d]
                    ok &= bool(abs(pair[0] - pair[1]) <= 1e-12 * abs(pair).max())
            if not exact:
                t1 = full.snv_table(lab, 2, sub); t2 = sh.snv_table(lab, 2, sub)
                ok &= np.array_equal(t1["Nobs"], t2["Nobs"]) and np.allclose(t1["S0"], t2["S0"], rtol=1e-12, atol=0) and np.allclose(t1["S1"], t2["S1"], rtol=1e-12, atol=0)
        if exchange == "peer":
            # the rest of the store API under cell sharding: K2 (local + gather), K3 and tree log-likelihood (scalar all-reduce)
            lab, _ = labels(rng, n)
            sl = np.zeros(m + 16, np.uint8)[:m]; sl[:] = rng.integers(0, 3, m)
            c1, c2 = full.cell_table(sl, 2), sh.cell_table(sl, 2)
            ok &= all(np.array_equal(c1[k], c2[k]) for k in c1)
            b1, b2 = full.block_sums(lab, 2, sl, 2), sh.block_sums(lab, 2, sl, 2)
            ok &= np.array_equal(b1[2], b2[2]) and np.array_equal(b1[3], b2[3])
            ok &= np.allclose(b1[0], b2[0], rtol=1e-12, atol=0) and np.allclose(b1[1], b2[1], rtol=1e-12, atol=0)
            cn = np.full(n + 16, 255, np.uint8)[:n]; cn[lab > 0] = lab[lab > 0] - 1
            sn = np.full(m + 16, 255, np.uint8)[:m]; sn[sl > 0] = sl[sl > 0] - 1
            pres = np.array([[1, 0], [1, 1]], np.uint8)
            ok &= np.allclose(full.tree_loglik(cn, sn, pres), sh.tree_loglik(cn, sn, pres), rtol=1e-12, atol=0)
            if not ok: print(f"[rank {rank}] K2/K3/tree mismatch", stem)
        if sh.peer is not None:
            ok &= not sh.peer.timed_out()
            sh.peer.close()
        sh.b.store.close()
    # SNV-sharded: the single-GPU kernel on a slice of the SNVs, outputs pushed to every rank -> exact equality.
    # Linear and branching steps alternate (they use 2 resp. 3 words per SNV of the same window).
    ss = SnvShardedStore.from_coo(n, m, cell, snv, var, total, rank, world, local)
    used.add("snv-sharded peer")
    rng = np.random.default_rng(12)
    for trial in range(trials + 1):
        if trial == trials:
            ss.store.set_group(4, 4)   # 4 lines per batch wherever the flat stream is walked
        lab, labA = labels(rng, n)
        l1, n1 = full.assign_linear(labA, float(labA.sum()))
        l2, n2 = ss.assign_linear(labA, float(labA.sum()))
        b1, m1 = full.assign_branching(lab)
        b2, m2 = ss.assign_branching(lab)
        l3, n3 = ss.assign_linear(labA, float(labA.sum()))
        good = (np.array_equal(l1, l2) and np.array_equal(n1, n2) and np.array_equal(b1, b2) and np.array_equal(m1, m2)
                and np.array_equal(l1, l3) and np.array_equal(n1, n3))
        if not good: print(f"[rank {rank}] SNV-sharded mismatch", stem, trial, int((l1 != l2).sum()), int((n1 != n2).sum()), int((b1 != b2).sum()))
        ok &= good
    ok &= not ss.peer.timed_out()
    ss.close()
    full.close()
# ---- the cell clustering sharded by rows (cluster.ShardedCopyDistance) against the single-GPU context: bit-identical
from phertilizer_b200.cluster import CopyDistance, ShardedCopyDistance

s = make_config("T1")
emb = s.embedding.numpy()
one = CopyDistance(emb, device=local)
shd = ShardedCopyDistance(emb, device=local, min_rows=64)
rng = np.random.default_rng(21)
for trial, (k, F, use_copy, strict) in enumerate(((s.n, 1, True, False), (777, 2, True, True), (1200, 1, False, False), (50, 1, True, False))):
    cells = np.sort(rng.choice(s.n, k, replace=False)).astype(np.int32)
    feats = rng.random((k, F))
    st1, dd1 = one.build(cells, feats, use_copy, 1.0, strict)
    st2, dd2 = shd.build(cells, feats, use_copy, 1.0, strict)
    good = all(st1[q] == st2[q] for q in ("kw", "kwc", "r", "has_nan")) and np.array_equal(dd1, dd2)
    X = rng.standard_normal((k, 3))
    good = good and np.array_equal(one.matmat(X), shd.matmat(X))
    l1 = one.min_cut(cells, feats, use_copy, 1.0, strict, np.random.RandomState(5))
    l2 = shd.min_cut(cells, feats, use_copy, 1.0, strict, np.random.RandomState(5))
    good = good and np.array_equal(l1, l2)
    if not good: print(f"[rank {rank}] sharded clustering mismatch, trial {trial}")
    ok &= bool(good)
mask = (rng.random(s.n) < 0.6).astype(np.uint8)
rows = np.sort(rng.choice(s.n, 300, replace=False)).astype(np.int32)
i1, t1 = one.nearest(rows, mask)
i2, t2 = shd.nearest(rows, mask)
ok &= bool(np.array_equal(i1, i2) and np.array_equal(t1, t2))
used.add("row-sharded clustering")
one.close(); shd.close()

t = torch.tensor([1 if ok else 0])
if nccl:
    t = t.cuda()
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("exchanges exercised:", sorted(used))
    print("sharded path == single GPU path on", world, "ranks:", bool(t.item()))
dist.destroy_process_group()
sys.exit(0 if ok else 1)
