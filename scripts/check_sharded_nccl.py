"""torchrun --nproc-per-node N scripts/check_sharded_nccl.py : the NCCL cell-sharded path equals the single-GPU path."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, ".")
from phertilizer_b200.sharded import ShardedStore
from phertilizer_b200.store import ObservationStore
from tests.golden_util import load_golden

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for stem in ("example", "synth_t0"):
    g = load_golden(stem); i = g.inputs
    full = ObservationStore.from_coo(g.n, g.m, i["cell"], i["snv"], i["var"], i["total"], device=local)
    sh = ShardedStore.from_coo(g.n, g.m, i["cell"], i["snv"], i["var"], i["total"], rank, world, local)
    rng = np.random.default_rng(11)
    for trial in range(4):
        lab = np.zeros(g.n + 16, np.uint8)[: g.n]; lab[:] = rng.integers(0, 3, g.n)
        labA = np.zeros(g.n + 16, np.uint8)[: g.n]; labA[:] = (lab == 1)
        sub = None if trial % 2 == 0 else np.sort(rng.choice(g.m, g.m // 3, replace=False)).astype(np.int32)
        l1, n1 = full.assign_linear(labA, float(labA.sum()), sub)
        l2, n2 = sh.assign_linear(labA, float(labA.sum()), sub)
        ok &= np.array_equal(l1, l2) and np.array_equal(n1, n2)
        if rank == 0 and not ok: print('linear mismatch', stem, trial)
        b1, m1 = full.assign_branching(lab, sub)
        b2, m2 = sh.assign_branching(lab, sub)
        t1 = full.snv_table(lab, 2, sub); t2 = sh.snv_table(lab, 2, sub)
        ok &= np.array_equal(m1, m2)
        diff = np.nonzero(b1 != b2)[0]
        if len(diff):
            # the all-reduce adds per-rank fp64 partials, so candidates that are MATHEMATICALLY equal (same multiset of
            # (var,total) pairs in both clusters) can differ in the last bit and resolve to another label; nothing else may
            S0, S1 = t1["S0"], t1["S1"]
            cand = np.stack([S1[:, 0] + S0[:, 1], S0[:, 0] + S1[:, 1], S1[:, 0] + S1[:, 1]])
            for d in diff:
                pair = cand[[b1[d] - 1, b2[d] - 1], d]
                tie = abs(pair[0] - pair[1]) <= 1e-12 * abs(pair).max()
                ok &= bool(tie)
            if rank == 0:
                print(stem, trial, "branching labels differing on exact ties:", len(diff), "of", len(b1), "ok so far", ok)
        ok &= np.array_equal(t1["Nobs"], t2["Nobs"]) and np.allclose(t1["S0"], t2["S0"], rtol=1e-12, atol=0) and np.allclose(t1["S1"], t2["S1"], rtol=1e-12, atol=0)
    full.close()
t = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("sharded NCCL path == single GPU path on", world, "ranks:", bool(t.item()))
dist.destroy_process_group()
sys.exit(0 if ok else 1)
