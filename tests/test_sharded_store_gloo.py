"""`CellShardedStore` (phertilizer_b200/sharded_store.py) on CPU: two gloo ranks, cells sharded, every method of the store
API must return on EVERY rank what one store over all cells returns.  The kernels are replaced by the dense-oracle test
double (tests/oracle_store.py) on the rank's block of cells; partitioning, label slicing, assembling the per-cell blocks,
the collectives and the fixed-order reductions on the assembled tables are the product code."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle_np as O
        from phertilizer_b200.sharded import ShardedStore, cell_block
        from phertilizer_b200.sharded_store import CellShardedStore
        from tests.golden_util import load_golden
        from tests.oracle_store import OracleStore
        from tests.test_sharded_gloo import CpuBackend

        g = load_golden("synth_t0")
        i = g.inputs
        n, m = g.n, g.m
        lo, hi = cell_block(n, rank, world)
        keep = (i["cell"] >= lo) & (i["cell"] < hi)
        local = OracleStore(hi - lo, m, i["cell"][keep] - lo, i["snv"][keep], i["var"][keep], i["total"][keep])
        full = OracleStore(n, m, i["cell"], i["snv"], i["var"], i["total"])
        k1 = ShardedStore(CpuBackend(local.D), n, m, lo, hi)
        sh = CellShardedStore(local, n, m, lo, hi, None, k1, dense=full)
        rng = np.random.default_rng(5)
        ok = True

        def same(a, b, what):
            nonlocal ok
            a, b = np.asarray(a), np.asarray(b)
            good = a.shape == b.shape and (np.array_equal(a, b) if a.dtype.kind in "iu" else np.allclose(a, b, rtol=1e-12, atol=1e-300, equal_nan=True))
            if not good:
                print(f"[rank {rank}] mismatch in {what}")
            ok &= bool(good)

        for trial in range(3):
            cl = rng.integers(0, 4, n).astype(np.uint8)          # 3 cell clusters + excluded
            sl = rng.integers(0, 4, m).astype(np.uint8)
            sub_s = np.sort(rng.choice(m, m // 3, replace=False))
            sub_c = np.sort(rng.choice(n, n // 2, replace=False))
            for lst in (None, sub_s):
                a, b = sh.snv_table(cl, 3, lst), full.snv_table(cl, 3, lst)
                for k in ("S0", "S1", "Nobs", "Nvar"):
                    same(a[k], b[k], f"snv_table {k}")
            for lst in (None, sub_c):
                a, b = sh.cell_table(sl, 3, lst), full.cell_table(sl, 3, lst)
                for k in ("S0", "S1", "Nobs", "Nvar"):
                    same(a[k], b[k], f"cell_table {k}")
            for C, Q in ((1, 1), (2, 2), (3, 3)):
                a = sh.block_sums(np.minimum(cl, C), C, np.minimum(sl, Q), Q)
                b = full.block_sums(np.minimum(cl, C), C, np.minimum(sl, Q), Q)
                for x, y, k in zip(a, b, ("sum0", "sum1", "nobs", "nvar")):
                    same(np.asarray(x).reshape(Q, C), np.asarray(y).reshape(Q, C), f"block_sums {k} C={C} Q={Q}")
            labA = (cl == 1).astype(np.uint8)
            a, b = sh.assign_linear(labA, float(labA.sum())), full.assign_linear(labA, float(labA.sum()))
            same(a[0], b[0], "assign_linear labels"); same(a[1], b[1], "assign_linear nobs")
            lab2 = np.where(cl > 2, 0, cl).astype(np.uint8)
            a, b = sh.assign_branching(lab2, sub_s), full.assign_branching(lab2, sub_s)
            same(np.asarray(a[1]).reshape(-1, 2), np.asarray(b[1]).reshape(-1, 2), "assign_branching nobs")
            diff = np.nonzero(np.asarray(a[0]) != np.asarray(b[0]))[0]
            if len(diff):   # fp64 all-reduce transport of the test double: exact ties only
                cand = np.stack(O.branching_candidates(full.D, np.nonzero(lab2 == 1)[0], np.nonzero(lab2 == 2)[0], sub_s))
                for d in diff:
                    pair = cand[[a[0][d] - 1, b[0][d] - 1], d]
                    ok &= bool(abs(pair[0] - pair[1]) <= 1e-12 * abs(pair).max())
            K = 3
            cn = np.where(cl == 0, 255, cl - 1).astype(np.uint8)
            sn = np.where(sl == 0, 255, sl - 1).astype(np.uint8)
            present = np.tril(np.ones((K, K), np.uint8))
            pa, pca = sh.tree_loglik(cn, sn, present, per_cell=True)
            pb, pcb = full.tree_loglik(cn, sn, present, per_cell=True)
            same(pa, pb, "tree_loglik per node"); same(pca, pcb, "tree_loglik per cell")
            same(sh.tree_loglik(cn, sn, present), pb, "tree_loglik")
            member = present.T.copy()
            same(sh.snv_node_table(cn, member, sub_s), full.snv_node_table(cn, member, sub_s), "snv_node_table")
            for lst in (None, sub_c):
                same(sh.cell_node_table(sn, member, lst), full.cell_node_table(sn, member, lst), "cell_node_table")
            X = rng.standard_normal((n, 2))
            same(sh.gauss_node_ll(X, cn, K), full.gauss_node_ll(X, cn, K), "gauss_node_ll")
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_cell_sharded_store():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
