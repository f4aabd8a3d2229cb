"""N > 1 path on CPU: two gloo ranks, cells sharded, per-SNV partial tables all-reduced, K1a on the sum.
The two kernels are replaced by a CPU stand-in (dense oracle on the rank's block of cells); everything else --
partition, label slicing, packing, the collective, the decision rules -- is the product code in
phertilizer_b200/sharded.py.  Results must equal the single-shard oracle on every rank."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class CpuBackend:
    """CPU stand-in for CudaBackend (tests only)."""

    def __init__(self, dense):
        self.D = dense

    def empty(self, n, dtype):
        return torch.empty(n, dtype=dtype)

    def to_device(self, arr):
        return torch.from_numpy(np.ascontiguousarray(arr))

    def snv_partial(self, label_t, C, packed_t, list_t=None):
        from oracle import oracle_np as O

        lab = label_t.numpy()[: self.D.n]
        muts = np.arange(self.D.m) if list_t is None else list_t.numpy().astype(np.int64)
        S0, S1, No, _ = O.snv_table(self.D, [np.nonzero(lab == c)[0] for c in range(1, C + 1)], muts)
        packed_t.copy_(torch.from_numpy(np.concatenate([S0.reshape(-1), S1.reshape(-1), No.reshape(-1).astype(np.float64)])))

    def finish_linear(self, packed_t, k, lenA, lab_t, nobs_t):
        p = packed_t.numpy()
        lab_t[:k] = torch.from_numpy(np.where(p[:k] / lenA >= p[k:2 * k] / lenA, 2, 1).astype(np.uint8))
        nobs_t.copy_(torch.from_numpy(p[2 * k:3 * k].astype(np.int32)))

    def finish_branching(self, packed_t, k, lab_t, nobs_t):
        p = packed_t.numpy().reshape(3, k, 2)
        cA, cB, cC = p[1, :, 0] + p[0, :, 1], p[0, :, 0] + p[1, :, 1], p[1, :, 0] + p[1, :, 1]
        mx = np.maximum(cA, np.maximum(cB, cC))
        lab_t[:k] = torch.from_numpy(np.where(cA == mx, 1, np.where(cB == mx, 2, 3)).astype(np.uint8))
        nobs_t.copy_(torch.from_numpy(p[2].reshape(-1).astype(np.int32)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle_np as O
        from phertilizer_b200.sharded import ShardedStore, cell_block
        from tests.golden_util import load_golden

        g = load_golden("synth_t0")
        i = g.inputs
        lo, hi = cell_block(g.n, rank, world)
        keep = (i["cell"] >= lo) & (i["cell"] < hi)
        local = O.DenseData(hi - lo, g.m, i["cell"][keep] - lo, i["snv"][keep], i["var"][keep], i["total"][keep])
        sh = ShardedStore(CpuBackend(local), g.n, g.m, lo, hi)
        full = O.DenseData(g.n, g.m, i["cell"], i["snv"], i["var"], i["total"])
        rng = np.random.default_rng(5)
        lab = rng.integers(0, 3, g.n).astype(np.uint8)
        cA, cB = np.nonzero(lab == 1)[0], np.nonzero(lab == 2)[0]
        muts = np.arange(g.m)
        ok = True
        # linear: labels 1 = cellsA only
        labA = (lab == 1).astype(np.uint8)
        out, nobs = sh.assign_linear(labA, len(cA))
        a, b = O.linear_mut_assignment(full, cA, muts, nobs_per_cluster=-1)
        ok &= np.array_equal(muts[out == 2], b) and np.array_equal(muts[out == 1], a)
        ok &= np.array_equal(nobs, np.count_nonzero(full.total[np.ix_(cA, muts)], axis=0))
        # branching, on a subset of SNVs
        sub = np.sort(rng.choice(g.m, g.m // 2, replace=False))
        out3, nobs2 = sh.assign_branching(lab, sub)
        ra, rb, rc = O.branching_mut_assignment(full, cA, cB, sub, nobs_per_cluster=-1)
        cand = np.stack(O.branching_candidates(full, cA, cB, sub))
        ref3 = np.zeros(len(sub), np.uint8); ref3[np.isin(sub, ra)] = 1; ref3[np.isin(sub, rb)] = 2; ref3[np.isin(sub, rc)] = 3
        for d in np.nonzero(ref3 != out3)[0]:   # only exact mathematical ties may differ
            pair = cand[[ref3[d] - 1, out3[d] - 1], d]
            ok &= abs(pair[0] - pair[1]) <= 1e-12 * abs(pair).max()
        S0, S1, No, _ = O.snv_table(full, [cA, cB], sub)
        ok &= np.array_equal(nobs2, No)
        t = sh.snv_table(lab, 2, sub)
        ok &= np.array_equal(t["Nobs"], No) and np.allclose(t["S0"], S0, rtol=1e-12, atol=0) and np.allclose(t["S1"], S1, rtol=1e-12, atol=0)
        # one (var,total) code table for all ranks (the peer exchange sums histograms indexed by code): the merged table of
        # the ranks' blocks must be the table of the whole data set, in the same order, on every rank
        from phertilizer_b200.sharded import global_pairs
        from phertilizer_b200.store import encode_with_table, order_pairs, pair_counts

        gv, gt = global_pairs(i["var"][keep], i["total"][keep])
        ev, et = order_pairs(*pair_counts(i["var"], i["total"]))
        ok &= np.array_equal(gv, ev) and np.array_equal(gt, et)
        code = encode_with_table(i["var"][keep], i["total"][keep], gv, gt)
        ok &= np.array_equal(gv[code], i["var"][keep]) and np.array_equal(gt[code], i["total"][keep])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_assignment():
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


def test_cell_blocks_cover_all_cells():
    from phertilizer_b200.sharded import cell_block, snv_block

    for n in (1, 7, 1427, 100000):
        for w in (1, 2, 3, 8):
            assert [snv_block(n, r, w) for r in range(w)] == [cell_block(n, r, w) for r in range(w)]
            blocks = [cell_block(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(w - 1))
