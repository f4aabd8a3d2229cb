"""The sparse C restatement (oracle/oracle_sparse.c, bench.py's CPU baseline) agrees with the dense numpy oracle,
which is pinned against the reference's recorded outputs."""
import numpy as np

from oracle import oracle_np as O
from oracle.oracle_sparse import SparseOracle
from phertilizer_b200.like_lut import encode_pairs, like_tables
from tests.golden_util import load_golden, rel_close


def _mk(stem):
    g = load_golden(stem)
    i = g.inputs
    pv, pt, _, code = encode_pairs(i["var"], i["total"])
    L0, L1 = like_tables(pv, pt)
    S = SparseOracle(g.n, g.m, i["cell"], i["snv"], code, L0, L1, pv > 0)
    D = O.DenseData(g.n, g.m, i["cell"], i["snv"], i["var"], i["total"])
    return g, S, D


def test_sparse_oracle_matches_dense_oracle_and_reference():
    for stem in ("example", "synth_t0"):
        g, S, D = _mk(stem)
        rng = np.random.default_rng(1)
        lab = rng.integers(0, 3, g.n).astype(np.uint8)
        groups = [np.nonzero(lab == c)[0] for c in (1, 2)]
        S0, S1, No, Nv = S.snv_table(lab, 2)
        d0, d1, dn, dv = O.snv_table(D, groups, np.arange(g.m))
        assert np.array_equal(No, dn) and np.array_equal(Nv, dv)
        assert rel_close(S0, d0, 1e-12) and rel_close(S1, d1, 1e-12)
        sl = rng.integers(0, 3, g.m).astype(np.uint8)
        A0, A1, No2, Nv2 = S.cell_table(sl, 2)
        e0, e1, en, ev = O.cell_table(D, np.arange(g.n), [np.nonzero(sl == q)[0] for q in (1, 2)])
        assert np.array_equal(No2, en) and np.array_equal(Nv2, ev)
        assert rel_close(A0, e0, 1e-12) and rel_close(A1, e1, 1e-12)
        # recorded reference assignments (all SNVs of the seed == all SNVs only for the first records; use subsets)
        for r in g.of("linear_mut_assignment"):
            if r.mutsB is None:
                continue
            cl = np.zeros(g.n, np.uint8); cl[r.cellsA] = 1
            out, _ = S.assign_linear(cl, len(r.cellsA))
            muts = r.muts.astype(np.int64)
            assert np.array_equal(muts[out[muts] == 2], r.mutsB)
        for r in g.of("branching_mut_assignment"):
            if len(r.mutsA) == 0 and len(r.mutsB) == 0:
                continue
            cl = np.zeros(g.n, np.uint8); cl[r.cellsA] = 1; cl[r.cellsB] = 2
            out, _ = S.assign_branching(cl)
            muts = r.muts.astype(np.int64)
            assert np.array_equal(muts[out[muts] == 1], r.mutsA) and np.array_equal(muts[out[muts] == 3], r.mutsC)
        for r in g.of("variant_likelihood_by_node"):
            cn = np.full(g.n, 255, np.uint8); sn = np.full(g.m, 255, np.uint8)
            cn[r.cells] = 0; sn[r.present_muts] = 0
            pc = S.tree_loglik(cn, sn, np.ones((1, 1), np.uint8))
            assert rel_close(pc.sum(), r.value, 1e-9)
