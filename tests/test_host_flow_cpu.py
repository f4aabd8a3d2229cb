"""Host logic without a GPU: phertilizer_b200's control flow (planting, splits, growing, scoring, post-processing,
outputs) driven through a CPU test double of the store (tests/oracle_store.py) must reproduce the reference's
recorded run on the same seed.  This isolates the host mirror from the CUDA kernels (which have their own -m gpu tests)."""
import io
from contextlib import redirect_stdout

import numpy as np
import pytest

from tests.golden_util import load_golden, rel_close, tree_from_record
from tests.oracle_store import patch_phertilizer
from tests.test_gpu_end_to_end import frames_from_golden


@pytest.mark.parametrize("stem,starts,iters,seed", [("synth_t0", 2, 6, 4242), ("example", 3, 10, 100 * 99059), ("synth_t1", 2, 5, 777)])
def test_host_flow_matches_reference(monkeypatch, stem, starts, iters, seed):
    patch_phertilizer(monkeypatch)
    from phertilizer_b200.phertilizer import Phertilizer

    g = load_golden(stem)
    vc, bc = frames_from_golden(g)
    ph = Phertilizer(vc, bc, alpha=0.001, max_copies=5, mode="rd", dim_reduce=False)
    with redirect_stdout(io.StringIO()):
        tree, tree_list, loglikes = ph.phertilize(max_iterations=iters, starts=starts, seed=seed, radius=1, gamma=0.95, d=10,
                                                  use_copy_kernel=True, post_process=True, low_cmb=0.05, high_cmb=0.15,
                                                  nobs_per_cluster=3)
    fin = g.of("final")[0]
    assert ph.params.minobs == fin.minobs and tree_list.size() == fin.n_trees
    assert rel_close(loglikes.to_numpy(), fin.loglikes, 1e-9), (loglikes.to_numpy(), fin.loglikes)
    nodes, edges, cells, muts = tree_from_record(fin)
    assert sorted(tree.tree.edges()) == sorted(edges)
    for k in nodes:
        assert np.array_equal(tree.get_tip_cells(k), cells[k])
        assert np.array_equal(np.asarray(tree.get_tip_muts(k)).astype(np.int64), muts[k])
    assert rel_close(tree.norm_loglikelihood, fin.norm, 1e-9)
