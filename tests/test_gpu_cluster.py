"""GPU cell clustering (csrc/ph_affinity.cu + phertilizer_b200/cluster.py) against the reference's host
arithmetic: scipy pdist / numpy exp + quantile for the affinity, scipy csgraph_laplacian for the Laplacian, sklearn
spectral_clustering for the cut (tests/oracle_store.py::HostCopyDistance restates linear_split.py:308-335,
branching_split.py:215-232, utils.py:131-150).  Matrices: 1e-12 relative (CUDA's exp() vs glibc's differ in the last ulp);
quantile, kernel widths and cut labels: exact."""
import numpy as np
import pytest

from tests.golden_util import load_golden, rel_close
from tests.oracle_store import HostCopyDistance

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["example", "synth_t1", "synth_t1_recompute"])
def env(request):
    """`_recompute`: the n x n distance matrix is NOT kept in HBM (PH_AFFINITY_STORE=0, what happens automatically for
    ~100k cells with a low-dimensional embedding); its entries are re-evaluated inside the affinity kernels."""
    import os

    from phertilizer_b200.cluster import CopyDistance

    stem = request.param.replace("_recompute", "")
    g = load_golden(stem)
    X = g.inputs["read_depth"].astype(np.float64)
    saved = os.environ.get("PH_AFFINITY_STORE")
    if request.param.endswith("_recompute"):
        os.environ["PH_AFFINITY_STORE"] = "0"
    try:
        gpu = CopyDistance(X)
    finally:
        if saved is None:
            os.environ.pop("PH_AFFINITY_STORE", None)
        else:
            os.environ["PH_AFFINITY_STORE"] = saved
    host = HostCopyDistance(X)
    yield g, gpu, host
    gpu.close()


def _features(rng, k, F):
    f = rng.integers(0, 6, size=(k, F)) / rng.integers(6, 12, size=(k, F))   # ratios of small counts, many ties
    return f.astype(np.float64)


def test_copy_distance_rows_and_full(env):
    g, gpu, host = env
    full = gpu.full()
    assert full.shape == host.full().shape and np.array_equal(full, full.T) and np.all(np.diag(full) == 0)
    assert rel_close(full, host.full(), 1e-9)
    idx = np.array([5, 0, g.n - 1, 17])
    assert np.array_equal(gpu.rows(idx), full[idx])


@pytest.mark.parametrize("F,strict,use_copy,radius", [(1, False, True, 1.0), (2, True, True, 1.0), (1, False, False, 1.0),
                                                      (2, True, True, 0.8), (1, False, True, 0.37)])
def test_laplacian_and_stats(env, F, strict, use_copy, radius):
    from scipy.sparse.csgraph import laplacian as csgraph_laplacian

    g, gpu, host = env
    rng = np.random.default_rng(F * 7 + int(strict))
    cells = np.sort(rng.choice(g.n, size=min(g.n, 700), replace=False))
    feats = _features(rng, len(cells), F)
    stats, dd = gpu.build(cells, feats, use_copy, radius, strict)
    assert not stats["has_nan"]
    # the scalars of the affinity are bit-exact: kernel widths (utils.py:75-84) and np.quantile(c_mat, radius)
    c_mat = gpu.full()[np.ix_(cells, cells)]
    from scipy.spatial.distance import pdist, squareform

    d_mat = squareform(pdist(feats))
    assert stats["kw"] == 0.2 * (np.max(d_mat) - np.min(d_mat))
    if use_copy:
        assert stats["kwc"] == 0.2 * (np.max(c_mat) - np.min(c_mat))
        assert stats["r"] == np.quantile(c_mat, radius)
    # Laplacian and degree vector vs scipy on the host affinity built from the SAME distances
    host._full = gpu.full()
    W = host.affinity(cells, feats, use_copy, radius, strict)
    L_ref, dd_ref = csgraph_laplacian(W, normed=True, return_diag=True)
    np.fill_diagonal(L_ref, 1)
    L = gpu.laplacian()
    assert rel_close(dd, dd_ref, 1e-12)
    assert np.allclose(L, L_ref, rtol=1e-12, atol=1e-300)
    assert np.array_equal(L == 0, L_ref == 0)            # the radius mask cuts exactly the same entries
    X = rng.standard_normal((len(cells), 3))
    assert np.allclose(gpu.matmat(X), L_ref @ X, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("F,strict,k", [(1, False, 0), (2, True, 0), (1, False, 300), (2, True, 40), (1, False, 12), (2, True, 7),
                                        (1, False, 2)])
def test_min_cut_labels_equal_sklearn(env, F, strict, k):
    g, gpu, host = env
    rng = np.random.default_rng(100 + F + k)
    cells = np.arange(g.n) if k == 0 else np.sort(rng.choice(g.n, size=k, replace=False))
    # two groups with different burdens so that the cut is meaningful
    feats = _features(rng, len(cells), F)
    feats[: len(cells) // 2] += 0.5
    host._full = gpu.full()
    for use_copy in (True, False):
        lab = gpu.min_cut(cells, feats, use_copy, 1.0, strict, np.random.RandomState(5))
        ref = host.min_cut(cells, feats, use_copy, 1.0, strict, np.random.RandomState(5))
        assert lab is not None and ref is not None
        assert np.array_equal(lab, ref), (int((lab != ref).sum()), len(lab))


def test_nan_affinity_terminates(env):
    g, gpu, host = env
    cells = np.arange(50)
    feats = np.full((50, 1), 0.25)            # all features equal -> kernel width 0 -> 0/0
    assert gpu.min_cut(cells, feats, True, 1.0, False, np.random.RandomState(1)) is None
    feats = np.linspace(0, 1, 50).reshape(-1, 1)
    feats[3] = np.nan
    assert gpu.min_cut(cells, feats, False, 1.0, False, np.random.RandomState(1)) is None


def test_nearest_donor_matches_reference_argsort(env):
    """impute_mut_features (utils.py:87-129): masked arg-min on the GPU == position 0 of the reference's argsort whenever
    the minimum is unique; the tie count flags the rows the host resolves."""
    g, gpu, host = env
    rng = np.random.default_rng(3)
    full = gpu.full()
    donor = rng.random(g.n) < 0.6
    rows = np.nonzero(~donor)[0][:200]
    idx, ties = gpu.nearest(rows, donor)
    for c, j, t in zip(rows, idx, ties):
        d = np.where(donor, full[c], np.nan)
        d[c] = np.nan
        dmin = np.nanmin(d)
        assert t == int(np.sum(d == dmin))
        assert d[j] == dmin and j == int(np.nonzero(d == dmin)[0][0])
        if t == 1:
            assert j == int(np.argsort(d)[0])
    # and through the host helper: identical features to the all-host path
    from phertilizer_b200.utils import impute_mut_features

    cells = np.sort(rng.choice(g.n, size=min(g.n, 500), replace=False))
    feats = rng.random((len(cells), 2))
    feats[rng.random(len(cells)) < 0.3] = np.nan
    a = impute_mut_features(gpu, feats.copy(), cells)
    host._full = full
    b = impute_mut_features(host, feats.copy(), cells)
    assert np.array_equal(a, b)
