"""Host helpers that replace numpy set algebra on the tree's index arrays (phertilizer_b200/utils.py)."""
import numpy as np
import pytest

from phertilizer_b200.utils import setdiff1d


@pytest.mark.parametrize("seed", range(6))
def test_setdiff1d_equals_numpy_on_index_sets(seed):
    rng = np.random.default_rng(seed)
    for trial in range(40):
        hi = int(rng.integers(1, 5000))
        a = rng.integers(0, hi, int(rng.integers(0, 3000)))
        b = rng.integers(0, hi, int(rng.integers(0, 3000)))
        if trial % 3 == 0:   # the usual case: sorted unique subsets, b inside a
            a = np.unique(a)
            b = rng.choice(a, len(a) // 2, replace=False) if len(a) else a
        for dt in (np.int64, np.int32):
            got, ref = setdiff1d(a.astype(dt), b.astype(dt)), np.setdiff1d(a.astype(dt), b.astype(dt))
            assert got.dtype == ref.dtype and np.array_equal(got, ref)


def test_setdiff1d_falls_back():
    """Floats, negative values, sparse ranges and the reference's float-typed empty arrays give numpy's answer."""
    cases = [(np.array([3, 1, 2]), np.array([])),                       # `np.array([])` is float64
             (np.array([], dtype=np.int64), np.array([1, 2])),
             (np.array([-5, 3, 7]), np.array([3])),
             (np.array([10 ** 12, 5]), np.array([5])),
             (np.array([0.5, 1.5]), np.array([1.5])),
             (np.array([[1, 2], [3, 4]]), np.array([4]))]
    for a, b in cases:
        got, ref = setdiff1d(a, b), np.setdiff1d(a, b)
        assert got.dtype == ref.dtype and np.array_equal(got, ref)
