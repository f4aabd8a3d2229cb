"""Per-observation binomial read-count log-likelihoods as a lookup table over distinct (var, total) pairs.

Replaces the reference's one-off numba pass over every observation (phertilizer.py:29-40 ``binom_pdf``,
52-67 ``likelihood1_numba``, 70-100 ``apply_like1_numba``, 102-119 ``apply_like0_numba``, 121-151
``compute_like0/1``): ``like0`` and ``like1`` depend only on ``(var, total)``, and ultra-low-coverage data has a
handful of distinct pairs (7 in the bundled example), so the device stores a small integer code per
observation and the fp64 table lives in shared memory.

Arithmetic parity:
  * ``coeff = scipy.special.comb(total, var)`` in float64 (phertilizer.py:295) -- scipy is called here as well,
    it is the reference's own third-party arithmetic;
  * numba lowers ``float64 ** int64`` to an LSB-first square-and-multiply loop, NOT libm ``pow``; the loop is
    restated in ``_ipow`` so the table is bit-identical to the jitted reference (verified in
    tests/test_like_lut.py against the reference's own functions);
  * evaluation order ``(coeff * p**k) * (1-p)**(n-k)``; the VAF mixture is summed in ascending-VAF order
    starting from 0.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.special


def _ipow(a: float, b: int) -> float:
    """float ** non-negative int exactly as numba evaluates it (square-and-multiply, LSB first)."""
    if b < 0:
        return 1.0 / _ipow(a, -b)
    if b > 0x10000:
        return math.pow(a, float(b))
    r = 1.0
    while b != 0:
        if b & 1:
            r *= a
        b >>= 1
        a *= a
    return r


def _binom_pdf(k: int, n: int, p: float, coeff: float) -> float:
    # phertilizer.py:39  prob = coeff * (p**k) * (1-p)**(n-k)
    return coeff * _ipow(p, k) * _ipow(1 - p, n - k)


def vaf_mixture(min_copies: int, max_copies: int, alpha: float):
    """phertilizer.py:83-93: unique VAFs i/c, their multiplicity weights, and error-adjusted VAFs."""
    vafs_all = np.array([i / c for c in np.arange(min_copies, max_copies + 1, 1) for i in np.arange(1, c + 1, 1)])
    vafs = np.unique(vafs_all)
    probs = np.zeros_like(vafs)
    prime = np.zeros_like(vafs)
    for i in range(len(vafs)):
        probs[i] = np.count_nonzero(vafs_all == vafs[i]) / len(vafs_all)
        prime[i] = vafs[i] * (1 - alpha) + (1 - vafs[i]) * (alpha / 3)
    return vafs, probs, prime


def like_tables(var: np.ndarray, total: np.ndarray, alpha: float = 0.001, max_copies: int = 5, min_copies: int = 1):
    """like0, like1 (float64 log-likelihoods) for each (var[i], total[i]) pair."""
    var = np.asarray(var, dtype=np.int64)
    total = np.asarray(total, dtype=np.int64)
    coeff = scipy.special.comb(total, var)
    _, probs, prime = vaf_mixture(min_copies, max_copies, alpha)
    l0 = np.empty(len(var), dtype=np.float64)
    l1 = np.empty(len(var), dtype=np.float64)
    for i in range(len(var)):
        k, n, c = int(var[i]), int(total[i]), float(coeff[i])
        l0[i] = _binom_pdf(k, n, alpha, c)
        prob = 0.0
        for v, w in zip(prime, probs):
            prob += _binom_pdf(k, n, float(v), c) * float(w)
        l1[i] = prob
    with np.errstate(divide="ignore"):
        return np.log(l0), np.log(l1)


def encode_pairs(var: np.ndarray, total: np.ndarray):
    """Distinct (var, total) pairs ordered by descending frequency (ties: ascending total, var).

    Returns (pair_var[T], pair_total[T], counts[T], code[nnz] int32).  Code 0 is the most frequent pair.
    """
    var = np.asarray(var, dtype=np.int64)
    total = np.asarray(total, dtype=np.int64)
    width = int(var.max()) + 1 if len(var) else 1
    key = total * width + var
    kmax = int(key.max()) + 1 if len(key) else 1
    if kmax <= (1 << 24):
        # read counts are small: a counting pass instead of a sort of all observations
        hist = np.bincount(key, minlength=kmax)
        uniq = np.nonzero(hist)[0]
        cnt = hist[uniq]
        order = np.lexsort((uniq, -cnt))
        lut = np.zeros(kmax, dtype=np.int32)
        lut[uniq[order]] = np.arange(len(order), dtype=np.int32)
        u = uniq[order]
        return (u % width).astype(np.int64), (u // width).astype(np.int64), cnt[order].astype(np.int64), lut[key]
    uniq, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
    order = np.lexsort((uniq, -cnt))
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    u = uniq[order]
    return (u % width).astype(np.int64), (u // width).astype(np.int64), cnt[order].astype(np.int64), rank[inv].astype(np.int32)
