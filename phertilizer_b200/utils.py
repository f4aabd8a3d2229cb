"""Host helpers of the tree operations.  Semantics follow phertilizer/utils.py of the reference (cited per function);
the observation counts come from the device store instead of dense matrices."""
from __future__ import annotations

import pickle

import numpy as np
from pandas import DataFrame, Series, concat
from scipy.special import comb


def snv_kernel_width(dmat):
    """utils.py:75-78."""
    return 0.2 * (np.max(dmat) - np.min(dmat))


def cnv_kernel_width(dmat):
    """utils.py:80-84."""
    return 0.2 * (np.max(dmat) - np.min(dmat))


def impute_mut_features(c_mat, mut_features, cells):
    """Fill NaN feature rows with the features of the closest (copy distance) cell of `cells` that has features
    (utils.py:87-129).  The reference argsorts a masked copy of the whole n x n matrix and reads column 0 of the
    rows it needs; only those rows are fetched (`c_mat.rows`, from the GPU-resident distance matrix) and sorted here.
    The row content (NaN on the diagonal, on cells outside `cells` and on cells that themselves need imputation) and
    the sort call are the reference's, so exact distance ties -- which do occur with integer bin counts -- resolve to
    the same neighbour.
    """
    cells = np.asarray(cells)
    na = np.isnan(mut_features.sum(axis=1).reshape(-1))
    if na.sum() > 0.5 * len(cells):
        return mut_features  # not enough cells to impute from (utils.py:96-98)
    if not na.any():
        return mut_features
    donor_mask = np.zeros(c_mat.shape[0], dtype=bool)
    donor_mask[cells[~na]] = True
    row_of = {int(c): i for i, c in enumerate(cells)}
    need = np.nonzero(na)[0]
    if hasattr(c_mat, "nearest"):
        # masked arg-min on the GPU; only rows whose minimum is shared by several donors (exact distance ties) go through
        # the reference's argsort below, which decides such ties
        near, ties = c_mat.nearest(cells[need], donor_mask)
        uniq = ties == 1
        src = np.fromiter((row_of[int(j)] for j in near[uniq]), dtype=np.int64, count=int(uniq.sum()))
        mut_features[need[uniq], :] = mut_features[src, :]
        need = need[~uniq]
    for lo in range(0, len(need), 2048):   # bounded slabs of distance rows
        part = need[lo:lo + 2048]
        dist = c_mat.rows(cells[part])
        for row, drow in zip(part, dist):
            c = cells[row]
            d = np.where(donor_mask, drow.astype(float), np.nan)
            d[c] = np.nan
            nearest = int(np.argsort(d)[0])
            mut_features[row, :] = mut_features[row_of[nearest], :]
    assert not np.any(np.isnan(mut_features))
    return mut_features


def split_by_labels(index, labels):
    """utils.py:146-150: the two sides of the cut as index arrays."""
    return index[labels == 0], index[labels == 1]


def success_prob(N, parts, d):
    """utils.py:195-199."""
    denom = comb(N + parts - 1, N)
    free = N - parts * d
    return comb(free + parts - 1, free) / denom


def power_calc(d, gamma, parts, Nmax):
    """Smallest number of observations with success probability >= gamma (utils.py:203-216)."""
    n = 0
    for n in range(1, Nmax + 1):
        if success_prob(n, parts, d) >= gamma:
            break
    return n


def check_obs(cells, muts, store):
    """Number of observed entries in the block cells x muts (utils.py:218-220), from the device store (K3)."""
    if len(cells) == 0 or len(muts) == 0:
        return 0
    _, _, nobs, _ = store.block_sums(store.cell_labels(cells), 1, store.snv_labels(muts), 1)
    return int(nobs[0, 0])


def get_next_label(T):
    """utils.py:222-228."""
    return np.max(np.array(list(T.nodes()))) + 1


def pickle_save(obj, fname):
    with open(fname, "wb") as handle:
        pickle.dump(obj, handle, protocol=pickle.HIGHEST_PROTOCOL)


def pickle_load(fname):
    with open(fname, "rb") as handle:
        return pickle.load(handle)


def mapping_to_dataframe(mapping, id_name, cell=False):
    """cluster mapping -> long dataframe sorted by id (utils.py:24-45)."""
    if len(mapping) == 0:
        return DataFrame(columns=[id_name, "cluster"])
    parts = []
    for k in mapping:
        if cell:
            for s in mapping[k]:
                part = DataFrame(mapping[k][s], columns=[id_name])
                part["cluster"] = k
                part["subcluster"] = s
                parts.append(part)
        else:
            part = DataFrame(mapping[k], columns=[id_name])
            part["cluster"] = k
            parts.append(part)
    return concat(parts).sort_values(by=[id_name])


def generate_cell_dataframe(cell_mapping, cell_lookup):
    """utils.py:309-316."""
    df = mapping_to_dataframe(cell_mapping, "cell_id", cell=True)
    df["cell"] = cell_lookup[df["cell_id"]].values
    return df.drop(["cell_id"], axis=1)


def generate_mut_dataframe(mut_mapping, mut_lookup):
    """utils.py:318-327."""
    df = mapping_to_dataframe(mut_mapping, "mutation_id")
    df["mutation"] = mut_lookup[df["mutation_id"]].values
    return df.drop(["mutation_id"], axis=1)


def dict_to_series(event_mapping):
    series = [Series(np.full(shape=len(event_mapping[s]), fill_value=s), index=event_mapping[s]) for s in event_mapping]
    return concat(series).sort_index()


def dict_to_dataframe(event_mapping):
    if event_mapping is None or len(event_mapping) == 0:
        return DataFrame()
    df = concat([dict_to_series(event_mapping[n]) for n in event_mapping], axis=1)
    df.columns = [f"node_{n}" for n in list(event_mapping.keys())]
    df.index.rename("bin", inplace=True)
    return df


def setdiff1d(ar1, ar2):
    """`np.setdiff1d(ar1, ar2)` -- sorted unique values of ar1 that are not in ar2 -- for the tree's index arrays.

    numpy sorts / hashes both arguments (3 ms per call on 1e5 indices, 9 s of a C3 tree); cell and SNV index sets are
    non-negative integers drawn from a dense range, so a mark array gives the same result in one linear pass.  Anything else
    (floats, negative values, sparse ranges) goes to numpy."""
    a, b = np.asarray(ar1), np.asarray(ar2)
    if a.dtype.kind in "iu" and a.size and (b.size == 0 or b.dtype.kind in "iu"):
        lo = int(a.min()) if b.size == 0 else min(int(a.min()), int(b.min()))
        hi = int(a.max()) if b.size == 0 else max(int(a.max()), int(b.max()))
        if lo >= 0 and hi < 4 * (a.size + b.size) + 4096:
            mark = np.zeros(hi + 1, dtype=bool)
            mark[a] = True
            if b.size:
                mark[b] = False
            return np.flatnonzero(mark).astype(a.dtype, copy=False)
    return np.setdiff1d(ar1, ar2)
