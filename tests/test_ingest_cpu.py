"""Host-side ingest (phertilizer_b200/phertilizer.py index_observations): labels as the reference builds them
(phertilizer.py:235-253) and its notion of an observation (total > 0: the dense matrices are unstack(fill_value=0))."""
import numpy as np
import pandas as pd
import pytest

from phertilizer_b200.phertilizer import index_observations


def frame(rows):
    return pd.DataFrame(rows, columns=["chr", "mutation_id", "cell_id", "base", "var", "total"])


def test_labels_sorted_like_the_reference():
    vc = frame([("chr2", 7, 30, "A", 1, 1), ("chr10", 5, 10, "A", 0, 2), ("chr2", 11, 20, "C", 0, 1), ("chr10", 5, 30, "A", 1, 1)])
    cells, muts, ci, mi, var, total = index_observations(vc)
    assert list(cells) == [10, 20, 30]
    assert list(muts) == sorted(["chr2_7", "chr10_5", "chr2_11"])          # string order, as np.sort(unique) gives
    assert [cells[c] for c in ci] == [30, 10, 20, 30]
    assert [muts[j] for j in mi] == ["chr2_7", "chr10_5", "chr2_11", "chr10_5"]
    assert list(var) == [1, 0, 0, 1] and list(total) == [1, 2, 1, 1]


def test_zero_total_rows_are_not_observations_but_keep_their_labels():
    """A row with total == 0 is a zero of the reference's dense matrices, i.e. unobserved (count_nonzero(total)); the cell
    and SNV it names still exist (labels come from all rows)."""
    rows = [("c", 1, 5, "A", 1, 2), ("c", 2, 6, "A", 0, 0), ("c", 3, 7, "A", 0, 1), ("c", 1, 7, "A", 0, 0)]
    cells, muts, ci, mi, var, total = index_observations(frame(rows))
    assert list(cells) == [5, 6, 7] and list(muts) == ["c_1", "c_2", "c_3"]
    assert len(ci) == 2 and list(total) == [2, 1] and list(var) == [1, 0]
    assert [cells[c] for c in ci] == [5, 7] and [muts[j] for j in mi] == ["c_1", "c_3"]
    # same store input as the frame without those rows, on the full label sets
    _, _, ci2, mi2, var2, total2 = index_observations(frame([rows[0], rows[2]]))
    assert list(var2) == list(var) and list(total2) == list(total)


def test_var_without_total_is_rejected():
    with pytest.raises(ValueError):
        index_observations(frame([("c", 1, 5, "A", 1, 0)]))
