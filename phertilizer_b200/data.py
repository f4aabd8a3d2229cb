"""`Data`: what the tree operations read.  The reference's `Data` (phertilizer/data.py:8-20) carries four dense
n x m matrices; here the observations live on the GPU inside an `ObservationStore` and only the lookups, the embedding
and the n x n copy distance (computed by the K5 kernel) stay on the host."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import pandas as pd

from .store import ObservationStore


@dataclass
class Data:
    cell_lookup: pd.Series
    mut_lookup: pd.Series
    store: ObservationStore
    read_depth: np.ndarray
    copy_distance: np.ndarray = None
    use_read_depth: bool = True
    like1_dict: dict = None   # kept for attribute compatibility (always None, as in the live reference path)
    cna_hmm: object = None    # idem ("depreciated" in the reference)

    @property
    def n_cells(self) -> int:
        return self.store.n_cells

    @property
    def n_snvs(self) -> int:
        return self.store.n_snvs
