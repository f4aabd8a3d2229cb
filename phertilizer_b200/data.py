"""`Data`: what the tree operations read.  The reference's `Data` (phertilizer/data.py:8-20) carries four dense
n x m matrices and the n x n copy distance; here the observations live on the GPU inside an `ObservationStore`, the copy
distance inside a `cluster.CopyDistance` (K5 kernel; `.full()` gives the host matrix), and only the lookups and the
embedding stay on the host."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import pandas as pd

from .store import ObservationStore


@dataclass
class Data:
    cell_lookup: pd.Series
    mut_lookup: pd.Series
    store: ObservationStore   # or sharded_store.CellShardedStore (same methods, cells sharded across GPUs)
    read_depth: np.ndarray
    copy_distance: object = None   # cluster.CopyDistance
    use_read_depth: bool = True
    like1_dict: dict = None   # kept for attribute compatibility (always None, as in the live reference path)
    cna_hmm: object = None    # idem ("depreciated" in the reference)

    @property
    def n_cells(self) -> int:
        return self.store.n_cells

    @property
    def n_snvs(self) -> int:
        return self.store.n_snvs
