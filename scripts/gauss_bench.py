import sys, time, numpy as np
sys.path.insert(0, ".")
from phertilizer_b200.store import ObservationStore
n, m, K = 50000, 100, 15
rng = np.random.default_rng(0)
cell = rng.integers(0, n, 5000).astype(np.int32); snv = rng.integers(0, m, 5000).astype(np.int32)
key = np.unique(cell.astype(np.int64) * m + snv); cell = (key // m).astype(np.int32); snv = (key % m).astype(np.int32)
st = ObservationStore.from_coo(n, m, cell, snv, np.ones(len(cell), np.int64), np.ones(len(cell), np.int64))
X = rng.standard_normal((n, 2))
lab = np.zeros(n + 16, np.uint8)[:n]; lab[:] = rng.integers(0, K, n)
for mode in ("uncached", "cached"):
    if mode == "cached": st.set_embedding(X)
    for _ in range(3): a = st.gauss_node_ll(X, lab, K)
    t = time.perf_counter()
    for _ in range(50): b = st.gauss_node_ll(X, lab, K)
    print(mode, round((time.perf_counter() - t) / 50 * 1e3, 3), "ms/call", np.array_equal(a, b))
