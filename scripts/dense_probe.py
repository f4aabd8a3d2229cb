"""Device timing + roofline numbers of the dense embedding kernels (K4 Gaussian node log-likelihood, K5 pairwise distance,
K6 affinity build / Laplacian product) at the shapes the tree operations use them (development aid; run under ncu for the
profiles).   python scripts/dense_probe.py [n] [D]"""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from phertilizer_b200 import _lib
from phertilizer_b200.cluster import CopyDistance
from phertilizer_b200.store import ObservationStore, _ptr

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
peak = 6544.3
try:
    peak = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
except Exception:
    pass
g = torch.Generator(device=dev); g.manual_seed(3)
K = 12
e = torch.randn(n, D, dtype=torch.float64, device=dev, generator=g) + torch.randint(0, K, (n, 1), device=dev, generator=g).double()
cn = torch.randint(0, K, (n + 16,), device=dev, generator=g, dtype=torch.uint8)
dense = ObservationStore.from_coo(n, 1, np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32),
                                  pairs=(np.array([0]), np.array([1])))
per_node = torch.empty(K, dtype=torch.float64, device=dev); per_cell = torch.empty(n, dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream().cuda_stream


def timeit(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {"n": n, "D": D}
ms = timeit(lambda: dense.gauss_node_ll_dev(e, D, cn, K, per_node, per_cell, stream=stream))
out["K4"] = {"ms": ms, "algorithmic_bytes": n * D * 8, "algorithmic_gbs": n * D * 8 / (ms * 1e-3) / 1e9, "frac": n * D * 8 / (ms * 1e-3) / 1e9 / peak}
R = min(n, 4096 if D <= 8 else 512)
dist_out = torch.empty(R * n, dtype=torch.float64, device=dev)
ms = timeit(lambda: _lib.check(_lib.lib().ph_pairwise_dist(_ptr(e), n, D, 0, R, _ptr(dist_out), 0, 1, stream)), reps=5)
out["K5"] = {"rows": R, "ms": ms, "write_gbs": R * n * 8 / (ms * 1e-3) / 1e9, "write_frac": R * n * 8 / (ms * 1e-3) / 1e9 / peak,
             "fp64_flops": 3.0 * R * n * D / (ms * 1e-3)}
if D >= 32:   # the Gram form on the FP64 tensor cores (opt-in path) beside the direct kernel
    sq = torch.empty(n, dtype=torch.float64, device=dev)
    ref = dist_out[: R * n].clone()
    ms = timeit(lambda: _lib.check(_lib.lib().ph_pairwise_dist_gram(_ptr(e), n, D, 0, R, _ptr(dist_out), _ptr(sq), 0, stream)), reps=5)
    got = dist_out[: R * n]
    nz = ref > 0
    out["K5_gram_dmma"] = {"rows": R, "ms": ms, "pair_dims_per_s": R * n * D / (ms * 1e-3), "speedup_vs_direct": out["K5"]["ms"] / ms,
                           "max_rel_diff_vs_direct": float(((got[nz] - ref[nz]).abs() / ref[nz]).max().item())}
dense.close()
if D <= 16:
    k = min(n, 60000)
    cd = CopyDistance(e.cpu().numpy(), device=0)
    cells = np.arange(k, dtype=np.int32)
    feats = np.random.default_rng(0).random((k, 1))
    t0 = time.perf_counter(); cd.build(cells, feats, True, 1.0, False); torch.cuda.synchronize(); tb = time.perf_counter() - t0
    X = np.random.default_rng(1).standard_normal((k, 3))
    cd.matmat(X)
    t0 = time.perf_counter()
    for _ in range(5): cd.matmat(X)
    tm = (time.perf_counter() - t0) / 5
    out["K6"] = {"k": k, "build_s": tb, "matmat_ms": tm * 1e3, "matmat_gbs": k * k * 8 / tm / 1e9, "matmat_frac": k * k * 8 / tm / 1e9 / peak}
    cd.close()
print(json.dumps(out))
