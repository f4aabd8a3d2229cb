"""Seeded synthetic ultra-low-coverage scDNA-seq data.

The reference ships no generator; the shapes are those named in BASELINE.json ``configs``.  A random
K-clone tree is drawn, cells are attached to clones, SNVs to tree edges, and read counts are sampled
directly in COO form (the dense ``n x m`` matrix of ``phertilizer.py:153-161`` is never built):

  total_ij ~ Poisson(cov) conditioned on total > 0        (density 1 - exp(-cov))
  var_ij   ~ Binomial(total_ij, 0.5(1-a) + 0.5 a/3)  if SNV j is on the root->clone(i) path else Binomial(total_ij, a)

Everything is torch so the same code runs on the CPU (tests) and on the B200 (bench, 4e8 observations).
torch is plumbing here (device memory + RNG), not the product path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

CONFIGS = {
    # name: n cells, m SNVs, coverage, clones, embedding dim, seed (1026 + config index)
    "C2": dict(n=10_000, m=50_000, cov=0.05, K=5, D=2, seed=1028),
    "C3": dict(n=100_000, m=200_000, cov=0.02, K=12, D=2, seed=1029),
    "C4": dict(n=50_000, m=100_000, cov=0.02, K=12, D=2, seed=1030),
    "C5": dict(n=250_000, m=500_000, cov=0.02, K=12, D=5000, seed=1031),
    # small shapes the dense oracle can hold (tests)
    "T0": dict(n=300, m=420, cov=0.08, K=3, D=2, seed=7),
    "T1": dict(n=2_000, m=3_000, cov=0.05, K=4, D=2, seed=11),
}


@dataclass
class Synthetic:
    n: int
    m: int
    cell: torch.Tensor  # int32 [nnz]  row (cell) index, sorted by (cell, snv)
    snv: torch.Tensor  # int32 [nnz]
    var: torch.Tensor  # int32 [nnz]
    total: torch.Tensor  # int32 [nnz]
    parent: np.ndarray  # [K] parent clone of each clone, -1 for the root
    cell_clone: torch.Tensor  # uint8 [n]
    snv_node: torch.Tensor  # uint8 [m]
    embedding: torch.Tensor | None  # float64 [n, D]

    @property
    def nnz(self) -> int:
        return int(self.cell.numel())

    def ancestor_matrix(self) -> np.ndarray:
        """anc[k, q] = 1 iff node q is an ancestor-or-self of clone k."""
        K = len(self.parent)
        anc = np.zeros((K, K), dtype=np.uint8)
        for k in range(K):
            q = k
            while q >= 0:
                anc[k, q] = 1
                q = int(self.parent[q])
        return anc


def _partition_sizes(total: int, K: int, rng: np.random.Generator, floor_frac: float = 0.02) -> np.ndarray:
    w = rng.dirichlet(np.full(K, 2.0))
    w = np.maximum(w, floor_frac)
    w /= w.sum()
    sizes = np.floor(w * total).astype(np.int64)
    sizes[np.argmax(sizes)] += total - sizes.sum()
    return sizes


def _sample_observations(n, m, cov, cell_clone, snv_node, anc, alpha, dev, g, chunk):
    """COO observations of n cells x m SNVs (cells = rows of `cell_clone`): positions, totals, variant counts."""
    # positions: Bernoulli(p) process over the flattened n*m grid, via geometric gaps
    p = 1.0 - math.exp(-cov)
    grid = n * m
    pos_chunks = []
    start = -1
    while True:
        gaps = torch.empty(chunk, dtype=torch.float64, device=dev).geometric_(p, generator=g)
        pos = torch.cumsum(gaps, 0).to(torch.int64) + start
        if int(pos[-1]) >= grid:
            pos = pos[pos < grid]
            pos_chunks.append(pos)
            break
        pos_chunks.append(pos)
        start = int(pos[-1])
    pos = torch.cat(pos_chunks) if len(pos_chunks) > 1 else pos_chunks[0]
    del pos_chunks
    cell = torch.div(pos, m, rounding_mode="floor").to(torch.int32)
    snv = (pos - cell.to(torch.int64) * m).to(torch.int32)
    del pos
    nnz = cell.numel()

    # zero-truncated Poisson totals by inverse CDF
    kmax = 24
    pk = np.array([math.exp(-cov) * cov ** k / math.factorial(k) for k in range(1, kmax + 1)])
    cdf = torch.from_numpy(np.cumsum(pk / pk.sum())).to(dev)
    u = torch.rand(nnz, dtype=torch.float64, device=dev, generator=g)
    total = (torch.bucketize(u, cdf, right=True) + 1).clamp_(max=kmax).to(torch.int32)
    del u

    present = anc[cell_clone[cell.long()].long(), snv_node[snv.long()].long()]
    p1 = 0.5 * (1 - alpha) + 0.5 * alpha / 3
    pv = torch.where(present, torch.tensor(p1, device=dev, dtype=torch.float32),
                     torch.tensor(alpha, device=dev, dtype=torch.float32))
    del present
    var = torch.zeros(nnz, dtype=torch.int32, device=dev)
    for k in range(int(total.max())):
        hit = (torch.rand(nnz, device=dev, generator=g) < pv) & (total > k)
        var += hit.to(torch.int32)
    del pv
    return cell, snv, var, total


def make_synthetic(n: int, m: int, cov: float, K: int, seed: int, D: int = 2, alpha: float = 0.001,
                   device: str | torch.device = "cpu", with_embedding: bool = True,
                   chunk: int = 1 << 26) -> Synthetic:
    rng = np.random.default_rng(seed)
    parent = np.full(K, -1, dtype=np.int64)
    for k in range(1, K):
        parent[k] = rng.integers(0, k)
    cell_sizes = _partition_sizes(n, K, rng)
    snv_sizes = _partition_sizes(m, K, rng)
    cell_clone_np = rng.permutation(np.repeat(np.arange(K, dtype=np.uint8), cell_sizes))
    snv_node_np = rng.permutation(np.repeat(np.arange(K, dtype=np.uint8), snv_sizes))

    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    cell_clone = torch.from_numpy(cell_clone_np).to(dev)
    snv_node = torch.from_numpy(snv_node_np).to(dev)

    anc = torch.zeros((K, K), dtype=torch.bool)
    for k in range(K):
        q = k
        while q >= 0:
            anc[k, q] = True
            q = int(parent[q])
    anc = anc.to(dev)
    cell, snv, var, total = _sample_observations(n, m, cov, cell_clone, snv_node, anc, alpha, dev, g, chunk)

    emb = None
    if with_embedding:
        if D <= 8:
            ang = 2 * math.pi * np.arange(K) / K
            centres = np.zeros((K, D))
            centres[:, 0] = 3 * np.cos(ang)
            if D > 1:
                centres[:, 1] = 3 * np.sin(ang)
            centres_t = torch.from_numpy(centres).to(dev)
            emb = centres_t[cell_clone.long()] + 0.3 * torch.randn(n, D, dtype=torch.float64, device=dev, generator=g)
        else:
            nseg = (D + 49) // 50
            cna = rng.integers(-1, 2, size=(K, nseg))
            cna[0] = 0
            prof = 333.0 * (1 + 0.5 * np.repeat(cna, 50, axis=1)[:, :D])
            prof_t = torch.from_numpy(prof).to(dev)
            emb = torch.poisson(prof_t[cell_clone.long()], generator=g)
    return Synthetic(n, m, cell, snv, var, total, parent, cell_clone, snv_node, emb)


def make_blocked(name: str, device="cuda", blocks: int = 8, with_embedding: bool = False, **over):
    """A config generated in `blocks` independent blocks of cells that share one clone tree (config 5: 2.48 G
    observations -- the one-shot generator above would need > 100 GB of scratch).  Same model, another random stream, so
    NOT the data set `make_config` yields.  Returns a Synthetic whose var / total are uint8 (totals are capped at 24)."""
    cfg = dict(CONFIGS[name])
    cfg.update(over)
    n, m, cov, K, seed, D = cfg["n"], cfg["m"], cfg["cov"], cfg["K"], cfg["seed"], cfg["D"]
    alpha = 0.001
    rng = np.random.default_rng(seed)
    parent = np.full(K, -1, dtype=np.int64)
    for k in range(1, K):
        parent[k] = rng.integers(0, k)
    cell_sizes = _partition_sizes(n, K, rng)
    snv_sizes = _partition_sizes(m, K, rng)
    cell_clone_np = rng.permutation(np.repeat(np.arange(K, dtype=np.uint8), cell_sizes))
    snv_node_np = rng.permutation(np.repeat(np.arange(K, dtype=np.uint8), snv_sizes))
    dev = torch.device(device)
    cell_clone = torch.from_numpy(cell_clone_np).to(dev)
    snv_node = torch.from_numpy(snv_node_np).to(dev)
    anc = torch.zeros((K, K), dtype=torch.bool)
    for k in range(K):
        q = k
        while q >= 0:
            anc[k, q] = True
            q = int(parent[q])
    anc = anc.to(dev)
    parts = []
    for b in range(blocks):
        lo, hi = (n * b) // blocks, (n * (b + 1)) // blocks
        g = torch.Generator(device=dev)
        g.manual_seed(seed * 1000 + b)
        c, s_, v, t = _sample_observations(hi - lo, m, cov, cell_clone[lo:hi], snv_node, anc, alpha, dev, g, 1 << 26)
        parts.append((c + lo, s_, v.to(torch.uint8), t.to(torch.uint8)))
        del c, s_, v, t
    if parts:
        cell, snv, var, total = (torch.cat([p[i] for p in parts]) for i in range(4))
    else:   # blocks = 0: the clone structure / embedding only
        cell = snv = torch.zeros(0, dtype=torch.int32, device=dev)
        var = total = torch.zeros(0, dtype=torch.uint8, device=dev)
    del parts
    emb = None
    if with_embedding:
        g = torch.Generator(device=dev)
        g.manual_seed(seed * 1000 + 999)
        if D <= 8:
            ang = 2 * math.pi * np.arange(K) / K
            centres = np.zeros((K, D))
            centres[:, 0] = 3 * np.cos(ang)
            if D > 1:
                centres[:, 1] = 3 * np.sin(ang)
            emb = torch.from_numpy(centres).to(dev)[cell_clone.long()] + 0.3 * torch.randn(n, D, dtype=torch.float64, device=dev, generator=g)
        else:
            nseg = (D + 49) // 50
            cna = rng.integers(-1, 2, size=(K, nseg))
            cna[0] = 0
            prof = torch.from_numpy(333.0 * (1 + 0.5 * np.repeat(cna, 50, axis=1)[:, :D])).to(dev)
            emb = torch.empty(n, D, dtype=torch.float64, device=dev)
            for lo in range(0, n, 16384):   # row blocks: torch.poisson needs the rate tensor materialised
                hi = min(n, lo + 16384)
                emb[lo:hi] = torch.poisson(prof[cell_clone[lo:hi].long()], generator=g)
    return Synthetic(n, m, cell, snv, var, total, parent, cell_clone, snv_node, emb)


def make_config(name: str, device="cpu", **over) -> Synthetic:
    cfg = dict(CONFIGS[name])
    cfg.update(over)
    return make_synthetic(device=device, **cfg)


def to_dataframes(s: Synthetic):
    """The reference's input frames (run.py:16-27): variant counts [chr, mutation_id, cell_id, base,
    var, total] and binned read counts [cell, bins...].  Only for sizes pandas can hold."""
    import pandas as pd

    vc = pd.DataFrame({
        "chr": "chr1",
        "mutation_id": s.snv.cpu().numpy().astype(np.int64),
        "cell_id": s.cell.cpu().numpy().astype(np.int64),
        "base": "A",
        "var": s.var.cpu().numpy().astype(np.int64),
        "total": s.total.cpu().numpy().astype(np.int64),
    })
    emb = s.embedding.cpu().numpy()
    bc = pd.DataFrame(emb, columns=[f"V{i + 1}" for i in range(emb.shape[1])])
    bc.insert(0, "cell", np.arange(s.n))
    return vc, bc
