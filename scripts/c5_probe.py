"""Config 5 (BASELINE.json configs[4]: 250 k cells x 500 k SNVs, ~2.48 G observations, 5 000-bin embedding) on the GPU.

  python scripts/c5_probe.py [C5] [--blocks 8] [--no-dense] [--out gpurun_out/c5.json]
  torchrun --nproc-per-node N scripts/c5_probe.py C5 --out ...     SNV-sharded K1 step over N GPUs (every rank: all cells)

Single GPU: builds the store from a block-wise generated data set (synth.make_blocked), times the sparse passes (the
label image of 250 k cells / 500 k SNVs exceeds shared memory: the passes walk it in phases, ph_pass.cu launch_pass),
checks them against oracle/oracle_sparse.c on every 16th SNV / cell (all of their observations), then the dense
embedding kernels at D = 5000: K4 (Gaussian node log-likelihood) against a float64 restatement in torch and the
scipy-rule oracle on one node, K5 (pairwise distance) on a block of rows with the projected cost of the full matrix.
Prints one JSON line.  The oracle is the checker here, never the thing measured."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("config", nargs="?", default="C5")
ap.add_argument("--blocks", type=int, default=8)
ap.add_argument("--n", type=int)
ap.add_argument("--m", type=int)
ap.add_argument("--D", type=int)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--no-dense", action="store_true")
ap.add_argument("--no-oracle", action="store_true")
ap.add_argument("--out")
args = ap.parse_args()

from phertilizer_b200.store import ObservationStore, encode_with_table, order_pairs, pair_counts, pairwise_dist
from phertilizer_b200.synth import CONFIGS, make_blocked

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=dev)
peak = 6544.3
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def log(*a):
    print(f"[rank {rank}]", *a, file=sys.stderr, flush=True)


over = {k: getattr(args, k) for k in ("n", "m", "D") if getattr(args, k)}
out = {"config": args.config, "gpus": world}
t0 = time.time()
s = make_blocked(args.config, device=dev, blocks=args.blocks, with_embedding=False, **over)
torch.cuda.synchronize()
n, m, nnz = s.n, s.m, s.nnz
out.update(n=n, m=m, nnz=nnz, synth_s=round(time.time() - t0, 1))
log(f"synthetic {args.config}: {n} x {m}, {nnz} observations in {out['synth_s']} s")

# one (var,total) table for the whole data set, then 16-bit codes (the int32 var / total of 2.5 G observations never exist)
t0 = time.time()
head = min(nnz, 1 << 28)   # frequency order from the first 2.7e8 observations, then every other (v, t <= 24) pair
pv, pt = order_pairs(*pair_counts(s.var[:head].to(torch.int32), s.total[:head].to(torch.int32)))
have = set(zip(pv.tolist(), pt.tolist()))
extra = [(v, t) for t in range(1, 25) for v in range(t + 1) if (v, t) not in have]
pv = np.concatenate([pv, np.array([p[0] for p in extra], np.int64)])
pt = np.concatenate([pt, np.array([p[1] for p in extra], np.int64)])
codes = torch.empty(nnz, dtype=torch.int16, device=dev)
step = 1 << 28
for lo in range(0, nnz, step):
    hi = min(nnz, lo + step)
    codes[lo:hi] = encode_with_table(s.var[lo:hi].to(torch.int32), s.total[lo:hi].to(torch.int32), pv, pt)
del s.var, s.total
torch.cuda.empty_cache()
out["n_codes"] = int(len(pv))
out["encode_s"] = round(time.time() - t0, 1)

snv_lo, snv_hi = (m * rank) // world, (m * (rank + 1)) // world
cell, snv = s.cell, s.snv
if world > 1:   # SNV-sharded: this rank keeps its block of SNVs (all cells)
    keep = (snv >= snv_lo) & (snv < snv_hi)
    cell, snv, codes = cell[keep].contiguous(), (snv[keep] - snv_lo).contiguous(), codes[keep].contiguous()
    del keep
    torch.cuda.empty_cache()
t0 = time.time()
st = ObservationStore.from_coo(n, snv_hi - snv_lo, cell, snv, None, None, device=local, pairs=(pv, pt), codes=codes)
torch.cuda.synchronize()
info = st.info()
out.update(build_s=round(time.time() - t0, 1), store=info, hbm_peak_gbs=peak)
log(f"store built in {out['build_s']} s: {info}")

g = torch.Generator(device=dev)
g.manual_seed(1)
P = 3
m_loc = snv_hi - snv_lo
cls = [(torch.rand(n + 16, device=dev, generator=g) < 0.5).to(torch.uint8) for _ in range(P)]
cl2 = [torch.randint(0, 3, (n + 16,), device=dev, generator=g, dtype=torch.uint8) for _ in range(P)]
sl = [torch.randint(0, 3, (m_loc + 16,), device=dev, generator=g, dtype=torch.uint8) for _ in range(P)]
K = len(s.parent)
cn = torch.full((n + 16,), 255, dtype=torch.uint8, device=dev)
cn[:n] = s.cell_clone
sn = torch.full((m_loc + 16,), 255, dtype=torch.uint8, device=dev)
sn[:m_loc] = s.snv_node[snv_lo:snv_hi]
present = torch.from_numpy(s.ancestor_matrix()).contiguous().to(dev)
big = max(n, m_loc)
lab_out = torch.empty(m_loc + 16, dtype=torch.uint8, device=dev)
nobs_out = torch.empty(2 * m_loc, dtype=torch.int32, device=dev)
S0 = torch.empty(big * 3, dtype=torch.float64, device=dev)
S1 = torch.empty_like(S0)
No = torch.empty(big * 3, dtype=torch.int32, device=dev)
Nv = torch.empty_like(No)
per_node = torch.empty(K, dtype=torch.float64, device=dev)
per_cell = torch.empty(n, dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream().cuda_stream


def timeit(fn, reps=args.reps):
    for i in range(2):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


by_snv = info["stream4_bytes_by_snv"] or info["stream_bytes_by_snv"]
by_cell = info["stream4_bytes_by_cell"] or info["stream_bytes_by_cell"]
nnz_loc = int(info["nnz"])
passes = {
    "assign_linear": (lambda i: st.assign_linear_dev(cls[i % P], n / 2, lab_out, nobs_out, stream=stream), by_snv),
    "assign_branching": (lambda i: st.assign_branching_dev(cl2[i % P], lab_out, nobs_out, stream=stream), by_snv),
    "snv_table_c2": (lambda i: st.snv_table_dev(cl2[i % P], 2, S0, S1, No, Nv, stream=stream), by_snv),
    "cell_table_q2": (lambda i: st.cell_table_dev(sl[i % P], 2, S0, S1, No, Nv, stream=stream), by_cell),
    "tree_loglik": (lambda i: st.tree_loglik_dev(cn, sn, present, K, per_node, per_cell, stream=stream), by_cell),
}
out["passes"] = {}
for name, (fn, byt) in passes.items():
    ms = timeit(fn)
    out["passes"][name] = {"ms": round(ms, 4), "nnz_per_s": nnz_loc / (ms * 1e-3), "stream_gbs": byt / (ms * 1e-3) / 1e9,
                           "frac_of_hbm_peak": byt / (ms * 1e-3) / 1e9 / peak}
    log(name, out["passes"][name])

if world > 1:
    # the SNV-sharded K1 + K1a step: one fused kernel per rank, results pushed to every rank (ph_peer_assign_linear_snv)
    from phertilizer_b200.sharded import SnvShardedStore

    ss = SnvShardedStore(st, n, m, snv_lo, snv_hi, dev, None)
    lab_all = torch.empty(m + 16, dtype=torch.uint8, device=dev)
    nobs_all = torch.empty(m, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for i in range(3):
        ss.assign_linear_dev(cls[i % P], n / 2, lab_all, nobs_all)
    torch.cuda.synchronize()
    dist.barrier()
    tot = 0.0
    for i in range(args.reps):
        flush.zero_()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ss.assign_linear_dev(cls[i % P], n / 2, lab_all, nobs_all)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    t = torch.tensor([tot / args.reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # every rank must hold the same assembled result; rank 0's own block must equal its plain single-GPU call
    chk = torch.tensor([int(lab_all[:m].to(torch.int64).sum().item()), int(nobs_all.to(torch.int64).sum().item())], dtype=torch.int64, device=dev)
    lo_, hi_ = chk.clone(), chk.clone()
    dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
    st.assign_linear_dev(cls[(args.reps - 1) % P], n / 2, lab_out, nobs_out, stream=stream)
    torch.cuda.synchronize()
    own_ok = bool(torch.equal(lab_out[:m_loc], lab_all[snv_lo:snv_hi]) and torch.equal(nobs_out[:m_loc], nobs_all[snv_lo:snv_hi]))
    out["snv_sharded_step"] = {"ms": round(ms, 4), "nnz_per_s": nnz / (ms * 1e-3), "ranks_agree": bool(torch.equal(lo_, hi_)),
                               "own_block_equals_single_gpu_call": own_ok, "timed_out": bool(ss.peer.timed_out())}
    log("snv_sharded_step", out["snv_sharded_step"])
    ss.peer.close()

# ---------------------------------------------------------------- parity against the sparse C oracle on a sample of lines
if not args.no_oracle and world == 1:
    from oracle.oracle_sparse import SparseOracle, use_all_cores

    use_all_cores()
    STEP = 16
    t0 = time.time()
    # every 16th SNV with ALL of its observations (CSC)
    keep = (s.snv % STEP) == 3
    c_s, s_s, k_s = s.cell[keep], torch.div(s.snv[keep], STEP, rounding_mode="floor"), codes[keep]
    m_s = int((m - 3 + STEP - 1) // STEP)
    order = torch.argsort(s_s.to(torch.int64) * n + c_s.to(torch.int64))
    colptr = np.zeros(m_s + 1, np.int64)
    np.cumsum(torch.bincount(s_s.to(torch.int64), minlength=m_s).cpu().numpy(), out=colptr[1:])
    row, ccode = c_s[order].cpu().numpy(), k_s[order].cpu().numpy().astype(np.uint16)
    del keep, c_s, s_s, k_s, order
    # every 16th cell with all of its observations (CSR)
    keep = (s.cell % STEP) == 5
    c_c, s_c, k_c = torch.div(s.cell[keep], STEP, rounding_mode="floor"), s.snv[keep], codes[keep]
    n_s = int((n - 5 + STEP - 1) // STEP)
    order = torch.argsort(c_c.to(torch.int64) * m + s_c.to(torch.int64))
    rowptr = np.zeros(n_s + 1, np.int64)
    np.cumsum(torch.bincount(c_c.to(torch.int64), minlength=n_s).cpu().numpy(), out=rowptr[1:])
    col, rcode = s_c[order].cpu().numpy(), k_c[order].cpu().numpy().astype(np.uint16)
    del keep, c_c, s_c, k_c, order
    torch.cuda.empty_cache()
    orc = SparseOracle.from_csc_csr(n_s, m_s, colptr, row, ccode, rowptr, col, rcode, st.L0, st.L1, np.asarray(st.pair_var) > 0)
    checks = {}
    # K1 + K1a linear / branching: oracle walks the sampled SNVs over ALL cells
    labA = cls[0][:n].cpu().numpy()
    st.assign_linear_dev(cls[0], n / 2, lab_out, nobs_out, stream=stream)
    torch.cuda.synchronize()
    g_lab, g_nobs = lab_out[:m].cpu().numpy()[3::STEP], nobs_out[:m].cpu().numpy()[3::STEP]
    o_lab, o_nobs = orc.assign_linear(labA, n / 2)
    checks["assign_linear_labels_equal"] = bool(np.array_equal(g_lab, o_lab))
    checks["assign_linear_nobs_equal"] = bool(np.array_equal(g_nobs, o_nobs))
    lab2 = cl2[0][:n].cpu().numpy()
    st.assign_branching_dev(cl2[0], lab_out, nobs_out, stream=stream)
    torch.cuda.synchronize()
    g_lab = lab_out[:m].cpu().numpy()[3::STEP]
    g_nobs = nobs_out[: 2 * m].cpu().numpy().reshape(m, 2)[3::STEP]
    o_lab, o_nobs = orc.assign_branching(lab2)
    checks["assign_branching_nobs_equal"] = bool(np.array_equal(g_nobs, o_nobs))
    bad = np.nonzero(g_lab != o_lab)[0]
    ties_only = True
    if len(bad):
        T0, T1, _, _ = orc.snv_table(lab2, 2)
        cand = np.stack([T1[:, 0] + T0[:, 1], T0[:, 0] + T1[:, 1], T1[:, 0] + T1[:, 1]])
        a_, b_ = cand[o_lab[bad] - 1, bad], cand[g_lab[bad] - 1, bad]
        ties_only = bool(np.all(np.abs(a_ - b_) <= 1e-12 * np.maximum(np.abs(a_), np.abs(b_))))
    checks["assign_branching_label_mismatches"] = int(len(bad))
    checks["assign_branching_mismatches_are_exact_ties"] = ties_only
    # K1 table
    st.snv_table_dev(cl2[0], 2, S0, S1, No, Nv, stream=stream)
    torch.cuda.synchronize()
    T0, T1, N0, N1 = orc.snv_table(lab2, 2)
    checks["snv_table_counts_equal"] = bool(np.array_equal(No[: 2 * m].cpu().numpy().reshape(m, 2)[3::STEP], N0) and
                                            np.array_equal(Nv[: 2 * m].cpu().numpy().reshape(m, 2)[3::STEP], N1))
    checks["snv_table_sums_1e-9"] = bool(np.allclose(S0[: 2 * m].cpu().numpy().reshape(m, 2)[3::STEP], T0, rtol=1e-9, atol=0) and
                                         np.allclose(S1[: 2 * m].cpu().numpy().reshape(m, 2)[3::STEP], T1, rtol=1e-9, atol=0))
    # K2 table and tree log-likelihood: oracle walks the sampled cells over ALL SNVs
    sl0 = sl[0][:m].cpu().numpy()
    st.cell_table_dev(sl[0], 2, S0, S1, No, Nv, stream=stream)
    torch.cuda.synchronize()
    A0, A1, M0, M1 = orc.cell_table(sl0, 2)
    checks["cell_table_counts_equal"] = bool(np.array_equal(No[: 2 * n].cpu().numpy().reshape(n, 2)[5::STEP], M0) and
                                             np.array_equal(Nv[: 2 * n].cpu().numpy().reshape(n, 2)[5::STEP], M1))
    checks["cell_table_sums_1e-9"] = bool(np.allclose(S0[: 2 * n].cpu().numpy().reshape(n, 2)[5::STEP], A0, rtol=1e-9, atol=0) and
                                          np.allclose(S1[: 2 * n].cpu().numpy().reshape(n, 2)[5::STEP], A1, rtol=1e-9, atol=0))
    st.tree_loglik_dev(cn, sn, present, K, per_node, per_cell, stream=stream)
    torch.cuda.synchronize()
    o_cell = orc.tree_loglik(cn[:n].cpu().numpy()[5::STEP], sn[:m].cpu().numpy(), present.cpu().numpy())
    checks["tree_loglik_per_cell_1e-9"] = bool(np.allclose(per_cell.cpu().numpy()[5::STEP], o_cell, rtol=1e-9, atol=0))
    checks["sample"] = f"every {STEP}th SNV ({m_s}) with all its observations; every {STEP}th cell ({n_s}) with all its observations"
    checks["oracle_s"] = round(time.time() - t0, 1)
    out["parity_vs_oracle_sparse_c"] = checks
    log("parity", checks)
    del orc, row, ccode, col, rcode

st.close()
del codes, cell, snv
s.cell = s.snv = None
torch.cuda.empty_cache()

# ---------------------------------------------------------------- dense embedding kernels at D bins
if not args.no_dense and world == 1:
    cfg = dict(CONFIGS[args.config]); cfg.update(over)
    D = cfg["D"]
    t0 = time.time()
    e = make_blocked(args.config, device=dev, blocks=0, with_embedding=True, **over).embedding
    torch.cuda.synchronize()
    out["embedding"] = {"n": n, "D": D, "gb": n * D * 8 / 1e9, "synth_s": round(time.time() - t0, 1)}
    # K4 needs a handle only for scratch / stream: a store without observations over the n cells
    dense = ObservationStore.from_coo(n, 1, np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32),
                                      device=local, pairs=(np.array([0]), np.array([1])))
    ms = timeit(lambda i: dense.gauss_node_ll_dev(e, D, cn, K, per_node, per_cell, stream=stream), reps=5)
    out["K4_gauss_node_ll"] = {"ms": round(ms, 3), "X_gb": n * D * 8 / 1e9, "algorithmic_gbs": n * D * 8 / (ms * 1e-3) / 1e9,
                               "frac_of_hbm_peak_algorithmic": n * D * 8 / (ms * 1e-3) / 1e9 / peak}
    got = per_node.cpu().numpy().copy()
    # float64 restatement in torch (per node: mean, population variance, diagonal Gaussian log-density; all variances are far
    # above scipy's singularity threshold for Poisson bins), and the scipy-rule oracle on the smallest node
    ref = np.zeros(K)
    cnl = cn[:n].long()
    for k in range(K):
        rows = torch.nonzero(cnl == k)[:, 0]
        if rows.numel() == 0:
            continue
        acc = 0.0
        for lo in range(0, D, 500):
            X = e[rows][:, lo:lo + 500]
            mu = X.mean(dim=0)
            var = ((X - mu) ** 2).mean(dim=0)
            acc += float((-0.5 * (torch.log(var).sum() * rows.numel() + (((X - mu) ** 2) / var).sum())).item())
        ref[k] = acc - 0.5 * D * rows.numel() * 1.8378770664093453
    out["K4_gauss_node_ll"]["max_rel_diff_vs_torch_f64"] = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)))
    try:
        from oracle import oracle_np as O

        sizes = torch.bincount(cnl, minlength=K).cpu().numpy()
        k = int(np.argmin(np.where(sizes > 0, sizes, 1 << 60)))
        rows = torch.nonzero(cnl == k)[:, 0].cpu().numpy()
        val = O.gauss_node_loglik(e[torch.from_numpy(rows).to(dev)].cpu().numpy(), np.arange(len(rows)))[0]
        out["K4_gauss_node_ll"]["rel_diff_vs_scipy_rule_oracle_smallest_node"] = float(abs(got[k] - val) / abs(val))
        out["K4_gauss_node_ll"]["smallest_node_cells"] = int(len(rows))
    except Exception as ex:  # pragma: no cover
        out["K4_gauss_node_ll"]["oracle_error"] = str(ex)
    log("K4", out["K4_gauss_node_ll"])
    dense.close()
    # K5: a block of rows of squareform(pdist(embedding)) (phertilizer.py:286)
    R = 1024
    dist_out = torch.empty(R * n, dtype=torch.float64, device=dev)
    from phertilizer_b200 import _lib
    from phertilizer_b200.store import _ptr

    def k5(i):
        _lib.check(_lib.lib().ph_pairwise_dist(_ptr(e), n, D, 0, R, _ptr(dist_out), local, 1, stream))
    ms = timeit(k5, reps=3)
    pair_dims = R * n * D
    out["K5_pairwise_dist"] = {"rows": R, "ms": round(ms, 2), "pair_dims_per_s": pair_dims / (ms * 1e-3),
                               "fp64_flop_per_s_3_per_pair_dim": 3 * pair_dims / (ms * 1e-3),
                               "projected_full_matrix_s": ms * 1e-3 * n / R, "full_matrix_bytes": n * n * 8}
    ref = torch.cdist(e[:64], e[:4096]).cpu().numpy()
    got5 = dist_out[: R * n].reshape(R, n)[:64, :4096].cpu().numpy()
    out["K5_pairwise_dist"]["max_rel_diff_vs_torch_cdist_f64"] = float(np.max(np.abs(got5 - ref) / np.maximum(ref, 1e-300)))
    log("K5", out["K5_pairwise_dist"])

if rank == 0:
    print(json.dumps(out))
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        open(args.out, "w").write(json.dumps(out) + "\n")
if world > 1:
    dist.destroy_process_group()
