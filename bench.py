#!/usr/bin/env python
"""bench.py -- likelihood hot path of Phertilizer on B200: nnz/s of one SNV-reassignment pass.

A "step" is one pass of the hot path over all observations: the SNV reassignment of the linear split
(reference linear_split.py:203-237, `Linear_split.mut_assignment`) = K1 (per-SNV like0/like1 sums + observation
counts over cellsA) fused with K1a (mean-likelihood decision), for a balanced random cellsA and all SNVs active.
Workload (BASELINE.json configs[2], the one the north-star target is quoted on; it fits one GPU): synthetic
100k cells x 200k SNVs at 0.02x coverage, 12 clones, ~396 M observations.  With --gpus N the problem is fixed
(strong scaling) and sharded by SNVs (default: per-SNV tables are independent, every rank runs ONE fused kernel per
step that also pushes its results to all ranks over NVLink peer memory and unpacks the assembled window) or by cells
(--shard cells: integer histograms over peer memory, or --exchange nccl: one all-reduce of the partial tables).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C3]

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "likelihood_nnz_per_sec"
UNIT = "nnz/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while a timed region runs (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows if len(r) > 3 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------ data
def make_workload(name, device, lo_frac=0.0, hi_frac=1.0):
    """Synthetic COO of the named config; keeps only the cells of this rank's block."""
    import torch

    from phertilizer_b200.synth import CONFIGS, make_config

    s = make_config(name, device=device, with_embedding=False)
    n = s.n
    lo, hi = int(round(lo_frac * n)), int(round(hi_frac * n))
    nnz_total = s.nnz
    if lo > 0 or hi < n:
        keep = (s.cell >= lo) & (s.cell < hi)
        cell, snv, var, total = (s.cell[keep] - lo).contiguous(), s.snv[keep].contiguous(), s.var[keep].contiguous(), s.total[keep].contiguous()
    else:
        cell, snv, var, total = s.cell, s.snv, s.var, s.total
    cfg = dict(CONFIGS[name])
    return dict(n=n, m=s.m, lo=lo, hi=hi, n_local=hi - lo, cell=cell, snv=snv, var=var, total=total, nnz_total=nnz_total,
                nnz_local=int(cell.numel()), cfg=cfg)


def alg_bytes(nnz, n, m, C=1):
    """nnz*(4 B index + 2 B (var,total) code) + (m+1)*8 offsets + n labels + m*C*24 outputs."""
    return nnz * 6 + (m + 1) * 8 + n + m * C * 24


def host_csc(w, code_t):
    """Device COO -> host CSC (colptr, row, code) for the CPU oracle; the sort runs on the GPU (plumbing)."""
    import torch

    key = w["snv"].to(torch.int64) * w["n_local"] + w["cell"].to(torch.int64)
    order = torch.argsort(key)
    del key
    row = w["cell"][order].cpu().numpy()
    code = code_t[order].cpu().numpy().astype(np.uint16)
    cnt = torch.bincount(w["snv"].to(torch.int64), minlength=w["m"]).cpu().numpy()
    colptr = np.zeros(w["m"] + 1, np.int64)
    np.cumsum(cnt, out=colptr[1:])
    del order
    return colptr, row, code


def cpu_linear_passes(orc, labels, lenA, budget_s, max_passes):
    t_tot, passes, out = 0.0, 0, None
    while passes < max_passes and (t_tot < budget_s or passes < 2):
        t0 = time.perf_counter()
        out = orc.assign_linear(labels[passes % len(labels)], lenA[passes % len(labels)])
        t_tot += time.perf_counter() - t0
        passes += 1
    return t_tot, passes, out


def extra_passes(st, w, info, peak, args):
    """Device-timed (CUDA events, back to back, rotating label sets; every pass streams more than L2) numbers for the other
    passes of the hot path on the same store: what `roofline` reports for the linear step, per pass."""
    import torch

    dev = torch.device("cuda", torch.cuda.current_device())
    n, m, nnz = w["n"], w["m"], w["nnz_total"]
    g = torch.Generator(device="cpu")
    g.manual_seed(11)
    P = 3
    cl = [torch.randint(0, 3, (n + 16,), generator=g, dtype=torch.uint8).to(dev) for _ in range(P)]
    sl = [torch.randint(0, 3, (m + 16,), generator=g, dtype=torch.uint8).to(dev) for _ in range(P)]
    K = 12
    cn = [torch.randint(0, K, (n + 16,), generator=g, dtype=torch.uint8).to(dev) for _ in range(P)]
    sn = [torch.randint(0, K, (m + 16,), generator=g, dtype=torch.uint8).to(dev) for _ in range(P)]
    present = torch.tril(torch.ones(K, K, dtype=torch.uint8)).contiguous().to(dev)
    big = max(n, m)
    lab_out = torch.empty(m + 16, dtype=torch.uint8, device=dev)
    nobs_out = torch.empty(2 * m, dtype=torch.int32, device=dev)
    S0 = torch.empty(big * 3, dtype=torch.float64, device=dev)
    S1 = torch.empty_like(S0)
    No = torch.empty(big * 3, dtype=torch.int32, device=dev)
    Nv = torch.empty_like(No)
    b0, b1 = torch.empty(16, dtype=torch.float64, device=dev), torch.empty(16, dtype=torch.float64, device=dev)
    bn, bv = torch.empty(16, dtype=torch.int64, device=dev), torch.empty(16, dtype=torch.int64, device=dev)
    per_node = torch.empty(K, dtype=torch.float64, device=dev)
    per_cell = torch.empty(n, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    by_snv = info["stream4_bytes_by_snv"] or info["stream_bytes_by_snv"]
    by_cell = info["stream4_bytes_by_cell"] or info["stream_bytes_by_cell"]
    nsub_s = info["dense_codes_by_snv"] * info["chunks_by_snv"]
    nsub_c = info["dense_codes_by_cell"] * info["chunks_by_cell"]
    ctas = info["sm_count"]
    fixed_s = (m * nsub_s + 1) * 4 + (m + 1) * 8 + n * ctas
    fixed_c = (n * nsub_c + 1) * 4 + (n + 1) * 8 + m * ctas
    passes = {
        "assign_branching": (lambda i: st.assign_branching_dev(cl[i % P], lab_out, nobs_out, stream=stream), by_snv + fixed_s + 9 * m,
                             "ph_assign_branching: K1 + 3-way K1a, 2 cell clusters (branching_split.py:293-348)"),
        "snv_table_c2": (lambda i: st.snv_table_dev(cl[i % P], 2, S0, S1, No, Nv, stream=stream), by_snv + fixed_s + 48 * m,
                         "ph_snv_table: K1, S0/S1/Nobs/Nvar for 2 cell clusters"),
        "cell_table_q2": (lambda i: st.cell_table_dev(sl[i % P], 2, S0, S1, No, Nv, stream=stream), by_cell + fixed_c + 48 * n,
                          "ph_cell_table: K2, per-cell table over 2 SNV clusters (calc_tmb / check_metrics)"),
        "block_sums_2x2": (lambda i: st.block_sums_dev(cl[i % P], 2, sl[i % P], 2, b0, b1, bn, bv, stream=stream), by_snv + fixed_s + 96 * m,
                           "ph_block_sums: K3 = K1 table + fixed-order block reduction (compute_norm_likelihood / check_obs)"),
        "tree_loglik_k12": (lambda i: st.tree_loglik_dev(cn[i % P], sn[i % P], present, K, per_node, per_cell, stream=stream), by_cell + fixed_c + 16 * n,
                            "ph_tree_loglik: variant log-likelihood of a 12-node tree (clonal_tree.py:932-944)"),
    }
    out = {}
    reps = max(5, min(args.steps, 20))
    for name, (fn, phys, what) in passes.items():
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[name] = {"ms": ms, "nnz_per_s": nnz / (ms * 1e-3), "physical_bytes": int(phys), "physical_gbs": phys / (ms * 1e-3) / 1e9,
                     "frac": phys / (ms * 1e-3) / 1e9 / peak, "what": what}
    return out


def e2e_tree_small():
    """BASELINE.json's metric names a second quantity: end-to-end `phertilize()` wall time.  A live measurement that fits a
    bench run: config 2 (10k cells x 50k SNVs, ~24 M observations), 3 starts x 10 iterations with post-processing, through the
    public `Phertilizer` class.  The north-star shape C3 at 1 / 2 / 4 / 8 GPUs is measured by scripts/e2e_tree.py (minutes) and
    committed under profiles/ (r2_e2e_tree.md)."""
    import io
    from contextlib import redirect_stdout

    import torch

    from phertilizer_b200.phertilizer import Phertilizer
    from phertilizer_b200.synth import make_config, to_dataframes

    try:
        s = make_config("C2", device="cuda")
        vc, bc = to_dataframes(s)
        t0 = time.time()
        ph = Phertilizer(vc, bc, mode="rd", dim_reduce=False)
        torch.cuda.synchronize()
        t_init = time.time() - t0
        t0 = time.time()
        with redirect_stdout(io.StringIO()):
            tree, trees, _ = ph.phertilize(max_iterations=10, starts=3, seed=9905900, radius=1, gamma=0.95, d=10,
                                           use_copy_kernel=True, post_process=True)
        t_run = time.time() - t0
        return {"config": "C2: 10000 cells x 50000 SNVs, coverage 0.05, 5 clones; -s 3 -j 10 --post_process --no-umap", "nnz": int(s.nnz),
                "init_s": round(t_init, 2), "phertilize_s": round(t_run, 2), "tree_nodes": int(tree.tree.number_of_nodes()),
                "candidate_trees": int(trees.size()), "norm_loglikelihood": float(tree.norm_loglikelihood),
                "larger_shapes": "profiles/r2_e2e_tree.md (C3 at 1 / 2 / 4 / 8 GPUs, C4 at the CLI defaults)"}
    except Exception as e:  # the tree is an extra: never take the bench line down with it
        return {"error": str(e)[:300]}


# ------------------------------------------------------------------------------------------------ arms
def run_reference(args):
    """CPU arm: the sparse oracle port of the reference path (OpenMP, all host cores) on the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle.oracle_sparse import SparseOracle, threads, use_all_cores
    from phertilizer_b200.like_lut import like_tables
    from phertilizer_b200.store import encode_pairs_device

    use_all_cores()   # torchrun exports OMP_NUM_THREADS=1: the reference arm still gets every host core
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    t0 = time.time()
    w = make_workload(args.workload, dev)
    pv, pt, code_t = encode_pairs_device(w["var"], w["total"])
    L0, L1 = like_tables(pv, pt)
    colptr, row, code = host_csc(w, code_t)
    orc = SparseOracle.from_csc(w["n"], w["m"], colptr, row, code, L0, L1, pv > 0)
    log(f"[reference] data ready in {time.time() - t0:.1f}s, nnz={w['nnz_total']}, threads={threads()}")
    rng = np.random.default_rng(7)
    labels = [(rng.random(w["n"]) < 0.5).astype(np.uint8) for _ in range(4)]
    lenA = [float(l.sum()) for l in labels]
    for i in range(args.warmup):
        orc.assign_linear(labels[i % 4], lenA[i % 4])
    t0 = time.perf_counter()
    for i in range(args.steps):
        orc.assign_linear(labels[i % 4], lenA[i % 4])
    dt = time.perf_counter() - t0
    value = args.steps * w["nnz_total"] / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, w, args.gpus),
        "notes": {"residency": "host memory (no GPU on this arm)", "l2_policy": "n/a (CPU arm; 0.8 GB of observations streamed per pass)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads(), "kind": "port",
                         "sample": f"all {w['nnz_total']} observations, {args.steps} timed passes; the dense reference "
                                   "cannot allocate this shape (4 x 160 GB), oracle/oracle_sparse.c restates its sums over CSC"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, w, world, exchange=None):
    """Names the workload only (identical in both arms); how each arm handles L2 / residency is reported under `notes`."""
    c = w["cfg"]
    return {"workload": f"{args.workload}: synthetic {c['n']} cells x {c['m']} SNVs, coverage {c['cov']}, {c['K']} clones "
                        f"(BASELINE.json configs[2]); step = linear-split SNV reassignment pass (K1+K1a) over all observations",
            "nnz": w["nnz_total"], "cells": c["n"], "snvs": c["m"], "coverage": c["cov"], "clones": c["K"],
            "l2_policy": l2_policy_for(w["nnz_total"], world)}


L2_NOMINAL = 134_217_728   # bytes; B200 reports 126-133 MB depending on the query: the rule below must not depend on it


def needs_l2_flush(nnz_total, world):
    """A rank streams ~2.1 B per observation per pass: below twice the L2 size part of it would stay cached between steps."""
    return 2.1 * nnz_total / world < 2 * L2_NOMINAL


def l2_policy_for(nnz_total, world):
    return ("L2 flushed (write of a buffer of twice its size, untimed) before every timed step" if needs_l2_flush(nnz_total, world)
            else "inputs larger than L2 (no flush)")


def sharding_note(world, exchange):
    if world == 1:
        return "single GPU"
    if exchange == "snv-peer":
        return (f"SNVs sharded over {world} rank(s), all cells on every rank: per-SNV tables are independent, no partial-table exchange; every batch's "
                "labels + counts leave the kernel as self-validating 8-byte words (step counter in the upper half) stored into every rank's window over "
                "NVLink peer memory while the kernel streams on; the same launch then polls its own window and unpacks it (one kernel per step, no fence, "
                "no flag, no NCCL)")
    return f"cells sharded over {world} rank(s); " + ("per-SNV integer histograms pushed to the owning rank over NVLink peer memory, labels pushed back (no NCCL)"
                                                      if exchange == "peer" else "NCCL all-reduce of the [S0|S1|Nobs] table")


def run_ours(args):
    import torch
    import torch.distributed as dist

    from phertilizer_b200 import _lib
    from phertilizer_b200.sharded import CudaBackend, ShardedStore
    from phertilizer_b200.store import ObservationStore, encode_pairs_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: phertilizer_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL announces its version on STDOUT when the first communicator comes up; stdout must carry the JSON line only
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    t0 = time.time()
    snv_sharded = world > 1 and args.shard == "snvs"
    if snv_sharded:
        w = make_workload(args.workload, dev)            # all cells; the SNV block is cut below
    else:
        w = make_workload(args.workload, dev, rank / world, (rank + 1) / world)
    n_loc, m, nnz_loc, nnz_tot = w["n_local"], w["m"], w["nnz_local"], w["nnz_total"]
    pairs = None
    if world > 1:  # the ranks must agree on the (var,total) codes: the exchanged histograms are indexed by code
        from phertilizer_b200.sharded import global_pairs

        pairs = global_pairs(w["var"], w["total"])
    snv_store = None
    if snv_sharded:
        from phertilizer_b200.sharded import SnvShardedStore

        try:
            snv_store = SnvShardedStore.from_coo(w["n"], m, w["cell"], w["snv"], w["var"], w["total"], rank, world, local)
        except _lib.PhError as e:   # raised on every rank alike (sharded.PeerExchange): fall back to cell sharding
            log(f"[rank {rank}] SNV-sharded peer path unavailable ({e}); using --shard cells")
            snv_sharded = False
            w = make_workload(args.workload, dev, rank / world, (rank + 1) / world)
            n_loc, m, nnz_loc, nnz_tot = w["n_local"], w["m"], w["nnz_local"], w["nnz_total"]
    if snv_store is not None:
        st = snv_store.store
        nnz_loc = int(st.info()["nnz"])
        sharded, exchange = None, "snv-peer"
    else:
        st = ObservationStore.from_coo(n_loc, m, w["cell"], w["snv"], w["var"], w["total"], device=local, pairs=pairs)
        sharded = ShardedStore(CudaBackend(st, dev), w["n"], m, w["lo"], w["hi"], exchange=args.exchange) if world > 1 else None
        exchange = sharded.exchange if sharded is not None else None
    torch.cuda.synchronize()
    info = st.info()
    log(f"[rank {rank}] store ready in {time.time() - t0:.1f}s: {info}")

    # L2 policy: a pass streams the rank's observations once.  On one GPU that is 0.8 GB against 126 MB of L2; a rank of
    # a multi-GPU run may hold little enough for L2 to keep part of it between steps -> flush L2 before every timed step
    l2_bytes = int(torch.cuda.get_device_properties(dev).L2_cache_size)
    flush_l2 = needs_l2_flush(nnz_tot, world)
    flush_buf = torch.empty(2 * l2_bytes, dtype=torch.uint8, device=dev) if flush_l2 else None
    if flush_l2:
        l2_policy = (f"per-rank stream of {info['stream_bytes_by_snv'] / 1e6:.0f} MB is below 2 x L2 ({l2_bytes / 1e6:.0f} MB): L2 flushed (write of a "
                     f"{2 * l2_bytes / 1e6:.0f} MB buffer, untimed) before every timed step, each step timed by its own CUDA event pair after a device-side rendez-vous of the ranks; label vectors rotate every step")
    else:
        l2_policy = (f"inputs larger than L2 ({info['stream_bytes_by_snv'] / 1e9:.2f} GB of observations streamed per pass and rank vs "
                     f"{l2_bytes / 1e6:.0f} MB L2); label vectors rotate every step")

    # label sets: balanced random cellsA (global draw, local slice), rotating so no step repeats its predecessor
    P = 4
    g = torch.Generator(device="cpu")
    g.manual_seed(7)
    labs_h, labs_d, lenA, full_labs_h = [], [], [], []
    for _ in range(P):
        full = (torch.rand(w["n"], generator=g) < 0.5).to(torch.uint8)
        full_labs_h.append(full)
        lenA.append(float(full.sum()))
        loc = torch.zeros(n_loc + 16, dtype=torch.uint8)
        loc[:n_loc] = full[w["lo"]:w["hi"]]          # SNV-sharded: lo..hi = all cells
        labs_h.append(loc.pin_memory())
        labs_d.append(loc.to(dev))
    lab_out = torch.empty(m + 16, dtype=torch.uint8, device=dev)
    nobs_out = torch.empty(m, dtype=torch.int32, device=dev)
    packed = torch.empty(3 * m, dtype=torch.float64, device=dev) if world > 1 else None
    lab_out_h = torch.empty(m + 16, dtype=torch.uint8).pin_memory()
    nobs_out_h = torch.empty(m, dtype=torch.int32).pin_memory()
    lab_out_np, nobs_out_np = lab_out_h.numpy()[:m], nobs_out_h.numpy()
    stream = torch.cuda.current_stream().cuda_stream

    kern_ev = []

    def step_device(i, record=False):
        if graphs is not None:
            graphs[i % P].replay()
            return
        if world == 1:
            st.assign_linear_dev(labs_d[i % P], lenA[i % P], lab_out, nobs_out, stream=torch.cuda.current_stream().cuda_stream)
        elif exchange == "snv-peer":
            # each rank: the fused single-GPU kernel on its SNVs, epilogue pushes label + counts to every rank; collect
            snv_store.assign_linear_dev(labs_d[i % P], lenA[i % P], lab_out, nobs_out)
        elif exchange == "peer":
            # pass kernel with peer-memory epilogue + owner finish + collect (csrc/ph_peer.cu); no NCCL in the step
            sharded.assign_linear_dev(labs_d[i % P], lenA[i % P], packed, lab_out, nobs_out)
        else:
            if record:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            # == ShardedStore.assign_linear_dev, unrolled here only to bracket the partial-table kernel with events
            sharded.b.snv_partial(labs_d[i % P], 1, packed)
            if record:
                e1.record()
                kern_ev.append((e0, e1))
            dist.all_reduce(packed)
            sharded.b.finish_linear(packed, m, lenA[i % P], lab_out, nobs_out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (value) ----------------
    graphs = None
    for i in range(args.warmup):
        step_device(i)
    barrier()
    l0 = _lib.kernel_launches()
    step_device(0)
    launches_per_step = _lib.kernel_launches() - l0 + (1 if exchange == "nccl" else 0)
    barrier()
    if args.cuda_graphs and exchange != "nccl":
        # One CUDA graph per label set (the step's kernel).  A step is one launch of tens of microseconds: graph replay
        # takes the host launch path out of the loop, but graph-to-graph hand-over costs a few microseconds that plain
        # back-to-back launches do not pay -- a short untimed trial of both decides (rank 0's verdict, for all ranks).
        def trial(n=40):
            barrier()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            for i in range(n):
                step_device(i)
            b_.record()
            barrier()
            return a_.elapsed_time(b_) / n

        try:
            t_eager = trial()
            captured = []
            for j in range(P):
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_):
                    step_device(j)
                captured.append(g_)
            graphs = captured
            for i in range(args.warmup):
                step_device(i)
            t_graph = trial()
            keep = torch.tensor([1 if t_graph <= t_eager else 0], dtype=torch.int32, device=dev)
            if world > 1:
                dist.broadcast(keep, 0)
            log(f"[rank {rank}] step trial: eager {1e3 * t_eager:.1f} us, graph replay {1e3 * t_graph:.1f} us")
            if int(keep.item()) == 0:
                graphs = None
            barrier()
        except Exception as e:  # capture not possible: stay eager
            log(f"[rank {rank}] CUDA graph capture failed ({e}); running eagerly")
            graphs = None
    # The timed region of K steps lasts a few ms -- shorter than one nvidia-smi poll -- so the sampler also runs over an
    # untimed ~1 s stretch of the very same steps right before it (clocks "under load"), and stays on through the
    # timed region.
    with ClockSampler(local) as clk:
        t_load = time.perf_counter()
        n_load = 0
        while True:
            # every rank must run the SAME number of (collective) steps: rank 0's clock decides, the others follow
            go = torch.tensor([1 if time.perf_counter() - t_load < args.clock_seconds else 0], dtype=torch.int32, device=dev)
            if world > 1:
                dist.broadcast(go, 0)
            if int(go.item()) == 0:
                break
            for i in range(50):
                step_device(n_load + i)
            n_load += 50
            torch.cuda.synchronize()
        barrier()
        barrier()
        if flush_l2:
            # the rank's share of the observations would (partly) stay in L2 from one step to the next: every timed step
            # is preceded by an untimed write of a buffer twice the size of L2 and bracketed by its own pair of events
            pairs_ev = []
            sync_t = torch.zeros(1, dtype=torch.int32, device=dev)
            for i in range(args.steps):
                flush_buf.zero_()
                if world > 1:
                    dist.all_reduce(sync_t)   # device-side rendez-vous: the ranks' flushes end at different times, the step starts together
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record()
                step_device(args.warmup + i, record=True)
                b_.record()
                pairs_ev.append((a_, b_))
            barrier()
            ms = sum(a_.elapsed_time(b_) for a_, b_ in pairs_ev)
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.steps):
                step_device(args.warmup + i, record=True)
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1)
    launches = launches_per_step * args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = args.steps * nnz_tot / (ms * 1e-3)
    kernel_ms = (sum(a.elapsed_time(b) for a, b in kern_ev) / len(kern_ev)) if kern_ev else ms / args.steps

    # ---------------- end to end through the C-ABI with HOST buffers ----------------
    def step_e2e(i):
        if world == 1:
            # ph_assign_linear(on_device=0): pinned staging, H2D labels, kernel, D2H labels + counts, sync
            return st.assign_linear(labs_h[i % P].numpy()[:n_loc], lenA[i % P], out=(lab_out_np, nobs_out_np))
        labs_d[i % P].copy_(labs_h[i % P], non_blocking=True)
        if snv_store is not None:
            # the collect phase unpacks in the peers' work order (scattered): into device arrays, then two D2H copies
            snv_store.assign_linear_dev(labs_d[i % P], lenA[i % P], lab_out, nobs_out)
        else:
            sharded.assign_linear_dev(labs_d[i % P], lenA[i % P], packed, lab_out, nobs_out)
        lab_out_h.copy_(lab_out, non_blocking=True)
        nobs_out_h.copy_(nobs_out, non_blocking=True)
        torch.cuda.synchronize()
        return lab_out_h, nobs_out_h

    for i in range(min(args.warmup, 3)):
        step_e2e(i)
    barrier()
    if flush_l2:
        e2e_s = 0.0
        for i in range(args.steps):
            flush_buf.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            step_e2e(i)
            e2e_s += time.perf_counter() - t0
        barrier()
    else:
        t0 = time.perf_counter()
        for i in range(args.steps):
            step_e2e(i)
        barrier()
        e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = args.steps * nnz_tot / e2e_s

    # ---------------- CPU baseline + parity on EVERY line (rank 0): the oracle port on the host cores ----------------
    # Every rank runs the N-rank step once per label set; rank 0 keeps the assembled outputs and compares each of them
    # with oracle/oracle_sparse.c on ALL observations (N > 1: a wrong rank offset or a lost word cannot hide).
    cpu = None
    peer = snv_store.peer if snv_store is not None else (sharded.peer if sharded is not None else None)
    if not args.no_cpu:
        graphs = None
        got = []
        for j in range(P):
            step_device(j)
            torch.cuda.synchronize()
            got.append((lab_out[:m].cpu().numpy().copy(), nobs_out.cpu().numpy().copy()))
        timed_out = bool(peer.timed_out()) if peer is not None else False
        verdict = torch.ones(1, dtype=torch.int32, device=dev)
        if rank == 0:
            from oracle.oracle_sparse import SparseOracle, threads, use_all_cores
            from phertilizer_b200.store import encode_with_table

            use_all_cores()
            t0 = time.time()
            wf = w if (world == 1 or snv_sharded) else make_workload(args.workload, dev)   # the oracle walks ALL observations
            code_t = encode_with_table(wf["var"], wf["total"], st.pair_var, st.pair_total)
            colptr, row, code = host_csc(wf, code_t)
            orc = SparseOracle.from_csc(wf["n"], m, colptr, row, code, st.L0, st.L1, np.asarray(st.pair_var) > 0)
            log(f"[cpu] host CSC ready in {time.time() - t0:.1f}s")
            full_labs = [l.numpy() for l in full_labs_h]
            t_tot, passes, _ = cpu_linear_passes(orc, full_labs, lenA, args.cpu_seconds if world == 1 else min(args.cpu_seconds, 3.0), 40)
            same = not timed_out
            for j in range(P):
                o_lab, o_nobs = orc.assign_linear(full_labs[j], lenA[j])
                same = same and bool(np.array_equal(got[j][0], o_lab) and np.array_equal(got[j][1], o_nobs))
            log(f"[cpu] {passes} passes in {t_tot:.2f}s on {threads()} threads; {world}-rank GPU labels + Nobs == oracle on {P} label sets: {same}"
                + (" (a peer wait TIMED OUT)" if timed_out else ""))
            cpu = {"value": passes * nnz_tot / t_tot, "unit": UNIT, "cores": threads(), "kind": "port",
                   "sample": f"all {nnz_tot} observations x {passes} passes of oracle/oracle_sparse.c orc_assign_linear (OpenMP); "
                             "the dense numpy reference cannot allocate this shape (4 x 160 GB)",
                   "gpu_matches_oracle": same, "label_sets_compared": P}
            verdict[0] = 1 if same else 0
        if world > 1:
            dist.broadcast(verdict, 0)
        if int(verdict.item()) == 0:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "error": "GPU assignment differs from the CPU oracle on the benchmark workload",
                                  "n_gpus": world, "cpu_baseline": cpu}), flush=True)
            raise SystemExit("bench: GPU assignment differs from the CPU oracle on the benchmark workload")

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        m_loc = int(info["n_snvs"])      # SNVs this rank walks (all of them unless SNV-sharded)
        ab = alg_bytes(nnz_loc, n_loc, m_loc, 1)
        # what the kernel really streams: 2 B per observation incl. 16-byte padding, tail words, sub-segment and tail
        # offsets, one label image per CTA, the outputs (labels + counts)
        nsub = info["dense_codes_by_snv"] * info["chunks_by_snv"]
        # single GPU: whole-orientation passes read the batched-4 copy of the stream when the store keeps one
        stream = info["stream4_bytes_by_snv"] if (exchange in (None, "snv-peer") and info.get("stream4_bytes_by_snv", 0)) else info["stream_bytes_by_snv"]
        phys = stream + (m_loc * nsub + 1) * 4 + (m_loc + 1) * 8 + n_loc * info["sm_count"] + 5 * m_loc
        achieved = ab / (kernel_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{args.workload}_assign_linear_dram_bytes")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, w, world, exchange),
            "notes": {"sharding": sharding_note(world, exchange), "l2_policy": l2_policy,
                      "residency": "observations resident in HBM; `value` also has labels resident, `e2e` copies them every step"},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n_loc) * world, "d2h_bytes_per_step": 5 * m * world,
                    "ms_per_step": 1e3 * e2e_s / args.steps,
                    "path": "ph_assign_linear(on_device=0): host labels -> pinned -> H2D, kernel storing SNV labels + Nobs into the page-locked host arrays, sync"
                            if world == 1 else ("pinned H2D labels, ph_peer_assign_linear_snv, D2H labels + Nobs" if exchange == "snv-peer" else
                                                "pinned H2D labels, ph_peer_assign_linear, D2H labels + Nobs" if exchange == "peer" else
                                                "pinned H2D labels, ph_snv_partial, NCCL all-reduce, ph_finish_linear, D2H labels + Nobs")},
            "gpu_launches": int(launches), "cuda_graphs": graphs is not None,
            # `achieved` / `frac_alg6` follow SURVEY 8d (ALG-6 bytes: 4 B index + 2 B code per observation); the store
            # moves a third of that (16-bit entries, code implied by the sub-segment), so ALG-6 bytes / time exceeds the
            # HBM peak and is NOT a utilisation.  `frac` = bytes the kernel really moves / time / measured peak.
            "roofline": {"bound": "hbm", "achieved": phys / (kernel_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": phys / (kernel_ms * 1e-3) / 1e9 / peak, "achieved_alg6": achieved, "frac_alg6": achieved / peak,
                         "traffic": traffic if world == 1 else None,
                         "kernel": "k_pass<MODE_LINEAR>" if world == 1 else ("k_pass<MODE_TABLE>" if exchange == "nccl" else
                                   ("k_pass<MODE_LINEAR, push + collect> (whole step = one launch)" if exchange == "snv-peer" else
                                    "k_pass<MODE_COUNTS> + k_peer_finish + k_peer_collect (whole step: kernel_ms includes the exchange)")),
                         "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": ab,
                         "physical_bytes_per_launch": phys, "physical_bytes_per_nnz": phys / max(nnz_loc, 1),
                         "physical_gbs": phys / (kernel_ms * 1e-3) / 1e9, "physical_frac": phys / (kernel_ms * 1e-3) / 1e9 / peak,
                         "peak_source": peak_src + ", of measured"},
            "cpu_baseline": cpu,
            "extra": {"passes": extra_passes(st, w, info, peak, args) if world == 1 else None,
                      "e2e_tree": e2e_tree_small() if (world == 1 and not args.no_tree) else None},
            "store": {k: info[k] for k in ("dense_codes_by_snv", "chunks_by_snv", "group_by_snv", "labels_in_smem_by_snv",
                                           "stream_bytes_by_snv", "stream4_bytes_by_snv", "device_bytes")},
        }
        print(json.dumps(line), flush=True)
    if peer is not None:
        if peer.timed_out():
            log(f"[rank {rank}] WARNING: a peer wait timed out during the run")
        peer.close()
    st.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU baseline budget (N=1 only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-tree", action="store_true", help="skip the small end-to-end tree (extra.e2e_tree)")
    ap.add_argument("--clock-seconds", type=float, default=1.0, help="untimed load stretch sampled for SM clocks")
    ap.add_argument("--no-cuda-graphs", dest="cuda_graphs", action="store_false",
                    help="launch every step eagerly instead of replaying a captured CUDA graph")
    ap.add_argument("--shard", default="snvs", choices=["snvs", "cells"],
                    help="N>1: which axis is split across ranks for the SNV-reassignment step (DESIGN.md section 5)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl"],
                    help="N>1: how the per-SNV partial tables are combined (peer = fused NVLink peer-memory kernels)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
