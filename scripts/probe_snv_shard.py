"""Device timing of ONE rank's share of an SNV-sharded K1 step, on one GPU (development aid, not the bench).

python scripts/probe_snv_shard.py C3 8   ->  C3 with m / 8 SNVs (what each rank of an 8-GPU run walks), world = 1 peer
context (local window): plain fused kernel vs. kernel with record push + unpack, eager and CUDA-graph replay."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
if os.environ.get("PH_LIB"):   # development aid: an alternative build of the library (e.g. -DPH_STAMPS)
    from phertilizer_b200 import _lib as _l
    _l.LIB_PATH = os.path.abspath(os.environ["PH_LIB"])
from phertilizer_b200.sharded import SnvShardedStore
from phertilizer_b200.synth import CONFIGS, make_config

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
groups = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [8]
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29547")
dist.init_process_group("gloo", rank=0, world_size=1)
dev = torch.device("cuda:0")
m = CONFIGS[name]["m"] // parts
s = make_config(name, device=dev, with_embedding=False, m=m)
torch.cuda.synchronize()
sh = SnvShardedStore.from_coo(s.n, s.m, s.cell, s.snv, s.var, s.total, 0, 1, 0)
st = sh.store
info = st.info()
g = torch.Generator(device=dev)
g.manual_seed(1)
labs = [(torch.rand(s.n + 16, device=dev, generator=g) < 0.5).to(torch.uint8) for _ in range(4)]
lab_out = torch.empty(s.m + 16, dtype=torch.uint8, device=dev)
nobs_out = torch.empty(2 * s.m, dtype=torch.int32, device=dev)


flush_buf = torch.empty(2 * torch.cuda.get_device_properties(dev).L2_cache_size, dtype=torch.uint8, device=dev)


def bench(fn, graph, cold=False):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    if graph:
        gs = []
        for j in range(4):
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                fn(j)
            gs.append(g_)
        run = lambda i: gs[i % 4].replay()
    else:
        run = fn
    for i in range(5):
        run(i)
    torch.cuda.synchronize()
    if cold:  # L2 flushed before every step, each step timed on its own
        evs = []
        for i in range(reps):
            flush_buf.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            run(i)
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        return 1e3 * sum(a.elapsed_time(b) for a, b in evs) / reps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


plain = lambda i: st.assign_linear_dev(labs[i % 4], s.n / 2, lab_out, nobs_out, stream=torch.cuda.current_stream().cuda_stream)
push = lambda i: sh.assign_linear_dev(labs[i % 4], s.n / 2, lab_out, nobs_out)
for grp in groups:
    st.set_group(grp, grp)
    out = {"config": name, "parts": parts, "group": grp, "cells": s.n, "snvs": s.m, "nnz": int(s.nnz), "stream_bytes": info["stream_bytes_by_snv"]}
    for label, fn in (("plain", plain), ("push+collect", push)):
        for graph in (False, True):
            us = bench(fn, graph)
            out[f"{label}{'_graph' if graph else ''}_us"] = round(us, 2)
        out[f"{label}_graph_cold_us"] = round(bench(fn, True, cold=True), 2)
    out["ideal_us_at_4TBps"] = round(info["stream_bytes_by_snv"] / 4.0e12 * 1e6, 2)
    print(json.dumps(out), flush=True)


def timeline(fn, label):
    """-DPH_STAMPS builds: per-CTA timeline of one launch (ns after the first CTA's entry)."""
    import ctypes
    import numpy as np
    from phertilizer_b200 import _lib
    L = _lib.lib()
    if not hasattr(L, "ph_debug_stamps"):
        return
    buf = np.zeros(256 * 8, np.uint64)
    rows = []
    for i in range(6):
        torch.cuda.synchronize()
        L.ph_debug_stamps(None, 1)
        fn(i)
        torch.cuda.synchronize()
        L.ph_debug_stamps(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_ulonglong)), 0)
        t = buf.reshape(256, 8).astype(np.int64)
        t = t[t[:, 0] > 0]
        t0 = t[:, 0].min()
        q = lambda v: [int(np.min(v)), int(np.median(v)), int(np.max(v))]
        rows.append({"ctas": int(len(t)), "entry_ns": q(t[:, 0] - t0), "staged_ns": q(t[:, 1] - t0),
                     "stream_end_ns": q(t[:, 2] - t0), "exit_ns": q(t[:, 3] - t0), "batches_per_cta": q(t[:, 4]),
                     "warp_ns_per_item": {"prologue": round(float(t[:, 5].sum() / t[:, 4].sum()), 1),
                                          "stream": round(float(t[:, 6].sum() / t[:, 4].sum()), 1),
                                          "holders": round(float(t[:, 7].sum() / t[:, 4].sum()), 1)}})
    print(json.dumps({"timeline": label, "launches": rows[2:]}), flush=True)


timeline(plain, "plain")
timeline(push, "push+collect")
sh.close()
dist.destroy_process_group()
