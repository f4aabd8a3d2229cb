"""Quick device timing probe of the sparse passes for several group sizes (development aid, not the bench)."""
import sys, time, json
import torch
sys.path.insert(0, ".")
import os
if os.environ.get("PH_LIB"):   # development aid: time an alternative build of the library (e.g. another ring depth)
    from phertilizer_b200 import _lib as _l
    _l.LIB_PATH = os.path.abspath(os.environ["PH_LIB"])
from phertilizer_b200.synth import make_config
from phertilizer_b200.store import ObservationStore

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
GS = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 32]
REPS = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device("cuda:0")
t = time.time()
over = {}
if len(sys.argv) > 4:
    over["n"] = int(sys.argv[4])   # e.g. C3 with n = 12500: the shard one rank of an 8-GPU run walks
s = make_config(name, device=dev, with_embedding=False, **over)
torch.cuda.synchronize()
print("synth", name, s.nnz, "nnz", round(time.time() - t, 2), "s", flush=True)
t = time.time()
st = ObservationStore.from_coo(s.n, s.m, s.cell, s.snv, s.var, s.total)
torch.cuda.synchronize()
print("build", round(time.time() - t, 2), "s", st.info(), flush=True)
g = torch.Generator(device=dev); g.manual_seed(1)
cl = (torch.rand(s.n + 16, device=dev, generator=g) < 0.5).to(torch.uint8) + 0
cl2 = torch.randint(0, 3, (s.n + 16,), device=dev, generator=g, dtype=torch.uint8)
sl = torch.randint(0, 3, (s.m + 16,), device=dev, generator=g, dtype=torch.uint8)
lab_out = torch.empty(s.m + 16, dtype=torch.uint8, device=dev)
nobs_out = torch.empty(2 * s.m, dtype=torch.int32, device=dev)
S0 = torch.empty(max(s.n, s.m) * 3, dtype=torch.float64, device=dev); S1 = torch.empty_like(S0)
No = torch.empty(max(s.n, s.m) * 3, dtype=torch.int32, device=dev); Nv = torch.empty_like(No)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, reps=REPS):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

stream = None
for G in GS:
    st.set_group(G, G)
    r = {}
    r["lin"] = timeit(lambda: st.assign_linear_dev(cl, s.n / 2, lab_out, nobs_out))
    r["br"] = timeit(lambda: st.assign_branching_dev(cl2, lab_out, nobs_out))
    r["tab2"] = timeit(lambda: st.snv_table_dev(cl2, 2, S0, S1, No, Nv))
    r["cell2"] = timeit(lambda: st.cell_table_dev(sl, 2, S0, S1, No, Nv))
    out = {k: (round(v, 4), round(s.nnz / v / 1e6, 1)) for k, v in r.items()}
    inf = st.info()
    print("batch", G, "ms, Gnnz/s:", out, " GB/s physical (lin):", round(inf["stream_bytes_by_snv"] / r["lin"] / 1e6, 1),
          " (cell2):", round(inf["stream_bytes_by_cell"] / r["cell2"] / 1e6, 1), flush=True)
