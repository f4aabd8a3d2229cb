"""Round-2 device probe (development aid): pass times on one GPU for several padded-round thresholds (PH_ROUND_FILL).

  python scripts/probe_r2.py [C3] [fills, e.g. 1.01,0.9,0.75,0.62] [reps]"""
import os, sys, time
import torch
sys.path.insert(0, ".")
if os.environ.get("PH_LIB"):   # development aid: time an alternative build of the library (e.g. another ring depth)
    from phertilizer_b200 import _lib as _l
    _l.LIB_PATH = os.path.abspath(os.environ["PH_LIB"])
from phertilizer_b200.synth import make_config
from phertilizer_b200.store import ObservationStore

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
fills = sys.argv[2].split(",") if len(sys.argv) > 2 else ["1.01", "0.9", "0.75", "0.62"]
REPS = int(sys.argv[3]) if len(sys.argv) > 3 else 15
dev = torch.device("cuda:0")
s = make_config(name, device=dev, with_embedding=False)
torch.cuda.synchronize()
print("synth", name, s.nnz, "nnz", flush=True)
g = torch.Generator(device=dev); g.manual_seed(1)
cls = [(torch.rand(s.n + 16, device=dev, generator=g) < 0.5).to(torch.uint8) for _ in range(3)]
cl2 = [torch.randint(0, 3, (s.n + 16,), device=dev, generator=g, dtype=torch.uint8) for _ in range(3)]
sl = [torch.randint(0, 3, (s.m + 16,), device=dev, generator=g, dtype=torch.uint8) for _ in range(3)]
lab_out = torch.empty(s.m + 16, dtype=torch.uint8, device=dev)
nobs_out = torch.empty(2 * s.m, dtype=torch.int32, device=dev)
big = max(s.n, s.m)
S0 = torch.empty(big * 3, dtype=torch.float64, device=dev); S1 = torch.empty_like(S0)
No = torch.empty(big * 3, dtype=torch.int32, device=dev); Nv = torch.empty_like(No)


def timeit(fn, reps=REPS):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ref = None
for f in fills:
    os.environ["PH_ROUND_FILL"] = f
    t = time.time()
    st = ObservationStore.from_coo(s.n, s.m, s.cell, s.snv, s.var, s.total)
    torch.cuda.synchronize()
    inf = st.info()
    r = {}
    r["lin"] = timeit(lambda i: st.assign_linear_dev(cls[i % 3], s.n / 2, lab_out, nobs_out))
    out = (lab_out[: s.m].clone(), nobs_out[: s.m].clone())
    if ref is None:
        ref = out
    same = bool(torch.equal(ref[0], out[0]) and torch.equal(ref[1], out[1]))
    r["br"] = timeit(lambda i: st.assign_branching_dev(cl2[i % 3], lab_out, nobs_out))
    r["tab2"] = timeit(lambda i: st.snv_table_dev(cl2[i % 3], 2, S0, S1, No, Nv))
    r["cell2"] = timeit(lambda i: st.cell_table_dev(sl[i % 3], 2, S0, S1, No, Nv))
    print(f"fill {f}: build {time.time() - t:.1f}s  stream4 snv {inf['stream4_bytes_by_snv'] / 1e6:.1f} MB cell {inf['stream4_bytes_by_cell'] / 1e6:.1f} MB "
          f"(flat {inf['stream_bytes_by_snv'] / 1e6:.1f} / {inf['stream_bytes_by_cell'] / 1e6:.1f})  "
          + "  ".join(f"{k} {v * 1e3:.1f} us" for k, v in r.items())
          + f"  lin GB/s {inf['stream4_bytes_by_snv'] / r['lin'] / 1e6:.0f}  cell2 GB/s {(inf['stream4_bytes_by_cell'] or inf['stream_bytes_by_cell']) / r['cell2'] / 1e6:.0f}  same_as_first {same}", flush=True)
    st.close()
