"""Development aid: run the GPU-backed Phertilizer on a golden input and report the first hot-path call whose
arguments/results differ from the reference's recorded call sequence."""
import io, sys
from contextlib import redirect_stdout
import numpy as np
sys.path.insert(0, ".")
from tests.golden_util import load_golden
from tests.test_gpu_end_to_end import frames_from_golden
from phertilizer_b200.phertilizer import Phertilizer
from phertilizer_b200 import linear_split as LS, branching_split as BS
from scipy.spatial.distance import pdist, squareform

stem = sys.argv[1] if len(sys.argv) > 1 else "example"
starts, iters, seed = (3, 10, 100 * 99059) if stem == "example" else (2, 6, 4242)
g = load_golden(stem)
vc, bc = frames_from_golden(g)
ph = Phertilizer(vc, bc, mode="rd", dim_reduce=False)
ref = squareform(pdist(np.asarray(ph.data.read_depth, float)))
cd = ph.data.copy_distance.full()
print("copy_distance exact fraction", np.mean(cd == ref), "max rel", np.max(np.abs(cd - ref) / np.maximum(ref, 1e-300)))
pos = {}
first = []

def nxt(kind):
    recs = g.of(kind)
    i = pos.get(kind, 0)
    pos[kind] = i + 1
    return recs[i] if i < len(recs) else None

def cmp(kind, **got):
    r = nxt(kind)
    if r is None:
        print("EXTRA call", kind, pos[kind]); return
    for k, v in got.items():
        e = getattr(r, k)
        ok = (e is None and v is None) or (e is not None and v is not None and np.array_equal(np.asarray(e), np.asarray(v)))
        if not ok:
            msg = f"MISMATCH {kind}#{pos[kind]-1} field {k}: got len {None if v is None else len(v)} exp len {None if e is None else len(e)}"
            if v is not None and e is not None:
                msg += f" symdiff {len(np.setxor1d(np.asarray(v), np.asarray(e)))}"
            print(msg)
            first.append(msg)

o1 = LS.Linear_split.mut_assignment
def ls_mut(self, cellsA, muts):
    out = o1(self, cellsA, muts)
    cmp("linear_mut_assignment", cellsA=cellsA, muts=muts, mutsA=out[0], mutsB=out[1])
    return out
LS.Linear_split.mut_assignment = ls_mut
o2 = LS.Linear_split.create_affinity
def ls_aff(self, cells, mutsB=None):
    out = o2(self, cells, mutsB)
    if mutsB is not None:
        cmp("linear_affinity_features", cells=cells, mutsB=mutsB)
    return out
LS.Linear_split.create_affinity = ls_aff
o3 = BS.Branching_split.mut_assignment
def bs_mut(self, cellsA, cellsB):
    out = o3(self, cellsA, cellsB)
    cmp("branching_mut_assignment", cellsA=cellsA, cellsB=cellsB, mutsA=out[0], mutsB=out[1], mutsC=out[2])
    return out
BS.Branching_split.mut_assignment = bs_mut
o4 = LS.Linear_split.random_weighted_init_array
def ls_init(self, cells, muts, p):
    out = o4(self, cells, muts, p)
    cmp("weighted_init_vaf", cells=cells, muts=muts)
    return out
LS.Linear_split.random_weighted_init_array = ls_init
o5 = LS.Linear_split.compute_norm_likelihood
def ls_norm(self, cellsA, cellsB, mutsA, mutsB):
    out = o5(self, cellsA, cellsB, mutsA, mutsB)
    r = nxt("linear_norm_likelihood")
    if r is not None and (not np.array_equal(r.cellsA, cellsA) or abs(out - r.value) > 1e-9 * abs(r.value)):
        print("MISMATCH linear_norm", out, r.value, len(cellsA), len(r.cellsA))
    return out
LS.Linear_split.compute_norm_likelihood = ls_norm
o6 = LS.Linear_split.check_metrics
def ls_met(self, ca, cb, ma, mb):
    out = o6(self, ca, cb, ma, mb)
    r = nxt("linear_check_metrics")
    if r is not None:
        print("metrics", out, (r.f1, r.f2, r.f3))
    return out
LS.Linear_split.check_metrics = ls_met

buf = io.StringIO()
with redirect_stdout(buf):
    tree, tl, ll = ph.phertilize(max_iterations=iters, starts=starts, seed=seed, radius=1, gamma=0.95, d=10, use_copy_kernel=True,
                                 post_process=True, low_cmb=0.05, high_cmb=0.15, nobs_per_cluster=3)
print("loglikes", ll.to_numpy(), "expected", g.of("final")[0].loglikes)
print("calls", pos, "golden", {k: len(g.of(k)) for k in pos})
print("first mismatches:", first[:5])
