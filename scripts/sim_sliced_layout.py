"""CPU simulation of a candidate store layout (design aid for round 2, not product code).

Today a warp streams 8 consecutive lines with lanes striding over the groups, so every lane crosses every sub-segment
boundary.  Candidate: slices of 32 consecutive lines interleaved at 16-byte granularity, lane L streams line L, every
(slice, sub-segment) padded to the slice's longest so that boundaries are warp-uniform.  Two unknowns are measured here
on C3-shaped synthetic lines: (1) the padding that costs in HBM bytes, (2) shared-memory wavefronts per LDS.U8 request
when 32 lanes gather labels of 32 different lines, with the round order rotated by the lane index.

  python scripts/sim_sliced_layout.py [n_slices]
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from phertilizer_b200.like_lut import encode_pairs
from phertilizer_b200.synth import CONFIGS, make_config

n_slices = int(sys.argv[1]) if len(sys.argv) > 1 else 100
CH, GE = 65520, 8
cfg = CONFIGS["C3"]
s = make_config("C3", with_embedding=False, m=32 * n_slices)
cell, snv, var, total = (x.numpy() for x in (s.cell, s.snv, s.var, s.total))
pv, pt, cnt, code = encode_pairs(var, total)
D = 2                                   # dense codes of C3's by_snv orientation
order = np.lexsort((cell, snv))
cell, snv, code = cell[order], snv[order], code[order]
starts = np.searchsorted(snv, np.arange(s.m + 1))
NCH = (s.n + CH - 1) // CH


def round_order(banks, rot):
    """positions of a sub-segment's entries in round order: round k = k-th entry of every bank, banks visited from `rot`."""
    key_bank = (banks - rot) % 32
    o = np.argsort(key_bank, kind="stable")
    kb = key_bank[o]
    first = np.searchsorted(kb, kb)          # rank within the bank
    rank = np.arange(len(kb)) - first
    return o[np.lexsort((kb, rank))]


def wavefronts(bank_matrix):
    """bank_matrix [steps, 32] (-1 = padding entry, hits the shared 'excluded' word: broadcast) -> wavefronts per request."""
    w = np.zeros(len(bank_matrix))
    for i, row in enumerate(bank_matrix):
        r = row[row >= 0]
        w[i] = max(np.bincount(r, minlength=32).max() if len(r) else 0, 1)
    return w


# variant: lines permuted so that a slice holds lines of similar shape (sorted by their sub-segment sizes, the rarest
# sub-segments most significant); the kernel's per-line results go back through the permutation
sizes_all = np.zeros((s.m, D * NCH), int)
for j in range(s.m):
    c, k = cell[starts[j]:starts[j + 1]], code[starts[j]:starts[j + 1]]
    for d in range(D):
        for ch in range(NCH):
            sizes_all[j, d * NCH + ch] = np.count_nonzero((k == d) & (c // CH == ch))
g_all = (sizes_all + GE - 1) // GE
perm = np.lexsort(tuple(g_all[:, q] for q in range(D * NCH)))   # last key (rarest sub-segment) most significant
g_sorted = g_all[perm].reshape(n_slices, 32, D * NCH)
groups_sorted = int((32 * g_sorted.max(axis=1)).sum())
# variant 2: sort by the line's TOTAL groups only, sub-segment boundaries per lane (divergent), slice padded to its longest line
tot = g_all.sum(axis=1)
groups_total_sorted = int((32 * np.sort(tot).reshape(n_slices, 32).max(axis=1)).sum())
groups_total_unsorted = int((32 * tot.reshape(n_slices, 32).max(axis=1)).sum())

groups_now = groups_sliced = 0
wf_now, wf_sliced, wf_sliced_norot = [], [], []
for sl in range(n_slices):
    per_line = []
    for L in range(32):
        j = sl * 32 + L
        c, k = cell[starts[j]:starts[j + 1]], code[starts[j]:starts[j + 1]]
        subs = []
        for d in range(D):
            for ch in range(NCH):
                sel = c[(k == d) & (c // CH == ch)]
                bank = ((16 + sel - ch * CH) // 4) % 32
                subs.append(bank)
        per_line.append(subs)
    for q in range(D * NCH):
        sizes = np.array([len(per_line[L][q]) for L in range(32)])
        g_now = (sizes + GE - 1) // GE
        groups_now += g_now.sum()
        groups_sliced += 32 * g_now.max()
        if sl < 12:  # bank behaviour: a few slices are enough
            # today: one line, 32 consecutive entries of its round-ordered sequence per request
            for L in range(0, 32, 8):
                b = per_line[L][q]
                seq = b[round_order(b, 0)]
                pad = (-len(seq)) % 32
                m_ = np.concatenate([seq, -np.ones(pad, int)]).reshape(-1, 32)
                wf_now.append(wavefronts(m_))
            # sliced: lane L is at the same position of ITS line
            T = g_now.max() * GE
            for rot, acc in ((True, wf_sliced), (False, wf_sliced_norot)):
                M = -np.ones((T, 32), int)
                for L in range(32):
                    b = per_line[L][q]
                    M[: len(b), L] = b[round_order(b, L if rot else 0)]
                acc.append(wavefronts(M))
print(f"lines {32 * n_slices}, observations {len(cell)}")
print(f"groups today (8-entry padding)       : {groups_now}  ({groups_now * GE / len(cell):.4f} entries per dense+tail observation)")
print(f"groups sliced (uniform sub-segments) : {groups_sliced}  (+{100 * (groups_sliced / groups_now - 1):.1f} % bytes)")
print(f"  ... lines sorted by sub-segment sizes  : {groups_sorted}  (+{100 * (groups_sorted / groups_now - 1):.1f} % bytes)")
print(f"  ... per-lane boundaries, slice padded to its longest line: unsorted +{100 * (groups_total_unsorted / groups_now - 1):.1f} %, "
      f"lines sorted by length +{100 * (groups_total_sorted / groups_now - 1):.1f} %")
print(f"LDS.U8 wavefronts per request  today: {np.concatenate(wf_now).mean():.2f}   sliced, rotated rounds: "
      f"{np.concatenate(wf_sliced).mean():.2f}   sliced, unrotated: {np.concatenate(wf_sliced_norot).mean():.2f}")

# ---- variant B: batches of 8 shape-sorted lines, 4 lanes per line, every sub-segment padded to the batch's longest in
# units of 4 groups (one warp step = 8 lines x 4 lanes); keeps today's task granularity (8 lines per warp batch)
g_sorted8 = g_all[perm][: (s.m // 8) * 8].reshape(-1, 8, D * NCH)
steps = (g_sorted8.max(axis=1) + 3) // 4                      # warp steps per (batch, sub-segment)
groups_b = int(32 * steps.sum())
g_unsorted8 = g_all[: (s.m // 8) * 8].reshape(-1, 8, D * NCH)
groups_b_unsorted = int(32 * ((g_unsorted8.max(axis=1) + 3) // 4).sum())
print(f"variant B (8 lines x 4 lanes, uniform sub-segments): sorted +{100 * (groups_b / groups_now - 1):.1f} % bytes, "
      f"consecutive lines +{100 * (groups_b_unsorted / groups_now - 1):.1f} %")
# bank behaviour of variant B: lane (a, c) reads entry (4 t + c) * 8 + e of line a, rounds rotated by a
wf_b = []
for b in range(min(24, len(g_sorted8))):
    lines = perm[b * 8:(b + 1) * 8]
    for q in range(D * NCH):
        d, ch = divmod(q, NCH)
        seqs = []
        for a, j in enumerate(lines):
            c, k = cell[starts[j]:starts[j + 1]], code[starts[j]:starts[j + 1]]
            sel = c[(k == d) & (c // CH == ch)]
            bank = ((16 + sel - ch * CH) // 4) % 32
            seqs.append(bank[round_order(bank, a)])
        T = int(steps[b, q])
        M = -np.ones((T * 8, 32), int)                      # row = (t, e), column = lane a * 4 + c
        for a in range(8):
            sq = seqs[a]
            idx = np.arange(len(sq))
            grp, e = idx // 8, idx % 8
            t, c4 = grp // 4, grp % 4
            M[t * 8 + e, a * 4 + c4] = sq
        wf_b.append(wavefronts(M))
print(f"variant B LDS.U8 wavefronts per request: {np.concatenate(wf_b).mean():.2f}")

# ---- variant B with "fuller rounds": a round in which at least `thr` of the 32 banks still have an entry is padded to all
# 32 slots (missing banks get padding entries), so the entry index keeps its phase with the bank; later rounds stay ragged
def padded_round_order(banks, rot, thr):
    kb = (banks - rot) % 32
    cnt = np.bincount(kb, minlength=32)
    seq = []
    k = 0
    while True:
        present = np.nonzero(cnt > k)[0]
        if len(present) == 0:
            break
        if len(present) >= thr * 32:
            row = np.full(32, -1)
            row[present] = (present + rot) % 32
            seq.extend(row.tolist())
        else:
            seq.extend(((present + rot) % 32).tolist())
        k += 1
    return np.array(seq, int)


for thr in (1.01, 0.9, 0.75, 0.5):
    tot_entries = 0
    wf = []
    for b in range(min(24, len(g_sorted8))):
        lines = perm[b * 8:(b + 1) * 8]
        for q in range(D * NCH):
            d, ch = divmod(q, NCH)
            seqs = []
            for a, j in enumerate(lines):
                c, k = cell[starts[j]:starts[j + 1]], code[starts[j]:starts[j + 1]]
                sel = c[(k == d) & (c // CH == ch)]
                bank = ((16 + sel - ch * CH) // 4) % 32
                seqs.append(padded_round_order(bank, a, thr))
            T = max((len(sq) + 31) // 32 for sq in seqs)
            tot_entries += T * 32 * 8
            M = -np.ones((T * 8, 32), int)
            for a in range(8):
                sq = seqs[a]
                idx = np.arange(len(sq))
                grp, e = idx // 8, idx % 8
                t, c4 = grp // 4, grp % 4
                M[t * 8 + e, a * 4 + c4] = sq
            wf.append(wavefronts(M))
    base = int(32 * steps[: min(24, len(g_sorted8))].sum()) * GE
    print(f"variant B, rounds with >= {thr:.2f} occupancy padded: {np.concatenate(wf).mean():.2f} wavefronts per request, "
          f"{100 * (tot_entries / base - 1):+.1f} % bytes vs unpadded variant B")
