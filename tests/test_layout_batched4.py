"""The arithmetic of the batched-4 stream (csrc/ph_build.cu k_round_keys / k_scatter, csrc/ph_pass.cu
LAY = 3), restated in numpy: where the r-th entry (round order, rotated by the line's slot) of a sub-segment goes, and that
a full round makes the 32 lanes of a warp step gather from 32 different shared-memory banks.  CPU only."""
import numpy as np


def rotated_round_order(banks: np.ndarray, rot: int) -> np.ndarray:
    """Entry order of one sub-segment: round k = the k-th entry of every bank, banks visited from `rot` (k_round_keys)."""
    kb = (banks - rot) % 32
    o = np.argsort(kb, kind="stable")
    first = np.searchsorted(kb[o], kb[o])
    rank = np.arange(len(o)) - first
    return o[np.lexsort((kb[o], rank))]


def place4(r: int, slot: int):
    """(warp step, lane, entry slot) of the r-th entry of the line in batch slot `slot` (k_scatter)."""
    return r >> 5, slot * 4 + ((r >> 3) & 3), r & 7


def test_full_rounds_are_conflict_free():
    rng = np.random.default_rng(0)
    R = 5                                            # every bank holds R entries of every line: all rounds are full
    steps = {}
    for slot in range(8):
        banks = np.repeat(np.arange(32), R)
        rng.shuffle(banks)
        order = rotated_round_order(banks, slot)
        assert sorted(order.tolist()) == list(range(len(banks)))
        for r, src in enumerate(order):
            t, lane, e = place4(r, slot)
            assert t < R and lane // 4 == slot
            steps.setdefault((t, e), {})[lane] = int(banks[src])
    assert len(steps) == R * 8
    for (t, e), lanes in steps.items():
        assert len(lanes) == 32, "every lane reads one entry per (step, slot)"
        assert len(set(lanes.values())) == 32, f"bank conflict at step {t} slot {e}"


def test_placement_is_a_bijection_onto_the_line_stream():
    # r -> (step, lane, slot) is the identity on the line's own stream: group 4 t + c, entry e  <->  r
    for slot in (0, 3, 7):
        seen = set()
        for r in range(32 * 7 + 13):
            t, lane, e = place4(r, slot)
            c = lane - slot * 4
            assert 0 <= c < 4 and 0 <= e < 8
            assert (4 * t + c) * 8 + e == r
            seen.add((t, lane, e))
        assert len(seen) == 32 * 7 + 13


def test_ragged_rounds_keep_every_entry():
    rng = np.random.default_rng(1)
    banks = rng.integers(0, 32, 777)
    for rot in (0, 5):
        order = rotated_round_order(banks, rot)
        assert sorted(order.tolist()) == list(range(777))
        # within a round the (rotated) banks increase strictly: one entry per bank and round
        kb = (banks[order] - rot) % 32
        starts = np.nonzero(np.diff(kb) <= 0)[0] + 1
        for seg in np.split(kb, starts):
            assert np.all(np.diff(seg) > 0)


# ---- padded rounds (csrc/ph_build.cu: k_round_plan, k_fill_rounds, k_scatter) restated in numpy
PAD = 128   # PH_PAD: excluded bytes at the start of a chunk of the label image = one 4-byte word per bank


def round_plan(banks: np.ndarray, min_fill: int):
    """K = rounds in which >= min_fill banks still hold an entry, R = entries in them (k_round_plan)."""
    cnt = np.bincount(banks, minlength=32)
    K = R = 0
    if len(banks) >= min_fill:
        while True:
            nk = int((cnt > K).sum())
            if nk < min_fill:
                break
            R += nk
            K += 1
    return K, R


def padded_positions(idx16: np.ndarray, rot: int, min_fill: int):
    """(position of every entry, padded length, positions and values of the padding entries of the padded rounds)."""
    banks = (idx16 >> 2) & 31
    K, R = round_plan(banks, min_fill)
    order = rotated_round_order(banks, rot)              # entry ids in round order
    kb = (banks - rot) % 32
    first = {}
    rank_in_bank = np.zeros(len(idx16), int)
    for e in np.argsort(idx16, kind="stable"):           # k_round_keys: rank among the entries of the same bank
        rank_in_bank[e] = first.get(int(banks[e]), 0)
        first[int(banks[e])] = rank_in_bank[e] + 1
    pos = np.zeros(len(idx16), int)
    for r, e in enumerate(order):
        k, b = rank_in_bank[e], kb[e]
        pos[e] = 32 * k + b if k < K else 32 * K + (r - R)
    holes = {32 * k + b: 4 * ((b + rot) % 32) for k in range(K) for b in range(32)}
    for p in pos:
        holes.pop(int(p), None)
    return pos, len(idx16) + 32 * K - R, holes


def place_flat(r: int, cnt: int):
    """csrc/ph_build.cu place(): position r of a padded sub-segment of cnt entries -> (group, slot) of the flat stream."""
    ng = (cnt + 7) // 8
    tf = ng >> 5
    if r < 256 * tf:
        t, j, ln = r // 256, (r % 256) >> 5, r & 31
        return (t << 5) + ln, j
    r2, gl = r - 256 * tf, ng - (tf << 5)
    return (tf << 5) + r2 % gl, r2 // gl


def test_padded_rounds_place_every_entry_once_and_pad_in_the_right_bank():
    rng = np.random.default_rng(3)
    for n_entries, min_fill, rot in ((500, 24, 0), (500, 24, 5), (37, 24, 3), (2000, 16, 7), (640, 29, 1), (10, 24, 0)):
        idx16 = PAD + np.sort(rng.choice(65408, n_entries, replace=False))
        pos, cnt, holes = padded_positions(idx16, rot, min_fill)
        assert len(set(pos.tolist())) == n_entries and pos.max() < cnt
        assert not (set(pos.tolist()) & set(holes))
        K, R = round_plan((idx16 >> 2) & 31, min_fill)
        assert cnt == n_entries + 32 * K - R and len(holes) == 32 * K - R
        # padded rounds: position 32 k + b holds an entry (or padding) whose label byte sits in bank (b + rot) % 32
        val = {int(p): int(v) for p, v in zip(pos, idx16)}
        val.update(holes)
        for p in range(32 * K):
            assert (val[p] >> 2) & 31 == ((p & 31) + rot) % 32
        assert all(v < PAD for v in holes.values())
        # flat stream: inside the full blocks of 32 groups the 32 positions of a padded round are the 32 lanes (groups)
        # of ONE gather, same slot (the ragged last block of a sub-segment has fewer than 32 lanes: no such guarantee)
        full_blocks = ((cnt + 7) // 8) >> 5
        for k in range(min(K, 8 * full_blocks)):
            gs = [place_flat(32 * k + b, cnt) for b in range(32)]
            assert len({g for g, _ in gs}) == 32 and len({e for _, e in gs}) == 1
        seen = {place_flat(r, cnt) for r in range(cnt)}
        assert len(seen) == cnt and max(g for g, _ in seen) < (cnt + 7) // 8
        # batched-4 stream: a padded round is one warp step; with 8 lines rotated by their slot the step is conflict free
        if K:
            lanes = {}
            for slot in range(8):
                p8, _, h8 = padded_positions(idx16, slot, min_fill)
                v8 = {int(p): int(v) for p, v in zip(p8, idx16)}
                v8.update(h8)
                for b in range(32):
                    t, lane, e = place4(b, slot)          # round 0
                    lanes.setdefault(e, {})[lane] = (v8[b] >> 2) & 31
            for e, m in lanes.items():
                assert len(m) == 32 and len(set(m.values())) == 32


def test_padded_rounds_cut_the_wavefronts_of_a_ragged_sub_segment():
    """Gather cost model (wavefronts per 32-lane request = largest number of lanes on one bank) on C3-like sub-segments."""
    rng = np.random.default_rng(4)

    def wavefronts(min_fill):
        tot = req = 0
        padded = plain = 0
        for _ in range(40):
            idx16 = PAD + np.sort(rng.choice(65408, int(rng.integers(380, 620)), replace=False))
            pos, cnt, holes = padded_positions(idx16, 0, min_fill)
            val = np.zeros(((cnt + 31) // 32) * 32, int)      # trailing padding: idx16 = 0 (bank 0)
            val[pos] = idx16
            for p, v in holes.items():
                val[p] = v
            words = val >> 2                                   # same word = broadcast, no conflict
            for row in words.reshape(-1, 32):
                worst = max(len(set(row[(row & 31) == b].tolist())) for b in range(32))
                tot += max(worst, 1)
                req += 1
            padded += cnt
            plain += len(idx16)
        return tot / req, padded / plain

    w_off, _ = wavefronts(33)
    w_on, grow = wavefronts(24)
    assert w_on < 0.9 * w_off and grow < 1.06, (w_off, w_on, grow)
