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
