"""Product-side likelihood table (phertilizer_b200/like_lut.py) vs the reference's numba functions (golden KATs)."""
import os

import numpy as np

from phertilizer_b200.like_lut import encode_pairs, like_tables
from tests.golden_util import GOLDEN_DIR, load_golden


def test_kat_bit_exact():
    k = np.load(os.path.join(GOLDEN_DIR, "kat_like.npz"))
    for s, (alpha, mc) in enumerate(k["settings"]):
        l0, l1 = like_tables(k["var"], k["total"], float(alpha), int(mc))
        assert np.array_equal(l0, k[f"like0_{s}"])
        assert np.array_equal(l1, k[f"like1_{s}"])


def test_survey_known_answers():
    # SURVEY.md section 8c table (alpha=0.001, copies 1..5)
    kat = {(0, 1): (-0.0010005003335835344, -1.0969470093490485), (1, 1): (-6.907755278982137, -0.4062987888567418),
           (2, 2): (-13.815510557964274, -0.6454732969698538), (2, 5): (-11.51592696597098, -1.8730342991022468),
           (7, 20): (-37.1090022106469, -3.09171276712331)}
    ks = np.array([k for k, _ in kat]); ns = np.array([n for _, n in kat])
    l0, l1 = like_tables(ks, ns)
    for i, key in enumerate(kat):
        assert l0[i] == kat[key][0] and l1[i] == kat[key][1]


def test_encode_pairs_example():
    g = load_golden("example")
    pv, pt, cnt, code = encode_pairs(g.inputs["var"], g.inputs["total"])
    # pair histogram of the example
    assert list(zip(pv.tolist(), pt.tolist(), cnt.tolist()))[:3] == [(0, 1, 15910), (1, 1, 7530), (0, 2, 89)]
    assert cnt.sum() == 23612 and np.all(np.diff(cnt) <= 0)
    l0, l1 = like_tables(pv, pt)
    assert np.array_equal(l0[code], g.inputs["like0"]) and np.array_equal(l1[code], g.inputs["like1"])
