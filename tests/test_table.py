"""Host-resolved pair table (psa_build_pair_table, psa_table.cpp) against the oracle's per-pair
primitives -- the restatement of cuda_funcs.cu:310-452 -- and against the golden matrices."""
import itertools
import math
import random

import pytest

from conftest import load_golden

ALPHA = [chr(65 + i) for i in range(26)] + ["-"]
WSETS = [[1, 3, 4, 2], [2, 1.5, 1.1, 1.3], [1.5, 2.6, 0.1, 0.2], [0.8, 0.54, 2.6, 13.7], [1, 1, 1, 1], [5, 4, 3, 2],
         [0, 0, 0, 0], [-1, 2, -3, 0.5], [0.25, 8, 0.5, 4], [1e6, 3, 1e-3, 2], [3, 3, 1, 1], [1, 1, 3, 3]]


def test_sign_matrix(psa):
    g = load_golden("sign_matrix.json")
    t = psa.build_pair_table([1, 1, 1, 1], True)
    for a in range(27):
        for b in range(27):
            assert t.sign[b][a] == g["rows"][a][b]        # table rows are the Seq2 symbol
            assert t.sign[a][b] == t.sign[b][a]


@pytest.mark.parametrize("w", WSETS)
@pytest.mark.parametrize("is_max", [0, 1])
def test_table_matches_oracle_primitives(psa, port, w, is_max):
    t = psa.build_pair_table(w, is_max, 500)
    diffs = set()
    for i2, c2 in enumerate(ALPHA):
        for i1, c1 in enumerate(ALPHA):
            sub = port.substitute(c1, c2, w, is_max)
            assert t.substitute[i2][i1] == sub and sub != ""
            d = port.weight(port.sign(c1, sub), w) - port.weight(port.sign(c1, c2), w)
            assert t.diff[i2][i1] == d
            diffs.add(d)
    order = sorted(diffs, reverse=not is_max)       # rank 1 = worst ... nranks = best
    assert t.nranks == len(order) <= 10
    for i2 in range(27):
        for i1 in range(27):
            assert t.rank[i2][i1] == order.index(t.diff[i2][i1]) + 1


def test_substitutes_golden(psa):
    g = load_golden("substitutes.json")
    for tb in g["tables"]:
        t = psa.build_pair_table(tb["weights"], tb["is_max"])
        for i1 in range(27):
            assert "".join(t.substitute[i2][i1] for i2 in range(27)) == tb["rows"][i1]


def test_exactness_analysis(psa):
    t = psa.build_pair_table([1, 3, 4, 2], False, 5000)
    assert t.exact and t.frac_bits == 0 and t.key_slack == 0
    t = psa.build_pair_table([1.5, 0.25, 4, 2], True, 5000)
    assert t.exact and t.frac_bits == 2
    t = psa.build_pair_table([1.5, 2.6, 0.1, 0.2], True, 500)
    assert not t.exact and t.key_slack > 0
    # huge integer weights stop being exactly summable
    assert psa.build_pair_table([2.0 ** 40, 1, 1, 1], True, 100).exact
    assert not psa.build_pair_table([2.0 ** 50, 1, 1, 1], True, 100).exact
    for bad in (float("nan"), float("inf")):
        with pytest.raises(psa.PsaError) as e:
            psa.build_pair_table([1, bad, 1, 1], True)
        assert e.value.status == psa.PSA_ERR_WEIGHTS


def test_key_slack_covers_double_rounding(psa, port):
    """For non-dyadic weights the fixed-point key and the reference's sequential double may order two
    offsets differently only within key_slack (psa_table.cpp); check the bound on real data."""
    rng = random.Random(3)
    for w in ([1.5, 2.6, 0.1, 0.2], [0.8, 0.54, 2.6, 13.7], [2, 1.5, 1.1, 1.3]):
        for is_max in (0, 1):
            n2 = 300
            t = psa.build_pair_table(w, is_max, n2)
            s1 = "".join(rng.choice(ALPHA[:26]) for _ in range(900))
            s2 = "".join(rng.choice(ALPHA[:26]) for _ in range(n2))
            scores = port.scores(w, is_max, s1, s2)
            scale = 2.0 ** t.frac_bits
            fixed = [round(x * scale) for x in (w[0], -w[1], -w[2], -w[3])]
            for off in range(0, 600, 7):
                r = port.offset_naive(w, is_max, s1, s2, off)
                i1, i2 = ALPHA.index(s1[off + r.char_offset]), ALPHA.index(s2[r.char_offset])
                after = "*:._".index(port.sign(s1[off + r.char_offset], r.ch))
                before = "*:._".index(t.sign[i2][i1])
                key = sum(n * f for n, f in zip(r.counts, fixed)) + fixed[after] - fixed[before]
                assert abs(key - scores[off] * scale) * 2 <= t.key_slack
