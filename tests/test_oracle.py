"""The oracle (oracle/psa_oracle.c) against everything the reference pins, and against the
reference itself (oracle/_ref, compiled from /root/reference) where that library is present."""
import random

import pytest

from conftest import load_golden, same_answer

ALPHA = [chr(65 + i) for i in range(26)] + ["-"]


def test_reference_output_txt(port, input_blocks):
    """input.txt block 1 -> output.txt: the only end-to-end answer the reference ships."""
    b = input_blocks[0]
    r = port.search(b["weights"], b["goal"] == "maximum", b["seq1"], b["seq2"])
    assert (r.offset, "%g" % r.score) == (4505, "-4879")
    assert r.mutant(b["seq2"]) + "\n%d %g" % (r.offset, r.score) == b["output_txt"]
    diffs = [i for i, (x, y) in enumerate(zip(b["seq2"], r.mutant(b["seq2"]))) if x != y]
    assert diffs == [144] and b["seq2"][144] == "R" and r.ch == "E"


def test_input_blocks(port, input_blocks):
    for b in input_blocks:
        r = port.search(b["weights"], b["goal"] == "maximum", b["seq1"], b["seq2"])
        assert same_answer(r, b["expect"]), b["weights_text"]
        assert "%g" % r.score == b["expect"]["score_g"]


def test_input_blocks_threads_and_naive(port, input_blocks):
    for b in input_blocks[1:4] + input_blocks[7:]:
        is_max = b["goal"] == "maximum"
        r1 = port.search(b["weights"], is_max, b["seq1"], b["seq2"], nthreads=1)
        r5 = port.search(b["weights"], is_max, b["seq1"], b["seq2"], nthreads=5)
        assert same_answer(r5, r1)
        n = port.offset_naive(b["weights"], is_max, b["seq1"], b["seq2"], r1.offset)
        assert (n.char_offset, n.ch, n.score, n.counts) == (r1.char_offset, r1.ch, r1.score, r1.counts)


def test_readme_pair_example(port):
    """README.md:29-34"""
    a, b = "PSEKHLQCLLQRHKGK", "HSKSHLQHLLQRHKSQ"
    assert "".join(port.sign(x, y) for x, y in zip(a, b)) == "_*:.***_******.:"


def _plain_score(port, w, s1, s2, off):
    signs = "".join(port.sign(s1[off + i], s2[i]) for i in range(len(s2)))
    return signs, sum(port.weight(c, w) for c in signs)


def test_readme_worked_examples(port):
    """README.md:51-59 and 62-70: plain alignment scores (no mutation)."""
    s1 = "LQRHKRTHTGEKPYEPSHLQYHERTHTGEKPYECHQCHQAFKKCSLLQRHKRTH"
    s2 = "HERTHTGEKPYECHQCRTAFKKCSLLQRHK"
    signs, score = _plain_score(port, [1.5, 2.6, 0.3, 0.2], s1, s2, 21)
    assert signs.replace("_", " ") == "****************: ************"
    assert abs(score - 39.2) < 1e-9
    s1b, s2b = "ELMVRTNMYTONEWVFNVJERVMKLWEMVKL", "MSKDVMSDLKWEV"
    signs, score = _plain_score(port, [5, 4, 3, 2], s1b, s2b, 3)
    assert signs.replace("_", " ") == ": .:: :  :* ."
    assert score == -31


def test_sign_matrix_golden(port):
    g = load_golden("sign_matrix.json")
    assert g["alphabet"] == "".join(ALPHA)
    for a, row in zip(ALPHA, g["rows"]):
        assert "".join(port.sign(a, b) for b in ALPHA) == row
    assert port.sign("a", "A") == "" and port.sign("A", "[") == "" and port.sign("@", "A") == ""


def test_substitutes_golden(port):
    g = load_golden("substitutes.json")
    for t in g["tables"]:
        for c1, row in zip(ALPHA, t["rows"]):
            got = "".join(port.substitute(c1, c2, t["weights"], t["is_max"]) or "?" for c2 in ALPHA)
            assert got == row, (t["weights"], t["is_max"], c1)
            assert "?" not in row          # every pair has a substitute (existence is weight independent)


def test_synthetic_golden(port, synthetic_cases):
    for c in synthetic_cases:
        r = port.search(c["weights"], c["is_max"], c["seq1"], c["seq2"])
        assert same_answer(r, c["expect"]), c["tag"]


def test_is_swapable_truth_table(port):
    for is_max in (0, 1):
        better, worse = (2.0, 1.0) if is_max else (1.0, 2.0)
        assert port.is_swapable(5, 5, 9, 9, worse, better, is_max)
        assert not port.is_swapable(5, 5, 1, 1, better, worse, is_max)
        assert port.is_swapable(5, 5, 4, 9, 1.0, 1.0, is_max)
        assert port.is_swapable(5, 5, 5, 4, 1.0, 1.0, is_max)
        assert not port.is_swapable(5, 5, 5, 5, 1.0, 1.0, is_max)
        assert not port.is_swapable(5, 5, 6, 0, 1.0, 1.0, is_max)


def test_rejects_bad_symbols_and_ranges(port):
    with pytest.raises(ValueError):
        port.search([1, 1, 1, 1], True, "ABCa", "AB")
    with pytest.raises(ValueError):
        port.search([1, 1, 1, 1], True, "AB", "ABC")
    with pytest.raises(ValueError):
        port.search([1, 1, 1, 1], True, "ABCD", "AB", first=2, last=2)


# ---- against the compiled reference (skipped where oracle/_ref is absent) ---------------------------

def test_vs_reference_random(port, ref):
    rng = random.Random(7)
    wsets = [[1, 3, 4, 2], [1, 1, 1, 1], [2, 1.5, 1.1, 1.3], [1.5, 2.6, 0.1, 0.2], [0.1, 0.7, 0.3, 0.9], [7, 0, 2, 0.5]]
    for trial in range(60):
        w = rng.choice(wsets)
        is_max = trial % 2
        n1 = rng.randint(1, 600)
        n2 = rng.randint(1, n1)
        alpha = ALPHA if trial % 5 == 0 else ALPHA[:26] if trial % 3 else "ACDG"
        s1 = "".join(rng.choice(alpha) for _ in range(n1))
        s2 = "".join(rng.choice(alpha) for _ in range(n2))
        a, b = port.search(w, is_max, s1, s2), ref.search(w, is_max, s1, s2)
        assert same_answer(a, b), (trial, w, is_max, n1, n2)


def test_vs_reference_per_offset_and_ranges(port, ref):
    rng = random.Random(11)
    s1 = "".join(rng.choice(ALPHA[:26]) for _ in range(400))
    s2 = "".join(rng.choice(ALPHA[:26]) for _ in range(37))
    for w in ([1, 3, 4, 2], [1.5, 2.6, 0.1, 0.2]):
        for is_max in (0, 1):
            scores = port.scores(w, is_max, s1, s2)
            for off in range(0, 364, 13):
                r = ref.offset_score(w, is_max, s1, s2, off)
                n = port.offset_naive(w, is_max, s1, s2, off)
                assert (r.score, r.char_offset, r.ch) == (n.score, n.char_offset, n.ch) and scores[off] == r.score
            for (f, l) in ((0, 364), (10, 11), (100, 300), (363, 364)):
                assert same_answer(port.search(w, is_max, s1, s2, f, l), ref.search(w, is_max, s1, s2, f, l))


def test_vs_reference_capacity_block(port, ref, input_blocks):
    """The largest in-spec problem (10000 / 5000, input.txt line 73) through the reference's own loop."""
    b = input_blocks[6]
    assert len(b["seq1"]) == ref.cap1 and len(b["seq2"]) == ref.cap2
    r = ref.search_omp(b["weights"], b["goal"] == "maximum", b["seq1"], b["seq2"], 8)
    assert same_answer(port.search(b["weights"], b["goal"] == "maximum", b["seq1"], b["seq2"], nthreads=4), r)


def test_emulated_two_rank_run_of_the_reference(ref, input_blocks):
    """oracle/ref_harness.cpp:ref_np2 -- two host threads standing in for `mpiexec -np 2` (divide_execute_tasks(&data, 2, pid)
    each, MAXLOC/MINLOC merge).  With every offset on the CPU loops (percentage 0, no GPU needed) the merged answer must be the
    single-rank answer; nothing in the reference pins the 2-rank result, so this is the emulation's own consistency check."""
    import bench
    for k in (1, 2, 3, 5, 8):
        b = input_blocks[k]
        with bench.silence_c_stdout():
            res, per_rank = ref.np2(b["weights"], b["goal"] == "maximum", b["seq1"], b["seq2"], pct=0, ndev=0, nthreads=1)       # 1 OpenMP thread per rank: the reference's fill_hash races with more
        e = b["expect"]
        assert (res.offset, res.char_offset, res.ch, res.score) == (e["offset"], e["char_offset"], e["ch"], e["score"]), (k, res)
        assert res.score in per_rank and len(per_rank) == 2
