#!/usr/bin/env python
"""Regenerates the golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE
(oracle/_ref/libpsa_ref.so, built from /root/reference by `make -C oracle ref`).

Only runs where /root/reference exists (the build container); the fixtures it writes are what
travels to the GPU box.  Usage:  python tests/golden/make_golden.py

  input_blocks.json   the 10 distinct problem blocks stacked in the reference's input.txt
                      (only block 0 is pinned by the reference's own output.txt: "4505 -4879");
                      answers from the reference's 1-thread path (the "-100" argument).
  sign_matrix.json    get_hashtable_sign over the 27x27 alphabet (cuda_funcs.cu:424-439),
                      the README's printed matrix.
  substitutes.json    get_substitute (cuda_funcs.cu:310-317) over 27x27 for several weight sets x goals.
  synthetic.json      seeded random problems (integer / dyadic / non-dyadic weights, ties, gaps)
                      with the reference's answer.
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF_INPUT = "/root/reference/input.txt"
REF_OUTPUT = "/root/reference/output.txt"
ALPHA = [chr(65 + i) for i in range(26)] + ["-"]


def main():
    oracle.build(ref=True)
    ref = oracle.Ref()

    # ---- input.txt blocks ----------------------------------------------------------------
    toks = open(REF_INPUT).read().split()
    assert len(toks) % 7 == 0
    seen, blocks = set(), []
    for k in range(0, len(toks), 7):
        key = tuple(toks[k:k + 7])
        if key in seen:
            continue
        seen.add(key)
        w = [float(x) for x in toks[k:k + 4]]
        s1, s2, goal = toks[k + 4], toks[k + 5], toks[k + 6]
        r = ref.search(w, goal == "maximum", s1, s2)
        blocks.append({"weights": w, "weights_text": toks[k:k + 4], "seq1": s1, "seq2": s2, "goal": goal,
                       "expect": {"offset": r.offset, "char_offset": r.char_offset, "ch": r.ch,
                                  "score": r.score, "score_g": "%g" % r.score}})
    out_lines = open(REF_OUTPUT).read().split("\n")
    b0 = blocks[0]
    mut = b0["seq2"][: b0["expect"]["char_offset"]] + b0["expect"]["ch"] + b0["seq2"][b0["expect"]["char_offset"] + 1:]
    assert out_lines[0] == mut and out_lines[1] == "%d %s" % (b0["expect"]["offset"], b0["expect"]["score_g"]), \
        "reference build does not reproduce its own output.txt"
    b0["output_txt"] = "\n".join(out_lines)
    json.dump(blocks, open(os.path.join(HERE, "input_blocks.json"), "w"), indent=0)

    # ---- sign matrix ------------------------------------------------------------------------------
    json.dump({"alphabet": "".join(ALPHA), "rows": ["".join(ref.sign(a, b) for b in ALPHA) for a in ALPHA]},
              open(os.path.join(HERE, "sign_matrix.json"), "w"), indent=0)

    # ---- substitutes --------------------------------------------------------------------------------
    wsets = [[1, 3, 4, 2], [2, 1.5, 1.1, 1.3], [1.5, 2.6, 0.1, 0.2], [0.8, 0.54, 2.6, 13.7], [1, 1, 1, 1],
             [5, 4, 3, 2], [1, 2, 2, 1], [1, 1, 2, 1], [0, 0, 0, 0], [3, 1, 2, 7], [0.25, 8, 0.5, 4],
             [1.5, 2.6, 0.3, 0.2], [10, 1, 1, 1], [1, 10, 1, 1], [1, 1, 10, 1], [1, 1, 1, 10]]
    subs = []
    for w in wsets:
        for is_max in (0, 1):
            subs.append({"weights": w, "is_max": is_max,
                         "rows": ["".join(ref.substitute(c1, c2, w, is_max) or "?" for c2 in ALPHA) for c1 in ALPHA]})
    json.dump({"alphabet": "".join(ALPHA), "index": "rows[c1][c2]", "tables": subs},
              open(os.path.join(HERE, "substitutes.json"), "w"), indent=0)

    # ---- synthetic problems -------------------------------------------------------------------------------
    rng = random.Random(20251018)
    cases = []

    def add(w, is_max, s1, s2, tag):
        r = ref.search(w, is_max, s1, s2)
        cases.append({"tag": tag, "weights": w, "is_max": int(is_max), "seq1": s1, "seq2": s2,
                      "expect": {"offset": r.offset, "char_offset": r.char_offset, "ch": r.ch,
                                 "score": r.score, "score_g": "%g" % r.score}})

    def rand_seq(n, alphabet=ALPHA[:26]):
        return "".join(rng.choice(alphabet) for _ in range(n))

    for w in ([1, 1, 1, 1], [1, 3, 4, 2], [2, 1.5, 1.1, 1.3], [1.5, 2.6, 0.1, 0.2], [0.8, 0.54, 2.6, 13.7],
              [0.25, 8, 0.5, 4], [0, 0, 0, 0]):
        for is_max in (0, 1):
            for (n1, n2) in ((60, 7), (300, 64), (700, 333), (1200, 1200), (257, 1)):
                add(w, is_max, rand_seq(n1), rand_seq(n2), "uniform26")
    # tie-heavy: tiny alphabets, planted copies, gaps
    for is_max in (0, 1):
        add([1, 1, 1, 1], is_max, rand_seq(500, "AB"), rand_seq(20, "AB"), "alphabet2")
        add([1, 3, 4, 2], is_max, rand_seq(800, "ACDEFGHIKLMNPQRSTVWY"), rand_seq(100, "ACDEFGHIKLMNPQRSTVWY"), "amino20")
        core = rand_seq(50)
        add([1, 3, 4, 2], is_max, rand_seq(100) + core + rand_seq(77) + core + rand_seq(31), core, "planted_twice")
        add([2, 1.5, 1.1, 1.3], is_max, rand_seq(100) + core + rand_seq(77) + core + rand_seq(31), core, "planted_twice_fp")
        add([1, 2, 2, 1], is_max, rand_seq(400, ALPHA), rand_seq(40, ALPHA), "with_gaps")
        add([1, 1, 1, 1], is_max, "A" * 300, "A" * 50, "all_equal")
        add([1.5, 2.6, 0.1, 0.2], is_max, "A" * 300, "A" * 50, "all_equal_fp")
        add([1, 3, 4, 2], is_max, rand_seq(33), rand_seq(33), "single_offset")
    json.dump(cases, open(os.path.join(HERE, "synthetic.json"), "w"), indent=0)
    print(f"wrote {len(blocks)} input blocks, {len(subs)} substitute tables, {len(cases)} synthetic cases")


if __name__ == "__main__":
    main()
