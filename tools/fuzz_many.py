#!/usr/bin/env python
"""Randomized parity run of psa_search_many: random lists of random problems (the generator of tools/fuzz_batches.py, plus
knobs stream_queries / zero_copy_results / stripe_mode), 1..4 lanes on 1..3 device slots of GPU 0 (PSA_FUZZ_GPUS=n: n real
GPUs), every record against the C oracle.      python tools/fuzz_many.py [lists] [seed]"""
import importlib
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
psa = importlib.import_module("parallel-sequence-alignment_b200")
import oracle  # noqa: E402

ALPHA = [chr(65 + i) for i in range(26)] + ["-"]


def same(g, e):
    return (g.offset, g.char_offset, g.ch) == (e.offset, e.char_offset, e.ch) and (g.score == e.score or (g.score != g.score and e.score != e.score)) \
        and tuple(g.counts) == tuple(e.counts)


def problem(rng):
    wsets = [[1, 3, 4, 2], [1, 1, 1, 1], [2, 1.5, 1.1, 1.3], [5, 1, 2, 3], [0.1, 0.7, 0.3, 0.9], [10, 2, 3, 4], [1.5, 2.6, 0.1, 0.2], [3, 3, 3, 3]]
    w = rng.choice(wsets)
    is_max = bool(rng.getrandbits(1))
    len1 = rng.choice([rng.randint(1, 400), rng.randint(400, 3000), rng.randint(3000, 6000)])
    nq = rng.choice([1, 1, 2, rng.randint(3, 40), rng.randint(40, 400)])
    alpha = rng.choice([ALPHA, ALPHA[:26], "ACDG", "AB", "A-"])
    s1 = "".join(rng.choice(alpha) for _ in range(len1))
    if rng.random() < 0.6:
        n2 = rng.choice([1, len1, rng.randint(1, len1), rng.randint(1, min(len1, 200)), rng.randint(1, min(len1, 700))])
        lens = [n2] * nq
    else:
        lens = [rng.choice([1, len1, rng.randint(1, len1), rng.randint(1, min(len1, 64))]) for _ in range(nq)]
    while sum((len1 - n + 1) * n for n in lens) > 25_000_000:
        lens = [max(1, n // 2) for n in lens]
    return (w, is_max, s1, ["".join(rng.choice(alpha) for _ in range(n)) for n in lens])


def main():
    lists = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = random.Random(seed)
    oracle.build()
    port = oracle.Port()
    gpus = int(os.environ.get("PSA_FUZZ_GPUS", "0"))
    bad = done = 0
    ctxs = [psa.Context(ngpus=gpus)] if gpus else [psa.Context(devices=[0] * k) for k in (1, 2, 3)]
    try:
        for trial in range(lists):
            ctx = rng.choice(ctxs)
            items = [problem(rng) for _ in range(rng.choice([1, 2, 5, rng.randint(6, 24)]))]
            lanes = rng.choice([0, 1, 2, 3, 4])
            for k, v in {"stream_queries": rng.choice([1, 1, 0]), "zero_copy_results": rng.choice([1, 1, 0]), "stripe_mode": rng.choice([-1, -1, 1, 0]),
                         "single_launch": rng.choice([1, 1, 0])}.items():
                ctx.set_option(k, v)
            got = ctx.search_many(items, lanes=lanes)
            for k, ((w, is_max, s1, qs), res) in enumerate(zip(items, got)):
                exp = port.search_batch(w, is_max, s1, qs)
                done += 1
                if len(res) != len(exp) or not all(same(g, e) for g, e in zip(res, exp)):
                    bad += 1
                    print("MISMATCH list", trial, "problem", k, w, is_max, len(s1), len(qs), "lanes", lanes, flush=True)
    finally:
        for c in ctxs:
            c.close()
    print("fuzz_many: %d lists, %d problems, seed %d, %d mismatching problems" % (lists, done, seed, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
