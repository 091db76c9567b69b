#!/usr/bin/env python
"""Longer randomized parity run than tests/test_gpu_parity.py::test_random_batches (same generator, any seed, any count,
random tuning knobs): python tools/fuzz_batches.py [trials] [seed]   (PSA_FUZZ_GPUS=n for a multi-GPU context,
PSA_FUZZ_SLOTS=n for n device slots on GPU 0).
Needs a B200; checks against the C oracle."""
import importlib
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
psa = importlib.import_module("parallel-sequence-alignment_b200")
import oracle  # noqa: E402

ALPHA = [chr(65 + i) for i in range(26)] + ["-"]


def same(g, e):
    return (g.offset, g.char_offset, g.ch) == (e.offset, e.char_offset, e.ch) and (g.score == e.score or (g.score != g.score and e.score != e.score))


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = random.Random(seed)
    oracle.build()
    port = oracle.Port()
    wsets = [[1, 3, 4, 2], [1, 1, 1, 1], [2, 1.5, 1.1, 1.3], [5, 1, 2, 3], [0.1, 0.7, 0.3, 0.9], [10, 2, 3, 4], [1.5, 2.6, 0.1, 0.2],
             [0, 0, 0, 0], [7, 0, 2, 0.5], [3, 3, 3, 3], [1e6, 1, 1e-3, 5]]
    bad = 0
    slots = int(os.environ.get("PSA_FUZZ_SLOTS", "0"))
    modes = {}
    with (psa.Context(devices=[0] * slots) if slots else psa.Context(ngpus=int(os.environ.get("PSA_FUZZ_GPUS", "1")))) as ctx:
        ctx.set_option("min_split_work", 0)            # small problems: exercise the split over the device slots all the same
        for trial in range(trials):
            w = rng.choice(wsets)
            is_max = bool(rng.getrandbits(1))
            len1 = rng.choice([rng.randint(1, 400), rng.randint(400, 3000), rng.randint(3000, 9000)])
            nq = rng.choice([1, 2, 3, rng.randint(4, 40), rng.randint(40, 300)])
            alpha = rng.choice([ALPHA, ALPHA[:26], "ACDG", "AB", "A-"])
            s1 = "".join(rng.choice(alpha) for _ in range(len1))
            if rng.getrandbits(1):
                n2 = rng.choice([1, len1, rng.randint(1, len1), rng.randint(1, min(len1, 200))])
                lens = [n2] * nq
            else:
                lens = [rng.choice([1, len1, rng.randint(1, len1), rng.randint(1, min(len1, 64))]) for _ in range(nq)]
            if sum((len1 - n + 1) * n for n in lens) > 60_000_000:
                lens = [min(n, 300) for n in lens]
            qs = ["".join(rng.choice(alpha) for _ in range(n)) for n in lens]
            knobs = {"rank_planes": rng.choice([-1, -1, 0, 1, 2, 4]), "sliced_keys": rng.choice([1, 1, 0]), "pack_queries": rng.choice([1, 1, 0, 2, 3, 8]),
                     "fused_finish": rng.choice([1, 1, 0]), "derive_rank": rng.choice([1, 1, 0]), "zero_copy_results": rng.choice([1, 1, 0]),
                     "batch_mode": rng.choice([-1, -1, 0, 1]), "scan_warps": rng.choice([0, 0, 1, 2, 3, 4]),
                     "stripe_mode": rng.choice([-1, -1, 1, 1, 0]), "single_launch": rng.choice([1, 1, 0])}
            for k, v in knobs.items():
                ctx.set_option(k, v)
            got = ctx.search_batch(w, is_max, s1, qs)
            mode = "stripe" if ctx.stat("stripe_mode") else "single" if ctx.stat("single_launch") else "batch" if ctx.stat("batch_mode") else \
                "packed" if ctx.stat("packed_queries") > 0 else "sliced" if ctx.stat("slices") > 1 else "scalar" if ctx.stat("engine") == 1 else "long"
            modes[mode] = modes.get(mode, 0) + 1
            exp = port.search_batch(w, is_max, s1, qs)
            for k, (g, e) in enumerate(zip(got, exp)):
                if not same(g, e) or tuple(g.counts) != tuple(e.counts):
                    bad += 1
                    print("MISMATCH trial", trial, w, is_max, len1, lens[k], nq, "".join(alpha) if isinstance(alpha, list) else alpha, knobs, k, g, e, flush=True)
                    break
    print("fuzz: %d trials, seed %d, %d mismatching batches; kernels used: %s" % (trials, seed, bad, sorted(modes.items())))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
