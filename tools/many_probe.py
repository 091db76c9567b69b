#!/usr/bin/env python
"""psa_search_many: end-to-end throughput of a list of independent batches (pinned host buffers, records written straight
into pinned host arrays) against the same list through one psa_search_batch call at a time.

    python tools/many_probe.py [--gpus N] [c3 c1 c5]"""
import importlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
psa = importlib.import_module("parallel-sequence-alignment_b200")
synth = importlib.import_module("parallel-sequence-alignment_b200.synth")


def main():
    args = sys.argv[1:]
    ngpus = 1
    if "--gpus" in args:
        ngpus = int(args[args.index("--gpus") + 1]); del args[args.index("--gpus"): args.index("--gpus") + 2]
    names = args or ["c3", "c1", "c5"]
    with psa.Context(ngpus) as c:
        for name in names:
            distinct, total = {"c3": (64, 512), "c1": (16, 512), "c2": (16, 512), "c5": (4, 32), "c4": (4, 64)}[name]
            total *= ngpus
            built = []
            for k in range(distinct):
                wl = bench.make_workload(synth, name, k) if name != "c1" else bench.make_workload(synth, name, 0)
                b = psa.Batch(wl.seq1, wl.queries, pinned=True)
                built.append((psa.c_weights(wl.weights), wl.is_max, b, c.new_result_array(b.nq, pinned=True)))
            items = [built[k % distinct] for k in range(total)]
            pe = sum(it[2].pair_evals for it in items)
            arr = c.make_problem_list(items)
            # one call at a time
            for it in items[:8]:
                c.search_batch_raw(it[0], it[1], it[2], it[3])
            t0 = time.perf_counter()
            for it in items:
                c.search_batch_raw(it[0], it[1], it[2], it[3])
            t_seq = time.perf_counter() - t0
            print(json.dumps({"workload": name, "gpus": ngpus, "problems": total, "how": "psa_search_batch, one call at a time",
                              "us_per_problem": round(t_seq / total * 1e6, 2), "pair_evals_per_s": pe / t_seq}), flush=True)
            for lanes in (1, 2, 3, 4, 6, 8):
                c.search_many_raw(arr, min(total, 16 * ngpus), lanes)
                best = None
                for _ in range(3):
                    t0 = time.perf_counter()
                    c.search_many_raw(arr, total, lanes)
                    dt = time.perf_counter() - t0
                    best = dt if best is None else min(best, dt)
                print(json.dumps({"workload": name, "gpus": ngpus, "problems": total, "how": f"psa_search_many, {lanes} lanes per GPU",
                                  "us_per_problem": round(best / total * 1e6, 2), "pair_evals_per_s": pe / best,
                                  "speedup_vs_one_at_a_time": round(t_seq / best, 3)}), flush=True)


if __name__ == "__main__":
    main()
