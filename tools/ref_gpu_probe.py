#!/usr/bin/env python
"""Times the REFERENCE's own CUDA path (gpu_run_program + its three kernels, cuda_funcs.cu:6-278, recompiled for
sm_100a inside oracle/_ref/libpsa_ref.so) on a few queries of a workload.  Baseline only: that path races across
blocks, so its answers are compared with the oracle and mismatches are merely counted."""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

import bench  # noqa: E402  (workload definitions shared with the benchmark)

synth = importlib.import_module("parallel-sequence-alignment_b200.synth")


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    wl = bench.make_workload(synth, name, 0, nq=nq if name in ("c3", "c5") else None)
    ref, port = oracle.Ref(), oracle.Port()
    seq1 = wl.seq1[: ref.cap1]
    qs = wl.queries[:nq]
    ref.divide_execute_tasks(wl.weights, wl.is_max, seq1, qs[0], 1, 0, 100, 1)        # warm-up (context, module load)
    dt = None
    for _ in range(3):                      # best of three passes: the first ones still pay allocator / clock warm-up
        t0 = time.perf_counter()
        got = [ref.divide_execute_tasks(wl.weights, wl.is_max, seq1, q, 1, 0, 100, 1) for q in qs]
        t = time.perf_counter() - t0
        dt = t if dt is None else min(dt, t)
    exp = [port.search(wl.weights, wl.is_max, seq1, q) for q in qs]
    bad = sum((g.offset, g.char_offset, g.score) != (e.offset, e.char_offset, e.score) for g, e in zip(got, exp))
    pe = sum((len(seq1) - len(q) + 1) * len(q) for q in qs)
    print(json.dumps({"workload": name, "queries": len(qs), "pair_evals_per_s": pe / dt, "seconds": dt,
                      "answers_differing_from_cpu_reference": bad}))


if __name__ == "__main__":
    main()
