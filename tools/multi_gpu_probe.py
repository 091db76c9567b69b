#!/usr/bin/env python
"""Single process driving 1..N GPUs (strong scaling, end to end through psa_search_batch, pinned host buffers)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
psa = importlib.import_module("parallel-sequence-alignment_b200")
synth = importlib.import_module("parallel-sequence-alignment_b200.synth")
have = psa.device_count()
for n in [g for g in (1, 2, 4, 8) if g <= have]:
    with psa.Context(ngpus=n) as c:
        for name in sys.argv[1:] or ["c3", "c4", "c5"]:
            wl = synth.workload(name)
            b = psa.Batch(wl.seq1, wl.queries, pinned=True)
            out = c.new_result_array(b.nq, pinned=True); wc = psa.c_weights(wl.weights)
            for _ in range(3): c.search_batch_raw(wc, wl.is_max, b, out)
            t0 = time.perf_counter()
            for _ in range(10): c.search_batch_raw(wc, wl.is_max, b, out)
            dt = (time.perf_counter() - t0) / 10
            print(f"single process, {n} GPUs, {name}: {dt*1e3:.3f} ms  {b.pair_evals/dt:.3e} pair-evals/s", flush=True)
