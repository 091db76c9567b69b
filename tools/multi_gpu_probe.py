#!/usr/bin/env python
"""Single process driving 1..N device slots (strong scaling, end to end through psa_search_batch, pinned host buffers),
with the library's host-side split of one call (psa_get_stat host_*_ns): planning on the calling thread, then per device
slot the H2D enqueue, the launches, and the wait for the stream + results.

    python tools/multi_gpu_probe.py [--dup] [--counts 1,2,4,8] [c3 c4 c5]

--dup repeats ordinal 0 (every slot still has its own stream, buffers and worker thread), which shows the host side of
an N-device context on a 1-GPU box; without it the slots are the first N visible GPUs."""
import importlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
psa = importlib.import_module("parallel-sequence-alignment_b200")
synth = importlib.import_module("parallel-sequence-alignment_b200.synth")


def main():
    args = sys.argv[1:]
    dup = "--dup" in args
    counts = [1, 2, 4, 8]
    if "--counts" in args:
        counts = [int(x) for x in args[args.index("--counts") + 1].split(",")]
        del args[args.index("--counts"): args.index("--counts") + 2]
    names = [a for a in args if not a.startswith("--")] or ["c3", "c4", "c5"]
    have = psa.device_count()
    base = {}
    for n in counts:
        if not dup and n > have:
            continue
        devices = [0] * n if dup else list(range(n))
        with psa.Context(devices=devices) as c:
            for name in names:
                wl = synth.workload(name)
                b = psa.Batch(wl.seq1, wl.queries, pinned=True)
                out = c.new_result_array(b.nq, pinned=True)
                wc = psa.c_weights(wl.weights)
                for _ in range(3):
                    c.search_batch_raw(wc, wl.is_max, b, out)
                reps = 10
                split = {"host_plan_ns": 0, "host_prepare_ns": 0, "host_enqueue_ns": 0, "host_wait_ns": 0, "host_total_ns": 0}
                t0 = time.perf_counter()
                for _ in range(reps):
                    c.search_batch_raw(wc, wl.is_max, b, out)
                    for k in split:
                        split[k] += c.stat(k)
                dt = (time.perf_counter() - t0) / reps
                base.setdefault(name, dt)
                rec = {"slots": n, "devices": "cuda:0 repeated" if dup else "distinct GPUs", "workload": name, "ms": dt * 1e3,
                       "pair_evals_per_s": b.pair_evals / dt, "speedup_vs_first": base[name] / dt,
                       "host_split_us": {k[5:-3]: v / reps * 1e-3 for k, v in split.items()}}
                print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
