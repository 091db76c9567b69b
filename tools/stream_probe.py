#!/usr/bin/env python
"""A/B of the one-shot call (psa_search_batch, pinned host buffers, wall clock): queries streamed on a second stream while
k_stripe runs vs one copy in front of the kernel, and records written straight into host memory vs copied back.

    python tools/stream_probe.py [c3 c5 ...]"""
import importlib
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
psa = importlib.import_module("parallel-sequence-alignment_b200")
synth = importlib.import_module("parallel-sequence-alignment_b200.synth")


def main():
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c3", "c5", "c1", "c2"]
    with psa.Context(1) as c:
        for name in names:
            wl = bench.make_workload(synth, name, 0)
            b = psa.Batch(wl.seq1, wl.queries, pinned=True)
            out = c.new_result_array(b.nq, pinned=True)
            wc = psa.c_weights(wl.weights)
            for stream, zc in ((0, 0), (0, 1), (1, 0), (1, 1), (0, 0), (1, 1)):
                c.set_option("stream_queries", stream)
                c.set_option("zero_copy_results", zc)
                for _ in range(5):
                    c.search_batch_raw(wc, wl.is_max, b, out)
                ts, split = [], {"host_plan_ns": [], "host_prepare_ns": [], "host_enqueue_ns": [], "host_wait_ns": []}
                for _ in range(40):
                    t0 = time.perf_counter()
                    c.search_batch_raw(wc, wl.is_max, b, out)
                    ts.append(time.perf_counter() - t0)
                    for k in split:
                        split[k].append(c.stat(k))
                print(json.dumps({"workload": name, "stream_queries": stream, "zero_copy_results": zc, "pieces": c.stat("streamed_chunks"),
                                  "e2e_us_median": round(statistics.median(ts) * 1e6, 1), "e2e_us_min": round(min(ts) * 1e6, 1),
                                  "host_split_us_median": {k[5:-3]: round(statistics.median(v) * 1e-3, 1) for k, v in split.items()}}), flush=True)


if __name__ == "__main__":
    main()
