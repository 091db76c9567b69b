#!/usr/bin/env python
"""Kernel throughput of a workload shape as the number of queries grows (is a small batch just too small?)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
psa = importlib.import_module("parallel-sequence-alignment_b200")
synth = importlib.import_module("parallel-sequence-alignment_b200.synth")
ctx = psa.Context(1)
ctx.set_option("kernel_events", 1)
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
for nq in [int(x) for x in (sys.argv[2:] or ["1024", "4096", "16384"])]:
    wl = synth.workload(name, nq=nq)
    b = psa.Batch(wl.seq1, wl.queries)
    for bm in (-1, 0, 1):
        ctx.set_option("batch_mode", bm)
        ctx.prepare(wl.weights, wl.is_max, b)
        ms = min(ctx.run() for _ in range(5)); k = ctx.stat("main_kernel_ns") / 1e6
        ctx.fetch()
        print(f"{name} nq={nq} batch_opt={bm} batch={ctx.stat('batch_mode')} warps={ctx.stat('scan_warps')} run {ms:.3f} ms  scan {k:.3f} ms  "
              f"kernel {b.pair_evals / k / 1e9:.1f} T pair-evals/s  run {b.pair_evals / ms / 1e9:.1f} T/s")
