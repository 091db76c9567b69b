#!/usr/bin/env python
"""Phase timeline of k_stripe (debug build: make -C parallel-sequence-alignment_b200 EXTRA=-DPSA_STRIPE_TRACE).
Prints, per phase, the SM-clock cycles between the marks warp 0 of each block leaves (median / max over blocks)."""
import ctypes as C
import importlib
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
psa = importlib.import_module("parallel-sequence-alignment_b200")
synth = importlib.import_module("parallel-sequence-alignment_b200.synth")
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
wl = synth.workload(name)
lib = C.CDLL(psa.LIB_PATH)
with psa.Context(1) as c:
    b = psa.Batch(wl.seq1, wl.queries, pinned=True)
    c.prepare(wl.weights, wl.is_max, b)
    for _ in range(5):
        ms = c.run()
    out = (C.c_longlong * (10 * 148))()
    assert lib.psa_debug_stripe_trace(out, 148) == 0
    rows = [list(out[10 * k: 10 * k + 8]) for k in range(148)]
    names = ["start->staged", "staged->window built", "built->row offsets (task 1)", "row offsets->counted (first pass)",
             "counted->keys+settle", "keys->team barrier", "barrier->finished"]
    print(f"{name}: device ms {ms:.4f}; cycles per phase as seen by warp 0 (median / max over 148 blocks)")
    tot = []
    for k, n in enumerate(names):
        d = [r[k + 1] - r[k] for r in rows if r[k + 1] > r[k] > 0]
        if d:
            print(f"  {n:40s} {statistics.median(d):9.0f} {max(d):9.0f}")
    d = [r[7] - r[0] for r in rows if r[7] > r[0] > 0]
    print(f"  {'whole block (warp 0)':40s} {statistics.median(d):9.0f} {max(d):9.0f}")
