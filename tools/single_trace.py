#!/usr/bin/env python
"""Phase timeline of k_single as seen by block 0 (debug build: make -C parallel-sequence-alignment_b200 EXTRA=-DPSA_SINGLE_TRACE)."""
import ctypes as C, importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
psa = importlib.import_module("parallel-sequence-alignment_b200")
synth = importlib.import_module("parallel-sequence-alignment_b200.synth")
name = sys.argv[1] if len(sys.argv) > 1 else "c1"
wl = bench.make_workload(synth, name, 0)
lib = C.CDLL(psa.LIB_PATH)
with psa.Context(1) as c:
    b = psa.Batch(wl.seq1, wl.queries, pinned=True)
    c.prepare(wl.weights, wl.is_max, b)
    for _ in range(5):
        ms = c.run()
    out = (C.c_longlong * 32)()
    assert lib.psa_debug_single_trace(out) == 0
    t = list(out)
    names = ["start", "unit staged", "window built", "unit counted (block 0's last unit)", "past the grid barrier", "combine tiles done"]
    print(f"{name}: device ms {ms:.4f} (events around the launch)")
    for k in range(1, 6):
        print(f"  -> {names[k]:40s} {t[k] - t[k - 1]:8d} cycles   {(t[16 + k] - t[16 + k - 1]) / 1e3:7.2f} us")
    print(f"  start -> combine done: {(t[16 + 5] - t[16]) / 1e3:.2f} us (block 0)")
    print(f"  block 0 start -> last block has the ticket: {(t[16 + 6] - t[16]) / 1e3:.2f} us; its finish step: {(t[16 + 7] - t[16 + 6]) / 1e3:.2f} us; "
          f"start -> end {(t[16 + 7] - t[16]) / 1e3:.2f} us")
