#!/usr/bin/env python
"""Phase timeline of k_finish on config 4 (debug build: make -C parallel-sequence-alignment_b200 EXTRA=-DPSA_FINISH_TRACE)."""
import ctypes as C, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
psa = importlib.import_module("parallel-sequence-alignment_b200")
synth = importlib.import_module("parallel-sequence-alignment_b200.synth")
wl = synth.workload(sys.argv[1] if len(sys.argv) > 1 else "c4")
lib = C.CDLL(psa.LIB_PATH)
with psa.Context(1) as c:
    b = psa.Batch(wl.seq1, wl.queries, pinned=True)
    c.prepare(wl.weights, wl.is_max, b)
    for _ in range(5):
        ms = c.run()
    out = (C.c_longlong * 16)()
    assert lib.psa_debug_finish_trace(out) == 0
    t = list(out)
    names = ["start -> tables in shared memory", "-> winner over the tile records", "-> (enter re-score)", "-> candidate tiles listed",
             "-> (count words)", "-> candidate words re-scored", "-> winner chosen", "-> final walk done"]
    print(f"device ms {ms:.4f}")
    for k in range(1, 8):
        if t[k] and t[k - 1]:
            print(f"  {names[k]:40s} {t[k] - t[k - 1]:8d} cycles")
    print(f"  total {t[7] - t[0]} cycles")
