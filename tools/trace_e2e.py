"""Host-side split of the one-shot call on config 3: PSA_TRACE=1 makes psa_search_batch print planning vs device time."""
import ctypes as C
import importlib
import os
import sys
import time

os.environ["PSA_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
psa = importlib.import_module("parallel-sequence-alignment_b200")
synth = importlib.import_module("parallel-sequence-alignment_b200.synth")
wl = synth.workload(sys.argv[1] if len(sys.argv) > 1 else "c3")
with psa.Context(ngpus=1) as ctx:
    batch = psa.Batch(wl.seq1, wl.queries, pinned=True)
    out = ctx.new_result_array(batch.nq, pinned=True)
    w = (C.c_double * 4)(*wl.weights)
    for k in range(8):
        t0 = time.perf_counter()
        ctx.search_batch_raw(w, wl.is_max, batch, out)
        print("call %d: %.1f us" % (k, 1e6 * (time.perf_counter() - t0)), file=sys.stderr)
