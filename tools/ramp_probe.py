"""how much of a short psa_search_many list is ramp: wall clock of lists of 2..80 config-3 batches, with and without an L2 flush + idle gap before the call"""
import importlib, os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
psa = importlib.import_module("parallel-sequence-alignment_b200")
synth = importlib.import_module("parallel-sequence-alignment_b200.synth")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
with psa.Context(1) as c:
    built = []
    for k in range(32):
        wl = bench.make_workload(synth, "c3", 0, variant=k)
        b = psa.Batch(wl.seq1, wl.queries, pinned=True)
        built.append((psa.c_weights(wl.weights), wl.is_max, b, c.new_result_array(b.nq, pinned=True)))
    for lanes in (2, 3):
        for n in (2, 4, 10, 20, 40, 80):
            items = [built[k % 32] for k in range(n)]
            arr = c.make_problem_list(items)
            c.search_many_raw(arr, n, lanes)
            for pre in ("none", "flush+sync", "sleep 5 ms"):
                t = []
                for _ in range(9):
                    if pre == "flush+sync":
                        flush.zero_(); torch.cuda.synchronize()
                    elif pre == "sleep 5 ms":
                        time.sleep(0.005)
                    t0 = time.perf_counter(); c.search_many_raw(arr, n, lanes); t.append(time.perf_counter() - t0)
                print(f"lanes {lanes}  list of {n:3d}  before the call: {pre:11s}  median {statistics.median(t) * 1e6:8.1f} us  = {statistics.median(t) * 1e6 / n:6.2f} us per batch", flush=True)
