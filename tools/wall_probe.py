"""wall clock around the split-phase run (resident batch: launch + kernel + wait, no copies) vs the one-shot call"""
import importlib, os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
psa = importlib.import_module("parallel-sequence-alignment_b200")
synth = importlib.import_module("parallel-sequence-alignment_b200.synth")
with psa.Context(1) as c:
    for name in sys.argv[1:] or ["c3", "c1"]:
        wl = bench.make_workload(synth, name, 0)
        b = psa.Batch(wl.seq1, wl.queries, pinned=True)
        out = c.new_result_array(b.nq, pinned=True)
        wc = psa.c_weights(wl.weights)
        for gate in (0, 1):
            c.set_option("gate_timed_runs", gate)
            c.prepare(wl.weights, wl.is_max, b)
            for _ in range(5):
                c.run()
            w, d = [], []
            for _ in range(40):
                t0 = time.perf_counter(); ms = c.run(); w.append(time.perf_counter() - t0); d.append(ms)
            print(name, "resident run, gate", gate, "wall median %.1f us, device (events) median %.1f us" % (statistics.median(w) * 1e6, statistics.median(d) * 1e3), flush=True)
        for _ in range(5):
            c.search_batch_raw(wc, wl.is_max, b, out)
        w = []
        for _ in range(40):
            t0 = time.perf_counter(); c.search_batch_raw(wc, wl.is_max, b, out); w.append(time.perf_counter() - t0)
        print(name, "one-shot wall median %.1f us" % (statistics.median(w) * 1e6), {k: c.stat(k) for k in ("host_plan_ns", "host_prepare_ns", "host_enqueue_ns", "host_wait_ns")}, flush=True)
        # empty call overhead of the wrapper
        w = []
        for _ in range(40):
            t0 = time.perf_counter(); c.stat("engine"); w.append(time.perf_counter() - t0)
        print("ctypes call overhead median %.2f us" % (statistics.median(w) * 1e6), flush=True)
