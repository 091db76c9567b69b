"""device time of one stripe-mode batch under a forced queries-per-task (PSA_STRIPE_Q, read once per process): python tools/plan_probe.py c5 8192"""
import importlib, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
psa = importlib.import_module("parallel-sequence-alignment_b200")
synth = importlib.import_module("parallel-sequence-alignment_b200.synth")
name, nq = sys.argv[1], int(sys.argv[2])
wl = bench.make_workload(synth, name, 0, nq=nq)
with psa.Context(1) as c:
    c.set_option("gate_timed_runs", 1)
    b = psa.Batch(wl.seq1, wl.queries, pinned=True)
    c.prepare(wl.weights, wl.is_max, b)
    for _ in range(4):
        c.run()
    t = [c.run() for _ in range(15)]
    print(name, nq, "PSA_STRIPE_Q=%s" % os.environ.get("PSA_STRIPE_Q", "auto"), "Q", c.stat("stripe_queries_per_task"), "T", c.stat("stripe_team_warps"),
          "split", c.stat("stripe_split"), "median %.1f us" % (statistics.median(t) * 1e3), flush=True)
