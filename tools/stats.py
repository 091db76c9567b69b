import importlib, sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
psa = importlib.import_module("parallel-sequence-alignment_b200")
synth = importlib.import_module("parallel-sequence-alignment_b200.synth")
ctx = psa.Context(1)
ctx.set_option("kernel_events", 1)
for name in ("c3","c5","c4","c1"):
    wl = synth.workload(name, nq=(4096 if name=="c5" else None)) if name!="c1" else None
    if wl is None:
        import json
        b=json.load(open('/root/repo/tests/golden/input_blocks.json'))[0]
        wl=synth.Workload("c1", b["weights"], b["goal"]=="maximum", b["seq1"].encode(), [b["seq2"].encode()])
    b = psa.Batch(wl.seq1, wl.queries)
    for planes, bm, sl in ((-1, -1, 1), (-1, -1, 0), (1, 1, 1), (2, 1, 1), (4, 1, 1), (2, 1, 0), (4, 1, 0), (1, 0, 1), (2, 0, 1), (4, 0, 1)):
        ctx.set_option("rank_planes", planes)
        ctx.set_option("batch_mode", bm)
        ctx.set_option("sliced_keys", sl)
        ctx.prepare(wl.weights, wl.is_max, b)
        ms = min(ctx.run() for _ in range(5)); ctx.fetch()
        t = psa.build_pair_table(wl.weights, wl.is_max, max(b.lens))
        print(name, "planes", ctx.stat("rank_planes"), "nranks", t.nranks, "exact", t.exact, "tiles", ctx.stat("tiles"), "rescored_words", ctx.stat("candidate_tiles"), "warps", ctx.stat("scan_warps"), "batch", ctx.stat("batch_mode"), "sliced", sl, "run_ms %.3f"%ms, "scan_ms %.3f"%(ctx.stat("main_kernel_ns")/1e6))
