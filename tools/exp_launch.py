import importlib, os, sys, time
sys.path.insert(0, "/root/repo")
import torch
psa = importlib.import_module("parallel-sequence-alignment_b200")
synth = importlib.import_module("parallel-sequence-alignment_b200.synth")
wl = synth.workload("c3")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
with psa.Context(1) as c:
    b = psa.Batch(wl.seq1, wl.queries, pinned=True)
    c.prepare(wl.weights, wl.is_max, b)
    for _ in range(5): c.run()
    def avg(n, pre):
        t = 0.0
        for _ in range(n):
            pre()
            t += c.run()
        return t / n * 1e3
    def f_flush(): flush.zero_(); torch.cuda.synchronize()
    def f_sync(): torch.cuda.synchronize()
    def f_none(): pass
    small = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    def f_small(): small.zero_(); torch.cuda.synchronize()
    def f_sleep(): torch.cuda.synchronize(); time.sleep(0.002)
    print("flush 512MB + sync : %.2f us" % avg(20, f_flush))
    print("sync only          : %.2f us" % avg(20, f_sync))
    print("nothing            : %.2f us" % avg(20, f_none))
    print("small torch fill   : %.2f us" % avg(20, f_small))
    print("sync + 2 ms idle   : %.2f us" % avg(20, f_sleep))
    print("flush 512MB + sync : %.2f us" % avg(20, f_flush))
