#!/usr/bin/env python
"""Times the REFERENCE's own CUDA path (gpu_run_program + its three kernels, cuda_funcs.cu:6-278, recompiled for
sm_100a inside oracle/_ref/libpsa_ref.so) on a few queries of a workload.

    python tools/ref_gpu_probe.py <workload> <queries> rank1|np2

rank1  one rank: divide_execute_tasks(&data, 1, 0) with the CUDA percentage at 100 (all offsets on one GPU)
np2    the EMULATED `mpiexec -np 2` CUDA+OpenMP run (MPI is not installed in the image): two host threads as the two ranks,
       each divide_execute_tasks(&data, 2, pid) on GPU pid % visible, MAXLOC/MINLOC merge (oracle/ref_harness.cpp: ref_np2)

Baselines only: that path races across blocks, so its answers are compared with the oracle and mismatches are
merely counted."""
import importlib.util
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

import bench  # noqa: E402  (workload definitions shared with the benchmark; does not import the product package)


def visible_gpus():
    import ctypes
    try:
        rt = ctypes.CDLL("libcudart.so")
    except OSError:
        try:
            import torch
            return torch.cuda.device_count()
        except Exception:
            return 1
    n = ctypes.c_int(0)
    return n.value if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 1


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    mode = sys.argv[3] if len(sys.argv) > 3 else "rank1"
    synth = bench.load_synth()
    wl = bench.make_workload(synth, name, 0, nq=nq if name in ("c3", "c5") else None)
    ref, port = oracle.Ref(), oracle.Port()
    seq1 = wl.seq1[: ref.cap1]
    qs = wl.queries[:nq]
    ndev = max(1, visible_gpus())

    def one(q):
        if mode == "np2":
            return ref.np2(wl.weights, wl.is_max, seq1, q, pct=100, ndev=ndev, nthreads=4)[0]
        return ref.divide_execute_tasks(wl.weights, wl.is_max, seq1, q, 1, 0, 100, 1)

    with bench.silence_c_stdout():
        one(qs[0])                          # warm-up (context, module load)
        dt = None
        for _ in range(3):                  # best of three passes: the first ones still pay allocator / clock warm-up
            t0 = time.perf_counter()
            got = [one(q) for q in qs]
            t = time.perf_counter() - t0
            dt = t if dt is None else min(dt, t)
    exp = [port.search(wl.weights, wl.is_max, seq1, q) for q in qs]
    bad = sum((g.offset, g.char_offset, g.score) != (e.offset, e.char_offset, e.score) for g, e in zip(got, exp))
    pe = sum((len(seq1) - len(q) + 1) * len(q) for q in qs)
    what = ("EMULATED mpiexec -np 2 of the reference's CUDA+OpenMP build: two host threads as the two ranks, each the reference's "
            f"divide_execute_tasks(&data, 2, pid) with all of its offsets on GPU pid % {ndev} (its own cudaMalloc + 4 kernels + cudaFree per "
            "call), MAXLOC/MINLOC merge" if mode == "np2" else
            "reference gpu_run_program (cudaMalloc + 4 kernels + cudaFree per query), one rank, one GPU")
    print(json.dumps({"workload": name, "mode": mode, "queries": len(qs), "value": pe / dt, "seconds": dt, "gpus_visible": ndev,
                      "answers_differing_from_cpu_reference": bad,
                      "what": what + "; Seq1 truncated to its 10000 capacity where longer"}))


if __name__ == "__main__":
    main()
