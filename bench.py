#!/usr/bin/env python
"""bench.py -- pair-evals/s of the mutant-offset search on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c5|c4|c2|c1] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic queries of the named BASELINE.json
shape.  Weak scaling: every rank owns a full batch (different seed per rank), no data-path collective;
`value` = pair-evals of all ranks / max-over-ranks device time.

  value      batch resident in HBM (psa_batch_prepare), per-step device time from CUDA events recorded
             by the library on the stream it launches on (psa_batch_run), L2 flushed between steps
  e2e        the public call psa_search_batch with HOST (pinned) buffers: H2D + kernels + D2H + host
             scoring, wall clock per step
  roofline   the dominant kernel alone (its own CUDA events) against the integer-issue roofline of
             SURVEY.md section 8(d); HBM is shown to be non-binding
  cpu_baseline / --impl reference
             the reference's own CPU loop (oracle/_ref, compiled from /root/reference) on a bounded
             sample of the same workload, all host threads
"""
from __future__ import annotations

import argparse
import contextlib
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "parallel-sequence-alignment_b200"

METRIC = "pair-evals/sec"
UNIT = "pair-evals/s"
# 2 int32 lane-ops per pair-eval, 128 lanes/clk/SM issue, 148 SMs at the measured max SM clock
SM_COUNT = 148
LANE_OPS_PER_PAIR_EVAL = 2.0
ALU_OPS_PER_WARP_STEP = 10.2    # k_scan class pass, from SASS: 265 LOP3 + 62 SHF per 32 steps
HBM_BYTES_PER_QUERY_FIXED = 56 + 40         # result record + one tile record (equal-length batches carry no per-query offsets)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "sm_max_mhz": d.get("sm_max_mhz", 1965.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}


def profiled_traffic(workload: str, kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu
    capture of this workload (profiles/r01_summary.json, written by tools/ncu_summary.py); None if not captured."""
    p = os.path.join(ROOT, "profiles", "r01_summary.json")
    if not os.path.exists(p):
        return None, None
    rec = json.load(open(p)).get(workload, {})
    for name, v in rec.items():
        if name == kernel or name.startswith(kernel + "<"):
            return v.get("dram_bytes_per_launch"), v
    return None, None


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm),
                "note": "sampled every 100 ms from the first timed step until the same step had run long enough for 6 samples"}


def golden_c1():
    with open(os.path.join(ROOT, "tests", "golden", "input_blocks.json")) as f:
        b = json.load(f)[0]
    return b["weights"], b["goal"] == "maximum", b["seq1"].encode(), [b["seq2"].encode()]


def make_workload(synth, name, rank, nq=None):
    if name == "c1":
        w, is_max, s1, qs = golden_c1()
        return synth.Workload("c1", w, is_max, s1, qs, "reference input.txt block 1 (9711/2131 MIN)")
    return synth.workload(name, nq=nq, seed_shift=1000 * rank)


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU loop on a bounded sample
# ------------------------------------------------------------------------------------------------
def reference_sample(synth, name, seconds=12.0):
    """Pick a sample (number of queries) of workload `name` worth ~`seconds` of reference CPU time."""
    import oracle
    threads = os.cpu_count() or 1
    kind = "reference" if oracle.ref_available() else "port"
    eng = oracle.Ref() if kind == "reference" else oracle.Port()
    wl = make_workload(synth, name, 0, nq=None if name in ("c1", "c2", "c4") else (1024 if name == "c3" else 8192))
    seq1 = wl.seq1
    cap_note = ""
    if kind == "reference" and len(seq1) > eng.cap1:          # c4: beyond the reference's static buffers
        seq1 = seq1[: eng.cap1]
        cap_note = f" (Seq1 truncated to the reference's capacity {eng.cap1})"

    def run(queries):
        t0 = time.perf_counter()
        for q in queries:
            if kind == "reference":
                eng.search_omp(wl.weights, wl.is_max, seq1, q, threads)
            else:
                eng.search(wl.weights, wl.is_max, seq1, q, nthreads=threads)
        return time.perf_counter() - t0

    pe = lambda qs: sum((len(seq1) - len(q) + 1) * len(q) for q in qs)
    probe = wl.queries[:1]
    dt = max(run(probe), 1e-4)
    n = max(1, min(len(wl.queries), int(seconds / dt)))
    sample = wl.queries[:n]
    return {"kind": kind, "threads": threads, "run": run, "sample": sample, "pair_evals": pe(sample),
            "desc": f"{n} of the workload's queries, {kind} find_best_mutant_cpu split over {threads} threads "
                    f"(-O3 build, table filled single-threaded){cap_note}", "wl": wl}


@contextlib.contextmanager
def silence_c_stdout():
    """The reference printf()s from C ("CUDA percentage set to ..."); keep that out of our one-JSON-line stdout,
    including whatever libc still holds in its buffer."""
    libc = ctypes.CDLL(None)
    sys.stdout.flush()
    libc.fflush(None)
    saved, devnull = os.dup(1), os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        yield
    finally:
        libc.fflush(None)
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)


def as_shipped_baseline(synth, name, seconds=3.0):
    """The reference exactly as its Makefile builds it (no -O flag) and as main.c runs it (4 OpenMP threads,
    divide_execute_tasks with the CUDA percentage forced to 0).  Timing only: with more than one thread the
    reference's fill_hash races, so its answers are not checked."""
    try:
        import oracle
        if not os.path.exists(oracle.REF_O0_SO):
            return {"unavailable": "oracle/_ref/libpsa_ref_O0.so not built"}
        eng = oracle.Ref(oracle.REF_O0_SO)
        wl = make_workload(synth, name, 0, nq=None if name in ("c1", "c2", "c4") else 256)
        seq1 = wl.seq1[: eng.cap1]
        with silence_c_stdout():
            t0 = time.perf_counter()
            done = 0
            for q in wl.queries:
                eng.divide_execute_tasks(wl.weights, wl.is_max, seq1, q, 1, 0, 0, 4)
                done += (len(seq1) - len(q) + 1) * len(q)
                if time.perf_counter() - t0 > seconds:
                    break
            dt = time.perf_counter() - t0
        return {"value": done / dt, "unit": UNIT, "cores": 4, "kind": "reference", "seconds": dt,
                "sample": "reference divide_execute_tasks, -O0 build, 4 OpenMP threads (as shipped), CUDA percentage 0"}
    except Exception as e:       # noqa: BLE001
        return {"unavailable": repr(e)[:200]}


def reference_gpu_probe(workload):
    """The reference's OWN CUDA path (gpu_run_program and its kernels from cuda_funcs.cu, recompiled for sm_100a in
    oracle/_ref) timed on a few queries of the workload, in a child process with a timeout -- "the reference's kernel
    on the same box" of SURVEY 8(d).  Reported baseline only; that path races across blocks."""
    try:
        p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_gpu_probe.py"), workload, "16"],
                           capture_output=True, text=True, timeout=180)
        last = [l for l in p.stdout.splitlines() if l.startswith("{")]
        if p.returncode == 0 and last:
            d = json.loads(last[-1])
            return {"value": d["pair_evals_per_s"], "unit": UNIT, "queries": d["queries"], "seconds": d["seconds"],
                    "answers_differing_from_cpu_reference": d["answers_differing_from_cpu_reference"],
                    "what": "reference gpu_run_program (cudaMalloc + 4 kernels + cudaFree per query), one GPU, "
                            "Seq1 truncated to its 10000 capacity where longer"}
        return {"unavailable": (p.stderr or "no output")[-200:]}
    except Exception as e:       # noqa: BLE001 - a baseline must never take the benchmark down
        return {"unavailable": repr(e)[:200]}


def run_reference_arm(args, synth, rank, world):
    if rank != 0:
        return
    s = reference_sample(synth, args.workload, seconds=max(2.0, 40.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        s["run"](s["sample"])
    t = 0.0
    for _ in range(args.steps):
        t += s["run"](s["sample"])
    value = s["pair_evals"] * args.steps / t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, s["wl"]), "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": s["threads"], "kind": s["kind"], "sample": s["desc"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


FULL_NOTE = {"c1": "reference input.txt block 1 (9711/2131 MIN)", "c2": "single pair len1=3000 len2=2000 MIN",
             "c3": "1024 queries len2=500 vs len1=3000 MAX", "c4": "len1=1000000 len2=2000 MAX",
             "c5": "65536 queries len2=64 vs len1=10000 MIN"}


def workload_name(name, wl):
    return f"{name}: {FULL_NOTE[name]}, weights {wl.weights}, uniform A-Z (splitmix64)"


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def quick_measure(psa, synth, ctx, name, flush_l2, steps):
    wl = make_workload(synth, name, 0)
    batch = psa.Batch(wl.seq1, wl.queries, pinned=True)
    ctx.prepare(wl.weights, wl.is_max, batch)
    for _ in range(3):
        ctx.run()
    ms = 0.0
    for _ in range(steps):
        flush_l2()
        ms += ctx.run()
    wc = psa.c_weights(wl.weights)
    out = ctx.new_result_array(batch.nq, pinned=True)
    ctx.search_batch_raw(wc, wl.is_max, batch, out)
    e2e = 0.0
    for _ in range(steps):
        flush_l2()
        t0 = time.perf_counter()
        ctx.search_batch_raw(wc, wl.is_max, batch, out)
        e2e += time.perf_counter() - t0
    return {"workload": workload_name(name, wl), "pair_evals": batch.pair_evals, "steps": steps,
            "value": batch.pair_evals * steps / (ms * 1e-3), "ms_per_step": ms / steps,
            "e2e": batch.pair_evals * steps / e2e, "e2e_ms_per_step": 1e3 * e2e / steps,
            "engine": {1: "scalar", 2: "bitsliced-scan"}.get(ctx.stat("engine")), "batch_mode": bool(ctx.stat("batch_mode")),
            "slices": ctx.stat("slices"), "exact_integer_keys": bool(ctx.stat("exact"))}


def run_ours(args, synth, rank, local_rank, world):
    import torch
    psa = importlib.import_module(PKG)
    if not torch.cuda.is_available() or psa.device_count() < 1:
        raise SystemExit("bench.py: no B200 visible; the product has no CPU path")
    torch.cuda.set_device(local_rank)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if not use_dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if not use_dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    peaks = load_peaks()
    wl = make_workload(synth, args.workload, rank)
    ctx = psa.Context(devices=[local_rank])
    if args.engine:
        ctx.set_option("engine", args.engine)
    for kv in args.opt:                                                   # library tuning knobs (psa_set_option), for A/B runs
        name, _, val = kv.partition("=")
        ctx.set_option(name, int(val))
    batch = psa.Batch(wl.seq1, wl.queries, pinned=True)
    pair_evals = batch.pair_evals
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def flush_l2():
        flush.zero_()
        torch.cuda.synchronize()

    # ---- value: resident batch, device time --------------------------------------------------------
    ctx.prepare(wl.weights, wl.is_max, batch)
    for _ in range(args.warmup):
        flush_l2()
        ctx.run()
    results = ctx.fetch()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    dev_ms, main_ns, launches = 0.0, 0, 0
    for _ in range(args.steps):
        flush_l2()
        dev_ms += ctx.run()
        launches += ctx.stat("kernel_launches")
    barrier()
    # the same K steps once more with the dominant kernel bracketed by its own events (roofline numerator);
    # kept out of the loop above because an event between two kernels stops them from overlapping
    ctx.set_option("kernel_events", 1)
    dev_ms_bracketed = 0.0
    for _ in range(args.steps):
        flush_l2()
        dev_ms_bracketed += ctx.run()
        main_ns += ctx.stat("main_kernel_ns")
        launches += ctx.stat("kernel_launches")
    ctx.set_option("kernel_events", 0)
    barrier()
    dev_ms_max = max_over_ranks(dev_ms)
    total_pe = sum_over_ranks(float(pair_evals))

    # ---- e2e: host buffers through the public call -------------------------------------------------
    # the C entry point itself (psa_search_batch) on pinned host buffers and a preallocated result array:
    # H2D + table resolution + kernels + D2H + host scoring; no per-result Python objects in the timed region
    wc = psa.c_weights(wl.weights)
    out = ctx.new_result_array(batch.nq, pinned=True)
    for _ in range(min(args.warmup, 3)):
        ctx.search_batch_raw(wc, wl.is_max, batch, out)
    barrier()
    e2e_s = 0.0
    for _ in range(args.steps):
        flush_l2()
        t0 = time.perf_counter()
        ctx.search_batch_raw(wc, wl.is_max, batch, out)
        e2e_s += time.perf_counter() - t0
        launches += ctx.stat("kernel_launches")
    barrier()
    if rank == 0:
        # the timed regions are a few ms in total, shorter than nvidia-smi's sampling period: keep running the same
        # step (untimed) until the sampler has seen the GPU under this load
        t_end = time.perf_counter() + 3.0
        while len(sampler.rows) < 6 and time.perf_counter() < t_end:
            ctx.search_batch_raw(wc, wl.is_max, batch, out)
    clocks = sampler.stop() if rank == 0 else None
    e2e_s_max = max_over_ranks(e2e_s)
    r2 = [ctx.result_from_array(out, i) for i in range(batch.nq)]
    assert [(a.offset, a.char_offset, a.score) for a in r2] == [(a.offset, a.char_offset, a.score) for a in results]

    if rank == 0:
        value = total_pe * args.steps / (dev_ms_max * 1e-3)
        e2e_value = total_pe * args.steps / e2e_s_max
        # roofline of the dominant kernel on this rank (its own events), per GPU
        k_s = main_ns * 1e-9 / args.steps
        achieved = pair_evals / k_s if k_s > 0 else 0.0
        clk = peaks["sm_max_mhz"] * 1e6
        peak = SM_COUNT * 128 * clk / LANE_OPS_PER_PAIR_EVAL
        alg_bytes = batch.len1 + sum(batch.lens) + batch.nq * HBM_BYTES_PER_QUERY_FIXED
        engine = ctx.stat("engine")
        kname = ("k_scan_batch" if ctx.stat("batch_mode") else "k_scan_packed" if ctx.stat("packed_queries") > 0 else "k_scan") if engine == 2 else "k_exact_tiles"
        traffic, prof = profiled_traffic(args.workload, kname)
        # the scan kernel's own bound: its inner loop issues ALU_OPS_PER_WARP_STEP integer-ALU instructions (LOP3/SHF,
        # 64 lanes/clk/SM) per warp per alignment step, and a warp step covers 1024 pair-evals (profiles/)
        kernel_model_peak = SM_COUNT * 64 * clk * 32.0 / ALU_OPS_PER_WARP_STEP
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, wl), "pair_evals_per_gpu_step": pair_evals,
                       "l2": "flushed between timed steps (512 MiB write)", "engine": {1: "scalar", 2: "bitsliced-scan"}.get(engine, engine),
                       "exact_integer_keys": bool(ctx.stat("exact")), "rank_planes": ctx.stat("rank_planes"),
                       "scan_warps": ctx.stat("scan_warps"), "packed_queries_per_block": ctx.stat("packed_queries"), "rescored_words": ctx.stat("candidate_tiles"), "sharding": "one full batch per rank, no collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": batch.h2d_bytes,
                    "d2h_bytes_per_step": 56 * batch.nq, "ms_per_step": 1e3 * e2e_s_max / args.steps,
                    "d2h_path": "result records stored by the finishing threads straight into the caller's page-locked array" if 56 * batch.nq <= 128 * 1024
                                else "one device-to-host copy into the caller's page-locked array"},
            "gpu_launches": launches,
            "roofline": {"bound": "int-alu-issue", "achieved": achieved, "peak": peak, "unit": UNIT, "frac": achieved / peak,
                         "traffic": traffic,
                         "kernel": kname, "kernel_ms": k_s * 1e3,
                         "ncu": None if prof is None else {"capture": "profiles/" + prof["capture"].replace(".ncu-rep", "_metrics.csv"),
                                                           "alu_pipe_pct_of_peak_active": prof["alu_pipe_pct_of_peak_active"],
                                                           "issue_active_pct": prof["issue_active_pct"],
                                                           "dram_throughput_pct": prof["dram_throughput_pct"]},
                         "kernel_share_of_step": (main_ns * 1e-6) / dev_ms_bracketed if dev_ms_bracketed else None,
                         "measured_int_lane_ops_per_clk_per_sm": {"LOP3": 64.0, "IADD3": 64.0, "SHF": 64.0, "IMAD(fma pipe)": 64.0,
                                                                  "source": "tools/probes/alu_rate_probe.cu on this pool's B200"},
                         "model": "SURVEY 8(d): 2 int32 lane-ops per pair-eval, 128 lanes/clk/SM (both integer-capable pipes) x 148 SMs x "
                                  f"{peaks['sm_max_mhz']:.0f} MHz ({peaks['source']} clock); not HBM, not tensor. frac > 1 is "
                                  "possible because the bit-sliced kernel spends 0.32 ALU lane-ops per pair-eval, not 2",
                         "kernel_model": {"alu_lane_ops_per_pair_eval": ALU_OPS_PER_WARP_STEP / 32.0, "peak": kernel_model_peak,
                                          "frac": achieved / kernel_model_peak,
                                          "note": "inner-loop bound only (no epilogue, no idle lanes); ncu "
                                                  "sm__inst_executed_pipe_alu of the same kernel is in profiles/"},
                         "hbm": {"algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / k_s / 1e9 if k_s else 0.0,
                                 "peak_gbs": peaks["hbm_gbs"], "frac": (alg_bytes / k_s / 1e9) / peaks["hbm_gbs"] if k_s else 0.0,
                                 "peak_source": peaks["source"]}},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            s = reference_sample(synth, args.workload, seconds=args.cpu_seconds)
            t = s["run"](s["sample"])
            line["cpu_baseline"] = {"value": s["pair_evals"] / t, "unit": UNIT, "cores": s["threads"], "kind": s["kind"],
                                    "sample": s["desc"], "seconds": t}
            line["cpu_baseline_as_shipped"] = as_shipped_baseline(synth, args.workload)
            line["reference_gpu"] = reference_gpu_probe(args.workload)
        if world == 1 and not args.no_others:
            # the other BASELINE.json configs, same method (resident value + host-buffer e2e), fewer steps
            line["other_workloads"] = {}
            for name in ("c1", "c2", "c4", "c5"):
                if name != args.workload:
                    line["other_workloads"][name] = quick_measure(psa, synth, ctx, name, flush_l2, steps=max(3, args.steps // 4))
        print(json.dumps(line), flush=True)
    ctx.close()
    if use_dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--engine", type=int, default=0, help="0 auto, 1 scalar, 2 bit-sliced scan")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE", help="psa_set_option knob, repeatable (A/B runs)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the quick measurements of the other BASELINE configs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, local_rank, world = dist_env()
    synth = importlib.import_module(PKG + ".synth")
    if args.impl == "reference":
        run_reference_arm(args, synth, rank, world)
    else:
        run_ours(args, synth, rank, local_rank, world)


if __name__ == "__main__":
    main()
