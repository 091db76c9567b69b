#!/usr/bin/env python
"""bench.py -- pair-evals/s of the mutant-offset search on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c5|c4|c2|c1] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic queries of the named BASELINE.json shape.

  value      one process per GPU, every rank owns a full batch (different seed per rank), no data-path collective (weak
             scaling): batch resident in HBM (psa_batch_prepare), per-step device time from CUDA events recorded by the
             library on the stream it launches on (psa_batch_run), L2 flushed between steps; pair-evals of all ranks /
             max-over-ranks time
  e2e        the public call psa_search_batch with HOST (pinned) buffers: H2D + kernels + results back, wall clock per step
  strong     north_star's split: ONE process (rank 0) drives all N GPUs through one psa_context -- query blocks for configs 3
             and 5, offset ranges for config 4 (the reference's rank split, cpu_funcs.c:128-133) -- end to end through
             psa_search_batch, against the same call on a 1-GPU context in the same process; 64 answers per workload are
             checked against the oracle; the library's own host-side split of a call names the limiter
  roofline   the dominant kernel alone (its own CUDA events) against the integer-issue roofline of SURVEY.md section 8(d) and
             against the kernel's own integer-ALU bound (inner-loop instruction mix and pipe rates read from profiles/);
             HBM is shown to be non-binding
  cpu_baseline / --impl reference
             the reference's own CPU loop (oracle/_ref, compiled from /root/reference) on a bounded sample of the same
             workload, all host threads; nested: the reference as shipped (-O0, 4 threads), its own CUDA path on one GPU,
             and the emulated `mpiexec -np 2` CUDA+OpenMP run
"""
from __future__ import annotations

import argparse
import contextlib
import ctypes
import importlib
import importlib.util
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
SM_COUNT = 148
LANE_OPS_PER_PAIR_EVAL = 2.0                 # one packed-counter add + one best-rank max per pair-eval
HBM_BYTES_PER_QUERY_FIXED = 56               # one result record per query (stripe mode leaves nothing else in HBM)
ROUND = "r02"


def load_synth():
    """synth.py by file path: the reference arm must not import the product package (its __init__ maps libpsa_b200.so)."""
    spec = importlib.util.spec_from_file_location("psa_synth", os.path.join(ROOT, PKG, "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["psa_synth"] = mod
    spec.loader.exec_module(mod)
    return mod


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "sm_max_mhz": d.get("sm_max_mhz", 1965.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}


def load_profile(workload: str, kernel: str):
    """Counters of the committed ncu capture of this workload's dominant kernel (profiles/rNN_summary.json, written by
    tools/ncu_summary.py + tools/ncu_regions.py): DRAM bytes per launch, pipe utilisation, and the instruction mix of the
    inner loop.  Newest round first; None if the kernel was never captured."""
    for rnd in (ROUND, "r01"):
        p = os.path.join(ROOT, "profiles", f"{rnd}_summary.json")
        if not os.path.exists(p):
            continue
        merged = {}
        for name, v in json.load(open(p)).get(workload, {}).items():
            if name == kernel or name.startswith(kernel + "<"):
                merged.update(v)
        if merged:
            merged["summary_file"] = f"profiles/{rnd}_summary.json"
            return merged
    return None


def load_pipe_rates():
    """Measured integer lane-ops per clock per SM (tools/probes/alu_rate_probe.cu via tools/run_probes.sh)."""
    for rnd in (ROUND, "r01"):
        p = os.path.join(ROOT, "profiles", f"{rnd}_probes.json")
        if os.path.exists(p):
            d = json.load(open(p))
            d["file"] = f"profiles/{rnd}_probes.json"
            return d
    return {"lane_ops_per_clk_per_sm": {"LOP3": 64.0, "IADD3": 64.0, "SHF": 64.0, "IMAD": 64.0},
            "file": None, "source": "round-1 run of tools/probes/alu_rate_probe.cu (no probes file committed)"}


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


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
FULL_NOTE = {"c1": "reference input.txt block 1 (len1=9711 len2=2131 MIN), the reference's own letters",
             "c2": "single pair len1=3000 len2=2000 MIN, uniform A-Z (splitmix64)",
             "c3": "1024 queries len2=500 vs len1=3000 MAX, uniform A-Z (splitmix64)",
             "c4": "len1=1000000 len2=2000 MAX, uniform A-Z (splitmix64)",
             "c5": "65536 queries len2=64 vs len1=10000 MIN, uniform A-Z (splitmix64)"}


def golden_c1():
    with open(os.path.join(ROOT, "tests", "golden", "input_blocks.json")) as f:
        b = json.load(f)[0]
    return b["weights"], b["goal"] == "maximum", b["seq1"].encode(), [b["seq2"].encode()]


def make_workload(synth, name, rank, nq=None, variant=0):
    """variant: another batch of the same shape and distribution (the pipelined end-to-end leg takes a different one per step)."""
    if name == "c1":
        w, is_max, s1, qs = golden_c1()
        return synth.Workload("c1", w, is_max, s1, qs, FULL_NOTE["c1"])
    return synth.workload(name, nq=nq, seed_shift=1000 * rank + variant)


def workload_name(name, wl):
    return f"{name}: {FULL_NOTE[name]}, weights {[float(x) for x in wl.weights]}"


def shared_config(name, wl):
    """The `config` object: the SAME in our arm and in the reference arm (what is measured, not how)."""
    return {"workload": workload_name(name, wl),
            "per_gpu": "our arm: every rank owns one full batch of the workload (weak scaling, no data-path collective); "
                       "reference arm: a bounded sample of the same batch on the host cores",
            "l2": "our arm: flushed between timed steps (512 MiB write); reference arm: CPU, not applicable",
            "timing": "our arm: `value` = per-step CUDA events on the library's stream; event + launch + event are enqueued behind a "
                      "gate the host opens afterwards (psa_set_option gate_timed_runs), so the bracket holds device time and no host "
                      "launch latency (tools/probes/launch_probe.cu, profiles/r02_launch_probe.txt); `e2e` = wall clock around "
                      "psa_search_batch on host buffers; reference arm: wall clock"}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baselines: the reference's own code on a bounded sample
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
    what = ("the reference's find_best_mutant_cpu (cpu_funcs.c:222-300) with its own thread split (cpu_funcs.c:192-197) over "
            f"{threads} threads, built -O3, table filled single-threaded -- FASTER than the reference as shipped (-O0, 4 threads: see "
            "cpu_baseline.as_shipped), so ratios against it are conservative" if kind == "reference" else
            f"oracle port over {threads} threads")
    return {"kind": kind, "threads": threads, "run": run, "sample": sample, "pair_evals": pe(sample),
            "desc": f"{n} of the workload's queries; {what}{cap_note}", "wl": wl}


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


def reference_gpu_probe(workload, mode):
    """The reference's OWN CUDA path (gpu_run_program and its kernels from cuda_funcs.cu, recompiled for sm_100a in
    oracle/_ref), in a child process with a timeout.  mode "rank1": one rank, all offsets on one GPU; mode "np2": the
    emulated `mpiexec -np 2` CUDA+OpenMP run (two host threads as the two ranks, divide_execute_tasks(&data, 2, pid) each,
    GPU pid % visible, MAXLOC/MINLOC merge).  Reported baselines only; that path races across blocks."""
    try:
        p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_gpu_probe.py"), workload, "16", mode],
                           capture_output=True, text=True, timeout=240)
        last = [l for l in p.stdout.splitlines() if l.startswith("{")]
        if p.returncode == 0 and last:
            d = json.loads(last[-1])
            d["unit"] = UNIT
            return d
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
            "config": shared_config(args.workload, s["wl"]),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": s["threads"], "kind": s["kind"], "sample": s["desc"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def oracle_check(wl, results, n=64):
    """The timed answers against the oracle on the first n queries (bit-exact: offset, char_offset, letter, score)."""
    import oracle
    port = oracle.Port()
    k = min(n, len(wl.queries))
    exp = port.search_batch(wl.weights, wl.is_max, wl.seq1, wl.queries[:k])
    bad = [i for i, (g, e) in enumerate(zip(results[:k], exp))
           if (g.offset, g.char_offset, g.ch, g.score) != (e.offset, e.char_offset, e.ch, e.score)]
    if bad:
        raise SystemExit(f"bench.py: {len(bad)} of {k} timed answers differ from the oracle (first: query {bad[0]}: "
                         f"{results[bad[0]]} vs {exp[bad[0]]})")
    return {"queries_checked": k, "mismatches": 0, "checker": "oracle/psa_oracle.c (pinned to the compiled reference), bit-exact"}


def kernel_name(ctx):
    if ctx.stat("engine") != 2:
        return "k_exact_tiles"
    if ctx.stat("stripe_mode"):
        return "k_stripe"
    if ctx.stat("single_launch"):
        return "k_single"
    return "k_scan_batch" if ctx.stat("batch_mode") else "k_scan_packed" if ctx.stat("packed_queries") > 0 else "k_scan"


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
    # the same problem as a pipelined list (psa_search_many): 8 lanes for the single-query configs, 2 for the batches
    lanes = 8 if batch.nq == 1 else 2
    n_list = 128 if batch.nq == 1 else max(8, 2 * steps)
    outs = [ctx.new_result_array(batch.nq, pinned=True) for _ in range(min(n_list, 16))]
    plist = ctx.make_problem_list([(wc, wl.is_max, batch, outs[k % len(outs)]) for k in range(n_list)])
    ctx.search_many_raw(plist, n_list, lanes)
    flush_l2()
    t0 = time.perf_counter()
    ctx.search_many_raw(plist, n_list, lanes)
    t_list = time.perf_counter() - t0
    return {"workload": workload_name(name, wl), "pair_evals": batch.pair_evals, "steps": steps,
            "value": batch.pair_evals * steps / (ms * 1e-3), "ms_per_step": ms / steps,
            "e2e": batch.pair_evals * steps / e2e, "e2e_ms_per_step": 1e3 * e2e / steps,
            "e2e_list": {"value": batch.pair_evals * n_list / t_list, "ms_per_step": 1e3 * t_list / n_list, "problems": n_list, "lanes": lanes,
                         "how": "the same problem n times as one psa_search_many list (every entry pays its own copies), wall clock around the list"},
            "kernel": kernel_name(ctx), "launches_per_step": ctx.stat("kernel_launches"),
            "slices": ctx.stat("slices"), "exact_integer_keys": bool(ctx.stat("exact"))}


def strong_scaling(psa, synth, torch, ngpus, steps, workloads=("c3", "c5", "c4")):
    """north_star's multi-GPU form, measured in THIS process: one psa_context over `ngpus` devices (worker thread, stream and
    buffers per device; contiguous query blocks, or offset ranges for a single query; host merge) against a 1-device context,
    end to end through psa_search_batch with pinned host buffers.  Efficiency = t(1) / (N t(N))."""
    flush = [torch.empty(512 << 20, dtype=torch.uint8, device=f"cuda:{g}") for g in range(ngpus)]

    def flush_l2():
        for g, f in enumerate(flush):
            f.zero_()
        for g in range(ngpus):
            torch.cuda.synchronize(g)

    rec = {"n_gpus": ngpus, "how": "rank 0 alone: psa_create over devices 0..N-1, psa_search_batch on pinned host buffers, wall clock per call, "
                                   "L2 of every GPU flushed between calls; t(1) from a 1-device context in the same process.  A call is spread "
                                   "over at most (its pair evaluations / 2.5e9) GPUs (`gpus_used`): a 40 us problem is not worth eight "
                                   "shards' copies and launches; streams of such problems scale as `list_of_batches` (psa_search_many)",
           "workloads": {}}
    ctxs = {1: psa.Context(devices=[0])}
    if ngpus > 1:
        ctxs[ngpus] = psa.Context(devices=list(range(ngpus)))
    try:
        for name in workloads:
            wl = make_workload(synth, name, 0)
            batch = psa.Batch(wl.seq1, wl.queries, pinned=True)
            wc = psa.c_weights(wl.weights)
            out = psa.Context.new_result_array(batch.nq, pinned=True)
            w = {"workload": workload_name(name, wl), "pair_evals": batch.pair_evals,
                 "split": "offset ranges of the one query (cpu_funcs.c:128-133 with GPUs for ranks)" if batch.nq == 1 else "contiguous query blocks"}
            for n, c in ctxs.items():
                for _ in range(3):
                    c.search_batch_raw(wc, wl.is_max, batch, out)
                t = 0.0
                split = {"host_plan_ns": 0, "host_prepare_ns": 0, "host_enqueue_ns": 0, "host_wait_ns": 0, "host_total_ns": 0}
                for _ in range(steps):
                    flush_l2()
                    t0 = time.perf_counter()
                    c.search_batch_raw(wc, wl.is_max, batch, out)
                    t += time.perf_counter() - t0
                    for k in split:
                        split[k] += c.stat(k)
                res = [c.result_from_array(out, i) for i in range(min(64, batch.nq))]
                chk = oracle_check(wl, res, 64) if name != "c4" else c4_check(psa, c, wl, res)
                # the kernels alone on the same context (resident batch, CUDA events, max over the GPUs)
                c.prepare(wl.weights, wl.is_max, batch)
                c.run()
                dev = 0.0
                for _ in range(steps):
                    flush_l2()
                    dev += c.run()
                w[f"n{n}"] = {"e2e_ms": 1e3 * t / steps, "e2e_pair_evals_per_s": batch.pair_evals * steps / t,
                              "device_ms_max_over_gpus": dev / steps, "device_pair_evals_per_s": batch.pair_evals * steps / (dev * 1e-3),
                              "host_split_us": {k[5:-3]: v / steps * 1e-3 for k, v in split.items()},
                              "launches_per_call": c.stat("kernel_launches"), "gpus_used": c.stat("devices_used"), "oracle_check": chk}
            # The same workload as a LIST of independent batches through psa_search_many (2 lanes per GPU, a batch is never
            # split): what a stream of such batches gets out of the N GPUs of one process, end to end on host buffers.
            if name != "c4":
                per_gpu = {"c3": 48, "c5": 12}.get(name, 8)
                distinct = {"c3": 16, "c5": 3}.get(name, 4)
                wls = [wl] + [make_workload(synth, name, 0, variant=k) for k in range(1, distinct)]
                built = [(wc, x.is_max, batch if k == 0 else psa.Batch(x.seq1, x.queries, pinned=True),
                          psa.Context.new_result_array(batch.nq, pinned=True)) for k, x in enumerate(wls)]
                many = {"how": "psa_search_many, 2 lanes per GPU, wall clock around the whole list (every batch pays its own copies)",
                        "distinct_batches": distinct}
                for n, c in ctxs.items():
                    items = [built[k % distinct] for k in range(per_gpu * n)]
                    plist = c.make_problem_list(items)
                    c.search_many_raw(plist, min(len(items), 4 * n), 2)
                    best = None
                    for _ in range(3):
                        flush_l2()
                        t0 = time.perf_counter()
                        c.search_many_raw(plist, len(items), 2)
                        dt = time.perf_counter() - t0
                        best = dt if best is None else min(best, dt)
                    last = (len(items) - 1) % distinct
                    chk = oracle_check(wls[last], [c.result_from_array(built[last][3], i) for i in range(min(32, batch.nq))], 32)
                    many[f"n{n}"] = {"batches": len(items), "ms_per_batch": 1e3 * best / len(items),
                                     "e2e_pair_evals_per_s": batch.pair_evals * len(items) / best, "oracle_check": chk}
                if ngpus > 1:
                    many["speedup"] = many[f"n{ngpus}"]["e2e_pair_evals_per_s"] / many["n1"]["e2e_pair_evals_per_s"]
                    many["efficiency"] = many["speedup"] / ngpus
                w["list_of_batches"] = many
            if ngpus > 1:
                a, b = w["n1"], w[f"n{ngpus}"]
                w["speedup_e2e"] = a["e2e_ms"] / b["e2e_ms"]
                w["efficiency_e2e"] = a["e2e_ms"] / (ngpus * b["e2e_ms"])
                w["efficiency_device"] = a["device_ms_max_over_gpus"] / (ngpus * b["device_ms_max_over_gpus"])
                hs = b["host_split_us"]
                kern = b["device_ms_max_over_gpus"] * 1e3
                other = max(hs["total"] - kern, 0.0)
                w["limiter"] = (f"of {hs['total']:.0f} us per call on {ngpus} GPUs the kernels are {kern:.0f} us; the rest ({other:.0f} us) is host side: "
                                f"planning {hs['plan']:.0f} us on the calling thread, then per GPU thread H2D enqueue {hs['prepare']:.0f} us, "
                                f"launches {hs['enqueue']:.0f} us, wait for the stream + results {hs['wait']:.0f} us (copies in both directions "
                                f"ride inside the wait)")
            rec["workloads"][name] = w
    finally:
        for c in ctxs.values():
            c.close()
    return rec


def c4_check(psa, ctx, wl, res):
    """Config 4 is one query over a million offsets: the oracle takes seconds on all host threads, so the check is the
    reference's own split instead -- the whole range on this context against two half ranges merged MAXLOC-style -- plus the
    oracle's score of the winning offset."""
    import oracle
    port = oracle.Port()
    r = res[0]
    n = port.offset_naive(wl.weights, wl.is_max, wl.seq1, wl.queries[0], r.offset)
    ok = (n.score, n.char_offset, n.ch) == (r.score, r.char_offset, r.ch)
    total = len(wl.seq1) - len(wl.queries[0]) + 1
    halves = [ctx.search_range(wl.weights, wl.is_max, wl.seq1, wl.queries[0], 0, total // 2),
              ctx.search_range(wl.weights, wl.is_max, wl.seq1, wl.queries[0], total // 2, total)]
    m = psa.merge_results(wl.is_max, halves)
    ok = ok and (m.offset, m.char_offset, m.score) == (r.offset, r.char_offset, r.score)
    if not ok:
        raise SystemExit(f"bench.py: config 4 answer {r} fails the oracle / partition check ({n}, {m})")
    return {"queries_checked": 1, "mismatches": 0,
            "checker": "oracle score + letter of the winning offset, and invariance under the reference's two-rank offset split"}


def run_ours(args, synth, rank, local_rank, world):
    import torch
    psa = importlib.import_module(PKG)
    if not torch.cuda.is_available() or psa.device_count() < 1:
        raise SystemExit("bench.py: no B200 visible; the product has no CPU path")
    torch.cuda.set_device(local_rank)
    use_dist = world > 1
    cpu_group = None
    if use_dist:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")       # host-side waits that leave the GPUs alone (strong-scaling leg)

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
    ctx.set_option("gate_timed_runs", 1)
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
    timed_results = ctx.fetch()                                           # the answers of the last timed step
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
    # ---- e2e, pipelined: the K steps as ONE list through psa_search_many ---------------------------------------------
    # Step k is its own batch of the workload's shape (a different one per step, up to `distinct`), in its own pinned host
    # buffers with its own pinned result array; psa_search_many runs the list over 2 lanes (stream + buffers + host thread
    # each), so the copies, launch and wake-up of one step overlap the kernel of another.  Every step still pays its own
    # H2D copies and gets its own records back; the wall clock is taken around the whole list.
    distinct = min(args.steps, {"c3": 32, "c1": 32, "c2": 32}.get(args.workload, 6))
    many_wl = [wl] + [make_workload(synth, args.workload, rank, variant=k) for k in range(1, distinct)]
    many_built = [(wc, w_.is_max, batch if k == 0 else psa.Batch(w_.seq1, w_.queries, pinned=True),
                   ctx.new_result_array(batch.nq, pinned=True)) for k, w_ in enumerate(many_wl)]
    many_items = [many_built[k % distinct] for k in range(args.steps)]
    many_list = ctx.make_problem_list(many_items)
    lanes = 2
    ctx.search_many_raw(many_list, args.steps, lanes)       # warm-up: the same list once (lanes created, kernels loaded, every buffer seen by the driver)
    flush_l2()
    barrier()
    t0 = time.perf_counter()
    ctx.search_many_raw(many_list, args.steps, lanes)
    many_s = time.perf_counter() - t0
    barrier()
    launches += args.steps * ctx.stat("kernel_launches")
    many_s_max = max_over_ranks(many_s)
    last = (args.steps - 1) % distinct
    many_last = (many_wl[last], [ctx.result_from_array(many_built[last][3], i) for i in range(min(64, batch.nq))])
    if rank == 0:
        # the timed regions are a few ms in total, shorter than nvidia-smi's sampling period: keep running the same
        # step (untimed) until the sampler has seen the GPU under this load
        t_end = time.perf_counter() + 3.0
        while len(sampler.rows) < 6 and time.perf_counter() < t_end:
            ctx.search_batch_raw(wc, wl.is_max, batch, out)
    clocks = sampler.stop() if rank == 0 else None
    e2e_s_max = max_over_ranks(e2e_s)
    r2 = [ctx.result_from_array(out, i) for i in range(batch.nq)]
    key = lambda a: (a.offset, a.char_offset, a.ch, a.score)
    assert [key(a) for a in r2] == [key(a) for a in results] == [key(a) for a in timed_results]
    kname = kernel_name(ctx)
    engine = ctx.stat("engine")
    launch_shape = {"engine": {1: "scalar", 2: "bitsliced-scan"}.get(engine, engine), "kernel": kname,
                    "launches_per_step": ctx.stat("kernel_launches"), "exact_integer_keys": bool(ctx.stat("exact")),
                    "rank_planes": ctx.stat("rank_planes"), "stripe_mode": bool(ctx.stat("stripe_mode")),
                    "stripe_lanes_per_query": ctx.stat("stripe_lanes"), "stripe_queries_per_task": ctx.stat("stripe_queries_per_task"),
                    "stripe_warps_per_team": ctx.stat("stripe_team_warps"), "stripe_teams_per_block": ctx.stat("stripe_teams"),
                    "packed_queries_per_block": ctx.stat("packed_queries"), "rescored_words": ctx.stat("candidate_tiles")}
    ctx.close()

    # ---- strong scaling: rank 0 alone drives all N GPUs through ONE context (north_star's split) -----------------------
    strong = None
    if not args.no_strong:
        if use_dist:
            dist.barrier(group=cpu_group)                 # everyone is done with its own GPU
        if rank == 0:
            strong = strong_scaling(psa, synth, torch, world, steps=max(5, args.steps // 2))
        if use_dist:
            dist.barrier(group=cpu_group)                 # ranks > 0 wait on the host, their GPUs stay idle meanwhile

    if rank == 0:
        value = total_pe * args.steps / (dev_ms_max * 1e-3)
        e2e_value = total_pe * args.steps / e2e_s_max
        many_value = total_pe * args.steps / many_s_max
        check = oracle_check(wl, timed_results, 64) if args.workload != "c4" else {"queries_checked": 0, "note": "see strong.workloads.c4"}
        many_check = oracle_check(many_last[0], many_last[1], 64) if args.workload != "c4" else {"queries_checked": 0, "note": "see strong.workloads.c4"}
        # roofline of the dominant kernel on this rank (its own events), per GPU
        k_s = main_ns * 1e-9 / args.steps
        achieved = pair_evals / k_s if k_s > 0 else 0.0
        clk = peaks["sm_max_mhz"] * 1e6
        peak = SM_COUNT * 128 * clk / LANE_OPS_PER_PAIR_EVAL
        alg_bytes = batch.len1 + sum(batch.lens) + batch.nq * HBM_BYTES_PER_QUERY_FIXED
        prof = load_profile(args.workload, kname)
        rates = load_pipe_rates()
        alu_rate = min(rates["lane_ops_per_clk_per_sm"].get("LOP3", 64.0), rates["lane_ops_per_clk_per_sm"].get("SHF", 64.0))
        # the scan kernel's own bound: its inner loop issues `alu_pipe_instr` integer-ALU instructions per warp per 32
        # alignment steps (measured mix of the committed capture), at the measured ALU-pipe rate; a warp step is 1024 pair-evals
        inner = (prof or {}).get("inner_loop")
        kernel_model = None
        if inner:
            per_step = inner["alu_pipe_instr"] / float(inner.get("steps_per_group", 32))
            km_peak = SM_COUNT * alu_rate * clk * 32.0 / per_step
            kernel_model = {"alu_warp_instr_per_32_steps": inner["alu_pipe_instr"], "alu_lane_ops_per_pair_eval": per_step / 32.0,
                            "alu_lane_ops_per_clk_per_sm": alu_rate, "peak": km_peak, "frac": achieved / km_peak,
                            "source": f"{prof['summary_file']} (inner_loop of {prof.get('capture')}) and {rates.get('file') or rates.get('source')}",
                            "note": "inner-loop bound only: no window build, no epilogue, no idle lanes, perfect balance over 592 schedulers"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic",
            "config": shared_config(args.workload, wl),
            "e2e": {"value": many_value, "unit": UNIT, "h2d_bytes_per_step": batch.h2d_bytes,
                    "d2h_bytes_per_step": 56 * batch.nq, "ms_per_step": 1e3 * many_s_max / args.steps,
                    "how": f"the K steps as ONE list through the C entry point psa_search_many ({lanes} lanes per GPU: stream, buffers and host "
                           f"thread each): step k is its own batch of the workload's shape ({distinct} distinct batches, each in its own pinned host "
                           "buffers) and pays its own host-to-device copies and gets its own records back; the copies, launch and wake-up of one "
                           "step overlap the kernel of another; wall clock around the whole list, L2 flushed before it",
                    "d2h_path": "result records stored by the finishing warps straight into the caller's page-locked arrays (stripe mode: one "
                                "56-byte write per record)",
                    "oracle_check_of_timed_answers": many_check,
                    "one_call_at_a_time": {"value": e2e_value, "ms_per_step": 1e3 * e2e_s_max / args.steps,
                                           "how": "psa_search_batch on the same pinned buffers, one synchronous call per step, L2 flushed "
                                                  "between calls, wall clock per call (the latency of a single batch)",
                                           "oracle_check_of_timed_answers": check},
                    "strong": strong},
            "gpu_launches": launches,
            "roofline": {"bound": "int-alu-issue", "achieved": achieved, "peak": peak, "unit": UNIT, "frac": achieved / peak,
                         "traffic": (prof or {}).get("dram_bytes_per_launch"),
                         "kernel": kname, "kernel_ms": k_s * 1e3, "pair_evals_per_launch": pair_evals, "launch": launch_shape,
                         "ncu": None if prof is None else {"capture": prof.get("capture"), "summary": prof["summary_file"],
                                                           "alu_pipe_pct_of_peak_active": prof.get("alu_pipe_pct_of_peak_active"),
                                                           "alu_pipe_pct_of_peak_active_busiest_sm": prof.get("alu_pipe_pct_of_peak_active_max_sm"),
                                                           "issue_active_pct": prof.get("issue_active_pct"),
                                                           "smem_wavefronts_pct_of_peak": prof.get("smem_wavefronts_pct_of_peak"),
                                                           "dram_throughput_pct": prof.get("dram_throughput_pct")},
                         "kernel_share_of_step": (main_ns * 1e-6) / dev_ms_bracketed if dev_ms_bracketed else None,
                         "measured_int_lane_ops_per_clk_per_sm": rates,
                         "model": "SURVEY 8(d): 2 int32 lane-ops per pair-eval, 128 lanes/clk/SM (both integer-capable pipes) x 148 SMs x "
                                  f"{peaks['sm_max_mhz']:.0f} MHz ({peaks['source']} clock); not HBM, not tensor. frac > 1 is possible because the "
                                  "bit-sliced kernel spends ~0.25 ALU lane-ops per pair-eval, not 2",
                         "kernel_model": kernel_model,
                         "hbm": {"algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / k_s / 1e9 if k_s else 0.0,
                                 "peak_gbs": peaks["hbm_gbs"], "frac": (alg_bytes / k_s / 1e9) / peaks["hbm_gbs"] if k_s else 0.0,
                                 "peak_source": peaks["source"]}},
            "clocks": clocks,
            "strong": strong,
        }
        if world == 1 and not args.no_cpu_baseline:
            s = reference_sample(synth, args.workload, seconds=args.cpu_seconds)
            t = s["run"](s["sample"])
            line["cpu_baseline"] = {"value": s["pair_evals"] / t, "unit": UNIT, "cores": s["threads"], "kind": s["kind"],
                                    "sample": s["desc"], "seconds": t,
                                    "as_shipped": as_shipped_baseline(synth, args.workload),
                                    "reference_gpu": reference_gpu_probe(args.workload, "rank1"),
                                    "reference_np2": reference_gpu_probe(args.workload, "np2")}
        if world == 1 and not args.no_others and kernel_model and args.workload == "c3":
            # The same kernel on 16 of these batches at once: the fixed cost of a launch (launch ramp, table, window build,
            # one partly filled wave of tasks) no longer weighs on a 40 us kernel, which shows what the inner loop reaches
            with psa.Context(devices=[local_rank]) as c3x:
                big = make_workload(synth, "c3", rank, nq=16 * batch.nq)
                bb = psa.Batch(big.seq1, big.queries, pinned=True)
                c3x.set_option("kernel_events", 1)
                c3x.set_option("gate_timed_runs", 1)
                c3x.prepare(big.weights, big.is_max, bb)
                for _ in range(3):
                    c3x.run()
                ns, n_big = 0, max(3, args.steps // 4)
                for _ in range(n_big):
                    flush_l2()
                    c3x.run()
                    ns += c3x.stat("main_kernel_ns")
                big_check = oracle_check(big, c3x.fetch(), 32)
                a_big = bb.pair_evals * n_big / (ns * 1e-9)
                kernel_model["at_scale"] = {"workload": f"{bb.nq} queries of the same shape in one call ({kernel_name(c3x)}, "
                                                        f"{c3x.stat('kernel_launches')} launch)",
                                            "kernel_ms": ns * 1e-6 / n_big, "achieved": a_big, "frac": a_big / kernel_model["peak"],
                                            "oracle_check": big_check}
        if world == 1 and not args.no_others:
            # the other BASELINE.json configs, same method (resident value + host-buffer e2e), fewer steps
            others = {}
            with psa.Context(devices=[local_rank]) as c2:
                c2.set_option("gate_timed_runs", 1)
                for name in ("c1", "c2", "c4", "c5"):
                    if name != args.workload:
                        others[name] = quick_measure(psa, synth, c2, name, flush_l2, steps=max(3, args.steps // 4))
            line["other_workloads"] = others
            line["roofline"]["other_workloads"] = {k: {"value": v["value"], "ms_per_step": v["ms_per_step"], "e2e_ms_per_step": v["e2e_ms_per_step"],
                                                       "kernel": v["kernel"]} for k, v in others.items()}
        print(json.dumps(line), flush=True)
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
    ap.add_argument("--no-strong", action="store_true", help="skip the single-process multi-GPU (strong scaling) leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, local_rank, world = dist_env()
    synth = load_synth()
    if args.impl == "reference":
        run_reference_arm(args, synth, rank, world)
    else:
        run_ours(args, synth, rank, local_rank, world)


if __name__ == "__main__":
    main()
