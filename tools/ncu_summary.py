#!/usr/bin/env python
"""Turn .ncu-rep captures (gpurun_out/) into the small text/CSV/JSON summaries committed under profiles/.

    python tools/ncu_summary.py gpurun_out/r01_c3_k_scan.ncu-rep c3 [more.ncu-rep workload ...]

Writes profiles/<stem>_details.txt, profiles/<stem>_metrics.csv and merges the headline counters of every
capture into profiles/r01_summary.json (read by bench.py for roofline.traffic)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__inst_executed_pipe_alu",
        "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_lsu", "smsp__issue_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared", "sm__throughput", "dram__throughput", "launch__", "sm__warps_active",
        "stall", "sm__cycles_elapsed.max", "sm__cycles_active.avg", "lts__t_bytes.sum")


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main(argv):
    out_dir = os.path.join(ROOT, "profiles")
    if argv and argv[0] == "--out":          # e.g. gpurun_out/profiles on the GPU box (only gpurun_out/ travels back)
        out_dir = argv[1]
        argv = argv[2:]
        os.makedirs(out_dir, exist_ok=True)
    summary_path = os.path.join(out_dir, os.environ.get("PSA_ROUND", "r02") + "_summary.json")
    summary = json.load(open(summary_path)) if os.path.exists(summary_path) else {}
    for rep, workload in zip(argv[0::2], argv[1::2]):
        stem = os.path.splitext(os.path.basename(rep))[0]
        details = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
        open(os.path.join(out_dir, stem + "_details.txt"), "w").write(details)
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        rec = dict(zip(hdr, vals))
        unit = dict(zip(hdr, units))
        with open(os.path.join(out_dir, stem + "_metrics.csv"), "w") as f:
            for h in hdr:
                if any(k in h for k in KEEP) and rec[h] not in ("", "0"):
                    f.write(f"{h},{unit[h]},{rec[h]}\n")
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        dram = sum((num(rec.get(k, "0")) or 0) * scale.get(unit.get(k, "byte"), 1)
                   for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        tscale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}
        name = rec.get("Kernel Name", stem).split("(")[0]
        name = name.replace("void ", "").replace("psa::", "").replace("<unnamed>::", "").replace("unnamed>::", "").strip()
        summary.setdefault(workload, {}).setdefault(name, {}).update({
            "capture": os.path.basename(rep),
            "dram_bytes_per_launch": dram,
            "duration_s_under_ncu": (num(rec.get("gpu__time_duration.sum", "0")) or 0) * tscale.get(unit.get("gpu__time_duration.sum", "ns"), 1e-9),
            "alu_pipe_pct_of_peak_active": num(rec.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "")),
            "issue_active_pct": num(rec.get("smsp__issue_active.avg.pct_of_peak_sustained_active", "")),
            "dram_throughput_pct": num(rec.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "")),
            "smem_wavefronts_pct_of_peak": num(rec.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "")),
            "registers_per_thread": num(rec.get("launch__registers_per_thread", "")),
            "warp_instructions": num(rec.get("smsp__inst_executed.sum", "")),
            "alu_pipe_pct_of_peak_active_max_sm": num(rec.get("sm__inst_executed_pipe_alu.max.pct_of_peak_sustained_active", "")),
            "sm_cycles_active_avg": num(rec.get("sm__cycles_active.avg", "")),
            "sm_cycles_elapsed_max": num(rec.get("sm__cycles_elapsed.max", "")),
        })
    json.dump(summary, open(summary_path, "w"), indent=1, sort_keys=True)
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main(sys.argv[1:])
