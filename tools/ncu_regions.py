#!/usr/bin/env python
"""Split an `ncu --page source --csv` export (optionally .gz) into runs of equal execution count (= code regions:
prologue, unrolled loop bodies, epilogue ...) and print each run's share of executed instructions and of the
warp-state samples, with its dominant stall reasons.

    python tools/ncu_regions.py gpurun_out/c3p_source.csv.gz [min_share_pct] [--json summary.json workload]

With --json the hottest region (largest share of executed instructions: the unrolled 32-step counting group) is also
recorded in the summary file under [workload][kernel]["inner_loop"]: its SASS instruction count split by issue pipe.
bench.py derives the kernel's own integer-ALU bound from that record instead of from a constant."""
import csv
import gzip
import io
import json
import os
import sys

ALU_PIPE = {"LOP3", "SHF", "IADD3", "VIADD", "ISETP", "SEL", "PRMT", "LEA", "IABS", "PLOP3", "VIMNMX", "IMNMX", "SGXT", "BMSK",
            "R2P", "P2R", "FSEL", "FMNMX", "IADD", "MOV", "VABSDIFF", "LOP", "SHL", "SHR", "VOTE"}
FMA_PIPE = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "HADD2", "HMUL2", "IDP"}
LSU_PIPE = {"LDS", "STS", "LDG", "STG", "LD", "ST", "LDL", "STL", "ATOMS", "ATOMG", "RED", "LDSM", "LDC"}


def clean_kernel_name(name):
    name = name.split("(psa::")[0].split("(")[0] if name.startswith("void") is False else name
    name = name.replace("void ", "").replace("psa::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    # drop the parameter list, keep the template arguments; C-style casts "(int)10" -> "10"
    depth, out = 0, []
    for ch in name:
        if ch == "<":
            depth += 1
        if ch == "(" and depth == 0:
            break
        out.append(ch)
        if ch == ">":
            depth -= 1
    name = "".join(out)
    for cast in ("(int)", "(bool)"):
        name = name.replace(cast, "")
    return name.strip()


def f(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def main(argv):
    path = argv[0]
    json_out = None
    if "--json" in argv:
        k = argv.index("--json")
        json_out = (argv[k + 1], argv[k + 2])
        argv = argv[:k] + argv[k + 3:]
    min_share = float(argv[1]) if len(argv) > 1 else 0.4
    fh = io.TextIOWrapper(gzip.open(path), newline="") if path.endswith(".gz") else open(path, newline="")
    rows = list(csv.reader(fh))
    hi = next(k for k, r in enumerate(rows) if "Instructions Executed" in r)
    h = rows[hi]
    ci, cs, cn = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
    stall = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    data = [r for r in rows[hi + 1:] if len(r) > cn]
    tot_i = sum(f(r[ci]) for r in data)
    tot_s = sum(f(r[cn]) for r in data) or 1.0
    print("kernel:", rows[0][1] if len(rows[0]) > 1 else "?")
    print("SASS instructions %d, executed warp-instructions %.0f, samples %.0f" % (len(data), tot_i, tot_s))
    runs, cur = [], None
    for k, r in enumerate(data):
        v = f(r[ci])
        if cur and abs(v - cur[2]) <= 0.02 * max(v, cur[2], 1.0):
            cur[1] = k
        else:
            if cur:
                runs.append(cur)
            cur = [k, k, v]
    runs.append(cur)
    if json_out:
        a, b, v = max(runs, key=lambda r: sum(f(x[ci]) for x in data[r[0]:r[1] + 1]))
        pipes = {"alu": 0, "fma": 0, "lsu": 0, "other": 0}
        ops = {}
        for r in data[a:b + 1]:
            t = r[cs].split()
            op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else t[0] if t else "?").split(".")[0]
            ops[op] = ops.get(op, 0) + 1
            pipes["alu" if op in ALU_PIPE else "fma" if op in FMA_PIPE else "lsu" if op in LSU_PIPE else "other"] += 1
        rec = {"sass_first": a, "sass_last": b, "sass_instr": b - a + 1, "exec_per_instr": v,
               "share_of_executed_pct": 100 * sum(f(x[ci]) for x in data[a:b + 1]) / tot_i,
               "alu_pipe_instr": pipes["alu"], "fma_pipe_instr": pipes["fma"], "lsu_instr": pipes["lsu"], "other_instr": pipes["other"],
               "ops": dict(sorted(ops.items(), key=lambda kv: -kv[1])[:8]), "steps_per_group": 32}
        summary = json.load(open(json_out[0])) if os.path.exists(json_out[0]) else {}
        kname = clean_kernel_name(rows[0][1] if len(rows[0]) > 1 else "?")
        summary.setdefault(json_out[1], {}).setdefault(kname, {})["inner_loop"] = rec
        summary[json_out[1]][kname]["executed_warp_instructions"] = tot_i
        json.dump(summary, open(json_out[0], "w"), indent=1, sort_keys=True)
    # merge tiny neighbours into "other"
    other_i = other_s = 0.0
    for a, b, v in runs:
        seg = data[a:b + 1]
        si, ss = sum(f(r[ci]) for r in seg), sum(f(r[cn]) for r in seg)
        if 100 * si / tot_i < min_share and 100 * ss / tot_s < min_share:
            other_i += si
            other_s += ss
            continue
        ops = {}
        for r in seg:
            t = r[cs].split()
            op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else t[0] if t else "?").split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        st = sorted(((sum(f(r[i]) for r in seg), c) for i, c in stall), reverse=True)[:4]
        print("sass %5d-%5d  n=%4d  exec/instr=%8.0f  instr %5.1f%%  samples %5.1f%%  ops %s  stalls %s" % (
            a, b, b - a + 1, v, 100 * si / tot_i, 100 * ss / tot_s,
            ",".join("%s:%d" % kv for kv in sorted(ops.items(), key=lambda kv: -kv[1])[:5]),
            ",".join("%s:%d" % (c.replace("stall_", ""), n) for n, c in st if n > 0)))
    print("other (runs below %.1f%%): instr %.1f%%  samples %.1f%%" % (min_share, 100 * other_i / tot_i, 100 * other_s / tot_s))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
