#!/usr/bin/env python
"""Split an `ncu --page source --csv` export (optionally .gz) into runs of equal execution count (= code regions:
prologue, unrolled loop bodies, epilogue ...) and print each run's share of executed instructions and of the
warp-state samples, with its dominant stall reasons.

    python tools/ncu_regions.py gpurun_out/c3p_source.csv.gz [min_share_pct]"""
import csv
import gzip
import io
import sys


def f(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def main(argv):
    path = argv[0]
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
