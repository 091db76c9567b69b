#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` export by source line: executed warp-instructions per line of our code.

    ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_source_hot.py src.csv [top]"""
import csv
import sys
from collections import defaultdict


def main(argv):
    rows = list(csv.reader(open(argv[0], newline="")))
    top = int(argv[1]) if len(argv) > 1 else 40
    hdr = None
    for k, r in enumerate(rows):
        if any("Instructions Executed" in c for c in r):
            hdr = k
            break
    if hdr is None:
        print("no 'Instructions Executed' column; columns:", rows[0][:20])
        return 1
    h = rows[hdr]
    ci = next(i for i, c in enumerate(h) if c.strip() == "Instructions Executed" or c.strip() == "# Instructions Executed" or "Warp Instructions Executed" in c or c.strip().startswith("Instructions Executed"))
    cs = next((i for i, c in enumerate(h) if c.strip() in ("Source", "Source Location", "Address Space")), None)
    cl = next((i for i, c in enumerate(h) if "Source" in c and "Loc" in c), cs)
    print("columns:", h)
    agg = defaultdict(float)
    total = 0.0
    for r in rows[hdr + 1:]:
        if len(r) <= ci:
            continue
        try:
            v = float(r[ci].replace(",", ""))
        except ValueError:
            continue
        key = r[cl] if cl is not None and cl < len(r) else "?"
        agg[key] += v
        total += v
    print("total executed warp-instructions: %.0f" % total)
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
        print("%12.0f %5.1f%%  %s" % (v, 100 * v / total, key))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
