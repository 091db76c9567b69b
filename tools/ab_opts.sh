#!/bin/bash
# A/B of library options on ONE box: tools/ab_opts.sh <workload> "<opts A>" "<opts B>" ...   (opts: "--opt name=value ..." or "")
w=$1; shift
for round in 1 2 3; do
  for o in "$@"; do
    python bench.py --workload $w --steps 40 --warmup 5 --no-cpu-baseline --no-others --no-strong $o 2>/dev/null | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$round [$o] $w ms=%.4f e2e_ms=%.4f kern_ms=%.4f' % (d['ms_per_step'], d['e2e']['ms_per_step'], r['kernel_ms']))"
  done
done
