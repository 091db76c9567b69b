#!/bin/bash
# A/B of library variants on ONE box: tools/ab_bench.sh <dir with *.so variants> [workloads...]
# Each variant is copied over the in-tree libpsa_b200.so (the box copy is scratch) and benched; two rounds, alternating.
dir=$1; shift
wls=${@:-c3 c5 c4}
pkg=parallel-sequence-alignment_b200
for round in 1 2; do
  for so in $dir/*.so; do
    cp $so $pkg/libpsa_b200.so
    for w in $wls; do
      python bench.py --workload $w --steps 40 --warmup 5 --no-cpu-baseline --no-others --no-strong 2>/dev/null | tail -1 | \
        python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$round $(basename $so) $w ms=%.4f e2e=%.4g kern_us=%s frac=%.3f' % (d['ms_per_step'], d['e2e']['value'], r.get('kernel_us', r.get('achieved')), r['frac']))"
    done
  done
done
