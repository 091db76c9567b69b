#!/bin/bash
# The two captures refresh_profiles.sh does not cover: config 4's sliced scan, and config 3 on the plain long-mode scan
# (packed mode switched off) for comparison with k_scan_packed.
set -u
out=gpurun_out/profiles; mkdir -p $out
[ -f $out/r01_summary.json ] || { [ -f profiles/r01_summary.json ] && cp profiles/r01_summary.json $out/; }   # keep what refresh_profiles.sh just wrote
cap() {   # workload, extra bench args, kernel regex, file stem
  B="python bench.py --workload $1 --steps 3 --warmup 3 --no-cpu-baseline --no-others --no-strong $2"
  $B > /dev/null 2>&1 || { echo "plain run of $1 failed"; return; }
  ncu --set full --import-source on --clock-control none -k regex:$3 -s 3 -c 1 -f -o /tmp/$4 $B > gpurun_out/ncu_$4.log 2>&1
  python tools/ncu_summary.py --out $out /tmp/$4.ncu-rep $1 >> gpurun_out/sum.log 2>&1
  ncu -i /tmp/$4.ncu-rep --page source --csv > /tmp/$4_src.csv 2>/dev/null
  python tools/ncu_regions.py /tmp/$4_src.csv 0.5 > $out/$4_regions.txt 2>&1
}
cap c4 "" 'k_scan' r01_c4_k_scan_slices
cap c3 "--opt pack_queries=0" 'k_scan' r01_c3_k_scan
ls -la $out
