#!/bin/bash
# Regenerate the ncu evidence under profiles/ on a GPU box (run through gpurun; outputs land in gpurun_out/profiles).
# Every command runs once WITHOUT ncu first and must exit 0 (B200_PROFILING.md); .ncu-rep files stay in /tmp.
set -u
out=gpurun_out/profiles; mkdir -p $out
[ -f profiles/r01_summary.json ] && cp profiles/r01_summary.json $out/
for w in c3 c5 c4 c1; do
  B="python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-others --no-strong"
  $B > gpurun_out/plain_$w.log 2>&1 || { echo "plain run of $w failed"; tail -3 gpurun_out/plain_$w.log; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/r01_${w}_launches.csv $B > gpurun_out/ncu_l$w.log 2>&1
done
cap() {   # workload, kernel regex, file stem
  B="python bench.py --workload $1 --steps 3 --warmup 3 --no-cpu-baseline --no-others --no-strong"
  $B > /dev/null 2>&1 || { echo "plain run of $1 failed"; return; }
  ncu --set full --import-source on --clock-control none -k regex:$2 -s 3 -c 1 -f -o /tmp/$3 $B > gpurun_out/ncu_$3.log 2>&1
  python tools/ncu_summary.py --out $out /tmp/$3.ncu-rep $1 >> gpurun_out/sum.log 2>&1
  ncu -i /tmp/$3.ncu-rep --page source --csv > /tmp/$3_src.csv 2>/dev/null
  python tools/ncu_regions.py /tmp/$3_src.csv 0.5 > $out/$3_regions.txt 2>&1
}
cap c3 'k_scan_packed' r01_c3_k_scan_packed
cap c5 'k_scan_batch' r01_c5_k_scan_batch
cap c4 '^.*k_scan<' r01_c4_k_scan_slices
ls -la $out
