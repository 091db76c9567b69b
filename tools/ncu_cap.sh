#!/bin/bash
# One `ncu --set full` capture of one kernel of one bench workload, summarised into gpurun_out/profiles/ (run through gpurun).
#   tools/ncu_cap.sh <workload> <kernel regex> <stem> [extra bench.py options...]
# The command runs once WITHOUT ncu first and must exit 0 (B200_PROFILING.md); the .ncu-rep stays in /tmp.
set -u
w=$1; k=$2; stem=$3; shift 3
out=gpurun_out/profiles; mkdir -p $out
for f in profiles/r0*_summary.json; do [ -f "$f" ] && [ ! -f $out/$(basename $f) ] && cp $f $out/; done
B="python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-others --no-strong --opt gate_timed_runs=0 --opt stream_queries=0 $*"
$B > gpurun_out/plain_$stem.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$stem.log; exit 1; }
ncu --set full --import-source on --clock-control none -k regex:$k -s 3 -c 1 -f -o /tmp/$stem $B > gpurun_out/ncu_$stem.log 2>&1
python tools/ncu_summary.py --out $out /tmp/$stem.ncu-rep $w >> gpurun_out/sum.log 2>&1
ncu -i /tmp/$stem.ncu-rep --page source --csv > /tmp/${stem}_src.csv 2>/dev/null
python tools/ncu_regions.py /tmp/${stem}_src.csv 0.5 --json $out/${PSA_ROUND:-r02}_summary.json $w > $out/${stem}_regions.txt 2>&1
gzip -c /tmp/${stem}_src.csv > gpurun_out/${stem}_src.csv.gz
cat $out/${stem}_regions.txt
grep -E "Duration|Registers Per|Shared Memory Config|Theoretical Occ|Achieved Occ|Executed Ipc Active|Issue Slots Busy|ALU|Warp Cycles Per Issued|One or More Eligible|No Eligible" $out/${stem}_details.txt | head -30
