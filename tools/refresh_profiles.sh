#!/bin/bash
# Regenerate the ncu evidence of this round on a GPU box (run through gpurun; outputs land in gpurun_out/profiles, copy what is
# to be judged into profiles/).  Every command runs once WITHOUT ncu first and must exit 0 (B200_PROFILING.md).
# Under ncu a kernel is profiled inside its launch call, so nothing the kernel or its stream waits for may be released by
# the host AFTER that call: the timed runs' gate and the streamed queries are switched off for these runs.
# PSA_REFRESH="c3 c5" limits the captures to those workloads' dominant kernels.
set -u
R=${PSA_ROUND:-r02}
out=gpurun_out/profiles; mkdir -p $out
for w in c3 c5 c4 c1; do
  B="python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-others --no-strong --opt gate_timed_runs=0 --opt stream_queries=0"
  $B > gpurun_out/plain_$w.log 2>&1 || { echo "plain run of $w failed"; tail -3 gpurun_out/plain_$w.log; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/${R}_${w}_launches.csv $B > gpurun_out/ncu_l$w.log 2>&1
done
want=${PSA_REFRESH:-c3 c5 c1 c4}
case " $want " in *" c3 "*) tools/ncu_cap.sh c3 k_stripe ${R}_c3_k_stripe > gpurun_out/cap_c3.log 2>&1;; esac
case " $want " in *" c5 "*) tools/ncu_cap.sh c5 k_stripe ${R}_c5_k_stripe > gpurun_out/cap_c5.log 2>&1;; esac
case " $want " in *" c1 "*) tools/ncu_cap.sh c1 k_single ${R}_c1_k_single > gpurun_out/cap_c1.log 2>&1;; esac
case " $want " in *" c4 "*) tools/ncu_cap.sh c4 'k_scan<' ${R}_c4_k_scan_slices > gpurun_out/cap_c4.log 2>&1
                            tools/ncu_cap.sh c4 k_finish ${R}_c4_k_finish > gpurun_out/cap_c4f.log 2>&1;; esac
ls -la $out
