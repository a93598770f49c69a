#!/bin/bash
# ncu evidence for the k-ary step kernel (1 GPU): launch list of a short step-only bench, then one full capture.
set -x
B="python bench.py --workload step --steps 20 --warmup 3 --cpu-seconds 0.5"
$B > gpurun_out/prof_plain_r1d.log 2>&1 || { tail -5 gpurun_out/prof_plain_r1d.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv $B > gpurun_out/ncu_l_r1d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:env_step_dense_kary -s 30 -c 1 -o gpurun_out/prof_r1d_step $B > gpurun_out/ncu_s_r1d.log 2>&1
tail -2 gpurun_out/ncu_l_r1d.log gpurun_out/ncu_s_r1d.log
