#!/bin/bash
set -x
I="python scripts/new_kernels_r1c.py"
$I > gpurun_out/prof_plain_r1c.log 2>&1 || { cat gpurun_out/prof_plain_r1c.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"gs_sparse_kernel|qlearning_steps" -c 3 -o gpurun_out/prof_r1c_new $I > gpurun_out/ncu_r1c.log 2>&1
tail -n 3 gpurun_out/ncu_r1c.log; cat gpurun_out/prof_plain_r1c.log
