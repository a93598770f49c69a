#!/bin/bash
set -x
I="python scripts/new_kernels_r1e.py"
$I > gpurun_out/prof_plain_r1e.log 2>&1 || { cat gpurun_out/prof_plain_r1e.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"episodic_batched_kernel|dirichlet_rows_kernel|psrl_steps_kernel|nig_rows_kernel" -s 4 -c 4 -o gpurun_out/prof_r1e_psrl $I > gpurun_out/ncu_r1e.log 2>&1
tail -n 3 gpurun_out/ncu_r1e.log; cat gpurun_out/prof_plain_r1e.log
