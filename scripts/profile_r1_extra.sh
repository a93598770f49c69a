#!/bin/bash
set -x
I="python scripts/extra_kernels.py"
$I > gpurun_out/prof_plain3.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"hitting_gemm|resident_solve|evi_rows|dsquare|dirichlet_rows|sparse_episodic|policy_chain" -c 9 -o gpurun_out/prof_r1b_extra $I > gpurun_out/ncu_extra.log 2>&1
tail -n 3 gpurun_out/ncu_extra.log; cat gpurun_out/prof_plain3.log
