#!/bin/bash
# round-2 ncu captures (one gpurun call; every profiled command first exits 0 without ncu).  Brings back
# gpurun_out/r2_*.ncu-rep; scripts/summarize_ncu.py turns them into profiles/r2_*.csv here.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python scripts/step_sweep_probe.py > gpurun_out/r2_step_sweep_probe.log 2>&1
for N in 65536 4194304; do
  python scripts/step_sweep_probe.py --ncu $N > gpurun_out/r2_plain_step_$N.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:env_step_dense_kary -s 3 -c 2 \
      -o gpurun_out/r2_step_$N -f python scripts/step_sweep_probe.py --ncu $N > gpurun_out/r2_ncu_step_$N.log 2>&1
done
