#!/bin/bash
# round-2 ncu captures of the SHIPPED kernels at the bench shapes (one gpurun call; every profiled command first exits 0
# without ncu).  Brings back gpurun_out/r2_*.ncu-rep; scripts/summarize_ncu_r2.py turns them into profiles/r2_*.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip, count, args...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  local t0=$SECONDS
  timeout 120 python scripts/ncu_targets_r2.py "$@" > gpurun_out/r2_plain_$name.log 2>&1 &&
  timeout 240 ncu --set full --clock-control none -k regex:$rx -s $skip -c $cnt -o gpurun_out/r2_$name -f \
      python scripts/ncu_targets_r2.py "$@" > gpurun_out/r2_ncu_$name.log 2>&1
  echo "$name rc=$? $((SECONDS - t0))s"
}
cap step_65536 env_step_kary_lean 3 2 step 65536
cap step_4194304 env_step_kary_lean 3 2 step 4194304
cap backup_c4 backup_kernel 2 2 backup_c4
cap backup_c5 backup_kernel 2 2 backup_c5 16384
cap gs gs_solve_tma 0 1 gs
cap umma hitting_umma_kernel 1 2 umma
cap epi_batched episodic_batched 1 1 epi_batched
# launch list of the default bench command (shares, not absolutes)
python bench.py --steps 20 --warmup 5 --no-sweep > gpurun_out/r2_bench_plain.json 2> gpurun_out/r2_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 20 --warmup 5 --no-sweep > gpurun_out/r2_bench_ncu.json 2> gpurun_out/r2_bench_ncu.err
echo "launch list rc=$?"
ls -la gpurun_out/r2_*.ncu-rep
# only 64 MiB of gpurun_out/ travel back: summarise on the box, keep the summaries, drop the reports
python scripts/summarize_ncu_r2.py gpurun_out/r2_profiles > gpurun_out/r2_summarize.log 2>&1
rm -f gpurun_out/r2_*.ncu-rep
du -sh gpurun_out
