#!/bin/bash
# round-2 (second pass) ncu captures of the kernels added after scripts/profile_r2.sh ran: the batched extended VI, the
# optimistic-sampling kernel of PSRLContinuous and the batched fp64 chain squaring.  Same protocol: every profiled
# command first exits 0 without ncu; summaries are made on the box (gpurun_out/r2b_profiles), reports dropped.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2_*.ncu-rep
cap() {  # name, kernel regex, skip, count, args...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  local t0=$SECONDS
  timeout 120 python scripts/ncu_targets_r2.py "$@" > gpurun_out/r2_plain_$name.log 2>&1 &&
  timeout 240 ncu --set full --clock-control none --profile-from-start off -k regex:$rx -s $skip -c $cnt -o gpurun_out/r2_$name -f \
      python scripts/ncu_targets_r2.py "$@" > gpurun_out/r2_ncu_$name.log 2>&1
  echo "$name rc=$? $((SECONDS - t0))s"
}
cap evi_batched evi_batched_kernel 1 1 evi_batched
cap psrlc_sample psrlc_sample_kernel 1 1 psrlc_sample
cap avg_rewards dsquare_kernel 6 1 avg_rewards
python scripts/summarize_ncu_r2.py gpurun_out/r2b_profiles > gpurun_out/r2b_summarize.log 2>&1
rm -f gpurun_out/r2_*.ncu-rep
ls gpurun_out/r2b_profiles
