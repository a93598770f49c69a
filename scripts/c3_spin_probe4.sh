#!/bin/bash
# C3 leg on FOUR host cores (what a rank has on an 8-GPU box with 32 cores): blocking-sync against spinning workers
cd "$(dirname "$0")/.."
for rep in 1 2; do
for cfg in "0 8" "1 8" "1 4" "1 3" "0 4" "0 12"; do
  set -- $cfg
  if [ $1 = 1 ]; then export COLO_SUITE_SPIN=1; else unset COLO_SUITE_SPIN; fi
  timeout 120 taskset -c 0-3 python bench.py --workload c3 --c3-instances 128 --c3-passes 3 --c3-workers $2 2>/tmp/c3_probe.err > /tmp/c3_probe.json || tail -5 /tmp/c3_probe.err
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open("/tmp/c3_probe.json"))
    print(f"4 cores, spin {sys.argv[1]} workers {sys.argv[2]}: {d['value']:.1f} inst/s passes {[round(x, 3) for x in d['passes_s']]}", d["config"]["rank0_seconds"]["slowest_hardness_s"], flush=True)
except Exception as e:
    print("failed", sys.argv[1:], e)
PY
done
done
