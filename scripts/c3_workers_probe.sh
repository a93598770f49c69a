#!/bin/bash
# C3 leg: instances/s against the runner (Python threads vs colo_suite_run's C++ threads) and the number of workers
cd "$(dirname "$0")/.."
for cfg in "python 4" "native 1" "native 2" "native 4" "native 8" "native 16" "native 32"; do
  set -- $cfg
  timeout 120 python bench.py --workload c3 --c3-instances 160 --c3-runner $1 --c3-workers $2 2>/tmp/c3_probe.err > /tmp/c3_probe.json || tail -5 /tmp/c3_probe.err
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open("/tmp/c3_probe.json"))
    print(f"runner {sys.argv[1]} workers {sys.argv[2]}: {d['value']:.1f} inst/s parity_ok={d['parity_ok']}", d["config"]["rank0_seconds"], flush=True)
except Exception as e:
    print("failed", sys.argv[1:], e)
PY
done
