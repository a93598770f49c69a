#!/bin/bash
# C3 leg: blocking-sync workers (default) against spinning ones (COLO_SUITE_SPIN=1), three runs each, passes listed
cd "$(dirname "$0")/.."
N=${1:-128}
for rep in 1 2 3; do
for spin in 0 1; do
  if [ $spin = 1 ]; then export COLO_SUITE_SPIN=1; else unset COLO_SUITE_SPIN; fi
  timeout 120 python bench.py --workload c3 --c3-instances $N --c3-passes 3 2>/tmp/c3_probe.err > /tmp/c3_probe.json || tail -5 /tmp/c3_probe.err
  python - "$spin" <<'PY'
import json, sys
try:
    d = json.load(open("/tmp/c3_probe.json"))
    print(f"spin {sys.argv[1]}: {d['value']:.1f} inst/s passes {[round(x, 3) for x in d['passes_s']]} parity_ok={d['parity_ok']}", d["config"]["rank0_seconds"], flush=True)
except Exception as e:
    print("failed", sys.argv[1:], e)
PY
done
done
