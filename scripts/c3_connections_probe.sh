#!/bin/bash
# C3 leg: instances/s against the number of worker threads (one stream each) and CUDA_DEVICE_MAX_CONNECTIONS (the number of
# hardware queues the streams of a context are spread over; default 8); three runs per configuration
cd "$(dirname "$0")/.."
N=${1:-256}
CFGS=${2:-"8:8 32:8 32:12 32:24"}
for rep in 1 2 3; do
for cfg in $CFGS; do
  set -- ${cfg/:/ }
  CUDA_DEVICE_MAX_CONNECTIONS=$1 timeout 120 python bench.py --workload c3 --c3-instances $N --c3-workers $2 2>/tmp/c3_probe.err > /tmp/c3_probe.json || tail -5 /tmp/c3_probe.err
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open("/tmp/c3_probe.json"))
    print(f"connections {sys.argv[1]} workers {sys.argv[2]}: {d['value']:.1f} inst/s parity_ok={d['parity_ok']}", d["config"]["rank0_seconds"], flush=True)
except Exception as e:
    print("failed", sys.argv[1:], e)
PY
done
done
