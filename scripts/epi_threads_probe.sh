#!/bin/bash
for t in 64 128 256; do
  echo "COLO_EPI_THREADS=$t"; COLO_EPI_THREADS=$t timeout 120 python scripts/psrl_parts_probe.py 2>&1 | grep -o "^[a-z0-9_]* *loops=[0-9]*\|episodic VI *[0-9.]* us" | paste - -
done
