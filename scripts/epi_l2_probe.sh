#!/bin/bash
for mb in 24 48 64 96 160 100000; do
  echo "COLO_EPI_L2_MB=$mb"; COLO_EPI_L2_MB=$mb timeout 120 python scripts/psrl_parts_probe.py 2>&1 | grep -o "^[a-z0-9_]* *loops=[0-9]*\|episodic VI *[0-9.]* us" | paste - - 
done
