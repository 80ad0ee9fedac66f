#!/bin/bash
# parameter sweep of the fast join on the full C2 shape; prints per-kernel ms (2 profiled iterations)
for tma in 0 1; do
for pb in 8 16 32 64; do
for sb in 268435456 1073741824; do
  echo "== NO_TMA=$tma PART_MB=$pb SUB=$sb"
  GSQL_JOIN_NO_TMA=$tma GSQL_JOIN_PART_BYTES=$((pb<<20)) GSQL_JOIN_SUB_BATCH=$sb timeout 120 python tools/steptime.py 1.0 prof 2>&1 | grep -E "^\{|iter 3"
done; done; done
