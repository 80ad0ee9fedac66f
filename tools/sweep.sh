#!/bin/bash
for tma in 0 1; do
  echo "== PARTITIONED NO_TMA=$tma"
  GSQL_JOIN_NO_TMA=$tma timeout 120 python tools/steptime.py 1.0 prof 2>&1 | grep -E "^\{|iter 3"
done
echo "== UNPARTITIONED"
GSQL_JOIN_PART_BYTES=1099511627776 timeout 120 python tools/steptime.py 1.0 prof 2>&1 | grep -E "^\{|iter 3"
