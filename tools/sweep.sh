#!/bin/bash
for tma in 0 1; do
for spr in 2 3 4; do
  echo "== NO_TMA=$tma SLOTS_PER_ROW=$spr"
  GSQL_JOIN_NO_TMA=$tma GSQL_JOIN_SLOTS_PER_ROW=$spr timeout 120 python tools/steptime.py 1.0 prof 2>&1 | grep -E "^\{|iter 3"
done; done
