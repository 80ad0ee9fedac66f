#!/bin/bash
echo "== PARTITIONED (plain tile kernel)"
timeout 120 python tools/steptime.py 1.0 prof 2>&1 | grep -E "^\{|iter 3"
echo "== PARTITIONED TMA"
GSQL_JOIN_TMA=1 timeout 120 python tools/steptime.py 1.0 prof 2>&1 | grep -E "^\{|iter 3"
echo "== UNPARTITIONED"
GSQL_JOIN_PART_BYTES=1099511627776 timeout 120 python tools/steptime.py 1.0 prof 2>&1 | grep -E "^\{|iter 3"
