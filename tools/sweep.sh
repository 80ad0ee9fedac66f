#!/bin/bash
run() { echo "== $*"; env "$@" timeout 120 python tools/steptime.py 1.0 prof 2>&1 | grep -E "^\{|iter 3" | sed -e "s/'join_fast_//g"; }
run GSQL_L2_FETCH_GRANULARITY=32
run GSQL_L2_FETCH_GRANULARITY=64
run GSQL_L2_FETCH_GRANULARITY=128
run GSQL_L2_FETCH_GRANULARITY=32 GSQL_JOIN_SLOTS_PER_ROW=2
run GSQL_L2_FETCH_GRANULARITY=32 GSQL_JOIN_PART_BYTES=16777216
