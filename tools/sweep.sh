#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "join" 2>&1 | tail -5
run() { echo "== $*"; env "$@" timeout 120 python tools/steptime.py 1.0 prof 2>&1 | grep -E "^\{|iter 3" | sed -e "s/'join_fast_//g"; }
run GSQL_X=1
run GSQL_JOIN_BUILD_GROUP_BYTES=8388608
run GSQL_JOIN_PART_BYTES=1099511627776
