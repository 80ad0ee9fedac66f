#!/bin/bash
cd /root/repo
run() { echo "== $*"; env "$@" timeout 120 python tools/steptime.py 1.0 prof 2>&1 | grep -E "^\{|iter 3" | sed -e "s/'join_fast_//g"; }
run GSQL_X=1
run GSQL_JOIN_PROBE_PIPE=1 GSQL_JOIN_LOOKUP_MODE=2
