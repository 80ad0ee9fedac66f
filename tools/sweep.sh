#!/bin/bash
cd /root/repo
run() { echo "== $*"; env "$@" timeout 300 python bench.py --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['clocks']['samples'], sum(d['roofline']['per_kernel_ms_per_step'].values()))"; }
run GSQL_BENCH_STEP_SYNC=consume
run GSQL_BENCH_STEP_SYNC=close
