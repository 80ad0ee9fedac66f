set -x
cd $GRAFT_REPO_ROOT
KPROF=push,scan,agg_q1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_xchg_push|k_scan|k_agg_reg" -c 6 -f -o gpurun_out/r02_kernels_b python tools/kprof.py > gpurun_out/kprof.log 2>&1
tail -5 gpurun_out/kprof.log
