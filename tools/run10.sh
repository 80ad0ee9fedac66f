set -x
cd $GRAFT_REPO_ROOT
KPROF=agg_q1,agg_c5 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_agg_reg|k_agg_consume" -c 3 -f -o gpurun_out/r02_agg_c python tools/kprof.py > gpurun_out/kprof_c.log 2>&1
tail -3 gpurun_out/kprof_c.log
ls -la gpurun_out/*.ncu-rep
